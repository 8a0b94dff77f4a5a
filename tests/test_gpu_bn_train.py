"""BATCH_NORM=true training (models.py:30-33, 49-52 with training=True): the layer-by-layer fp32 path of csrc/bn_train.cu
against the oracle's batch-statistics forward + autograd (stop-gradient on the fine sample positions)."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle as O
from oracle import models_ref as MR

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nk():
    import nerf_keras_b200 as nk
    return nk


def _setup(seed=0, B=96, Nc=16, Nf=32):
    rng = np.random.default_rng(seed)
    wc, wf = O.init_weights(seed=seed + 1, bias_range=0.1), O.init_weights(seed=seed + 2, bias_range=0.1)
    bns = []
    for _ in range(2):
        bn = MR.init_bn()
        for st in bn.values():
            n = st["gamma"].numel()
            st["gamma"] = torch.from_numpy(rng.uniform(0.7, 1.3, n).astype(np.float32))
            st["beta"] = torch.from_numpy(rng.uniform(-0.2, 0.2, n).astype(np.float32))
            st["mean"] = torch.from_numpy(rng.uniform(-0.1, 0.1, n).astype(np.float32))
            st["var"] = torch.from_numpy(rng.uniform(0.5, 1.5, n).astype(np.float32))
        bns.append(bn)
    o, d = O.get_rays(12, 12, 15.0, torch.from_numpy(np.asarray(O.pose_spherical(25.0, -35.0, 4.0))))
    o, d = o.reshape(-1, 3)[:B].contiguous(), d.reshape(-1, 3)[:B].contiguous()
    t = O.generate_t_vals(2.0, 6.0, B, Nc, True, u=torch.from_numpy(rng.random(Nc, dtype=np.float32)))
    u = torch.from_numpy(rng.random((B, Nf), dtype=np.float32))
    img = torch.from_numpy(rng.random((B, 3), dtype=np.float32))
    return wc, wf, bns, o, d, t, u, img, Nc, Nf


def _oracle_grads(wc, wf, bns, o, d, t, u, img, Nf):
    bns = [{r: {k: v.clone() for k, v in st.items()} for r, st in bn.items()} for bn in bns]
    params = MR._params(wc) + MR._params(wf)
    bnp = [st[k] for bn in bns for st in bn.values() for k in ("gamma", "beta")]
    for p in params + bnp:
        p.requires_grad_(True)
    rgbs = O.forward_pass(wc, wf, o, d, t, 10, 4, Nf, u, training=True, stop_grad_samples=True, bn_coarse=bns[0],
                          bn_fine=bns[1])[0]
    loss_c, loss_f = MR.mse(img, rgbs[0]), MR.mse(img, rgbs[1])
    grads = torch.autograd.grad(loss_c + loss_f, params + bnp)
    for p in params + bnp:
        p.requires_grad_(False)
    n = len(params)
    return [g.detach() for g in grads[:n]], [g.detach() for g in grads[n:]], float(loss_c.detach()), float(loss_f.detach()), bns


def _flat_bn(bn):
    roles = list(bn.keys())
    return np.concatenate([np.concatenate([bn[r][k].detach().numpy() for r in roles]) for k in ("gamma", "beta", "mean", "var")])


def test_bn_forward_backward_matches_oracle_autograd(nk):
    from nerf_keras_b200 import _lib
    wc, wf, bns, o, d, t, u, img, Nc, Nf = _setup()
    g_w, g_bn, loss_c, loss_f, bns_after = _oracle_grads(wc, wf, bns, o, d, t, u, img, Nf)
    L = _lib.lib()
    B = o.shape[0]
    cfg = _lib.NerfConfig(8, 256, 4, 10, 4, Nc, Nf, B, 0, 0, 5e-4, 1)   # the batch_norm flag of nerf_create stays 0
    n, nbn = int(L.nerf_param_count(C.byref(cfg))), int(L.nerf_bn_param_count(C.byref(cfg)))
    assert nbn == 8 * 256 + 128
    params = torch.from_numpy(np.concatenate([O.flatten_weights(wc), O.flatten_weights(wf)])).cuda()
    bn = torch.from_numpy(np.concatenate([_flat_bn(bns[0]), _flat_bn(bns[1])]).astype(np.float32)).cuda()
    grads, bng = torch.empty(2 * n, device="cuda"), torch.empty(4 * nbn, device="cuda")
    metrics = torch.empty(3, device="cuda")
    ws = torch.empty(int(L.nerf_bn_workspace_bytes(C.byref(cfg), B)), dtype=torch.uint8, device="cuda")
    dev = lambda x: x.cuda().contiguous()
    io = [dev(x) for x in (img, o, d, t, u)]
    _lib.check(L.nerf_bn_forward_backward(C.byref(cfg), params.data_ptr(), bn.data_ptr(), *[x.data_ptr() for x in io], B,
                                          grads.data_ptr(), bng.data_ptr(), metrics.data_ptr(), ws.data_ptr(), ws.numel(),
                                          torch.cuda.current_stream().cuda_stream), "bn fb")
    torch.cuda.synchronize()
    m = metrics.cpu().numpy()
    assert abs(m[0] - loss_c) <= 2e-5 and abs(m[1] - loss_f) <= 2e-5
    ref = np.concatenate([g.reshape(-1).numpy() for g in g_w])
    got = grads.cpu().numpy()
    cos = float(np.dot(ref, got) / (np.linalg.norm(ref) * np.linalg.norm(got)))
    print('BN gradient vs oracle: cosine', cos, 'norm ratio', np.linalg.norm(got) / np.linalg.norm(ref))
    assert cos >= 0.9995 and abs(np.linalg.norm(got) / np.linalg.norm(ref) - 1) <= 3e-3, cos
    # Differences are single ReLU decisions: the forward GEMMs (gemm_tc.cu, three-way split bf16 operands on the tensor
    # cores) reproduce a pre-activation to ~1e-6 of its scale, the CPU BLAS of the oracle sums in another order, and among
    # 3.3 M (sample, unit) entries per net a handful sit that close to the threshold (overall cosine 0.9998; 0.9992 with a
    # two-way split, 0.9999 with fp32 FMAs).  One flipped entry changes ONE column of that layer's dW by a few per cent (observed: unit 218 of d6 off by
    # 2.2 %, every other column exact) and shows up as ~0.1-1 % noise in every layer below; the heads and the layers above the
    # first flip agree to 1e-5.  Hence: tight overall direction and norm, per-tensor relative L2 loose enough for a few flips.
    off, worst = 0, []
    for net in ("c", "f"):
        for role, fi, fo in MR.layer_shapes():
            for kind, k in (("W", fi * fo), ("b", fo)):
                a_, b_ = got[off:off + k], ref[off:off + k]
                off += k
                if np.linalg.norm(b_) <= 1e-6 * np.linalg.norm(ref):     # biases in front of a batch norm: it cancels them
                    assert np.linalg.norm(a_) <= 1e-5 * np.linalg.norm(ref)
                    continue
                worst.append((float(np.linalg.norm(a_ - b_) / np.linalg.norm(b_)), net + "/" + role + "/" + kind))
    by_name = {nm: e for e, nm in worst}
    assert max(by_name["c/rgb/W"], by_name["c/sigma/W"], by_name["c/rgb/b"]) <= 1e-4       # above every batch norm: no flips possible
    assert max(e for e, _ in worst) <= 8e-2, sorted(worst, reverse=True)[:5]
    # gamma / beta gradients: oracle order is per layer (gamma, beta); ours is [all gamma | all beta] per net
    roles = list(bns[0].keys())
    k = 0
    for net in range(2):
        gam = np.concatenate([g_bn[k + 2 * i].numpy() for i in range(len(roles))])
        bet = np.concatenate([g_bn[k + 2 * i + 1].numpy() for i in range(len(roles))])
        k += 2 * len(roles)
        mine = bng.cpu().numpy()[net * 2 * nbn:(net + 1) * 2 * nbn]
        scale = max(np.abs(gam).max(), np.abs(bet).max())
        print('BN gamma / beta gradient rel. L2 error', np.linalg.norm(mine[:nbn] - gam) / np.linalg.norm(gam), np.linalg.norm(mine[nbn:] - bet) / np.linalg.norm(bet))
        assert np.linalg.norm(mine[:nbn] - gam) <= 3e-2 * np.linalg.norm(gam) and np.linalg.norm(mine[nbn:] - bet) <= 3e-2 * np.linalg.norm(bet)
        assert np.abs(mine[:nbn] - gam).max() <= 0.2 * scale and np.abs(mine[nbn:] - bet).max() <= 0.2 * scale
    # moving statistics (momentum 0.99) updated in place like Keras does in training mode
    after = bn.cpu().numpy()
    for net in range(2):
        want = _flat_bn(bns_after[net])
        np.testing.assert_allclose(after[net * 4 * nbn + 2 * nbn:(net + 1) * 4 * nbn], want[2 * nbn:], rtol=2e-4, atol=2e-5)
        np.testing.assert_array_equal(after[net * 4 * nbn:net * 4 * nbn + 2 * nbn], want[:2 * nbn])   # gamma, beta untouched


def test_bn_trainer_trains_and_renders(nk, tmp_path):
    wc, wf, bns, o, d, t, u, img, Nc, Nf = _setup(seed=3)
    to_np = lambda bn: {r: {k: v.numpy() for k, v in st.items()} for r, st in bn.items()}
    mc = nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=True)
    mf = nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=True)
    mc.set_flat_weights(O.flatten_weights(wc)); mf.set_flat_weights(O.flatten_weights(wf))
    mc.set_bn_params(to_np(bns[0])); mf.set_bn_params(to_np(bns[1]))
    tr = nk.NeRFTrainer(mc, mf, o.shape[0], Nc, Nf, 10, 4)
    tr.compile(nk.Adam(5e-4), nk.MeanSquaredError())
    batch = (img.cuda(), (o.cuda(), d.cuda(), t.cuda()))
    # the first step's metrics equal the oracle's training-mode forward
    _, _, loss_c, loss_f, _ = _oracle_grads(wc, wf, bns, o, d, t, u, img, Nf)
    first = tr.train_step(batch, u_pdf=u.cuda())
    assert abs(float(first["loss"]) - loss_f) <= 2e-5 and abs(float(first["loss_coarse"]) - loss_c) <= 2e-5
    tr.reset_metrics()
    losses = [float(tr.train_step(batch, u_pdf=u.cuda())["loss"]) for _ in range(40)]
    tr.reset_metrics()
    last = float(tr.train_step(batch, u_pdf=u.cuda())["loss"])
    assert np.isfinite(losses).all() and last < loss_f * 0.7                      # it learns the batch
    # rendering after training uses the folded, UPDATED parameters (moving statistics, not batch statistics)
    tr.reset_metrics()
    val = tr.test_step(batch, u_pdf=u.cuda())
    assert np.isfinite(float(val["loss"]))
    w1 = mc.get_weights()["d3"]["W"]
    assert np.abs(w1 - wc["d3"]["W"].numpy()).max() > 1e-4                          # weights moved
    assert np.abs(mc.get_bn_params()["d3"]["mean"] - bns[0]["d3"]["mean"].numpy()).max() > 1e-5   # so did the statistics
    path = str(tmp_path / "bn_trained.npz")
    tr.save_weights(path)
    m2c = nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=True); m2f = nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=True)
    tr2 = nk.NeRFTrainer(m2c, m2f, o.shape[0], Nc, Nf, 10, 4); tr2.build(); tr2.load_weights(path)
    a = tr.forward_pass(o.cuda(), d.cuda(), t.cuda(), u_pdf=u.cuda())[0][1]
    b = tr2.forward_pass(o.cuda(), d.cuda(), t.cuda(), u_pdf=u.cuda())[0][1]
    assert torch.equal(a, b)


def test_bn_trainer_state_follows_weight_setters_and_getters(nk, tmp_path):
    """with BATCH_NORM=true the training weights live in a device-side training state.  Setters called
    after compile() (load_weights, set_weights, set_bn_params) must become the starting point of the next step, and
    getters / model calls must see the trained values without an explicit save."""
    wc, wf, bns, o, d, t, u, img, Nc, Nf = _setup(seed=5)
    to_np = lambda bn: {r: {k: v.numpy() for k, v in st.items()} for r, st in bn.items()}

    def make(wc_, wf_):
        mc = nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=True)
        mf = nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=True)
        mc.set_flat_weights(O.flatten_weights(wc_)); mf.set_flat_weights(O.flatten_weights(wf_))
        mc.set_bn_params(to_np(bns[0])); mf.set_bn_params(to_np(bns[1]))
        tr = nk.NeRFTrainer(mc, mf, o.shape[0], Nc, Nf, 10, 4, stop_grad_samples=True)
        tr.compile(nk.Adam(5e-4), nk.MeanSquaredError())
        return tr

    batch = (img.cuda(), (o.cuda(), d.cuda(), t.cuda()))
    a = make(wc, wf)
    path = str(tmp_path / "bn_ckpt.npz")
    a.save_weights(path)
    first_a = float(a.train_step(batch, u_pdf=u.cuda())["loss"])
    # a trainer compiled on DIFFERENT weights, then handed the checkpoint: its first step must equal a's first step
    other_c, other_f = O.init_weights(991, 0.1), O.init_weights(992, 0.1)
    b = make(other_c, other_f)
    b.load_weights(path)
    first_b = float(b.train_step(batch, u_pdf=u.cuda())["loss"])
    assert abs(first_a - first_b) <= 1e-6 * max(1.0, abs(first_a)), (first_a, first_b)
    # ... and the update started from the loaded weights (not from the compile-time ones)
    wa, wb = a.coarse_model.get_flat_weights(), b.coarse_model.get_flat_weights()       # getters sync the training state
    assert np.abs(wa - wb).max() <= 1e-6
    assert np.abs(wa - O.flatten_weights(wc)).max() > 1e-5                               # trained values, not the stale host copy
    ga, gb = a.coarse_model.get_bn_params(), b.coarse_model.get_bn_params()
    assert all(np.allclose(ga[r][k], gb[r][k], atol=1e-6) for r in ga for k in ga[r])
    # a model call after training uses the trained, folded weights
    enc_x, enc_d = np.zeros((3, 63), np.float32), np.zeros((3, 27), np.float32)
    ya, yb = a.coarse_model([enc_x, enc_d]).cpu().numpy(), b.coarse_model([enc_x, enc_d]).cpu().numpy()
    assert np.abs(ya - yb).max() <= 1e-5
