"""GPU: edge cases of the public surface -- empty and ragged batches, small / odd sample counts, batches larger than
the workspace, error behaviour (same exception types as the reference: TypeError / ValueError / KeyError)."""
import numpy as np
import pytest
import torch

import oracle as O
from oracle import models_ref as MR
from tests.util import golden_weights, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nk():
    import nerf_keras_b200 as nk
    return nk


def _trainer(nk, wc, wf, batch, Nc, Nf, precision=None, compile_=False):
    mc = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    mf = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    mc.set_flat_weights(O.flatten_weights(wc)); mf.set_flat_weights(O.flatten_weights(wf))
    tr = nk.NeRFTrainer(mc, mf, batch, Nc, Nf, 10, 4, precision=nk.PRECISION_BF16_TC if precision is None else precision)
    if compile_:
        tr.compile(nk.Adam(learning_rate=5e-4), nk.MeanSquaredError())
    else:
        tr.build()
    return tr


def test_empty_inputs(nk):
    z3 = np.zeros((0, 3), np.float32)
    assert nk.encode_position(z3, 10).shape == (0, 63)
    assert nk.generate_t_vals(2.0, 6.0, 0, 64, False).shape == (0, 64)
    r, d, w = nk.volume_render(np.zeros((0, 16, 4), np.float32), np.zeros((0, 16), np.float32))
    assert r.shape == (0, 3) and d.shape == (0,) and w.shape == (0, 16)
    rays, dirs = nk.sample_rays(z3, z3, np.zeros((0, 8), np.float32))
    assert rays.shape == (0, 8, 3)
    assert nk.sample_pdf(np.zeros((0, 15), np.float32), np.zeros((0, 16), np.float32), 32).shape == (0, 32)


@pytest.mark.parametrize("B", [1, 3, 127, 129, 300])
def test_ragged_batches_match_fp32_kernel(nk, B):
    """Batches that do not fill a 128-row tile / a 256-row tile pair (and a single ray)."""
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    Nc, Nf = 16, 32
    rng = np.random.default_rng(B)
    o, d = O.get_rays(20, 20, 27.8, O.pose_spherical(11.0, -30.0, 4.0))
    sel = rng.choice(400, B, replace=False)
    o, d = o.reshape(-1, 3)[sel].numpy(), d.reshape(-1, 3)[sel].numpy()
    t = O.generate_t_vals(2.0, 6.0, B, Nc, True, u=g["u_t"]).numpy()
    u = rng.random((B, Nf), dtype=np.float32)
    tc = _trainer(nk, wc, wf, max(B, 4), Nc, Nf)
    f32 = _trainer(nk, wc, wf, max(B, 4), Nc, Nf, precision=nk.PRECISION_FP32)
    a = tc.mlp_forward_rays("coarse", o, d, t)
    b = f32.mlp_forward_rays("coarse", o, d, t)
    assert a.shape == (B, Nc, 4) and torch.isfinite(a).all()
    assert (a - b).abs().max().item() <= 5e-2
    full = tc.forward_pass(o, d, t, u_pdf=u)
    assert full[0][1].shape == (B, 3) and full[2][1].shape == (B, Nc + Nf) and torch.isfinite(full[0][1]).all()


@pytest.mark.parametrize("Nc,Nf", [(2, 1), (16, 32), (33, 7), (64, 128), (100, 60)])
def test_sample_counts_fp32_forward_vs_oracle(nk, Nc, Nf):
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    B = 24
    rng = np.random.default_rng(Nc * 1000 + Nf)
    o, d = g["o"][:B], g["d"][:B]
    t = O.generate_t_vals(2.0, 6.0, B, Nc, True, u=rng.random(Nc, dtype=np.float32)).numpy()
    u = rng.random((B, Nf), dtype=np.float32)
    tr = _trainer(nk, wc, wf, B, Nc, Nf, precision=nk.PRECISION_FP32)
    rgbs, depths, ws, preds = tr.forward_pass(o, d, t, u_pdf=u)
    with torch.no_grad():
        ref = O.forward_pass(wc, wf, torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(t), 10, 4, Nf, torch.from_numpy(u))
    np.testing.assert_allclose(rgbs[0].cpu().numpy(), ref[0][0].numpy(), atol=1e-5)
    np.testing.assert_allclose(ws[0].cpu().numpy(), ref[2][0].numpy(), atol=1e-5)
    np.testing.assert_allclose(rgbs[1].cpu().numpy(), ref[0][1].numpy(), atol=2e-4)
    # same counts through the tensor-core kernel
    tc = _trainer(nk, wc, wf, B, Nc, Nf)
    rgbs_tc = tc.forward_pass(o, d, t, u_pdf=u)[0]
    assert torch.isfinite(rgbs_tc[0]).all() and (rgbs_tc[0] - rgbs[0]).abs().median().item() < 1e-3


def test_batch_larger_than_workspace_is_tiled(nk):
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    tr_small = _trainer(nk, wc, wf, 32, 16, 32)        # workspace of 32 rays
    tr_big = _trainer(nk, wc, wf, 96, 16, 32)
    a = tr_small.forward_pass(g["o"], g["d"], g["t"], u_pdf=g["u_pdf"])
    b = tr_big.forward_pass(g["o"], g["d"], g["t"], u_pdf=g["u_pdf"])
    assert a[0][1].shape == (96, 3)
    assert (a[0][1] - b[0][1]).abs().max().item() <= 1e-6
    # train_step grows the workspace when it sees a bigger batch than batch_size
    tr = _trainer(nk, wc, wf, 32, 16, 32, compile_=True)
    m = tr.train_step((g["img"], (g["o"], g["d"], g["t"])), u_pdf=g["u_pdf"])
    assert np.isfinite(float(m["loss"]))


def test_error_behaviour(nk):
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    with pytest.raises(ValueError):
        nk.encode_position(np.zeros((4, 2), np.float32), 10)
    with pytest.raises(ValueError):
        nk.volume_render(np.zeros((4, 8, 4), np.float32), np.zeros((4, 7), np.float32))
    with pytest.raises(ValueError):
        nk.get_rays(4, 4, 1.0, np.eye(3, dtype=np.float32))
    with pytest.raises(ValueError):
        nk.generate_t_vals(2.0, 6.0, 4, 8, True, u=np.zeros(7, np.float32))
    with pytest.raises(ValueError):
        nk.sample_pdf(np.zeros((4, 8), np.float32), np.zeros((4, 8), np.float32), 4)
    with pytest.raises(TypeError):
        nk.NeRFTrainer("not a model", nk.create_nerf_complete_model(8, 256, 4, 10, 4), 8, 4, 8, 10, 4)
    tr = _trainer(nk, wc, wf, 8, 16, 32)
    with pytest.raises(RuntimeError):
        tr.train_step((g["img"][:8], (g["o"][:8], g["d"][:8], g["t"][:8])))       # compile() not called
    with pytest.raises(ValueError):
        tr.forward_pass(g["o"][:8], g["d"][:8], g["t"][:8, :5])                   # wrong number of coarse samples
    with pytest.raises(ValueError):
        tr.forward_pass(g["o"][:8], g["d"][:8], g["t"][:8], 9, 4)                 # l_xyz differs from the trainer's
    # a non 8x256 architecture has no tensor-core path: fp32 works, the tcgen05 request is refused
    small = nk.create_nerf_complete_model(4, 64, 2, 6, 2)
    out = small([np.zeros((5, 39), np.float32), np.zeros((5, 15), np.float32)])
    assert out.shape == (5, 4)
    tr2 = nk.NeRFTrainer(small, nk.create_nerf_complete_model(4, 64, 2, 6, 2), 8, 4, 8, 6, 2)
    tr2.build()
    with pytest.raises(ValueError):
        tr2.forward_pass(g["o"][:8], g["d"][:8], g["t"][:8, :4], u_pdf=g["u_pdf"][:8, :8])


def test_weights_roundtrip_and_save_load(nk, tmp_path):
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    tr = _trainer(nk, wc, wf, 96, 16, 32, compile_=True)
    tr.train_step((g["img"], (g["o"], g["d"], g["t"])), u_pdf=g["u_pdf"])
    path = str(tmp_path / "w.npz")
    tr.save_weights(path)
    before = tr.forward_pass(g["o"], g["d"], g["t"], u_pdf=g["u_pdf"])[0][1].clone()
    tr2 = _trainer(nk, wc, wf, 96, 16, 32)
    tr2.load_weights(path)
    after = tr2.forward_pass(g["o"], g["d"], g["t"], u_pdf=g["u_pdf"])[0][1]
    assert (before - after).abs().max().item() <= 1e-6
    assert np.abs(tr.coarse_model.get_weights()["d0"]["W"] - wc["d0"]["W"].numpy()).max() > 0


def test_batch_norm_checkpoints_render_by_folding(nk, tmp_path):
    """BATCH_NORM=true (models.py:30-33, 49-52; three of the reference's six configs): inference folds the moving
    statistics into the Dense weights.  fp32 model call vs the oracle's BN forward <= 2e-4; bf16 render within the
    north_star bounds of the oracle's; save/load keeps the BN parameters (training: tests/test_gpu_bn_train.py)."""
    rng = np.random.default_rng(11)
    wc, wf = O.init_weights(seed=5, bias_range=0.1), O.init_weights(seed=6, bias_range=0.1)
    bns = []
    for s in (1, 2):
        bn = MR.init_bn()
        for st in bn.values():
            n = st["gamma"].numel()
            st["gamma"] = torch.from_numpy(rng.uniform(0.5, 1.5, n).astype(np.float32))
            st["beta"] = torch.from_numpy(rng.uniform(-0.2, 0.2, n).astype(np.float32))
            st["mean"] = torch.from_numpy(rng.uniform(-0.3, 0.3, n).astype(np.float32))
            st["var"] = torch.from_numpy(rng.uniform(0.2, 2.0, n).astype(np.float32))
        bns.append(bn)
    to_np = lambda bn: {r: {k: v.numpy() for k, v in st.items()} for r, st in bn.items()}
    mc = nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=True)
    mf = nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=True)
    mc.set_flat_weights(O.flatten_weights(wc)); mf.set_flat_weights(O.flatten_weights(wf))
    mc.set_bn_params(to_np(bns[0])); mf.set_bn_params(to_np(bns[1]))
    # the Keras-style model call (fp32 kernels) against the oracle's BN forward, inference mode
    x = torch.from_numpy(rng.uniform(-1, 1, (300, 63)).astype(np.float32))
    dd = torch.from_numpy(rng.uniform(-1, 1, (300, 27)).astype(np.float32))
    ref = O.nerf_mlp(wc, x, dd, bn=bns[0], training=False)
    got = mc([x.cuda(), dd.cuda()])
    assert np.abs(got.cpu().numpy() - ref.numpy()).max() <= 2e-4
    with pytest.raises(NotImplementedError):
        mc([x.cuda(), dd.cuda()], training=True)
    # trainer: render on the tcgen05 path vs the oracle's forward pass with the same BN parameters
    B, Nc, Nf = 200, 32, 64
    o, d = O.get_rays(20, 20, 25.0, torch.from_numpy(np.asarray(O.pose_spherical(30.0, -30.0, 4.0))))
    o, d = o.reshape(-1, 3)[:B], d.reshape(-1, 3)[:B]
    t = O.generate_t_vals(2.0, 6.0, B, Nc, False)
    u = torch.from_numpy(rng.random((B, Nf), dtype=np.float32))
    rgbs, depths, ws, preds = O.forward_pass(wc, wf, o, d, t, 10, 4, Nf, u, bn_coarse=bns[0], bn_fine=bns[1])[:4]
    tr = nk.NeRFTrainer(mc, mf, B, Nc, Nf, 10, 4)
    tr.build()
    got = tr.forward_pass(o.cuda(), d.cuda(), t.cuda(), u_pdf=u.cuda())
    assert np.abs(got[3][0].cpu().numpy() - preds[0].numpy()).max() <= 5e-2       # coarse raw predictions, bf16 MLP
    stable = np.abs(preds[0].numpy()[:, -1, 3]) > 0.06
    assert stable.mean() > 0.5
    assert np.abs(got[0][0].cpu().numpy() - rgbs[0].numpy())[stable].max() <= 2e-3
    with pytest.raises(NotImplementedError):      # a training-mode forward outside train_step is not offered (see test_gpu_bn_train.py)
        tr.forward_pass(o.cuda(), d.cuda(), t.cuda(), u_pdf=u.cuda(), training=True)
    # weights + BN parameters survive save / load
    path = str(tmp_path / "bn.npz")
    tr.save_weights(path)
    m2c = nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=True)
    m2f = nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=True)
    tr2 = nk.NeRFTrainer(m2c, m2f, B, Nc, Nf, 10, 4)
    tr2.build()
    tr2.load_weights(path)
    again = tr2.forward_pass(o.cuda(), d.cuda(), t.cuda(), u_pdf=u.cuda())
    assert torch.equal(again[0][1], got[0][1])
    np.testing.assert_array_equal(m2f.get_bn_params()["ddir"]["var"], bns[1]["ddir"]["var"].numpy())


def test_host_prefetcher_delivers_batches_in_order(nk):
    """`HostPrefetcher`: pinned-host batches copied on a side stream one step ahead; values, order, slot reuse."""
    from nerf_keras_b200.synthetic import HostPrefetcher
    rng = np.random.default_rng(0)
    host = [tuple(torch.from_numpy(rng.random(shape, dtype=np.float32)).pin_memory() for shape in ((64, 3), (64, 7)))
            for _ in range(7)]
    seen = []
    acc = torch.zeros((), device="cuda")
    for a, b in HostPrefetcher(host):
        assert a.is_cuda and b.shape == (64, 7)
        acc = acc + a.sum() * 1e-3            # some work on the compute stream that reads the slot
        seen.append((a.clone(), b.clone()))
    assert len(seen) == 7
    for (a, b), (ha, hb) in zip(seen, host):
        assert torch.equal(a.cpu(), ha) and torch.equal(b.cpu(), hb)
    assert list(HostPrefetcher([])) == []
    # feeds train_step like any other batch source
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    tr = _trainer(nk, wc, wf, 32, int(g["Nc"]), int(g["Nf"]), compile_=True)
    assert g["o"].shape[0] >= 96
    hb = [tuple(torch.from_numpy(np.ascontiguousarray(g[k][i * 32:(i + 1) * 32])).pin_memory() for k in ("img", "o", "d", "t", "u_pdf"))
          for i in range(3)]
    losses = []
    for img, o, d, t, u in HostPrefetcher(hb):
        losses.append(float(tr.train_step((img, (o, d, t)), u_pdf=u)["loss"]))
    assert len(losses) == 3 and all(np.isfinite(losses))


def test_create_batched_dataset_pipeline_epoch_semantics(nk):
    """data_utils.py:140-170: every ray once per epoch (drop_remainder), locally shuffled, one shared jitter vector."""
    n, B, N = 10_000, 512, 16
    img = np.arange(n * 3, dtype=np.float32).reshape(n, 3)
    o = np.stack([np.arange(n, dtype=np.float32)] * 3, 1)
    d = np.ones((n, 3), np.float32)
    ds = nk.create_batched_dataset_pipeline(img, o, d, N, B, None, near=2.0, far=6.0, shuffle=True, rand_sampling=True)
    assert len(ds) == n // B
    ids, first_t = [], None
    for images, (ro, rd, t) in ds:
        assert images.shape == (B, 3) and ro.shape == (B, 3) and t.shape == (B, N)
        assert torch.equal(images[:, 0], ro[:, 0] * 3)                  # pixels travel with their rays
        first_t = t[0] if first_t is None else first_t
        assert torch.equal(t, first_t.expand(B, N))                      # Q1: one jitter vector for the whole dataset
        ids.append(ro[:, 0].long())
    ids = torch.cat(ids).cpu().numpy()
    assert len(np.unique(ids)) == len(ids) == (n // B) * B               # no ray twice
    pos = np.arange(len(ids))
    assert np.abs(ids - pos).max() <= 5 * B + B and np.abs(ids - pos).mean() > B / 4   # a local shuffle, not sequential
    seq = nk.create_batched_dataset_pipeline(img, o, d, N, B, None, shuffle=False, rand_sampling=False)
    first = next(iter(seq))
    assert torch.equal(first[1][0][:, 0].cpu(), torch.arange(B, dtype=torch.float32))
    assert torch.equal(first[1][2][0].cpu(), O.generate_t_vals(2.0, 6.0, 1, N, False)[0])


def test_randomised_shapes_against_the_oracle(nk):
    """Seeded fuzz over ragged sizes: ray generation (bit-exact), compositing (<= 1e-5) and resample + merge (== sort of
    the concatenation) on 40 random shape combinations, including 1-row / 1-sample / non-multiple-of-32 cases."""
    rng = np.random.default_rng(2024)
    for case in range(40):
        H, W = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        focal = float(rng.uniform(0.5, 900.0))
        pose = np.asarray(O.pose_spherical(float(rng.uniform(-180, 180)), float(rng.uniform(-90, 0)), float(rng.uniform(1, 6))))
        o, d = nk.get_rays(H, W, focal, pose)
        o_ref, d_ref = O.get_rays(H, W, float(np.float32(focal)), torch.from_numpy(pose.astype(np.float32)))
        assert np.array_equal(o.cpu().numpy(), o_ref.numpy()) and np.array_equal(d.cpu().numpy(), d_ref.numpy()), (case, H, W)
        B, N = int(rng.integers(1, 300)), int(rng.integers(1, 260))
        preds = torch.from_numpy(rng.normal(0, 2.5, (B, N, 4)).astype(np.float32))
        t = torch.from_numpy(np.sort(rng.uniform(0.5, 9.0, (B, N)).astype(np.float32), axis=1))
        rgb_r, dep_r, w_r = O.volume_render(preds, t)
        rgb, dep, w = nk.volume_render(preds, t)
        assert np.abs(rgb.cpu().numpy() - rgb_r.numpy()).max() <= 1e-5, (case, B, N)
        assert np.abs(w.cpu().numpy() - w_r.numpy()).max() <= 1e-5, (case, B, N)
        Nc, Nf = int(rng.integers(2, 130)), int(rng.integers(1, 270))
        tc = torch.from_numpy(np.sort(rng.uniform(2.0, 6.0, (B, Nc)).astype(np.float32), axis=1)).cuda()
        wc = torch.from_numpy(rng.random((B, Nc), dtype=np.float32) ** 3).cuda()
        u = torch.from_numpy(rng.random((B, Nf), dtype=np.float32)).cuda()
        t_all, src = nk.resample_merge(tc, wc, Nf, u=u, return_index=True)
        fine = nk.sample_pdf(0.5 * (tc[:, 1:] + tc[:, :-1]), wc, Nf, u=u) if Nc >= 2 else None
        cat = torch.cat([tc, fine], dim=1)
        assert torch.equal(t_all, torch.sort(cat, dim=1).values), (case, B, Nc, Nf)
        assert torch.equal(torch.gather(cat, 1, src.long()), t_all), (case, B, Nc, Nf)
