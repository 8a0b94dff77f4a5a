"""GPU: round-2 additions -- CUDA-graph replay of the training step, in-kernel Philox draws, the single-net benchmark
shape, micro-batch accumulation, optimiser-state / learning-rate plumbing, test_step parity (models.py:122-145), and
the CUDA path against the fixtures produced by the REFERENCE'S OWN SOURCE (tests/golden/ref_source.npz)."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle as O
from tests.util import cuda, golden_weights, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nk():
    import nerf_keras_b200 as nk
    return nk


def _trainer(nk, wc, wf, B, Nc, Nf, compile_=True, lr=5e-4, **kw):
    mc = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    mf = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    mc.set_flat_weights(O.flatten_weights(wc))
    mf.set_flat_weights(O.flatten_weights(wf))
    tr = nk.NeRFTrainer(mc, mf, B, Nc, Nf, 10, 4, **kw)
    if compile_:
        tr.compile(nk.Adam(learning_rate=lr), nk.MeanSquaredError())
    else:
        tr.build()
    return tr


def _dev_batch(g):
    return tuple(cuda(g[k]) for k in ("img", "o", "d", "t", "u_pdf"))


def _weights(tr):
    return np.concatenate([tr.coarse_model.get_flat_weights(), tr.fine_model.get_flat_weights()])


def test_cuda_graph_replay_matches_eager_steps(nk):
    """The captured step reads step count, learning rate and draws from device memory: replays == eager steps."""
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    img, o, d, t, u = _dev_batch(g)
    runs = {}
    for mode in (True, False):
        tr = _trainer(nk, wc, wf, 96, 16, 32, use_cuda_graph=mode, stop_grad_samples=True)
        losses = []
        for _ in range(6):
            losses.append(float(tr.train_step((img, (o, d, t)), u_pdf=u)["loss_coarse"]))
            tr.reset_metrics()
        assert len(tr._graphs) == (1 if mode else 0)
        m, v, step = tr._ctx.optimizer_state()
        assert step == 6
        runs[mode] = (np.array(losses), _weights(tr))
    # same kernels, same order; only the atomics of the weight-gradient reduction reorder sums
    np.testing.assert_allclose(runs[True][0], runs[False][0], rtol=2e-3)
    assert runs[True][0][-1] < runs[True][0][0]
    dw = np.abs(runs[True][1] - runs[False][1])
    assert np.quantile(dw, 0.999) < 3e-4 and np.median(dw) < 2e-5, (np.quantile(dw, 0.999), np.median(dw))


def test_inkernel_draws_are_uniform_and_shared_by_forward_and_backward(nk):
    from nerf_keras_b200 import _lib
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    B, nf = 4096, 128
    a, b, c = (torch.empty((B, nf), device="cuda") for _ in range(3))
    _lib.check(L.nerf_debug_pdf_draws(42, 0, B, nf, a.data_ptr(), st), "draws")
    _lib.check(L.nerf_debug_pdf_draws(42, 1, B, nf, b.data_ptr(), st), "draws")
    _lib.check(L.nerf_debug_pdf_draws(43, 0, B, nf, c.data_ptr(), st), "draws")
    x = a.cpu().numpy().astype(np.float64)
    assert x.min() >= 0.0 and x.max() < 1.0
    assert abs(x.mean() - 0.5) < 2e-3 and abs(x.var() - 1 / 12) < 1e-3
    assert abs(np.corrcoef(x[:, :-1].ravel(), x[:, 1:].ravel())[0, 1]) < 5e-3          # neighbouring draws of a ray
    assert abs(np.corrcoef(x[:-1].ravel(), x[1:].ravel())[0, 1]) < 5e-3                # neighbouring rays
    hist = np.histogram(x, bins=64, range=(0, 1))[0]
    assert hist.min() > 0.9 * x.size / 64 and hist.max() < 1.1 * x.size / 64
    assert not np.array_equal(a.cpu().numpy(), b.cpu().numpy()) and not np.array_equal(a.cpu().numpy(), c.cpu().numpy())

    # a training step with u_pdf=None uses exactly the draws of (seed, optimiser step): its gradient buffer equals the one
    # of the explicit-draw step, including the un-stopped term whose BACKWARD kernel regenerates the numbers
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    img, o, d, t, _ = _dev_batch(g)
    nk.set_random_seed(7)
    tr = _trainer(nk, wc, wf, 96, 16, 32, use_cuda_graph=False, stop_grad_samples=False)
    u0 = torch.empty((96, 32), device="cuda")
    _lib.check(L.nerf_debug_pdf_draws(7, 0, 96, 32, u0.data_ptr(), st), "draws")
    metrics = torch.empty(3, device="cuda")
    grads = []
    for up in (0, u0.data_ptr()):
        _lib.check(L.nerf_train_phases(tr._ctx.handle, img.data_ptr(), o.data_ptr(), d.data_ptr(), t.data_ptr(), up, 96,
                                       metrics.data_ptr(), 3, st), "train_phases")
        grads.append(tr._ctx.grad_tensor().clone())
    assert torch.isfinite(grads[0]).all() and grads[0].abs().max() > 0
    rel = (grads[0] - grads[1]).norm() / grads[1].norm()
    assert rel < 1e-4, float(rel)
    # forward passes draw fresh numbers on every call (data_utils.py:196 is random at inference too)
    tr2 = _trainer(nk, wc, wf, 96, 16, 32, compile_=False)
    t1 = tr2.forward_pass(o, d, t, return_t_all=True)[4]
    t2 = tr2.forward_pass(o, d, t, return_t_all=True)[4]
    assert not torch.equal(t1, t2) and bool((t1[:, 1:] >= t1[:, :-1]).all())


def test_single_net_shape_ns_fine_0(nk):
    """NS_FINE = 0: the 64-samples-per-ray single-net point of the ray-batch sweep (BASELINE configs[4])."""
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    img, o, d, t, u = _dev_batch(g)
    two = _trainer(nk, wc, wf, 96, 16, 32, compile_=False)
    one = _trainer(nk, wc, wf, 96, 16, 0, use_cuda_graph=True)
    r2 = two.forward_pass(o, d, t, u_pdf=u)
    r1 = one.forward_pass(o, d, t)
    assert torch.equal(r1[0][0], r2[0][0]) and r1[0][1] is None and r1[3][1] is None
    first = float(one.train_step((img, (o, d, t)))["loss"])
    assert abs(first - float(g["metrics"][0])) <= 2e-3            # `loss` is the only net's loss
    n = one._ctx.n_params
    for _ in range(30):
        one.reset_metrics()
        last = float(one.train_step((img, (o, d, t)))["loss"])
    assert last < 0.7 * first
    assert np.array_equal(one.fine_model.get_flat_weights(), O.flatten_weights(wf))     # the second net is untouched


def test_microbatch_accumulation_equals_one_big_batch(nk):
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    img, o, d, t, u = _dev_batch(g)
    big = _trainer(nk, wc, wf, 96, 16, 32, use_cuda_graph=False, stop_grad_samples=True)
    small = _trainer(nk, wc, wf, 32, 16, 32, use_cuda_graph=False, stop_grad_samples=True)   # 32-ray workspace: three micro-batches per step
    for tr in (big, small):
        for _ in range(2):
            tr.train_step((img, (o, d, t)), u_pdf=u)
    assert small._ctx.max_rays == 32 and small._ctx.optimizer_state()[2] == 2
    # Adam's first steps move a weight by ~lr in the direction of its gradient's sign: weights whose tiny gradient changes
    # sign with the summation order differ by up to 2 lr per step, the bulk agrees to a few 1e-6
    w0 = np.concatenate([O.flatten_weights(wc), O.flatten_weights(wf)])
    ub, us = _weights(big) - w0, _weights(small) - w0
    dw = np.abs(ub - us)
    cos = float(ub @ us) / (np.linalg.norm(ub) * np.linalg.norm(us))
    assert dw.max() <= 2.1e-3 and np.median(dw) < 1e-4 and cos > 0.9, (dw.max(), np.median(dw), cos)
    lb, ls = float(big.loss_tracker.result()), float(small.loss_tracker.result())
    assert abs(lb - ls) <= 2e-3 * max(lb, 1e-3)


def test_optimizer_state_and_learning_rate_plumbing(nk):
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    img, o, d, t, u = _dev_batch(g)
    tr = _trainer(nk, wc, wf, 96, 16, 32, stop_grad_samples=True)
    for _ in range(3):
        tr.train_step((img, (o, d, t)), u_pdf=u)
    m, v, step = tr._ctx.optimizer_state()
    assert step == 3 and float(m.abs().max()) > 0 and float(v.max()) > 0
    # moving to a larger workspace keeps the moments and the step
    tr._rebuild_ctx(max_rays=128)
    m2, v2, step2 = tr._ctx.optimizer_state()
    assert step2 == 3 and torch.equal(m, m2) and torch.equal(v, v2)
    # learning rate 0 after compile(): the step runs, the weights stay (also through a replayed graph)
    w0 = _weights(tr)
    tr.optimizer.learning_rate = 0.0
    for _ in range(3):
        tr.train_step((img, (o, d, t)), u_pdf=u)
    assert np.array_equal(_weights(tr), w0) and tr._ctx.optimizer_state()[2] == 6
    tr.optimizer.learning_rate = 5e-4
    tr.train_step((img, (o, d, t)), u_pdf=u)
    assert not np.array_equal(_weights(tr), w0)
    with pytest.raises(TypeError):
        tr.compile(nk.Adam(5e-4), lambda a, b: 0.0)                # only the reference's MeanSquaredError is implemented


@pytest.mark.parametrize("name", ["lego_small", "fern_small"])
def test_test_step_matches_oracle(nk, name):
    """models.py:122-145: loss_coarse, loss (fine only) and psnr of a validation batch, fp32 path, running means."""
    g = load_golden(name)
    wc, wf = golden_weights(g)
    B, Nc, Nf = g["o"].shape[0], int(g["Nc"]), int(g["Nf"])
    tr = _trainer(nk, wc, wf, B, Nc, Nf, compile_=False, precision=nk.PRECISION_FP32)
    img, o, d, t, u = _dev_batch(g)
    ref = O.test_step(wc, wf, *(torch.from_numpy(g[k]) for k in ("img", "o", "d", "t")), 10, 4, Nf, torch.from_numpy(g["u_pdf"]))
    got = {k: float(v) for k, v in tr.test_step((img, (o, d, t)), u_pdf=u).items()}
    for k in ("loss_coarse", "loss", "psnr"):
        assert abs(got[k] - ref[k]) <= 2e-5 * max(1.0, abs(ref[k])), (k, got[k], ref[k])
    # second validation batch: half the rays -> the trackers report the mean over the two batches
    h = B // 2
    ref2 = O.test_step(wc, wf, *(torch.from_numpy(g[k][:h]) for k in ("img", "o", "d", "t")), 10, 4, Nf,
                       torch.from_numpy(g["u_pdf"][:h]))
    got2 = {k: float(v) for k, v in tr.test_step((img[:h], (o[:h], d[:h], t[:h])), u_pdf=u[:h]).items()}
    for k in ("loss_coarse", "loss", "psnr"):
        assert abs(got2[k] - 0.5 * (ref[k] + ref2[k])) <= 3e-5 * max(1.0, abs(ref[k])), k
    # the bf16 tensor-core path reports the same metrics within its precision
    trb = _trainer(nk, wc, wf, B, Nc, Nf, compile_=False)
    gb = {k: float(v) for k, v in trb.test_step((img, (o, d, t)), u_pdf=u).items()}
    assert abs(gb["loss_coarse"] - ref["loss_coarse"]) <= 2e-3 and abs(gb["psnr"] - ref["psnr"]) <= 0.3


# ---------------------------------------------------------------------------------------------------------------------
# the CUDA path against what the reference's own source computes (tests/golden/ref_source.npz)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def R():
    return load_golden("ref_source")


@pytest.mark.parametrize("name", ["lego800", "fern378", "odd"])
def test_get_rays_bit_exact_vs_reference_source(nk, R, name):
    import hashlib
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    H, W = int(R[f"rays_{name}_H"]), int(R[f"rays_{name}_W"])
    o, d = nk.get_rays(H, W, R[f"rays_{name}_focal"], R[f"rays_{name}_pose"])
    assert sha(o.cpu().numpy()) == str(R[f"rays_{name}_o_sha"]) and sha(d.cpu().numpy()) == str(R[f"rays_{name}_d_sha"])


def test_ops_vs_reference_source(nk, R):
    for name in ("lego", "fern", "dbg"):
        near, far, N = R[f"tv_{name}_args"]
        assert np.array_equal(nk.generate_t_vals(near, far, 5, int(N), True, u=R[f"tv_{name}_u"]).cpu().numpy(), R[f"tv_{name}_jit"])
        assert np.array_equal(nk.generate_t_vals(near, far, 5, int(N), False).cpu().numpy(), R[f"tv_{name}_nojit"])
    for case, ref in zip(R["pose_spherical_cases"], R["pose_spherical"]):
        assert np.array_equal(np.asarray(nk.pose_spherical(*map(float, case))), ref)
    pts, dirs = nk.sample_rays(R["op_o"], R["op_d"], R["op_t"])
    assert np.array_equal(pts.cpu().numpy(), R["op_pts"]) and np.array_equal(dirs.cpu().numpy(), R["op_dirs"])
    assert np.abs(nk.encode_position(pts, 10).cpu().numpy()[:12] - R["op_enc_x"]).max() <= 1e-6
    assert np.abs(nk.encode_position(dirs, 4).cpu().numpy()[:12] - R["op_enc_d"]).max() <= 1e-6
    rgb, depth, w = nk.volume_render(R["vr_preds"], R["op_t"])
    assert np.abs(rgb.cpu().numpy() - R["vr_rgb"]).max() <= 1e-5 and np.abs(w.cpu().numpy() - R["vr_w"]).max() <= 1e-5
    assert np.abs(depth.cpu().numpy() - R["vr_depth"]).max() <= 5e-5
    s = nk.sample_pdf(R["sp_t_mid"], R["sp_w"], R["sp_u"].shape[1], u=R["sp_u"])
    assert np.abs(s.cpu().numpy() - R["sp_samples"]).max() <= 3e-5


@pytest.mark.parametrize("tag", ["m_lego", "m_fern"])
def test_forward_pass_fp32_vs_reference_source(nk, R, tag):
    """NeRFTrainer.forward_pass on the fp32 path against the reference source's outputs: north_star's 1e-5 on rgb."""
    sc, sf = (int(x) for x in R[f"{tag}_seeds"])
    wc, wf = O.init_weights(sc, 0.1), O.init_weights(sf, 0.1)
    B, Nc, Nf = (int(x) for x in R[f"{tag}_dims"])
    tr = _trainer(nk, wc, wf, B, Nc, Nf, compile_=False, precision=nk.PRECISION_FP32)
    o, d, t, u = (cuda(R[f"{tag}_{k}"]) for k in ("o", "d", "t", "u1"))
    rgbs, depths, ws, preds = tr.forward_pass(o, d, t, u_pdf=u)
    assert np.abs(rgbs[0].cpu().numpy() - R[f"{tag}_rgb_c"]).max() <= 1e-5
    assert np.abs(preds[0].cpu().numpy() - R[f"{tag}_pred_c"]).max() <= 1e-4
    assert np.abs(ws[0].cpu().numpy() - R[f"{tag}_w_c"]).max() <= 1e-5
    # the fine pass resamples through the inverse CDF (ill-conditioned in empty bins): 1e-4 end to end
    assert np.abs(rgbs[1].cpu().numpy() - R[f"{tag}_rgb_f"]).max() <= 1e-4
    # the model call on explicit encodings (models.py:24-62)
    pts, dirs = nk.sample_rays(o, d, t)
    mlp = tr.coarse_model([nk.encode_position(pts, 10), nk.encode_position(dirs, 4)])
    assert np.abs(mlp.cpu().numpy() - R[f"{tag}_mlp_c"]).max() <= 1e-4
    # test_step metrics
    got = {k: float(v) for k, v in tr.test_step((cuda(R[f"{tag}_img"]), (o, d, t)), u_pdf=u).items()}
    np.testing.assert_allclose([got["loss_coarse"], got["loss"], got["psnr"]], R[f"{tag}_test_metrics"], rtol=1e-4)


@pytest.mark.parametrize("tag", ["m_lego", "m_fern"])
def test_train_step_vs_reference_source(nk, R, tag):
    """First train_step of the reference (tape.gradient of the literal graph + Adam, models.py:94-107): metrics, and the
    weights after the update (Adam's first step is -lr * g / (|g| + eps'): compared where |g| is above the bf16 noise)."""
    sc, sf = (int(x) for x in R[f"{tag}_seeds"])
    wc, wf = O.init_weights(sc, 0.1), O.init_weights(sf, 0.1)
    B, Nc, Nf = (int(x) for x in R[f"{tag}_dims"])
    tr = _trainer(nk, wc, wf, B, Nc, Nf, use_cuda_graph=False)          # reference gradient semantics (Q5) by default
    img, o, d, t, u2 = (cuda(R[f"{tag}_{k}"]) for k in ("img", "o", "d", "t", "u2"))
    logs = {k: float(v) for k, v in tr.train_step((img, (o, d, t)), u_pdf=u2).items()}
    ref = R[f"{tag}_train_logs"][0]
    assert abs(logs["loss_coarse"] - ref[0]) <= 2e-3 * max(1.0, ref[0])
    assert abs(logs["loss"] - ref[1]) <= 0.05 * ref[1] + 1e-3
    w0 = np.concatenate([O.flatten_weights(wc), O.flatten_weights(wf)])[::53]
    w1 = _weights(tr)[::53]
    ref1 = R[f"{tag}_weights_after1"]
    half = ref1.size // 2
    # fine net: the direction of every first update matches for the overwhelming majority of weights
    mv_ref, mv = ref1[half:] - w0[half:], w1[half:] - w0[half:]
    big = np.abs(mv_ref) > 4.5e-4                                       # |g| well above eps: a full-size +-lr step
    agree = np.mean(np.sign(mv[big]) == np.sign(mv_ref[big]))
    assert big.mean() > 0.3 and agree > 0.97, (big.mean(), agree)


def test_graph_cache_follows_buffers_and_batch_sizes(nk):
    """One graph per distinct set of input buffers (and batch size); eager steps for buffers seen once; release_graphs()."""
    from nerf_keras_b200.synthetic import HostPrefetcher
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    tr = _trainer(nk, wc, wf, 96, 16, 32, stop_grad_samples=True)
    full = _dev_batch(g)[:4]
    half = tuple(x[:48].contiguous() for x in full)
    for _ in range(3):
        tr.train_step((full[0], full[1:4]))
        tr.train_step((half[0], half[1:4]))
    assert len(tr._graphs) == 2 and tr._ctx.optimizer_state()[2] == 6
    # temporaries (new device tensors every call) never match a captured graph: plain eager steps, same result path
    for _ in range(3):
        tr.train_step((g["img"], (g["o"], g["d"], g["t"])))
    assert tr._ctx.optimizer_state()[2] == 9
    # a host prefetcher hands out two alternating staging slots, also after being re-iterated: two more graphs, not more
    host = [tuple(torch.from_numpy(np.ascontiguousarray(g[k])).pin_memory() for k in ("img", "o", "d", "t")) for _ in range(2)]
    pf = HostPrefetcher((host[i % 2] for i in range(1 << 20)), torch.device("cuda"))
    n0 = len(tr._graphs)
    for rep in range(2):
        it = iter(pf)
        for _ in range(5):
            img, o, d, t = next(it)
            tr.train_step((img, (o, d, t)))
    assert len(tr._graphs) - n0 <= 2 and tr._ctx.optimizer_state()[2] == 19
    losses = float(tr.loss_tracker.result())
    assert np.isfinite(losses)
    tr.release_graphs()
    assert len(tr._graphs) == 0
    tr.train_step((full[0], full[1:4]))
    assert tr._ctx.optimizer_state()[2] == 20


@pytest.mark.gpu
def test_backward_overlap_matches_sequential():
    """nerf_set_backward_overlap: the weight-gradient kernel consuming dZ images NEXT TO the dX chain (progress counters,
    side stream) gives the gradients of the default schedule (only the split-K summation order differs)."""
    import ctypes as C
    import nerf_keras_b200 as nk
    from nerf_keras_b200 import _lib
    from nerf_keras_b200.models import _ptr, _stream
    L = _lib.lib()
    B, Nc, Nf = 2048, 64, 128      # 512 / 1536 tile pairs: both nets are above the one-pair-per-SM threshold
    nk.set_random_seed(7)
    c = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    f = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    tr = nk.NeRFTrainer(c, f, B, Nc, Nf, 10, 4)
    tr.compile(nk.Adam(learning_rate=5e-4), nk.MeanSquaredError())
    o, d = nk.get_rays(64, 32, 60.0, nk.pose_spherical(20.0, -30.0, 4.0))
    o, d = o.reshape(-1, 3).contiguous(), d.reshape(-1, 3).contiguous()
    t = nk.generate_t_vals(2.0, 6.0, B, Nc, True, u=np.random.default_rng(5).random(Nc, dtype=np.float32))
    u = torch.rand(B, Nf, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    img = torch.rand(B, 3, device="cuda", generator=torch.Generator("cuda").manual_seed(4))
    metrics = torch.empty(3, device="cuda")
    grads = {}
    for sms in (0, 52, 0, 74):
        _lib.check(L.nerf_set_backward_overlap(tr._ctx.handle, sms), "overlap")
        _lib.check(L.nerf_train_forward_backward(tr._ctx.handle, _ptr(img), _ptr(o), _ptr(d), _ptr(t), _ptr(u), B,
                                                 _ptr(metrics), _stream()), "fb")
        torch.cuda.synchronize()
        grads.setdefault(sms, []).append(tr._ctx.grad_tensor().cpu().numpy().copy())
    ref = grads[0][0]
    assert np.isfinite(ref).all() and np.abs(ref).max() > 0
    rerun = np.abs(grads[0][1] - ref).max()                    # atomics: run-to-run noise of the default schedule
    for sms in (52, 74):
        err = np.abs(grads[sms][0] - ref).max()
        assert err <= max(4 * rerun, 1e-5 * np.abs(ref).max()), (sms, err, rerun)
    assert L.nerf_set_backward_overlap(tr._ctx.handle, 5) != 0     # below the 13 jobs: refused
    _lib.check(L.nerf_set_backward_overlap(tr._ctx.handle, 0), "overlap off")
