"""GPU: training path (tcgen05 forward with saved activations, compositing backward, dX chain, weight
gradient GEMMs, Keras-form Adam) against torch.autograd on the oracle."""
import numpy as np
import pytest
import torch

import oracle as O
from oracle.models_ref import _params
from tests.util import cuda, golden_weights, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nk():
    import nerf_keras_b200 as nk
    return nk


def _trainer(nk, g, wc, wf, batch=None, lr=5e-4, stop_grad=True):
    mc = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    mf = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    mc.set_flat_weights(O.flatten_weights(wc))
    mf.set_flat_weights(O.flatten_weights(wf))
    tr = nk.NeRFTrainer(mc, mf, batch or g["o"].shape[0], int(g["Nc"]), int(g["Nf"]), 10, 4, stop_grad_samples=stop_grad)
    tr.compile(nk.Adam(learning_rate=lr), nk.MeanSquaredError())
    return tr


def _split(flat, shapes):
    out, off = {}, 0
    for role, fi, fo in shapes:
        out[role + "/W"] = flat[off:off + fi * fo]; off += fi * fo
        out[role + "/b"] = flat[off:off + fo]; off += fo
    return out


def test_adam_kernel_matches_keras_form(nk):
    from nerf_keras_b200 import _lib
    n = 100003
    gen = torch.Generator().manual_seed(0)
    p = torch.randn(n, generator=gen); g1 = torch.randn(n, generator=gen) * 1e-2; g2 = torch.randn(n, generator=gen) * 1e-3
    p_ref = p.clone()
    opt = O.KerasAdam([p_ref], learning_rate=5e-4)
    pc, m, v = p.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step, g in enumerate((g1, g2, g1), start=1):
        opt.apply_gradients([g])
        gc = (g * 4.0).cuda()   # grad_scale 0.25 folds the 1/world mean into the kernel
        _lib.check(_lib.lib().nerf_adam_flat(pc.data_ptr(), gc.data_ptr(), m.data_ptr(), v.data_ptr(), n, step, 5e-4,
                                             0.25, torch.cuda.current_stream().cuda_stream), "adam")
    np.testing.assert_allclose(pc.cpu().numpy(), p_ref.numpy(), atol=2e-6, rtol=1e-5)


def _bf16_emulated_mlp(w, enc_x, enc_d):
    """fp32 torch math with weights / activations rounded to bf16 where the kernels round them
    (straight-through rounding): separates precision effects from kernel bugs."""
    rnd = lambda x: x + (x.bfloat16().float() - x).detach()
    x = rnd(enc_x)
    for i in range(8):
        p = w[f"d{i}"]
        x = rnd(torch.relu(x @ rnd(p["W"]) + p["b"]))
        if i == 4:
            x = torch.cat([x, rnd(enc_x)], -1)
    sigma = x @ w["sigma"]["W"] + w["sigma"]["b"]
    feat = rnd(x @ rnd(w["feature"]["W"]) + w["feature"]["b"])
    Wd = w["ddir"]["W"]
    hd = torch.relu(feat @ rnd(Wd[:256]) + enc_d @ Wd[256:] + w["ddir"]["b"])
    return torch.cat([hd @ w["rgb"]["W"] + w["rgb"]["b"], sigma], -1)


@pytest.mark.parametrize("name,net", [("lego_small", "coarse"), ("fern_small", "coarse"), ("fern_small", "fine")])
@pytest.mark.parametrize("upstream", ["coherent", "random"])
def test_mlp_backward_matches_autograd(nk, name, net, upstream):
    """d(sum(preds * d_preds))/dW through the tcgen05 fwd+bwd kernels vs torch.autograd on the fp32 oracle at
    identical sample positions.

    The residual against fp32 is dominated by ReLU sign flips of near-zero pre-activations (any bf16 forward
    flips ~1% of them; a torch bf16 emulation of the same net deviates from fp32 by the same 0.4% (heads) to
    13% (first layer)), so per tensor: cosine >= 0.985, norm within 5%, and the kernel must be as close to
    fp32 as the bf16 emulation is (within 1.5x).  Checked for a same-sign and a random upstream gradient."""
    g = load_golden(name)
    wc, wf = golden_weights(g)
    w = wc if net == "coarse" else wf
    t = g["t"] if net == "coarse" else g["t_all"]
    gen = torch.Generator().manual_seed(5)
    if upstream == "random":
        d_preds = torch.randn(t.shape + (4,), generator=gen) * 0.1
    else:
        d_preds = torch.tensor([0.05, -0.03, 0.04, 0.02]).expand(t.shape + (4,)).contiguous()
    o, d, tt = map(torch.from_numpy, (g["o"], g["d"], t))
    params = _params(w)
    for p in params:
        p.requires_grad_(True)
    rays, dirs = O.sample_rays(o, d, tt)
    ex, ed = O.encode_position(rays, 10), O.encode_position(dirs, 4)
    pred_ref = O.nerf_mlp(w, ex, ed)
    grads_ref = torch.autograd.grad((pred_ref * d_preds).sum(), params)
    grads_emu = torch.autograd.grad((_bf16_emulated_mlp(w, ex, ed) * d_preds).sum(), params)
    for p in params:
        p.requires_grad_(False)
    flat = lambda gs: np.concatenate([x.numpy().reshape(-1) for x in gs])
    ref, emu = flat(grads_ref), flat(grads_emu)

    tr = _trainer(nk, g, wc, wf)
    preds, grads = tr.debug_mlp_grads(net, g["o"], g["d"], t, d_preds.numpy())
    np.testing.assert_allclose(preds.cpu().numpy(), pred_ref.detach().numpy(), atol=5e-2)
    got = grads.cpu().numpy()
    assert np.isfinite(got).all()
    shapes = O.layer_shapes()
    a, b, e = _split(got, shapes), _split(ref, shapes), _split(emu, shapes)
    for k in a:
        nb = np.linalg.norm(b[k]) + 1e-20
        rel = np.linalg.norm(a[k] - b[k]) / nb
        rel_emu = np.linalg.norm(e[k] - b[k]) / nb
        cos = float(a[k] @ b[k]) / (np.linalg.norm(a[k]) * nb + 1e-20)
        assert cos >= 0.985 and abs(np.linalg.norm(a[k]) / nb - 1.0) <= 0.05, (k, cos, rel)
        assert rel <= 1.5 * rel_emu + 5e-3, (k, rel, rel_emu)


def test_train_step_metrics_and_coarse_grads(nk):
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    tr = _trainer(nk, g, wc, wf)
    m = {k: float(v) for k, v in tr.train_step((g["img"], (g["o"], g["d"], g["t"])), u_pdf=g["u_pdf"]).items()}
    # the coarse net sees identical inputs; the fine net's sample positions are re-drawn through the
    # ill-conditioned inverse CDF from bf16 coarse weights, so its loss is compared loosely
    assert abs(m["loss_coarse"] - float(g["metrics"][0])) <= 2e-3 * max(1.0, float(g["metrics"][0]))
    assert abs(m["loss"] - float(g["metrics"][1])) <= 0.05 * float(g["metrics"][1]) + 1e-3
    assert abs(m["psnr"] + 10 * np.log10(m["loss"])) < 1e-3


def test_coarse_gradients_match_golden_stop_grad(nk):
    """Gradient buffer after forward+backward (before Adam) vs the oracle's stop-grad gradients."""
    from nerf_keras_b200 import _lib
    for name in ("lego_small", "fern_small"):
        g = load_golden(name)
        wc, wf = golden_weights(g)
        tr = _trainer(nk, g, wc, wf)
        img, o, d, t, u = (cuda(g[k]) for k in ("img", "o", "d", "t", "u_pdf"))
        metrics = torch.empty(3, device="cuda")
        _lib.check(_lib.lib().nerf_train_forward_backward(tr._ctx.handle, img.data_ptr(), o.data_ptr(), d.data_ptr(),
                                                          t.data_ptr(), u.data_ptr(), o.shape[0], metrics.data_ptr(),
                                                          torch.cuda.current_stream().cuda_stream), "fwd_bwd")
        got = tr._ctx.grad_tensor().cpu().numpy()
        n = tr._ctx.n_params
        coarse = got[:n][::61] if False else got[::61][: (n + 60) // 61]
        ref = g["grads_stop_sample"][: coarse.size]
        cos = float(coarse @ ref) / (np.linalg.norm(coarse) * np.linalg.norm(ref))
        assert cos >= 0.99 and abs(np.linalg.norm(coarse) / np.linalg.norm(ref) - 1.0) <= 0.03, (name, cos)
        assert np.isfinite(got).all()


def test_training_reduces_loss_like_the_oracle(nk):
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    tr = _trainer(nk, g, wc, wf)
    batch = (g["img"], (g["o"], g["d"], g["t"]))
    fl = lambda logs: {k: float(v) for k, v in logs.items()}
    first = fl(tr.train_step(batch, u_pdf=g["u_pdf"]))
    tr.reset_metrics()
    for _ in range(40):
        last = fl(tr.train_step(batch, u_pdf=g["u_pdf"]))
        tr.reset_metrics()
    assert last["loss_coarse"] < 0.6 * first["loss_coarse"], (first, last)
    assert last["loss"] < 0.8 * first["loss"], (first, last)
    # oracle run of the same schedule (stop-grad variant) ends in the same regime
    o, d, t, u, img = map(torch.from_numpy, (g["o"], g["d"], g["t"], g["u_pdf"], g["img"]))
    opt = O.KerasAdam(_params(wc) + _params(wf), learning_rate=5e-4)
    for _ in range(41):
        mo = O.train_step(wc, wf, opt, img, o, d, t, 10, 4, int(g["Nf"]), u, stop_grad_samples=True)
    assert abs(last["loss_coarse"] - mo["loss_coarse"]) <= 0.15 * mo["loss_coarse"] + 2e-3, (last, mo)
    # weights round-trip through the trainer
    w_after = tr.coarse_model.get_flat_weights()
    assert np.isfinite(w_after).all() and np.abs(w_after - O.flatten_weights(golden_weights(g)[0])).max() > 1e-4


# ---------------------------------------------------------------- reference semantics: no stop-gradient (quirk Q5)
@pytest.mark.parametrize("name", ["lego_small", "fern_small"])
def test_sample_pdf_backward_matches_autograd(nk, name):
    """d(sum(t_all * g)) / d(weights) through sort(concat([t, sample_pdf(t_mid, w, Nf)])) (models.py:165-167)."""
    from nerf_keras_b200 import _lib
    g = load_golden(name)
    Nc, Nf = int(g["Nc"]), int(g["Nf"])
    t = torch.from_numpy(g["t"]); u = torch.from_numpy(g["u_pdf"])
    w = torch.from_numpy(g["wt_c"]).clone().requires_grad_(True)
    gen = torch.Generator().manual_seed(9)
    gt = torch.randn(t.shape[0], Nc + Nf, generator=gen)
    t_mid = 0.5 * (t[:, 1:] + t[:, :-1])
    t_all, _ = torch.sort(torch.cat([t, O.sample_pdf(t_mid, w, Nf, u=u)], -1), -1)
    (t_all * gt).sum().backward()
    t_all_c, idx = nk.resample_merge(g["t"], g["wt_c"], Nf, u=g["u_pdf"], return_index=True)
    d_w = torch.empty(t.shape[0], Nc, device="cuda")
    tc, wc_, uc, gc = cuda(g["t"]), cuda(g["wt_c"]), cuda(g["u_pdf"]), cuda(gt.numpy())
    _lib.check(_lib.lib().nerf_sample_pdf_bwd(tc.data_ptr(), wc_.data_ptr(), uc.data_ptr(), idx.data_ptr(), gc.data_ptr(), 0,
                                              t.shape[0], Nc, Nf, d_w.data_ptr(), torch.cuda.current_stream().cuda_stream))
    ref, got = w.grad.numpy(), d_w.cpu().numpy()
    # per ray: direction and scale (knot gradients are ill-conditioned where a bin is nearly empty: 1/den)
    num = (ref * got).sum(-1); den = np.linalg.norm(ref, axis=-1) * np.linalg.norm(got, axis=-1) + 1e-30
    assert np.median(num / den) > 0.999 and (num / den > 0.98).mean() > 0.95
    assert abs(np.linalg.norm(got) / np.linalg.norm(ref) - 1.0) < 0.05


@pytest.mark.parametrize("name", ["lego_small", "fern_small"])
def test_delta_path_through_fine_compositing_matches_autograd(nk, name):
    """Coarse weights -> sample_pdf -> sort -> fine compositing deltas -> colour, with FIXED fine predictions (fp32 end
    to end, identical inputs): d(sum(rgb_f * g)) / d(w_c) via volume_render_bwd's d_delta and the resampling backward."""
    from nerf_keras_b200 import _lib
    g = load_golden(name)
    Nc, Nf = int(g["Nc"]), int(g["Nf"])
    B = g["t"].shape[0]
    t = torch.from_numpy(g["t"]); u = torch.from_numpy(g["u_pdf"]); pred_f = torch.from_numpy(g["pred_f"])
    w = torch.from_numpy(g["wt_c"]).clone().requires_grad_(True)
    gr = torch.randn(B, 3, generator=torch.Generator().manual_seed(3))
    t_mid = 0.5 * (t[:, 1:] + t[:, :-1])
    t_all, _ = torch.sort(torch.cat([t, O.sample_pdf(t_mid, w, Nf, u=u)], -1), -1)
    rgb, _, _ = O.volume_render(pred_f, t_all)
    (rgb * gr).sum().backward()
    ref = w.grad.numpy()
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    t_all_c, idx = nk.resample_merge(g["t"], g["wt_c"], Nf, u=g["u_pdf"], return_index=True)
    pf, grc = cuda(g["pred_f"]), cuda(gr.numpy())
    d_preds = torch.empty(B, Nc + Nf, 4, device="cuda"); d_delta = torch.empty(B, Nc + Nf, device="cuda")
    _lib.check(L.nerf_volume_render_bwd(pf.data_ptr(), t_all_c.data_ptr(), grc.data_ptr(), 0, B, Nc + Nf, d_preds.data_ptr(),
                                        d_delta.data_ptr(), st))
    tc, wc_, uc = cuda(g["t"]), cuda(g["wt_c"]), cuda(g["u_pdf"])
    d_w = torch.empty(B, Nc, device="cuda")
    _lib.check(L.nerf_sample_pdf_bwd(tc.data_ptr(), wc_.data_ptr(), uc.data_ptr(), idx.data_ptr(), 0, d_delta.data_ptr(), B, Nc,
                                     Nf, d_w.data_ptr(), st))
    got = d_w.cpu().numpy()
    num = (ref * got).sum(-1); den = np.linalg.norm(ref, axis=-1) * np.linalg.norm(got, axis=-1) + 1e-30
    assert np.median(num / den) > 0.999 and (num / den > 0.98).mean() > 0.9, (np.median(num / den), (num / den > 0.98).mean())
    assert abs(np.linalg.norm(got) / np.linalg.norm(ref) - 1.0) < 0.1


@pytest.mark.parametrize("name", ["lego_small", "fern_small"])
def test_input_gradient_matches_autograd(nk, name):
    """< d_ray, dL/dpts > of the fine net (tcgen05 dZ0 W0^T + dZ5 W5b^T, positional-encoding backward) vs autograd."""
    g = load_golden(name)
    wc, wf = golden_weights(g)
    o, d = torch.from_numpy(g["o"]), torch.from_numpy(g["d"])
    t = torch.from_numpy(g["t_all"]).clone().requires_grad_(True)
    d_preds = torch.tensor([0.05, -0.03, 0.04, 0.02]).expand(t.shape + (4,)).contiguous()
    rays, dirs = O.sample_rays(o, d, t)
    pred = O.nerf_mlp(wf, O.encode_position(rays, 10), O.encode_position(dirs.detach(), 4))
    (pred * d_preds).sum().backward()
    ref = t.grad.numpy()
    tr = _trainer(nk, g, wc, wf, stop_grad=False)
    _, _, dtp = tr.debug_mlp_grads("fine", g["o"], g["d"], g["t_all"], d_preds.numpy(), return_input_grad=True)
    got = dtp.cpu().numpy()
    assert np.isfinite(got).all()
    cos = float((ref * got).sum() / (np.linalg.norm(ref) * np.linalg.norm(got)))
    assert cos > 0.97 and abs(np.linalg.norm(got) / np.linalg.norm(ref) - 1.0) < 0.1, (cos, np.linalg.norm(got), np.linalg.norm(ref))
    # the training step takes the same quantity from the weight-gradient kernel (the jobs that stream dZ0 / dZ5 multiply
    # them with W0^T / W5b^T on the way): same bf16 operands, fp32 accumulation split over two jobs
    _, _, fused = tr.debug_mlp_grads("fine", g["o"], g["d"], g["t_all"], d_preds.numpy(), return_input_grad="fused")
    fused = fused.cpu().numpy()
    assert np.abs(fused - got).max() <= 2e-5 * np.abs(got).max() + 1e-9, (np.abs(fused - got).max(), np.abs(got).max())


@pytest.mark.parametrize("name", ["lego_small", "fern_small"])
def test_unstopped_gradient_into_coarse_weights_at_identical_inputs(nk, name):
    """The whole extra path of the reference (fine loss -> fine compositing -> fine MLP input gradient ->
    sort/sample_pdf -> dL/dw_coarse), CUDA kernels vs autograd on the oracle, starting from the SAME coarse weights
    w_c.  (End to end the term is chaotic: a 1e-4 relative change of the coarse weights decorrelates it in the fp32
    oracle itself, see DESIGN.md -- so it is pinned here with w_c held identical.)"""
    from nerf_keras_b200 import _lib
    g = load_golden(name)
    wc, wf = golden_weights(g)
    Nc, Nf = int(g["Nc"]), int(g["Nf"])
    B = g["o"].shape[0]
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    tr = _trainer(nk, g, wc, wf, stop_grad=False)
    # CUDA forward (bf16 tensor-core MLPs)
    rgbs, depths, ws, preds, t_all = tr.forward_pass(g["o"], g["d"], g["t"], u_pdf=g["u_pdf"], return_t_all=True)
    w_c = ws[0].contiguous()
    _, idx = nk.resample_merge(g["t"], w_c, Nf, u=g["u_pdf"], return_index=True)
    img = cuda(g["img"])
    d_rgb_f = (2.0 * (rgbs[1] - img) / (3 * B)).contiguous()
    d_pred_f = torch.empty(B, Nc + Nf, 4, device="cuda"); d_delta = torch.empty(B, Nc + Nf, device="cuda")
    _lib.check(L.nerf_volume_render_bwd(preds[1].data_ptr(), t_all.data_ptr(), d_rgb_f.data_ptr(), 0, B, Nc + Nf,
                                        d_pred_f.data_ptr(), d_delta.data_ptr(), st))
    _, _, dtp = tr.debug_mlp_grads("fine", g["o"], g["d"], t_all, d_pred_f, return_input_grad=True)
    tc, uc = cuda(g["t"]), cuda(g["u_pdf"])
    d_w = torch.empty(B, Nc, device="cuda")
    _lib.check(L.nerf_sample_pdf_bwd(tc.data_ptr(), w_c.data_ptr(), uc.data_ptr(), idx.data_ptr(), dtp.data_ptr(),
                                     d_delta.data_ptr(), B, Nc, Nf, d_w.data_ptr(), st))
    got = d_w.cpu().numpy()
    # oracle from the same w_c (fp32 fine net)
    o, d, t, u = map(torch.from_numpy, (g["o"], g["d"], g["t"], g["u_pdf"]))
    w = w_c.cpu().clone().requires_grad_(True)
    t_mid = 0.5 * (t[:, 1:] + t[:, :-1])
    ta, _ = torch.sort(torch.cat([t, O.sample_pdf(t_mid, w, Nf, u=u)], -1), -1)
    rays, dirs = O.sample_rays(o, d, ta)
    pf = O.nerf_mlp(wf, O.encode_position(rays, 10), O.encode_position(dirs, 4))
    rgb_f, _, _ = O.volume_render(pf, ta)
    O.mse(torch.from_numpy(g["img"]), rgb_f).backward()
    ref = w.grad.numpy()
    assert np.isfinite(got).all()
    num = (ref * got).sum(-1); den = np.linalg.norm(ref, axis=-1) * np.linalg.norm(got, axis=-1) + 1e-30
    rn, gn = np.linalg.norm(ref, axis=-1), np.linalg.norm(got, axis=-1)
    print(name, "per-ray cosine median", np.median(num / den), "frac > 0.9:", (num / den > 0.9).mean(),
          "median norm ratio", np.median(gn / (rn + 1e-30)), "total norm", np.linalg.norm(got), np.linalg.norm(ref))
    assert np.median(num / den) > 0.95 and (num / den > 0.8).mean() > 0.8
    assert 0.8 < np.median(gn / (rn + 1e-30)) < 1.25


def test_training_with_reference_semantics_reduces_loss(nk):
    """40 Adam steps with the reference's (un-stopped) gradient.  The fine loss -- the one the reference reports as
    `loss` -- goes down as in the oracle run of the same schedule; the coarse loss is driven by the chaotic
    sample-position term in both implementations (it stalls or rises), so it is only required to stay finite."""
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    tr = _trainer(nk, g, wc, wf, stop_grad=False)
    batch = (g["img"], (g["o"], g["d"], g["t"]))
    fl = lambda logs: {k: float(v) for k, v in logs.items()}
    first = fl(tr.train_step(batch, u_pdf=g["u_pdf"]))
    tr.reset_metrics()
    for _ in range(40):
        last = fl(tr.train_step(batch, u_pdf=g["u_pdf"]))
        tr.reset_metrics()
    assert np.isfinite(list(last.values())).all()
    assert last["loss"] < 0.6 * first["loss"], (first, last)          # oracle: 0.1334 -> 0.0447
    assert last["loss_coarse"] < 1.0


def test_direction_rows_and_input_gradient_with_ragged_rays(nk):
    """Ray boundaries that fall INSIDE the 16-row blocks of the chain kernel's per-ray dZ_ddir sums (23 samples per ray)
    and a ragged last tile (37 rays x 23 = 851 rows): the direction rows of dW_ddir (rows 256..282, models.py:48-54), every
    other tensor and the fused input gradient against torch.autograd on the fp32 oracle."""
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    B, N = 37, 23
    gen = torch.Generator().manual_seed(11)
    o = torch.randn(B, 3, generator=gen) * 0.3 + torch.tensor([0.0, 0.0, 3.5])
    d = torch.nn.functional.normalize(torch.randn(B, 3, generator=gen) * 0.2 + torch.tensor([0.0, 0.0, -1.0]), dim=-1)
    t = (torch.sort(torch.rand(B, N, generator=gen), -1).values * 4.0 + 2.0).requires_grad_(True)
    d_preds = torch.randn(B, N, 4, generator=gen) * 0.1
    params = _params(wf)
    for p in params:
        p.requires_grad_(True)
    rays, dirs = O.sample_rays(o, d, t)
    pred = O.nerf_mlp(wf, O.encode_position(rays, 10), O.encode_position(dirs.detach(), 4))
    grads_ref = torch.autograd.grad((pred * d_preds).sum(), params + [t])
    for p in params:
        p.requires_grad_(False)
    ref = np.concatenate([x.numpy().reshape(-1) for x in grads_ref[:-1]])
    dtp_ref = grads_ref[-1].numpy()

    mc = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    mf = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    mc.set_flat_weights(O.flatten_weights(wc)); mf.set_flat_weights(O.flatten_weights(wf))
    tr = nk.NeRFTrainer(mc, mf, B, 10, 13, 10, 4, stop_grad_samples=False)
    tr.compile(nk.Adam(learning_rate=5e-4), nk.MeanSquaredError())
    tn = t.detach().numpy()
    _, grads, dtp = tr.debug_mlp_grads("fine", o.numpy(), d.numpy(), tn, d_preds.numpy(), return_input_grad="fused")
    _, _, dtp_alone = tr.debug_mlp_grads("fine", o.numpy(), d.numpy(), tn, d_preds.numpy(), return_input_grad=True)
    got, shapes = grads.cpu().numpy(), O.layer_shapes()
    a, b = _split(got, shapes), _split(ref, shapes)
    wd_got, wd_ref = a["ddir/W"].reshape(283, 128)[256:], b["ddir/W"].reshape(283, 128)[256:]
    cos = float((wd_got * wd_ref).sum() / (np.linalg.norm(wd_got) * np.linalg.norm(wd_ref)))
    assert cos >= 0.995 and abs(np.linalg.norm(wd_got) / np.linalg.norm(wd_ref) - 1.0) <= 0.03, cos
    for k in a:
        nb = np.linalg.norm(b[k]) + 1e-20
        c = float(a[k] @ b[k]) / (np.linalg.norm(a[k]) * nb + 1e-20)
        assert c >= 0.97 and abs(np.linalg.norm(a[k]) / nb - 1.0) <= 0.08, (k, c)
    dtp, dtp_alone = dtp.cpu().numpy(), dtp_alone.cpu().numpy()
    assert np.abs(dtp - dtp_alone).max() <= 2e-5 * np.abs(dtp_alone).max() + 1e-9
    c = float((dtp * dtp_ref).sum() / (np.linalg.norm(dtp) * np.linalg.norm(dtp_ref)))
    assert c > 0.97, c
