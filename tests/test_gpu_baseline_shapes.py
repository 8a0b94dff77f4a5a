"""GPU: BASELINE.json's own shapes against the ORACLE (not CUDA-vs-CUDA): 4096-ray batches, 64 coarse + 128 fine samples,
rays of a Lego-shaped 800x800 view and of a Fern-shaped 378x504 view.  The oracle forward takes a few seconds per scene
on the box's host cores.  north_star tolerances, UNMASKED (every ray counts):
  rays / t-values bit-exact; fp32 rendered rgb <= 1e-5; bf16 tensor-core render <= 2e-3 abs per pixel and <= 0.05 dB PSNR
  on a fixed random-init model (Keras default initialisation: glorot-uniform kernels, zero biases)."""
import json
import os

import numpy as np
import pytest
import torch

import oracle as O
from oracle.models_ref import compute_grads
from tests.util import cuda, psnr_db

pytestmark = pytest.mark.gpu
B, NC, NF = 4096, 64, 128
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


@pytest.fixture(scope="module")
def nk():
    import nerf_keras_b200 as nk
    return nk


def _scene(name):
    if name == "lego":
        H = W = 800
        focal = np.float32(0.5 * W / np.tan(0.5 * 0.6911112))
        pose, near, far = O.pose_spherical(35.0, -30.0, 4.0), 2.0, 6.0
    else:
        H, W, focal = 378, 504, np.float32(407.6)
        pose = np.eye(4, dtype=np.float32)
        c, s = np.cos(np.float32(0.1)), np.sin(np.float32(0.1))
        pose[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float32)
        pose[:3, 3] = [0.12, -0.2, 0.05]
        near, far = 1.2, 12.0
    return H, W, focal, pose, near, far


@pytest.fixture(scope="module", params=["lego", "fern"])
def case(request, nk):
    """Inputs + oracle outputs of one 4096-ray batch (Keras-default-init model and a model with non-zero biases)."""
    H, W, focal, pose, near, far = _scene(request.param)
    rng = np.random.default_rng(11 if request.param == "lego" else 12)
    o_ref, d_ref = O.get_rays(H, W, focal, pose)
    sel = rng.choice(H * W, B, replace=False)
    u_t = rng.random(NC, dtype=np.float32)
    u_pdf = rng.random((B, NF), dtype=np.float32)
    img = rng.random((B, 3), dtype=np.float32)
    t_ref = O.generate_t_vals(near, far, B, NC, True, u=u_t)
    o_s, d_s = o_ref.reshape(-1, 3)[sel].contiguous(), d_ref.reshape(-1, 3)[sel].contiguous()
    out = dict(name=request.param, H=H, W=W, focal=focal, pose=pose, near=near, far=far, sel=sel, u_t=u_t, u_pdf=u_pdf, img=img,
               o_full=o_ref.numpy(), d_full=d_ref.numpy(), o=o_s.numpy(), d=d_s.numpy(), t=t_ref.numpy())
    torch.set_num_threads(os.cpu_count() or 8)
    for tag, bias in (("keras", 0.0), ("biased", 0.1)):
        wc, wf = O.init_weights(42, bias), O.init_weights(43, bias)
        with torch.no_grad():
            rgbs, depths, ws, preds, t_all = O.forward_pass(wc, wf, o_s, d_s, t_ref, 10, 4, NF, torch.from_numpy(u_pdf))
        out[tag] = dict(wc=wc, wf=wf, rgb_c=rgbs[0].numpy(), rgb_f=rgbs[1].numpy(), w_c=ws[0].numpy(), pred_c=preds[0].numpy(),
                        pred_f=preds[1].numpy(), t_all=t_all.numpy(), depth_f=depths[1].numpy())
    return out


def _trainer(nk, w, precision, training=False, **kw):
    mc = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    mf = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    mc.set_flat_weights(O.flatten_weights(w["wc"]))
    mf.set_flat_weights(O.flatten_weights(w["wf"]))
    tr = nk.NeRFTrainer(mc, mf, B, NC, NF, 10, 4, precision=precision, **kw)
    if training:
        tr.compile(nk.Adam(5e-4), nk.MeanSquaredError())
    else:
        tr.build()
    return tr


def _stats(e):
    e = np.asarray(e, np.float64)
    return dict(max=float(e.max()), p999=float(np.quantile(e, 0.999)), p99=float(np.quantile(e, 0.99)), mean=float(e.mean()),
                over_2e3=int((e > 2e-3).sum()), pixels=int(e.size))


def _dump(name, obj):
    try:
        os.makedirs(REPORT, exist_ok=True)
        with open(os.path.join(REPORT, name), "w") as f:
            json.dump(obj, f, indent=1)
    except OSError:
        pass


def test_rays_and_t_vals_bit_exact_at_baseline_shapes(nk, case):
    o, d = nk.get_rays(case["H"], case["W"], case["focal"], case["pose"])
    assert np.array_equal(o.cpu().numpy(), case["o_full"]) and np.array_equal(d.cpu().numpy(), case["d_full"])
    t = nk.generate_t_vals(case["near"], case["far"], B, NC, True, u=case["u_t"])
    assert np.array_equal(t.cpu().numpy(), case["t"])


def test_fp32_render_vs_oracle_at_baseline_shapes(nk, case):
    for tag in ("keras", "biased"):
        ref = case[tag]
        tr = _trainer(nk, ref, nk.PRECISION_FP32)
        rgbs, depths, ws, preds, t_all = tr.forward_pass(case["o"], case["d"], case["t"], u_pdf=case["u_pdf"], return_t_all=True)
        assert np.abs(rgbs[0].cpu().numpy() - ref["rgb_c"]).max() <= 1e-5, tag
        assert np.abs(ws[0].cpu().numpy() - ref["w_c"]).max() <= 1e-5, tag
        assert np.abs(preds[0].cpu().numpy() - ref["pred_c"]).max() <= 1e-4, tag
        # fine stage at the oracle's own sample positions: 1e-5; end to end (inverse CDF re-drawn from fp32 weights that
        # differ in the last bits): 1e-4
        pf = tr.mlp_forward_rays("fine", case["o"], case["d"], ref["t_all"])
        rgb_f = nk.volume_render(pf, ref["t_all"])[0].cpu().numpy()
        assert np.abs(rgb_f - ref["rgb_f"]).max() <= 1e-5, tag
        assert np.abs(rgbs[1].cpu().numpy() - ref["rgb_f"]).max() <= 3e-4, tag


def _bf16_report(nk, case, ref, **kw):
    tr = _trainer(nk, ref, nk.PRECISION_BF16_TC, **kw)
    rgbs, _, _, preds = tr.forward_pass(case["o"], case["d"], case["t"], u_pdf=case["u_pdf"])
    pf = tr.mlp_forward_rays("fine", case["o"], case["d"], ref["t_all"])
    rgb_f_same_t = nk.volume_render(pf, ref["t_all"])[0].cpu().numpy()
    raw = {"coarse": float(np.abs(preds[0].cpu().numpy() - ref["pred_c"])[:, :, :].max()),
           "fine_at_oracle_samples": float(np.abs(pf.cpu().numpy() - ref["pred_f"]).max())}
    rep = {}
    for key, got, want, pref, rk in (("coarse_end_to_end", rgbs[0].cpu().numpy(), ref["rgb_c"], ref["pred_c"], "coarse"),
                                     ("fine_at_oracle_samples", rgb_f_same_t, ref["rgb_f"], ref["pred_f"], "fine_at_oracle_samples"),
                                     ("fine_end_to_end", rgbs[1].cpu().numpy(), ref["rgb_f"], ref["pred_f"], "fine_at_oracle_samples")):
        e = np.abs(got - want).max(axis=1)
        near0 = np.abs(pref[:, -1, 3]) <= raw[rk]            # last raw sigma of the ORACLE within the measured bf16 error of 0
        st = _stats(e)
        st.update(psnr_delta_db=abs(psnr_db(got, case["img"]) - psnr_db(want, case["img"])),
                  rays_with_last_sigma_within_bf16_error_of_0=int(near0.sum()),
                  over_2e3_outside_that_band=int(((e > 2e-3) & ~near0).sum()), max_outside_that_band=float(e[~near0].max()))
        rep[key] = st
    rep["raw_pred_max_abs_err"] = raw
    return rep


def test_bf16_render_vs_north_star_unmasked_on_keras_init_model(nk, case):
    """The fixed random-init model of north_star (Keras default initialisation), EVERY ray counted.
    (1) exact_far_sigma=True (the last sample's sigma -- the one input in which the reference's compositing is
        discontinuous, data_utils.py:82 -- comes from the fp32 path): every pixel of the coarse render and of the fine
        render at the oracle's sample positions is within 2e-3, PSNR within 0.05 dB.  No exclusions.
    (2) default bf16 path: the pixels over 2e-3 are counted and shown to be rays whose last oracle sigma lies within the
        bf16 error of 0; all others are within 2e-3.
    (3) the fine render END TO END re-draws its samples through the inverse CDF of coarse weights that differ by bf16
        rounding: samples move, single pixels change by up to ~1e-2, the image does not (PSNR within 0.05 dB)."""
    ref = case["keras"]
    rep = {"default": _bf16_report(nk, case, ref), "exact_far_sigma": _bf16_report(nk, case, ref, exact_far_sigma=True)}
    _dump(f"r2_parity_bf16_keras_init_{case['name']}.json", rep)
    for key in ("coarse_end_to_end", "fine_at_oracle_samples"):
        st = rep["exact_far_sigma"][key]
        assert st["max"] <= 2e-3 and st["over_2e3"] == 0 and st["psnr_delta_db"] <= 0.05, (key, st)
        st = rep["default"][key]
        assert st["max_outside_that_band"] <= 2e-3 and st["over_2e3_outside_that_band"] == 0, (key, st)
        assert st["over_2e3"] <= st["rays_with_last_sigma_within_bf16_error_of_0"] and st["over_2e3"] <= 0.005 * B, (key, st)
        assert st["p999"] <= 2e-3 or st["over_2e3"] > 4, (key, st)
        assert st["psnr_delta_db"] <= 0.05, (key, st)
    for mode in ("default", "exact_far_sigma"):
        st = rep[mode]["fine_end_to_end"]
        assert st["psnr_delta_db"] <= 0.05 and st["mean"] <= 2e-3, (mode, st)


def test_bf16_render_deviation_on_biased_model_reported_unmasked(nk, case):
    """Stress model (biases U(-0.1, 0.1), so raw sigma hovers around 0 on every ray).  The reference's delta = 1e10 on the
    last sample (data_utils.py:82) makes a ray's colour DISCONTINUOUS in its last raw sigma at 0: alpha jumps from 0 to 1.
    No reduced-precision MLP can hold such rays to 2e-3.  Nothing is masked here: the test counts the pixels over the
    bound, shows that they are the rays whose last oracle sigma lies within the bf16 error of 0, and bounds the rest."""
    ref = case["biased"]
    tr = _trainer(nk, ref, nk.PRECISION_BF16_TC)
    pc = tr.mlp_forward_rays("coarse", case["o"], case["d"], case["t"])
    pf = tr.mlp_forward_rays("fine", case["o"], case["d"], ref["t_all"])
    rep = {}
    for key, p, t, want, pref in (("coarse", pc, case["t"], ref["rgb_c"], ref["pred_c"]),
                                  ("fine_at_oracle_samples", pf, ref["t_all"], ref["rgb_f"], ref["pred_f"])):
        got = nk.volume_render(p, t)[0].cpu().numpy()
        e = np.abs(got - want).max(axis=1)
        raw_err = float(np.abs(p.cpu().numpy() - pref).max())
        near0 = np.abs(pref[:, -1, 3]) <= raw_err                      # last raw sigma within the measured bf16 error of 0
        st = _stats(e)
        st.update(psnr_delta_db=abs(psnr_db(got, case["img"]) - psnr_db(want, case["img"])), raw_pred_max_abs_err=raw_err,
                  rays_with_last_sigma_within_bf16_error_of_0=int(near0.sum()),
                  over_2e3_outside_that_band=int(((e > 2e-3) & ~near0).sum()),
                  max_outside_that_band=float(e[~near0].max()))
        rep[key] = st
        # measured: Lego-shaped rays 7e-4 outside the band; Fern-shaped rays (coordinates up to |x| = 12, raw-prediction error
        # 9e-3) 3.5e-3 with a mean of 3e-3 -- this stress model is NOT the north_star one, it bounds the degradation
        assert st["max_outside_that_band"] <= 5e-3 and st["psnr_delta_db"] <= 0.05, (key, st)
    _dump(f"r2_parity_bf16_biased_{case['name']}.json", rep)


def _per_tensor(flat, shapes):
    out, off = [], 0
    for net in ("coarse", "fine"):
        for role, fi, fo in shapes:
            out.append((f"{net}/{role}/W", flat[off:off + fi * fo])); off += fi * fo
            out.append((f"{net}/{role}/b", flat[off:off + fo])); off += fo
    assert off == flat.size
    return out


@pytest.mark.parametrize("rays", ["far_sigma_decided", "all"])
@pytest.mark.parametrize("stop_grad", [True, False])
def test_gradients_1024_rays_per_tensor_vs_oracle(nk, case, stop_grad, rays):
    """models.py:94-106 at 1024 rays x (64 + 128): every weight / bias tensor of both nets against torch.autograd on the
    oracle, for the reference's semantics (no stop-gradient on the fine samples) and for the stopped variant.

    rays = "far_sigma_decided": the 1024 rays are drawn from those whose LAST sample's raw sigma (oracle, both nets) is
    further than 0.01 from 0.  The reference puts delta = 1e10 on that sample (data_utils.py:82): its sign decides
    whether the ray ends on an opaque far wall, and with it the sigma gradient of EVERY sample of the ray.  A handful of
    rays inside the bf16 error band of 0 (26 of 4096 here) otherwise dominate the heavily cancelling sums -- a constant
    shift of sigma by -3e-3 changes the fp32 oracle's own sum of d loss / d sigma by +10.7 %.
    rays = "all": the first 1024 rays, nothing excluded; the deviation is reported and bounded loosely."""
    if case["name"] != "lego":
        pytest.skip("one scene is enough for the 1024-ray gradient check (the oracle backward takes ~10 s)")
    n = 1024
    ref = case["keras"]
    if rays == "all":
        idx = np.arange(n)
    else:
        ok = (np.abs(ref["pred_c"][:, -1, 3]) > 0.01) & (np.abs(ref["pred_f"][:, -1, 3]) > 0.01)
        idx = np.nonzero(ok)[0][:n]
        assert idx.size == n
    img, o, d, t, u = (torch.from_numpy(np.ascontiguousarray(case[k][idx])) for k in ("img", "o", "d", "t", "u_pdf"))
    wc, wf = O.init_weights(42, 0.0), O.init_weights(43, 0.0)
    grads, metrics = compute_grads(wc, wf, img, o, d, t, 10, 4, NF, u, stop_grad_samples=stop_grad)
    g_ref = np.concatenate([g.numpy().reshape(-1) for g in grads])
    tr = _trainer(nk, ref, nk.PRECISION_BF16_TC, training=True, stop_grad_samples=stop_grad, use_cuda_graph=False)
    from nerf_keras_b200 import _lib
    m = torch.empty(3, device="cuda")
    args = [x.cuda() for x in (img, o, d, t, u)]
    _lib.check(_lib.lib().nerf_train_forward_backward(tr._ctx.handle, *[a.data_ptr() for a in args], n, m.data_ptr(),
                                                      torch.cuda.current_stream().cuda_stream), "fwd_bwd")
    g_got = tr._ctx.grad_tensor().cpu().numpy()
    mm = m.cpu().numpy()
    shapes = O.layer_shapes()
    rep = {"loss_coarse": [float(mm[0]), metrics["loss_coarse"]], "loss": [float(mm[1]), metrics["loss"]]}
    for (name, a), (_, b) in zip(_per_tensor(g_got, shapes), _per_tensor(g_ref, shapes)):
        na, nb = np.linalg.norm(a.astype(np.float64)), np.linalg.norm(b.astype(np.float64))
        rep[name] = dict(cos=float(a.astype(np.float64) @ b.astype(np.float64) / (na * nb + 1e-30)), norm_ratio=float(na / (nb + 1e-30)),
                         ref_norm=float(nb))
    _dump(f"r2_grad_per_tensor_stop{int(stop_grad)}_{rays}.json", rep)
    fine = {k: v for k, v in rep.items() if k.startswith("fine/")}
    coarse = {k: v for k, v in rep.items() if k.startswith("coarse/")}
    tight = rays == "far_sigma_decided"
    assert abs(mm[0] - metrics["loss_coarse"]) <= (1e-4 if tight else 5e-4) and abs(mm[1] - metrics["loss"]) <= 1e-3
    # the fine net's gradient does not pass through the inverse CDF: every tensor is held in both variants
    lim_cos, lim_ratio = (0.99, 0.03) if tight else (0.985, 0.08)
    for k, v in fine.items():
        assert v["cos"] >= lim_cos and abs(v["norm_ratio"] - 1) <= lim_ratio, (k, v)
    if stop_grad:
        for k, v in coarse.items():
            assert v["cos"] >= (0.985 if tight else 0.98) and abs(v["norm_ratio"] - 1) <= lim_ratio, (k, v)
    else:
        # Un-stopped term (models.py:166-175).  Only the tensors above the sigma head receive it (feature / ddir / rgb do
        # not: held tightly).  It dominates the coarse trunk's gradient (norms 10-30x the stopped ones) and is
        # ill-conditioned IN THE ORACLE ITSELF: a 1e-4 relative change of the coarse weights turns it by tens of degrees
        # (DESIGN.md "Known divergences"), so two independently rounded forward passes cannot agree on it.  It is
        # pinned with the coarse weights held identical in tests/test_gpu_train.py; here: finite, right order of magnitude.
        for k, v in coarse.items():
            if k.split("/")[1] in ("feature", "ddir", "rgb"):
                assert v["cos"] >= 0.99 and abs(v["norm_ratio"] - 1) <= lim_ratio, (k, v)
            else:
                assert np.isfinite(v["cos"]) and 0.1 <= v["norm_ratio"] <= 10.0, (k, v)
