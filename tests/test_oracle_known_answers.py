"""Known-answer tests that pin the oracle restatement.

The reference ships no tests, so these analytic cases -- each derived from the cited
reference lines -- are what anchors the oracle ("parity unpinned" otherwise)."""
import math

import numpy as np
import torch

import oracle as O


def test_get_rays_2x2_identity():
    # data_utils.py:41-45 with H=W=2, focal=1, pose=I
    o, d = O.get_rays(2, 2, 1.0, np.eye(4, dtype=np.float32))
    exp = np.array([[[-1, 1, -1], [0, 1, -1]], [[-1, 0, -1], [0, 0, -1]]], dtype=np.float32)
    assert np.array_equal(d.numpy(), exp)
    assert np.array_equal(o.numpy(), np.zeros((2, 2, 3), np.float32))


def test_get_rays_rotation_translation():
    pose = O.pose_spherical(30.0, -30.0, 4.0)
    o, d = O.get_rays(5, 7, 3.0, pose)
    assert o.shape == (5, 7, 3) and d.shape == (5, 7, 3)
    assert np.array_equal(o.numpy()[2, 3], pose[:3, 3])
    # float64 check of one pixel (h=1,w=5)
    dc = np.array([(5 - 3.5) / 3.0, -(1 - 2.5) / 3.0, -1.0])
    np.testing.assert_allclose(d.numpy()[1, 5], pose[:3, :3].astype(np.float64) @ dc, rtol=1e-6, atol=1e-7)
    # origin lies at radius 4
    assert abs(np.linalg.norm(pose[:3, 3]) - 4.0) < 1e-5


def test_encode_position_zero_and_widths():
    e = O.encode_position(torch.zeros(2, 3), 10).numpy()
    assert e.shape == (2, 63)
    exp = np.concatenate([np.zeros(3)] + [np.array([0, 0, 0, 1, 1, 1.0])] * 10)
    assert np.array_equal(e[0], exp.astype(np.float32))
    assert O.encode_position(torch.zeros(2, 3), 4).shape == (2, 27)
    x = torch.tensor([[0.3, -1.2, 2.5]])
    e = O.encode_position(x, 10).numpy()[0]
    np.testing.assert_allclose(e[3 + 6 * 3:3 + 6 * 3 + 3], np.sin(8.0 * x.numpy()[0].astype(np.float64)), atol=2e-6)
    np.testing.assert_allclose(e[3 + 6 * 3 + 3:3 + 6 * 4], np.cos(8.0 * x.numpy()[0].astype(np.float64)), atol=2e-6)


def test_generate_t_vals():
    t = O.generate_t_vals(2.0, 6.0, 5, 64, rand_sampling=False).numpy()
    assert t.shape == (5, 64)
    assert (t == t[0]).all()
    assert t[0, 0] == np.float32(2.0) and t[0, 63] == np.float32(6.0)
    np.testing.assert_allclose(np.diff(t[0]), 4.0 / 63, rtol=1e-5)
    u = np.random.default_rng(3).random(64, dtype=np.float32)
    tj = O.generate_t_vals(2.0, 6.0, 5, 64, rand_sampling=True, u=u).numpy()
    np.testing.assert_allclose(tj[0] - t[0], u * 4.0 / 64.0, atol=5e-7)
    assert (tj == tj[0]).all()  # one jitter vector shared by every ray (Q1)


def test_volume_render_known_answers():
    B, N = 3, 16
    t = O.generate_t_vals(2.0, 6.0, B, N, rand_sampling=False)
    # preds = 0 -> sigma = 0 -> nothing absorbed
    rgb, depth, w = O.volume_render(torch.zeros(B, N, 4), t)
    assert float(w.abs().max()) == 0 and float(rgb.abs().max()) == 0 and float(depth.abs().max()) == 0
    # one opaque sample k
    k = 5
    preds = torch.full((B, N, 4), -1.0)
    preds[:, k, 3] = 1e4
    preds[:, k, :3] = torch.tensor([0.5, -0.25, 2.0])
    rgb, depth, w = O.volume_render(preds, t)
    np.testing.assert_allclose(rgb.numpy()[0], torch.sigmoid(torch.tensor([0.5, -0.25, 2.0])).numpy(), atol=1e-6)
    np.testing.assert_allclose(depth.numpy(), t.numpy()[:, k], rtol=1e-6)
    assert abs(float(w[0, k]) - 1.0) < 1e-6
    # sigma > 0 on the last sample -> alpha = 1 there (delta = 1e10), sum(w) <= 1
    preds = torch.zeros(B, N, 4)
    preds[:, -1, 3] = 1e-3
    _, _, w = O.volume_render(preds, t)
    assert float(w[0, -1]) == 1.0
    r = torch.randn(B, N, 4, generator=torch.Generator().manual_seed(0))
    _, _, w = O.volume_render(r, t)
    assert float(w.sum(-1).max()) <= 1.0 + 1e-6


def test_sample_pdf_uniform_weights():
    Nc, Nf, B = 64, 128, 4
    t = O.generate_t_vals(2.0, 6.0, B, Nc, rand_sampling=False)
    t_mid = 0.5 * (t[:, 1:] + t[:, :-1])
    u = torch.from_numpy(np.random.default_rng(4).random((B, Nf), dtype=np.float32))
    s = O.sample_pdf(t_mid, torch.ones(B, Nc), Nf, u=u).numpy()
    tm = t_mid.numpy().astype(np.float64)
    un = u.numpy().astype(np.float64)
    k = np.floor(un * Nc).astype(int)
    frac = un * Nc - k
    kb = np.minimum(k, Nc - 2)
    ka = np.minimum(k + 1, Nc - 2)
    exp = np.take_along_axis(tm, kb, 1) + frac * (np.take_along_axis(tm, ka, 1) - np.take_along_axis(tm, kb, 1))
    # away from bin edges (cdf rounding may move an edge by an ulp)
    ok = (frac > 1e-3) & (frac < 1 - 1e-3)
    np.testing.assert_allclose(s[ok], exp[ok], atol=2e-5)
    assert s.min() >= tm[0, 0] - 1e-6 and s.max() <= tm[0, Nc - 2] + 1e-6


def test_mlp_shapes_and_param_count():
    shapes = dict((r, (i, o)) for r, i, o in O.layer_shapes())
    assert shapes["d0"] == (63, 256) and shapes["d5"] == (319, 256) and shapes["d4"] == (256, 256)
    assert shapes["sigma"] == (256, 1) and shapes["feature"] == (256, 256)
    assert shapes["ddir"] == (283, 128) and shapes["rgb"] == (128, 3)
    assert O.param_count() == 595844
    w = O.init_weights(42, bias_range=0.1)
    blob = O.flatten_weights(w)
    assert blob.size == 595844
    w2 = O.unflatten_weights(blob)
    for r in w:
        assert torch.equal(w[r]["W"], w2[r]["W"]) and torch.equal(w[r]["b"], w2[r]["b"])


def test_mlp_channel_order_and_skip_layout():
    # sigma is output channel 3 (models.py:59); W5 rows 0..255 hidden, 256..318 encoding (:38-39)
    w = O.init_weights(1)
    for r in w:
        w[r]["W"].zero_(); w[r]["b"].zero_()
    w["sigma"]["b"][0] = 7.0
    w["rgb"]["b"][:] = torch.tensor([1.0, 2.0, 3.0])
    out = O.nerf_mlp(w, torch.randn(2, 5, 63), torch.randn(2, 5, 27))
    assert out.shape == (2, 5, 4)
    assert torch.equal(out[0, 0], torch.tensor([1.0, 2.0, 3.0, 7.0]))
    # only the encoding rows of d5 are non-zero -> h5 depends on enc only
    w = O.init_weights(2)
    enc = torch.randn(1, 1, 63)
    w["d5"]["W"][:256].zero_()
    x = enc
    out_a = O.nerf_mlp(w, enc, torch.zeros(1, 1, 27))
    for i in range(5):
        w[f"d{i}"]["W"].mul_(-3.0)  # changing the trunk below the skip must not matter now
    out_b = O.nerf_mlp(w, enc, torch.zeros(1, 1, 27))
    assert torch.allclose(out_a, out_b)


def test_metrics_definitions():
    a = torch.rand(8, 3, generator=torch.Generator().manual_seed(1))
    b = torch.rand(8, 3, generator=torch.Generator().manual_seed(2))
    m = float(O.mse(a, b))
    assert abs(m - float(((a - b) ** 2).sum() / 24)) < 1e-7
    assert abs(float(O.psnr(a, b)) - (-10 * math.log10(m))) < 1e-4


def test_keras_adam_first_step():
    g = torch.tensor([1e-2, -3.0, 5e-2])
    p = torch.zeros(3)
    opt = O.KerasAdam([p], learning_rate=5e-4)
    opt.apply_gradients([g])
    # step 1 from zero state ~ -lr*sign(g) for |g| >> eps/sqrt(1-b2)
    np.testing.assert_allclose(p.numpy(), -5e-4 * np.sign(g.numpy()), rtol=2e-3)
    exp = -5e-4 * math.sqrt(1 - 0.999) / (1 - 0.9) * (0.1 * 1e-2) / (math.sqrt(0.001 * 1e-4) + 1e-7)
    assert abs(float(p[0]) - exp) < 1e-9


def test_forward_pass_shapes_and_q5_gradient_path():
    B, Nc, Nf = 6, 16, 32
    wc, wf = O.init_weights(42, 0.1), O.init_weights(43, 0.1)
    pose = O.pose_spherical(20.0, -30.0, 4.0)
    o, d = O.get_rays(4, 4, 5.0, pose)
    o, d = o.reshape(-1, 3)[:B], d.reshape(-1, 3)[:B]
    t = O.generate_t_vals(2.0, 6.0, B, Nc, True, u=np.random.default_rng(3).random(Nc, dtype=np.float32))
    u = torch.from_numpy(np.random.default_rng(4).random((B, Nf), dtype=np.float32))
    rgbs, depths, ws, preds, t_all = O.forward_pass(wc, wf, o, d, t, 10, 4, Nf, u)
    assert rgbs[0].shape == (B, 3) and rgbs[1].shape == (B, 3)
    assert ws[0].shape == (B, Nc) and ws[1].shape == (B, Nc + Nf)
    assert preds[1].shape == (B, Nc + Nf, 4)
    assert bool((t_all[:, 1:] >= t_all[:, :-1]).all())
    img = torch.rand(B, 3, generator=torch.Generator().manual_seed(5))
    from oracle.models_ref import compute_grads
    g_ref, m = compute_grads(wc, wf, img, o, d, t, 10, 4, Nf, u, stop_grad_samples=False)
    g_stop, _ = compute_grads(wc, wf, img, o, d, t, 10, 4, Nf, u, stop_grad_samples=True)
    n = len(g_ref) // 2
    # fine-net grads are identical; coarse-net grads differ (the reference has no stop_gradient, Q5)
    for a, b in zip(g_ref[n:], g_stop[n:]):
        assert torch.allclose(a, b, atol=1e-7)
    assert any(not torch.allclose(a, b, atol=1e-9) for a, b in zip(g_ref[:n], g_stop[:n]))
    assert m["psnr"] == m["psnr"]
