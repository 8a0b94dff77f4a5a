"""BASELINE.json's full benchmark shape (config/lego_batch_h256.json: 4096-ray batches, 64 + 128 samples per ray, 800x800
views) through size-independent properties -- the oracle takes minutes at this size, these checks take seconds:
bit-exact ray generation against the oracle, sortedness / multiset inclusion of the merged samples, compositing
invariants, idempotence, invariance to how the batch is split, linearity of the mean-loss gradient in the batch."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu
B, NC, NF, H, W = 4096, 64, 128, 800, 800
FOCAL = 0.5 * W / np.tan(0.5 * 0.6911112)


@pytest.fixture(scope="module")
def nk():
    import nerf_keras_b200 as nk
    return nk


@pytest.fixture(scope="module")
def scene(nk):
    pose = nk.pose_spherical(35.0, -30.0, 4.0)
    o, d = nk.get_rays(H, W, FOCAL, pose)
    g = torch.Generator(device="cpu").manual_seed(5)
    pick = torch.randperm(H * W, generator=g)[:B].cuda()
    o, d = o.reshape(-1, 3)[pick].contiguous(), d.reshape(-1, 3)[pick].contiguous()
    u_t = np.random.default_rng(3).random(NC, dtype=np.float32)
    t = nk.generate_t_vals(2.0, 6.0, B, NC, True, u=u_t)
    u_pdf = torch.from_numpy(np.random.default_rng(4).random((B, NF), dtype=np.float32)).cuda()
    img = torch.from_numpy(np.random.default_rng(1).random((B, 3), dtype=np.float32)).cuda()
    return dict(pose=pose, o=o, d=d, t=t, u=u_pdf, img=img)


def _trainer(nk, training=False, seed=42):
    nk.set_random_seed(seed)
    c, f = nk.create_nerf_complete_model(8, 256, 4, 10, 4), nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    for m, s in ((c, 1), (f, 2)):                          # non-zero biases so that every bias path is live
        w = m.get_flat_weights()
        m.set_flat_weights(w + np.random.default_rng(s).uniform(-0.02, 0.02, w.shape).astype(np.float32))
    tr = nk.NeRFTrainer(c, f, B, NC, NF, 10, 4)
    if training:
        tr.compile(nk.Adam(5e-4), nk.MeanSquaredError())
    else:
        tr.build()
    return tr


def test_full_frame_rays_and_t_vals_bit_exact(nk, scene):
    o, d = nk.get_rays(H, W, FOCAL, scene["pose"])
    o_ref, d_ref = O.get_rays(H, W, float(np.float32(FOCAL)), torch.as_tensor(np.asarray(scene["pose"], dtype=np.float32)))
    assert np.array_equal(o.cpu().numpy(), o_ref.numpy()) and np.array_equal(d.cpu().numpy(), d_ref.numpy())
    u_t = np.random.default_rng(3).random(NC, dtype=np.float32)
    t = nk.generate_t_vals(2.0, 6.0, H * W, NC, True, u=u_t)           # the 164 MB array the reference materialises
    t_ref = O.generate_t_vals(2.0, 6.0, 4, NC, True, u=torch.from_numpy(u_t)).numpy()
    tt = t.cpu().numpy()
    assert tt.shape == (H * W, NC) and np.array_equal(tt[:4], t_ref)
    assert np.array_equal(tt, np.broadcast_to(tt[0], tt.shape))       # one shared jitter vector (Q1)
    assert np.all(np.diff(tt[0]) > 0)


def test_forward_pass_invariants_at_full_size(nk, scene):
    tr = _trainer(nk)
    s = scene
    (rgb_c, rgb_f), (dep_c, dep_f), (w_c, w_f), (p_c, p_f), t_all = tr.forward_pass(s["o"], s["d"], s["t"], u_pdf=s["u"],
                                                                                    return_t_all=True)
    assert t_all.shape == (B, NC + NF) and p_f.shape == (B, NC + NF, 4)
    # sortedness and multiset inclusion: the merged samples are sorted and contain every coarse sample (models.py:167)
    ta = t_all.cpu().numpy()
    assert np.all(np.diff(ta, axis=1) >= 0)
    tc = s["t"].cpu().numpy()
    pos = np.array([np.searchsorted(ta[i], tc[i]) for i in range(0, B, 97)])
    assert np.array_equal(np.take_along_axis(ta[::97], pos, 1), tc[::97])
    mids = 0.5 * (tc[:, 1:] + tc[:, :-1])
    assert ta.min() >= tc.min() and np.all(ta.max(1) <= np.maximum(tc[:, -1], mids[:, -1]) + 1e-6)
    # compositing invariants (data_utils.py:75-98): weights in [0,1], sum <= 1, colours in [0,1], depth inside the samples
    for w, rgb, dep, t in ((w_c, rgb_c, dep_c, s["t"]), (w_f, rgb_f, dep_f, t_all)):
        w, rgb, dep, t = w.cpu().numpy(), rgb.cpu().numpy(), dep.cpu().numpy(), t.cpu().numpy()
        assert w.min() >= 0 and w.max() <= 1 + 1e-6 and np.all(w.sum(1) <= 1 + 1e-4)
        assert rgb.min() >= 0 and rgb.max() <= 1 + 1e-5
        assert np.all(dep >= -1e-5) and np.all(dep <= t[:, -1] * (1 + 1e-5))
        assert np.isfinite(w).all() and np.isfinite(rgb).all()
    # idempotence: same inputs, same bits
    again = tr.forward_pass(s["o"], s["d"], s["t"], u_pdf=s["u"])
    assert torch.equal(again[0][1], rgb_f) and torch.equal(again[3][1], p_f)
    # invariance to the batch split: every ray is independent of its tile (forward_pass_with_minibatch, models.py:178-225)
    split = tr.forward_pass_with_minibatch(s["o"], s["d"], s["t"], batch_size=1000, u_pdf=s["u"])
    assert torch.equal(split[0][1], rgb_f) and torch.equal(split[1][1], dep_f) and torch.equal(split[3][0], p_c)
    # a permutation of the rays permutes the outputs
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0)).cuda()
    pr = tr.forward_pass(s["o"][perm].contiguous(), s["d"][perm].contiguous(), s["t"][perm].contiguous(),
                         u_pdf=s["u"][perm].contiguous())
    assert torch.equal(pr[0][1], rgb_f[perm]) and torch.equal(pr[2][1], w_f[perm])


def _grads(tr, img, o, d, t, u):
    from nerf_keras_b200 import _lib
    from nerf_keras_b200.models import _ptr, _stream
    n = o.shape[0]
    metrics = torch.empty(3, device="cuda")
    _lib.check(_lib.lib().nerf_train_forward_backward(tr._ctx.handle, _ptr(img), _ptr(o), _ptr(d), _ptr(t), _ptr(u), n,
                                                      _ptr(metrics), _stream()), "train_forward_backward")
    return tr._ctx.grad_tensor().clone(), metrics.cpu().numpy()


def test_mean_loss_gradient_is_linear_in_the_batch(nk, scene):
    """grad(mean loss over 4096 rays) = mean of the gradients of the two 2048-ray halves; the metrics average the same
    way.  Exercises the whole training path (both nets, reference gradient semantics) at the benchmark shape."""
    tr = _trainer(nk, training=True)
    s = scene
    c = lambda x, a, b: x[a:b].contiguous()
    g_full, m_full = _grads(tr, s["img"], s["o"], s["d"], s["t"], s["u"])
    h = B // 2
    g1, m1 = _grads(tr, *(c(s[k], 0, h) for k in ("img", "o", "d", "t", "u")))
    g2, m2 = _grads(tr, *(c(s[k], h, B) for k in ("img", "o", "d", "t", "u")))
    g_sum = 0.5 * (g1 + g2)
    assert torch.isfinite(g_full).all() and float(g_full.norm()) > 0
    rel = float((g_full - g_sum).norm() / g_full.norm())
    assert rel <= 2e-3, rel                               # fp32 atomics order + bf16 dZ rounding of the 1/B scale
    np.testing.assert_allclose(m_full[:2], 0.5 * (m1[:2] + m2[:2]), rtol=1e-5)
    # determinism of the value path: same batch twice gives the same metrics bit for bit
    _, m_again = _grads(tr, s["img"], s["o"], s["d"], s["t"], s["u"])
    assert np.array_equal(m_full, m_again)


def test_train_step_at_full_size_moves_weights_by_about_lr(nk, scene):
    """First Adam step from zero state moves every parameter with a non-negligible gradient by ~lr (Keras form,
    train_lego.py:149-151): a checksum over all 1.19 M parameters."""
    tr = _trainer(nk, training=True)
    s = scene
    w0 = np.concatenate([tr.coarse_model.get_flat_weights(), tr.fine_model.get_flat_weights()])
    g, _ = _grads(tr, s["img"], s["o"], s["d"], s["t"], s["u"])
    out = tr.train_step((s["img"], (s["o"], s["d"], s["t"])), u_pdf=s["u"])
    w1 = np.concatenate([tr.coarse_model.get_flat_weights(), tr.fine_model.get_flat_weights()])
    g = g.cpu().numpy()
    step = w1 - w0
    # step 1 from zero state: m = 0.1 g, v = 0.001 g^2, alpha_1 = lr sqrt(0.001) / 0.1  =>  -lr g / (|g| + 1e-7 / sqrt(0.001))
    want = -5e-4 * g / (np.abs(g) + 1e-7 / np.sqrt(1e-3))
    big = np.abs(g) > 1e-5
    assert big.mean() > 0.3
    np.testing.assert_allclose(step[big], want[big], rtol=5e-3, atol=2e-7)
    np.testing.assert_allclose(step, want, atol=2e-5)      # tiny gradients: atomics-order noise moves g itself
    assert np.all(np.abs(step) <= 5e-4 * 1.001)
    assert np.isfinite(float(out["loss"])) and np.isfinite(float(out["psnr"]))
