"""CPU: the oracle reproduces the committed golden fixtures (regression pin of the restatement)."""
import numpy as np
import pytest
import torch

import oracle as O
from tests.util import golden_weights, load_golden


@pytest.mark.parametrize("name", ["lego_small", "fern_small"])
def test_oracle_matches_golden(name):
    g = load_golden(name)
    H, W, Nc, Nf = int(g["H"]), int(g["W"]), int(g["Nc"]), int(g["Nf"])
    o, d = O.get_rays(H, W, g["focal"], g["pose"])
    assert np.array_equal(o.numpy(), g["rays_o_full"]) and np.array_equal(d.numpy(), g["rays_d_full"])
    oo, dd = o.reshape(-1, 3)[g["sel"]], d.reshape(-1, 3)[g["sel"]]
    B = oo.shape[0]
    t = O.generate_t_vals(float(g["near"]), float(g["far"]), B, Nc, True, u=g["u_t"])
    assert np.array_equal(t.numpy(), g["t"])
    assert np.array_equal(O.generate_t_vals(float(g["near"]), float(g["far"]), B, Nc, False).numpy(), g["t_nojit"])
    rays, dirs = O.sample_rays(oo, dd, t)
    assert np.array_equal(rays.numpy(), g["pts"])
    np.testing.assert_allclose(O.encode_position(rays, 10).numpy(), g["enc_x"], atol=1e-6)
    wc, wf = golden_weights(g)
    with torch.no_grad():
        rgbs, depths, ws, preds, t_all = O.forward_pass(wc, wf, oo, dd, t, 10, 4, Nf, torch.from_numpy(g["u_pdf"]))
    np.testing.assert_allclose(preds[0].numpy(), g["pred_c"], atol=2e-5)
    np.testing.assert_allclose(rgbs[0].numpy(), g["rgb_c"], atol=2e-6)
    np.testing.assert_allclose(rgbs[1].numpy(), g["rgb_f"], atol=2e-5)
    np.testing.assert_allclose(t_all.numpy(), g["t_all"], atol=2e-5)
    np.testing.assert_allclose(ws[0].numpy(), g["wt_c"], atol=2e-6)


def test_oracle_minibatch_equals_full():
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    o, d, t, u = map(torch.from_numpy, (g["o"], g["d"], g["t"], g["u_pdf"]))
    with torch.no_grad():
        full = O.forward_pass(wc, wf, o, d, t, 10, 4, int(g["Nf"]), u)
        tiled = O.forward_pass_with_minibatch(wc, wf, o, d, t, 10, 4, int(g["Nf"]), u, batch_size=17)
    np.testing.assert_allclose(full[0][1].numpy(), tiled[0][1].numpy(), atol=1e-6)
    np.testing.assert_allclose(full[3][1].numpy(), tiled[3][1].numpy(), atol=2e-4)  # MKL blocking differs per batch shape


def test_oracle_train_step_reduces_loss():
    g = load_golden("lego_small")
    wc, wf = golden_weights(g)
    o, d, t, u, img = map(torch.from_numpy, (g["o"], g["d"], g["t"], g["u_pdf"], g["img"]))
    from oracle.models_ref import _params
    opt = O.KerasAdam(_params(wc) + _params(wf), learning_rate=5e-4)
    m0 = O.train_step(wc, wf, opt, img, o, d, t, 10, 4, int(g["Nf"]), u)
    np.testing.assert_allclose([m0["loss_coarse"], m0["loss"], m0["psnr"]], g["metrics"], rtol=1e-4)
    for _ in range(5):
        m = O.train_step(wc, wf, opt, img, o, d, t, 10, 4, int(g["Nf"]), u)
    assert m["loss"] < m0["loss"] and m["loss_coarse"] < m0["loss_coarse"]
