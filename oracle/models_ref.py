"""Oracle restatement of the reference's `models.py` (torch-CPU fp32) + the Keras Adam it uses.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PARITY UNPINNED (no reference tests exist).

Weights are an explicit dict keyed by layer role, in creation order
d0..d{L-1}, sigma, feature, ddir, rgb, each {"W": (in,out), "b": (out,)} -- the Keras
`Dense` convention `y = x @ W + b` (models.py:29-59).  The flat blob used at the C-ABI is
the concatenation, role by role, of W (row-major (in,out)) then b.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .data_utils_ref import encode_position, sample_pdf, sample_rays, volume_render

F32 = torch.float32
Weights = Dict[str, Dict[str, torch.Tensor]]


def LAYER_ROLES(num_layers: int = 8) -> List[str]:
    return [f"d{i}" for i in range(num_layers)] + ["sigma", "feature", "ddir", "rgb"]


def layer_shapes(num_layers=8, hidden_dim=256, skip_layer=4, lxyz=10, ldir=4) -> List[Tuple[str, int, int]]:
    """(role, fan_in, fan_out) in creation order -- models.py:24-62.

    The skip concat is `[h, ray_input]` AFTER layer i's activation when i % skip == 0 and
    i > 0 (:38-39), so the layer FOLLOWING such an i has fan_in = hidden + (3+6*lxyz)."""
    exyz = 3 + 6 * lxyz
    edir = 3 + 6 * ldir
    shapes = []
    fan_in = exyz
    for i in range(num_layers):
        shapes.append((f"d{i}", fan_in, hidden_dim))
        fan_in = hidden_dim
        if i % skip_layer == 0 and i > 0:
            fan_in = hidden_dim + exyz
    trunk_out = fan_in
    shapes.append(("sigma", trunk_out, 1))                 # :42
    shapes.append(("feature", trunk_out, hidden_dim))      # :45
    shapes.append(("ddir", hidden_dim + edir, hidden_dim // 2))  # :48,54
    shapes.append(("rgb", hidden_dim // 2, 3))             # :57
    return shapes


def param_count(**kw) -> int:
    return sum(i * o + o for _, i, o in layer_shapes(**kw))


def init_weights(seed: int, bias_range: float = 0.0, **kw) -> Weights:
    """Keras defaults restated: glorot-uniform kernels U(+-sqrt(6/(in+out))), zero biases;
    `bias_range` > 0 draws biases U(-r, r) instead so the bias paths are exercised."""
    rng = np.random.default_rng(seed)
    w: Weights = {}
    for role, fi, fo in layer_shapes(**kw):
        lim = math.sqrt(6.0 / (fi + fo))
        W = rng.uniform(-lim, lim, size=(fi, fo)).astype(np.float32)
        if bias_range > 0:
            b = rng.uniform(-bias_range, bias_range, size=(fo,)).astype(np.float32)
        else:
            b = np.zeros((fo,), dtype=np.float32)
        w[role] = {"W": torch.from_numpy(W), "b": torch.from_numpy(b)}
    return w


def flatten_weights(w: Weights) -> np.ndarray:
    parts = []
    for role in w:
        parts.append(w[role]["W"].detach().numpy().reshape(-1))
        parts.append(w[role]["b"].detach().numpy().reshape(-1))
    return np.concatenate(parts).astype(np.float32)


def unflatten_weights(blob: np.ndarray, **kw) -> Weights:
    w: Weights = {}
    off = 0
    for role, fi, fo in layer_shapes(**kw):
        W = blob[off:off + fi * fo].reshape(fi, fo); off += fi * fo
        b = blob[off:off + fo]; off += fo
        w[role] = {"W": torch.from_numpy(np.array(W, dtype=np.float32)),
                   "b": torch.from_numpy(np.array(b, dtype=np.float32))}
    assert off == blob.size
    return w


def nerf_mlp(w: Weights, ray_enc: torch.Tensor, dir_enc: torch.Tensor, num_layers=8, skip_layer=4,
             bn: Optional[dict] = None, training: bool = False) -> torch.Tensor:
    """models.py:24-62 forward.  Output (..., 4) = [r, g, b, sigma] raw (:59).

    bn: optional dict role -> {"gamma","beta","mean","var"} for the BATCH_NORM=true variant
    (:30-33, :49-52; Keras defaults momentum 0.99, eps 1e-3).  Training mode uses batch
    statistics over all leading axes and updates the moving stats in place."""
    x = ray_enc
    for i in range(num_layers):
        p = w[f"d{i}"]
        x = x @ p["W"] + p["b"]
        if bn is not None:
            x = _batch_norm(x, bn[f"d{i}"], training)
        x = torch.relu(x)
        if i % skip_layer == 0 and i > 0:
            x = torch.cat([x, ray_enc], dim=-1)           # [h, enc]  (:38-39)
    sigma = x @ w["sigma"]["W"] + w["sigma"]["b"]          # :42
    feature = x @ w["feature"]["W"] + w["feature"]["b"]    # :45 (linear)
    feature = torch.cat([feature, dir_enc], dim=-1)        # :48
    x = feature @ w["ddir"]["W"] + w["ddir"]["b"]          # :54
    if bn is not None:
        x = _batch_norm(x, bn["ddir"], training)
    x = torch.relu(x)
    rgb = x @ w["rgb"]["W"] + w["rgb"]["b"]                # :57
    return torch.cat([rgb, sigma], dim=-1)                 # :59


def _batch_norm(x, st, training, momentum=0.99, eps=1e-3):
    if training:
        red = tuple(range(x.dim() - 1))
        mean = x.mean(dim=red)
        var = x.var(dim=red, unbiased=False)
        with torch.no_grad():
            st["mean"].mul_(momentum).add_((1 - momentum) * mean)
            st["var"].mul_(momentum).add_((1 - momentum) * var)
    else:
        mean, var = st["mean"], st["var"]
    return (x - mean) * torch.rsqrt(var + eps) * st["gamma"] + st["beta"]


def init_bn(num_layers=8, hidden_dim=256) -> dict:
    bn = {}
    for role, n in [(f"d{i}", hidden_dim) for i in range(num_layers)] + [("ddir", hidden_dim // 2)]:
        bn[role] = {"gamma": torch.ones(n), "beta": torch.zeros(n), "mean": torch.zeros(n), "var": torch.ones(n)}
    return bn


def forward_pass(w_coarse: Weights, w_fine: Weights, ray_origins, ray_directions, t_vals, l_xyz, l_dir,
                 ns_fine: int, u_pdf, training: bool = False, stop_grad_samples: bool = False,
                 num_layers=8, skip_layer=4, bn_coarse=None, bn_fine=None):
    """models.py:151-176.  Returns ((rgb_c,rgb_f),(depth_c,depth_f),(w_c,w_f),(pred_c,pred_f)).

    stop_grad_samples=False is the reference (no stop_gradient on t_fine, quirk Q5)."""
    rays, dirs = sample_rays(ray_origins, ray_directions, t_vals)                  # :152
    rays_enc = encode_position(rays, l_xyz)                                        # :153
    dirs_enc = encode_position(dirs, l_dir)                                        # :154
    pred_c = nerf_mlp(w_coarse, rays_enc, dirs_enc, num_layers, skip_layer, bn_coarse, training)  # :157
    rgb_c, depth_c, weights_c = volume_render(pred_c, t_vals)                      # :164
    t_mid = 0.5 * (t_vals[..., 1:] + t_vals[..., :-1])                             # :165
    t_fine = sample_pdf(t_mid, weights_c, ns_fine, u=u_pdf)                        # :166
    if stop_grad_samples:
        t_fine = t_fine.detach()
    t_all, _ = torch.sort(torch.cat([t_vals, t_fine], dim=-1), dim=-1)            # :167
    rays_f, dirs_f = sample_rays(ray_origins, ray_directions, t_all)               # :169
    rays_f_enc = encode_position(rays_f, l_xyz)                                    # :170
    dirs_f_enc = encode_position(dirs_f, l_dir)                                    # :171
    pred_f = nerf_mlp(w_fine, rays_f_enc, dirs_f_enc, num_layers, skip_layer, bn_fine, training)  # :173
    rgb_f, depth_f, weights_f = volume_render(pred_f, t_all)                       # :175
    return (rgb_c, rgb_f), (depth_c, depth_f), (weights_c, weights_f), (pred_c, pred_f), t_all


def forward_pass_with_minibatch(w_coarse, w_fine, ray_origins, ray_directions, t_vals, l_xyz, l_dir,
                                ns_fine, u_pdf, batch_size=512, **kw):
    """models.py:178-225 -- ray-tile loop over forward_pass + concat of the 8 outputs."""
    outs = []
    n = ray_origins.shape[0]
    for s in range(0, n, batch_size):
        e = min(n, s + batch_size)
        outs.append(forward_pass(w_coarse, w_fine, ray_origins[s:e], ray_directions[s:e], t_vals[s:e],
                                 l_xyz, l_dir, ns_fine, u_pdf[s:e], **kw))
    cat = lambda i, j: torch.cat([o[i][j] for o in outs], dim=0)
    return ((cat(0, 0), cat(0, 1)), (cat(1, 0), cat(1, 1)), (cat(2, 0), cat(2, 1)), (cat(3, 0), cat(3, 1)),
            torch.cat([o[4] for o in outs], dim=0))


def mse(a, b):
    """keras.losses.MeanSquaredError (train_lego.py:154): mean over every element."""
    return torch.mean((a - b) ** 2)


def psnr(a, b, max_val=1.0):
    """keras.ops.psnr (models.py:110): 20 log10(max) - 10 log10(mse)."""
    m = torch.mean((a - b) ** 2)
    return 20.0 * math.log10(max_val) - 10.0 * torch.log10(m)


class KerasAdam:
    """keras.optimizers.Adam(learning_rate) as built at train_lego.py:149-151, Keras-3
    `update_step` form (restated, unverifiable offline): beta1 .9, beta2 .999, eps 1e-7,
    alpha_t = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g^2-v)(1-b2);
    theta -= alpha_t * m / (sqrt(v) + eps)."""

    def __init__(self, params: List[torch.Tensor], learning_rate=5e-4, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.params = params
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = 0

    @torch.no_grad()
    def apply_gradients(self, grads: List[torch.Tensor]):
        self.t += 1
        b1p = self.b1 ** self.t
        b2p = self.b2 ** self.t
        alpha = np.float32(self.lr * math.sqrt(1.0 - b2p) / (1.0 - b1p))
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            m.add_((g - m) * np.float32(1.0 - self.b1))
            v.add_((g * g - v) * np.float32(1.0 - self.b2))
            p.sub_(float(alpha) * m / (torch.sqrt(v) + np.float32(self.eps)))


def _params(w: Weights) -> List[torch.Tensor]:
    out = []
    for role in w:
        out += [w[role]["W"], w[role]["b"]]
    return out


def compute_grads(w_coarse, w_fine, images, ray_origins, ray_directions, t_vals, l_xyz, l_dir, ns_fine, u_pdf,
                  stop_grad_samples=False, **kw):
    """models.py:94-106 -- loss = MSE(rgb_c) + MSE(rgb_f); grads wrt coarse then fine variables."""
    params = _params(w_coarse) + _params(w_fine)
    for p in params:
        p.requires_grad_(True)
        p.grad = None
    rgbs, _, _, _, _ = forward_pass(w_coarse, w_fine, ray_origins, ray_directions, t_vals, l_xyz, l_dir, ns_fine,
                                    u_pdf, training=True, stop_grad_samples=stop_grad_samples, **kw)
    loss_c = mse(images, rgbs[0])
    loss_f = mse(images, rgbs[1])
    loss = loss_c + loss_f
    grads = torch.autograd.grad(loss, params)
    for p in params:
        p.requires_grad_(False)
    ps = psnr(images, rgbs[1].detach())
    return [g.detach() for g in grads], {"loss_coarse": float(loss_c.detach()), "loss": float(loss_f.detach()),
                                          "psnr": float(ps)}


def train_step(w_coarse, w_fine, opt: KerasAdam, images, ray_origins, ray_directions, t_vals, l_xyz, l_dir,
               ns_fine, u_pdf, stop_grad_samples=False, **kw):
    """models.py:88-120.  `loss` is the FINE loss only (:114)."""
    grads, metrics = compute_grads(w_coarse, w_fine, images, ray_origins, ray_directions, t_vals, l_xyz, l_dir,
                                   ns_fine, u_pdf, stop_grad_samples, **kw)
    opt.apply_gradients(grads)
    return metrics


@torch.no_grad()
def test_step(w_coarse, w_fine, images, ray_origins, ray_directions, t_vals, l_xyz, l_dir, ns_fine, u_pdf, **kw):
    """models.py:122-145."""
    rgbs, _, _, _, _ = forward_pass(w_coarse, w_fine, ray_origins, ray_directions, t_vals, l_xyz, l_dir, ns_fine,
                                    u_pdf, training=False, **kw)
    return {"loss_coarse": float(mse(images, rgbs[0])), "loss": float(mse(images, rgbs[1])),
            "psnr": float(psnr(images, rgbs[1]))}
