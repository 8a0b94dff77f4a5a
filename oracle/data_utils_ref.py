"""Oracle restatement of the reference's `data_utils.py` (torch-CPU, fp32, op for op).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PARITY UNPINNED (no reference tests exist).

Every function cites the reference lines it restates (paths relative to the reference
root).  All arithmetic is float32 with the reference's operation order; torch-CPU eager
elementwise ops are individually IEEE-rounded (no FMA contraction), which is what the
TensorFlow eager/graph CPU kernels do for separately-dispatched ops.  Random draws are
explicit inputs (`u`), because the reference's Keras/TF generators cannot be reproduced.
"""
from __future__ import annotations

import math

import numpy as np
import torch

F32 = torch.float32


def _t(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(F32)
    return torch.as_tensor(np.asarray(x, dtype=np.float32))


def encode_position(x, pos_encode_dims: int) -> torch.Tensor:
    """data_utils.py:7-21 -- [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]."""
    x = _t(x)
    positions = [x]
    for i in range(pos_encode_dims):
        scaled = torch.tensor(2.0 ** i, dtype=F32) * x  # exact power-of-two scale
        positions.append(torch.sin(scaled))
        positions.append(torch.cos(scaled))
    return torch.cat(positions, dim=-1)


def get_rays(height: int, width: int, focal, pose):
    """data_utils.py:23-52 -- pinhole rays, no +0.5 pixel centre, un-normalised directions.

    dir_cam = [(u - W*0.5)/focal, -(v - H*0.5)/focal, -1]  (:41-45, subtract then divide)
    ray_d[i] = ((dc0*R[i,0]) + (dc1*R[i,1])) + (dc2*R[i,2])  (:48-50, products materialised,
    then a 3-element reduce in index order -- no FMA)
    ray_o = pose[:3,-1] broadcast (:47,51).
    """
    pose = _t(pose)
    focal32 = torch.tensor(float(np.float32(focal)), dtype=F32)
    u = torch.arange(width, dtype=F32).view(1, width).expand(height, width)
    v = torch.arange(height, dtype=F32).view(height, 1).expand(height, width)
    half_w = torch.tensor(width * 0.5, dtype=F32)
    half_h = torch.tensor(height * 0.5, dtype=F32)
    tu = (u - half_w) / focal32
    tv = (v - half_h) / focal32
    directions = torch.stack([tu, -tv, -torch.ones_like(tu)], dim=-1)  # (H,W,3)
    camera_matrix = pose[:3, :3]
    translations = pose[:3, -1]
    camera_dirs = directions[..., None, :] * camera_matrix  # (H,W,3,3) [h,w,i,j]=dc[j]*R[i,j]
    ray_directions = (camera_dirs[..., 0] + camera_dirs[..., 1]) + camera_dirs[..., 2]
    ray_origins = translations.expand(ray_directions.shape).contiguous()
    return ray_origins, ray_directions.contiguous()


def ndc_rays(height: int, width: int, focal, near: float, rays_o, rays_d):
    """EXTENSION (not in the reference): original-NeRF `ndc_rays`.

    Shift the origin to the z=-near plane, then project.  Same op order the CUDA kernel
    uses (separately rounded ops); parity for this function is pinned only by this file.
    """
    o = _t(rays_o)
    d = _t(rays_d)
    near32 = torch.tensor(near, dtype=F32)
    focal32 = torch.tensor(float(np.float32(focal)), dtype=F32)
    t = -(near32 + o[..., 2]) / d[..., 2]
    o = o + t[..., None] * d
    sx = -(focal32 / torch.tensor(width * 0.5, dtype=F32))
    sy = -(focal32 / torch.tensor(height * 0.5, dtype=F32))
    two_n = torch.tensor(2.0, dtype=F32) * near32
    oxz = o[..., 0] / o[..., 2]
    oyz = o[..., 1] / o[..., 2]
    o0 = sx * oxz
    o1 = sy * oyz
    o2 = torch.tensor(1.0, dtype=F32) + two_n / o[..., 2]
    d0 = sx * (d[..., 0] / d[..., 2] - oxz)
    d1 = sy * (d[..., 1] / d[..., 2] - oyz)
    d2 = -two_n / o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def sample_rays(ray_origins, ray_directions, t_vals):
    """data_utils.py:55-73 -- pts = o + (d * t) (separate mul, add); dirs broadcast."""
    o = _t(ray_origins)
    d = _t(ray_directions)
    t = _t(t_vals)
    rays = o[..., None, :] + (d[..., None, :] * t[..., :, None])
    dirs = d[..., None, :].expand(rays.shape)
    return rays, dirs


def volume_render(preds, t_vals):
    """data_utils.py:75-98 -- alpha compositing; sigma is the LAST channel (:78)."""
    preds = _t(preds)
    t_vals = _t(t_vals)
    rgb = torch.sigmoid(preds[..., :-1])                              # :77
    sigma_a = torch.relu(preds[..., -1])                              # :78
    delta = t_vals[..., 1:] - t_vals[..., :-1]                        # :81
    const = torch.full((delta.shape[0], 1), 1e10, dtype=F32)          # :82
    delta = torch.cat([delta, const], dim=-1)                         # :83
    alpha = 1.0 - torch.exp(-sigma_a * delta)                         # :85
    exp_term = 1.0 - alpha                                            # :86
    epsilon = 1e-10                                                   # :87
    tm = torch.cumprod(exp_term + epsilon, dim=-1)                    # :90
    tm = torch.roll(tm, shifts=1, dims=-1)                            # :91
    transmittance = torch.cat([torch.ones((tm.shape[0], 1), dtype=F32), tm[:, 1:]], dim=-1)  # :92
    weights = alpha * transmittance                                   # :95
    rgb_w = torch.sum(weights[..., None] * rgb, dim=-2)               # :96
    depth_map = torch.sum(weights * t_vals, dim=-1)                   # :97
    return rgb_w, depth_map, weights


def split_data(images, poses, split_ratio=0.8):
    """data_utils.py:100-117."""
    n = images.shape[0]
    k = int(n * split_ratio)
    return images[:k], images[k:], poses[:k], poses[k:]


def tf_linspace_f32(start: float, stop: float, num: int) -> torch.Tensor:
    """TF `linspace` as called from data_utils.py:131 (restated from TF 2.16 math_ops
    `linspace_nd`, unverifiable offline): exact `start`, `start + delta*i` for the interior
    (separately rounded mul then add), exact `stop`; delta = (stop-start)/(num-1) in f32."""
    s = torch.tensor(start, dtype=F32)
    e = torch.tensor(stop, dtype=F32)
    if num == 1:
        return s.view(1)
    delta = (e - s) / torch.tensor(float(num - 1), dtype=F32)
    i = torch.arange(1, num - 1, dtype=F32)
    interior = s + delta * i
    return torch.cat([s.view(1), interior, e.view(1)])


def generate_t_vals(near, far, batch_size, num_samples, rand_sampling=True, u=None):
    """data_utils.py:119-138.  `u` (shape (N,), the reference's shared jitter vector, or
    (B,N)) replaces `keras.random.uniform` (:133); evaluation order (u*(far-near))/N."""
    t_vals = tf_linspace_f32(float(near), float(far), int(num_samples))
    if rand_sampling:
        if u is None:
            raise ValueError("oracle.generate_t_vals needs explicit uniform draws `u` when rand_sampling=True")
        u = _t(u)
        noise = u * torch.tensor(float(far) - float(near), dtype=F32) / torch.tensor(float(num_samples), dtype=F32)
        t_vals = t_vals + noise
    t_vals = torch.broadcast_to(t_vals, (int(batch_size), int(num_samples)))
    return t_vals.contiguous()


def sample_pdf(t_vals_mid, weights, ns_fine, u=None):
    """data_utils.py:172-223 (2-D case).  Quirks kept: Nc weights vs Nc-1 mids (half-bin
    shift + clamp, Q4); `u` always random in the reference (:196) -> explicit input here."""
    t_vals_mid = _t(t_vals_mid)
    weights = _t(weights)
    weights = weights + 1e-5                                          # :179
    pdf = weights / torch.sum(weights, dim=-1, keepdim=True)          # :182
    cdf = torch.cumsum(pdf, dim=-1)                                   # :185
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)    # :188
    if u is None:
        raise ValueError("oracle.sample_pdf needs explicit uniform draws `u` of shape (B, ns_fine)")
    u = _t(u)
    assert u.shape == (weights.shape[0], ns_fine)
    indices = torch.searchsorted(cdf.detach().contiguous(), u.contiguous(), right=True)  # :200
    below = torch.clamp(indices - 1, min=0)                           # :203
    above = torch.clamp(indices, max=cdf.shape[-1] - 1)               # :204
    cdf_b = torch.gather(cdf, -1, below)                              # :208
    cdf_a = torch.gather(cdf, -1, above)
    nm = t_vals_mid.shape[-1] - 1
    t_b = torch.gather(t_vals_mid, -1, torch.clamp(below, max=nm))    # :211-213
    t_a = torch.gather(t_vals_mid, -1, torch.clamp(above, max=nm))
    denom = cdf_a - cdf_b                                             # :216
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)  # :217
    t = (u - cdf_b) / denom                                           # :218
    samples = t_b + t * (t_a - t_b)                                   # :219-220
    return samples


def get_translation_t(t):
    """data_utils.py:225-233."""
    return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]], dtype=np.float32)


def get_rotation_phi(phi):
    """data_utils.py:236-244 (cos/sin evaluated in f32 like keras.ops on a python float)."""
    c = np.cos(np.float32(phi)).astype(np.float32)
    s = np.sin(np.float32(phi)).astype(np.float32)
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], dtype=np.float32)


def get_rotation_theta(theta):
    """data_utils.py:247-255."""
    c = np.cos(np.float32(theta)).astype(np.float32)
    s = np.sin(np.float32(theta)).astype(np.float32)
    return np.array([[c, 0, -s, 0], [0, 1, 0, 0], [s, 0, c, 0], [0, 0, 0, 1]], dtype=np.float32)


def pose_spherical(theta, phi, t):
    """data_utils.py:258-267 -- camera-to-world from (theta, phi, radius)."""
    c2w = get_translation_t(t)
    c2w = get_rotation_phi(phi / 180.0 * math.pi) @ c2w
    c2w = get_rotation_theta(theta / 180.0 * math.pi) @ c2w
    c2w = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32) @ c2w
    return c2w.astype(np.float32)
