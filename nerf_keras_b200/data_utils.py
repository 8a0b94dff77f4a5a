"""Drop-in for the reference's `data_utils.py`: same function names and argument meaning, executed by
hand-written sm_100a CUDA kernels behind the C-ABI of libnerf_b200.so.

Inputs may be CUDA torch tensors (used in place) or anything `torch.as_tensor` accepts (copied to
the current CUDA device); outputs are float32 CUDA tensors.  Random draws are explicit optional
inputs (`u`) so results are reproducible against the reference given the same draws; when omitted
they are drawn with torch's CUDA generator (the reference draws with keras/tf at the same places).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib


def _dev() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("nerf_keras_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _f32(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x, dtype=np.float32))
    return t.to(device=_dev(), dtype=torch.float32).contiguous()


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def encode_position(x, pos_encode_dims):
    """data_utils.py:7-21 -- (..., 3) -> (..., 3 + 6*pos_encode_dims)."""
    x = _f32(x)
    if x.shape[-1] != 3:
        raise ValueError("encode_position expects a trailing dimension of 3")
    n = x.numel() // 3
    out = torch.empty(x.shape[:-1] + (3 + 6 * pos_encode_dims,), device=x.device, dtype=torch.float32)
    if n == 0:
        return out
    _lib.check(_lib.lib().nerf_encode_position(_ptr(x), n, int(pos_encode_dims), _ptr(out), _stream()),
               "encode_position")
    return out


def get_rays(height, width, focal, pose):
    """data_utils.py:23-52 -- (ray_origins, ray_directions), each (H, W, 3)."""
    pose_h = np.asarray(pose.detach().cpu().numpy() if isinstance(pose, torch.Tensor) else pose, dtype=np.float32)
    if pose_h.ndim != 2 or pose_h.shape[0] < 3 or pose_h.shape[1] != 4:
        raise ValueError("pose must be a (3..4, 4) camera-to-world matrix")
    p12 = np.ascontiguousarray(pose_h[:3, :4]).reshape(-1)
    dev = _dev()
    o = torch.empty((height, width, 3), device=dev, dtype=torch.float32)
    d = torch.empty((height, width, 3), device=dev, dtype=torch.float32)
    arr = (C.c_float * 12)(*[float(v) for v in p12])
    focal32 = float(np.float32(focal))
    _lib.check(_lib.lib().nerf_get_rays(int(height), int(width), focal32, arr, _ptr(o), _ptr(d), _stream()), "get_rays")
    return o, d


def ndc_rays(height, width, focal, near, rays_o, rays_d):
    """EXTENSION (not in the reference): original-NeRF NDC ray transform for forward-facing scenes."""
    o, d = _f32(rays_o), _f32(rays_d)
    oo, do = torch.empty_like(o), torch.empty_like(d)
    _lib.check(_lib.lib().nerf_ndc_rays(int(height), int(width), float(np.float32(focal)), float(near), _ptr(o),
                                        _ptr(d), _ptr(oo), _ptr(do), o.numel() // 3, _stream()), "ndc_rays")
    return oo, do


def sample_rays(ray_origins, ray_directions, t_vals):
    """data_utils.py:55-73 -- rays = o + d*t, dirs = broadcast(d); (B,N,3) each."""
    o, d, t = _f32(ray_origins), _f32(ray_directions), _f32(t_vals)
    B, N = t.shape
    rays = torch.empty((B, N, 3), device=o.device, dtype=torch.float32)
    dirs = torch.empty((B, N, 3), device=o.device, dtype=torch.float32)
    if B == 0:
        return rays, dirs
    _lib.check(_lib.lib().nerf_sample_rays(_ptr(o), _ptr(d), _ptr(t), B, N, _ptr(rays), _ptr(dirs), _stream()),
               "sample_rays")
    return rays, dirs


def volume_render(preds, t_vals, return_acc=False):
    """data_utils.py:75-98 -- (rgb (B,3), depth (B,), weights (B,N)); optional acc map (extension)."""
    p, t = _f32(preds), _f32(t_vals)
    if t.dim() != 2 or p.shape != t.shape + (4,):
        raise ValueError("volume_render expects preds (B,N,4) and static 2-D t_vals (B,N)")
    B, N = t.shape
    rgb = torch.empty((B, 3), device=p.device, dtype=torch.float32)
    depth = torch.empty((B,), device=p.device, dtype=torch.float32)
    w = torch.empty((B, N), device=p.device, dtype=torch.float32)
    acc = torch.empty((B,), device=p.device, dtype=torch.float32) if return_acc else None
    if B == 0:
        return (rgb, depth, w, acc) if return_acc else (rgb, depth, w)
    _lib.check(_lib.lib().nerf_volume_render(_ptr(p), _ptr(t), B, N, _ptr(rgb), _ptr(depth), _ptr(w), _ptr(acc),
                                             _stream()), "volume_render")
    return (rgb, depth, w, acc) if return_acc else (rgb, depth, w)


def split_data(images, poses, split_ratio=0.8):
    """data_utils.py:100-117 (host slicing)."""
    n = images.shape[0]
    k = int(n * split_ratio)
    return images[:k], images[k:], poses[:k], poses[k:]


def generate_t_vals(near, far, batch_size, num_samples, rand_sampling=True, u=None):
    """data_utils.py:119-138 -- (batch_size, num_samples) t-values.

    u: optional uniform draws, shape (num_samples,) (the reference's single shared jitter vector)
    or (batch_size, num_samples); drawn on the device when omitted and rand_sampling is True."""
    dev = _dev()
    B, N = int(batch_size), int(num_samples)
    per_ray = 0
    uu = None
    if rand_sampling:
        uu = torch.rand((N,), device=dev, dtype=torch.float32) if u is None else _f32(u)
        if uu.shape == (N,):
            per_ray = 0
        elif uu.shape == (B, N):
            per_ray = 1
        else:
            raise ValueError("u must have shape (num_samples,) or (batch_size, num_samples)")
    t = torch.empty((B, N), device=dev, dtype=torch.float32)
    if B == 0:
        return t
    _lib.check(_lib.lib().nerf_generate_t_vals(float(near), float(far), B, N, _ptr(uu), per_ray, _ptr(t), _stream()),
               "generate_t_vals")
    return t


def sample_pdf(t_vals_mid, weights, ns_fine, u=None):
    """data_utils.py:172-223 -- inverse-CDF samples (B, ns_fine); `u` (B, ns_fine) explicit draws."""
    tm, w = _f32(t_vals_mid), _f32(weights)
    if w.dim() != 2 or tm.shape != (w.shape[0], w.shape[1] - 1):
        raise ValueError("sample_pdf expects weights (B,Nc) and t_vals_mid (B,Nc-1)")
    B, nc = w.shape
    uu = torch.rand((B, ns_fine), device=w.device, dtype=torch.float32) if u is None else _f32(u)
    if uu.shape != (B, ns_fine):
        raise ValueError("u must have shape (B, ns_fine)")
    out = torch.empty((B, ns_fine), device=w.device, dtype=torch.float32)
    if B == 0:
        return out
    _lib.check(_lib.lib().nerf_sample_pdf(_ptr(tm), _ptr(w), _ptr(uu), B, nc, int(ns_fine), _ptr(out), _stream()),
               "sample_pdf")
    return out


def resample_merge(t_vals, weights, ns_fine, u=None, return_index=False):
    """Fused models.py:165-167: sort(concat([t, sample_pdf(mid(t), w, ns_fine)])) -> (B, Nc+ns_fine)."""
    t, w = _f32(t_vals), _f32(weights)
    B, nc = t.shape
    uu = torch.rand((B, ns_fine), device=t.device, dtype=torch.float32) if u is None else _f32(u)
    out = torch.empty((B, nc + ns_fine), device=t.device, dtype=torch.float32)
    idx = torch.empty((B, nc + ns_fine), device=t.device, dtype=torch.int32) if return_index else None
    if B == 0:
        return (out, idx) if return_index else out
    _lib.check(_lib.lib().nerf_resample_merge(_ptr(t), _ptr(w), _ptr(uu), B, nc, int(ns_fine), _ptr(out), _ptr(idx),
                                              _stream()), "resample_merge")
    return (out, idx) if return_index else out


def create_batched_dataset_pipeline(images_s, ray_oris_s, ray_dirs_s, num_samples, batch_size, auto=None, near=2.0, far=6.0,
                                    shuffle=True, rand_sampling=True):
    """data_utils.py:140-170 without tf.data: an iterable of (images, (ray_origins, ray_directions, t_vals)) batches.
    Same arguments (`auto`, tf.data.AUTOTUNE in the reference, is ignored); every ray once per epoch, drop_remainder, a
    local shuffle equivalent to the 5 * batch_size shuffle buffer, ONE jitter vector shared by all rays (the reference
    materialises the (n_rays, num_samples) t-value array; here each batch expands the shared row).  The data stays on
    the device; see synthetic.HostPrefetcher for host-resident ray sets."""
    from .synthetic import BatchedRayDataset
    dev = _dev()
    to_dev = lambda x: _f32(x).reshape(-1, x.shape[-1]).to(dev).contiguous()
    return BatchedRayDataset(to_dev(images_s), to_dev(ray_oris_s), to_dev(ray_dirs_s), num_samples, batch_size, near, far,
                             shuffle=shuffle, rand_sampling=rand_sampling, order="windowed")


# ---- host 4x4 camera helpers (data_utils.py:225-267); negligible work, stays on the host ----
def get_translation_t(t):
    return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]], dtype=np.float32)


def get_rotation_phi(phi):
    c, s = np.cos(np.float32(phi)), np.sin(np.float32(phi))
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], dtype=np.float32)


def get_rotation_theta(theta):
    c, s = np.cos(np.float32(theta)), np.sin(np.float32(theta))
    return np.array([[c, 0, -s, 0], [0, 1, 0, 0], [s, 0, c, 0], [0, 0, 0, 1]], dtype=np.float32)


def pose_spherical(theta, phi, t):
    c2w = get_translation_t(t)
    c2w = get_rotation_phi(phi / 180.0 * math.pi) @ c2w
    c2w = get_rotation_theta(theta / 180.0 * math.pi) @ c2w
    c2w = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32) @ c2w
    return c2w.astype(np.float32)
