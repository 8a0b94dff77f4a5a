"""Real-data loaders with the return structure of the reference's, without
TensorFlow / imageio / ImageMagick / network access:

  * tiny-NeRF `tiny_nerf_data.npz` (images, poses, focal)          -- `prepare_lego_data`, lego_data_utils.py:8-51
  * LLFF scenes (`poses_bounds.npy` + `images[_<factor>]/`)         -- `load_fern_data`, `prepare_fern_data`,
                                                                      fern_data_utils.py:133-189, 369-520
  * Blender / NeRF-synthetic `transforms_<split>.json` + PNGs       -- extension (the reference has no Blender loader)

Pose pre-processing is host-side float64 NumPy like the reference's (it runs once per scene); images are decoded with
Pillow; the per-pixel rays come from the CUDA `get_rays` kernel and everything the trainer consumes is returned as
flattened fp32 CUDA tensors: ((images, origins, directions) train, (...) val, (near, far), focal).

Differences from the reference, all deliberate:
  * nothing is downloaded: a missing file raises FileNotFoundError naming the expected path (lego_data_utils.py:11-14
    fetches from the network);
  * LLFF images are down-sampled in memory with Pillow (box filter) when `images_<factor>/` does not exist; the
    reference shells out to ImageMagick `mogrify` and writes that directory (fern_data_utils.py:8-57);
  * `path_zflat=True` halves the view count with an integer division (`N_views /= 2` in the reference makes a float,
    which current NumPy rejects in `np.linspace`);
  * the Lego validation rays keep the reference's quirk of being generated at (target_height, target_height)
    (lego_data_utils.py:34).
"""
from __future__ import annotations

import json
import os
from typing import Optional, Tuple

import numpy as np
import torch

from . import data_utils as du

_IMG_EXT = ("JPG", "jpg", "png")


# --------------------------------------------------------------------------------------------------
# pose maths (vectorised; compare oracle/llff_ref.py which restates the reference loop by loop)
# --------------------------------------------------------------------------------------------------
def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def look_at_frame(forward, up_hint, position):
    """Camera-to-world (…,3,4) whose z axis is `forward`, x = up_hint × z, y = z × x (fern_data_utils.py:254-260)."""
    z = _unit(np.asarray(forward, dtype=np.float64))
    x = _unit(np.cross(np.asarray(up_hint, dtype=np.float64), z))
    y = _unit(np.cross(z, x))
    return np.stack([x, y, z, np.broadcast_to(np.asarray(position, dtype=np.float64), z.shape)], axis=-1)


def average_pose(poses):
    """(N,3,5) -> (3,5): mean position, summed z / y axes, hwf column of the first pose (fern_data_utils.py:268-278)."""
    frame = look_at_frame(poses[:, :3, 2].sum(0), poses[:, :3, 1].sum(0), poses[:, :3, 3].mean(0))
    return np.concatenate([frame, poses[0, :3, 4:5]], axis=1)


def recenter(poses):
    """Express every pose in the frame of the average pose (fern_data_utils.py:296-309).  The average frame is
    orthonormal by construction, so its inverse is (R^T, -R^T t)."""
    avg = average_pose(poses)
    Rt = avg[:3, :3].T
    out = np.array(poses, copy=True)
    out[:, :3, :3] = Rt @ poses[:, :3, :3]
    out[:, :3, 3] = (poses[:, :3, 3] - avg[:3, 3]) @ Rt.T
    return out


def spiral_path(c2w, up, rads, focal, zrate, rots, n_views):
    """(n_views,3,5) poses on a spiral around `c2w`, all looking at the point `focal` in front of it
    (fern_data_utils.py:282-292)."""
    theta = np.linspace(0.0, 2.0 * np.pi * rots, n_views + 1)[:-1]
    local = np.stack([np.cos(theta), -np.sin(theta), -np.sin(theta * zrate), np.ones_like(theta)], -1) * np.append(rads, 1.0)
    centers = local @ c2w[:3, :4].T
    focus = c2w[:3, :4] @ np.array([0.0, 0.0, -focal, 1.0])
    frames = look_at_frame(centers - focus, np.broadcast_to(up, centers.shape), centers)
    return np.concatenate([frames, np.broadcast_to(c2w[:, 4:5], (n_views, 3, 1))], axis=-1)


def spherify(poses, bds):
    """Inward-facing capture: recentre on the point closest to all optical axes, scale the cameras onto the unit
    sphere and return a 120-pose circular render path (fern_data_utils.py:315-366)."""
    d = poses[:, :3, 2]
    o = poses[:, :3, 3]
    A = np.eye(3)[None] - d[:, :, None] * d[:, None, :]
    AtA = np.einsum("nji,njk->nik", A, A).mean(0)
    center = np.linalg.solve(AtA, (A @ o[:, :, None]).mean(0))[:, 0]
    v0 = _unit((o - center).mean(0))
    v1 = _unit(np.cross([0.1, 0.2, 0.3], v0))
    v2 = _unit(np.cross(v0, v1))
    R = np.stack([v1, v2, v0], 1)                                # orthonormal: inverse = transpose
    reset = np.array(poses[:, :3, :4], dtype=np.float64, copy=True)
    reset[:, :, :3] = R.T @ poses[:, :3, :3]
    reset[:, :, 3] = (o - center) @ R
    scale = 1.0 / np.sqrt(np.mean(np.sum(np.square(reset[:, :, 3]), -1)))
    reset[:, :, 3] *= scale
    zh = reset[:, 2, 3].mean()
    radcircle = np.sqrt(1.0 - zh ** 2)
    th = np.linspace(0.0, 2.0 * np.pi, 120)
    origin = np.stack([radcircle * np.cos(th), radcircle * np.sin(th), np.full_like(th, zh)], -1)
    a2 = _unit(origin)
    a0 = _unit(np.cross(a2, np.broadcast_to(np.array([0.0, 0.0, -1.0]), a2.shape)))
    a1 = _unit(np.cross(a2, a0))
    ring = np.stack([a0, a1, a2, origin], -1)
    hwf = poses[0, :3, 4:5]
    ring = np.concatenate([ring, np.broadcast_to(hwf, (ring.shape[0], 3, 1))], -1)
    reset = np.concatenate([reset, np.broadcast_to(hwf, (reset.shape[0], 3, 1))], -1)
    return reset, ring, bds * scale


def llff_pose_pipeline(poses_arr, image_hw, factor, recenter_poses=True, bd_factor=0.75, spherify_poses=False,
                       path_zflat=False):
    """`poses_bounds.npy` rows (N,17) -> poses (N,3,5) f32 [R | t | hwf], bds (N,2) f32, render_poses (M,3,5) f32 and
    the hold-out index: LLFF axis fix [y, -x, z], bound rescale, recentring, spiral (or spherical) render path,
    hold-out = the view closest to the average pose (fern_data_utils.py:135-137,176-177,393-457)."""
    raw = np.asarray(poses_arr, dtype=np.float64)
    p = raw[:, :15].reshape(-1, 3, 5)
    p[:, 0, 4], p[:, 1, 4] = image_hw[0], image_hw[1]
    p[:, 2, 4] = p[:, 2, 4] / factor
    poses = np.concatenate([p[:, :, 1:2], -p[:, :, 0:1], p[:, :, 2:]], axis=2).astype(np.float32)
    bds = raw[:, 15:17].astype(np.float32)
    sc = np.float32(1.0) if bd_factor is None else np.float32(1.0 / (bds.min() * bd_factor))
    poses[:, :3, 3] *= sc
    bds = bds * sc
    if recenter_poses:
        poses = recenter(poses.astype(np.float64)).astype(np.float32)
    if spherify_poses:
        poses, render_poses, bds = spherify(poses, bds)
    else:
        c2w = average_pose(poses)
        up = _unit(poses[:, :3, 1].sum(0))
        close_depth, inf_depth = bds.min() * 0.9, bds.max() * 5.0
        focus = 1.0 / (0.25 / close_depth + 0.75 / inf_depth)
        rads = np.percentile(np.abs(poses[:, :3, 3]), 90, 0)
        n_views, n_rots = 120, 2
        if path_zflat:
            c2w[:3, 3] = c2w[:3, 3] - close_depth * 0.1 * c2w[:3, 2]
            rads[2] = 0.0
            n_views, n_rots = 60, 1
        render_poses = spiral_path(c2w, up, rads, focus, 0.5, n_rots, n_views)
    avg = average_pose(poses)
    i_test = int(np.argmin(np.sum(np.square(avg[:3, 3] - poses[:, :3, 3]), -1)))
    return poses.astype(np.float32), np.asarray(bds, dtype=np.float32), np.asarray(render_poses, dtype=np.float32), i_test


# --------------------------------------------------------------------------------------------------
# images
# --------------------------------------------------------------------------------------------------
def _read_image(path, size_wh=None):
    from PIL import Image
    with Image.open(path) as im:
        if size_wh is not None and im.size != tuple(size_wh):
            im = im.resize(tuple(size_wh), Image.BOX)
        a = np.asarray(im)
    return a


def resize_images(images, height, width):
    """Bilinear, half-pixel centres, no antialiasing: the arithmetic of `tf.image.resize(images, (H, W))` used at
    lego_data_utils.py:23 / fern_data_utils.py:474.  images (N,H0,W0,C) float32 -> (N,H,W,C) float32."""
    x = torch.as_tensor(np.ascontiguousarray(images), dtype=torch.float32)
    if x.shape[1] == height and x.shape[2] == width:
        return x
    y = torch.nn.functional.interpolate(x.permute(0, 3, 1, 2), size=(height, width), mode="bilinear",
                                        align_corners=False, antialias=False)
    return y.permute(0, 2, 3, 1).contiguous()


def _rays_for(poses, height, width, focal):
    oris, dirs = [], []
    for pose in poses:
        o, d = du.get_rays(height, width, float(focal), np.asarray(pose, dtype=np.float32))
        oris.append(o.reshape(-1, 3)); dirs.append(d.reshape(-1, 3))
    if not oris:
        e = torch.empty((0, 3), device=du._dev(), dtype=torch.float32)
        return e, e.clone()
    return torch.cat(oris), torch.cat(dirs)


def _flat_images(images):
    return torch.as_tensor(images, dtype=torch.float32).reshape(-1, images.shape[-1]).to(du._dev()).contiguous()


# --------------------------------------------------------------------------------------------------
# tiny-NeRF (Lego 100x100) -- lego_data_utils.py:8-51
# --------------------------------------------------------------------------------------------------
def prepare_lego_data(target_height, target_width, npz_path="data/tiny_nerf_data.npz", split_ratio=0.8):
    if not os.path.exists(npz_path):
        raise FileNotFoundError(f"{npz_path}: tiny_nerf_data.npz not found (the reference downloads it from "
                                "cseweb.ucsd.edu/~viscomp/projects/LF/papers/ECCV20/nerf/; no network here)")
    data = np.load(npz_path)
    images, poses, focal = data["images"], data["poses"], data["focal"]
    images_r = resize_images(images, target_height, target_width)
    tr_img, va_img, tr_pose, va_pose = du.split_data(images_r, poses, split_ratio)
    tr_o, tr_d = _rays_for(tr_pose, target_height, target_width, focal)
    va_o, va_d = _rays_for(va_pose, target_height, target_height, focal)      # sic: lego_data_utils.py:34
    return (_flat_images(tr_img), tr_o, tr_d), (_flat_images(va_img), va_o, va_d), (2.0, 6.0), focal


# --------------------------------------------------------------------------------------------------
# LLFF (Fern) -- fern_data_utils.py:133-189, 369-520
# --------------------------------------------------------------------------------------------------
def load_fern_data(basedir, factor=8, recenter=True, bd_factor=0.75, spherify=False, path_zflat=False):
    """-> images (N,H,W,3) f32 in [0,1], poses (N,3,5), bds (N,2), render_poses (M,3,5), i_test."""
    pb = os.path.join(basedir, "poses_bounds.npy")
    if not os.path.exists(pb):
        raise FileNotFoundError(f"{pb} not found")
    poses_arr = np.load(pb)
    full_dir = os.path.join(basedir, "images")
    factor = 1 if factor is None else factor
    small_dir = os.path.join(basedir, f"images_{factor}") if factor != 1 else full_dir
    src_dir = small_dir if os.path.isdir(small_dir) else full_dir
    files = [os.path.join(src_dir, f) for f in sorted(os.listdir(src_dir)) if f.endswith(_IMG_EXT)]
    if len(files) != poses_arr.shape[0]:
        raise ValueError(f"Mismatch between imgs {len(files)} and poses {poses_arr.shape[0]}")
    size_wh = None
    if src_dir == full_dir and factor != 1:                      # minify in memory instead of `mogrify`
        from PIL import Image
        with Image.open(files[0]) as im0:
            size_wh = (im0.size[0] // factor, im0.size[1] // factor)
    imgs = np.stack([_read_image(f, size_wh)[..., :3].astype(np.float32) / np.float32(255.0) for f in files], 0)
    poses, bds, render_poses, i_test = llff_pose_pipeline(poses_arr, imgs.shape[1:3], factor, recenter, bd_factor,
                                                         spherify, path_zflat)
    return imgs, poses, bds, render_poses, i_test


def prepare_fern_data(target_height, target_width, datadir="data/nerf_example_data/nerf_llff_data/fern", factor=8):
    images, poses_ori, bds, render_poses, i_test = load_fern_data(datadir, factor=factor, recenter=True, bd_factor=0.75,
                                                                  spherify=False)
    images_r = resize_images(images, target_height, target_width)
    focal = poses_ori[0, 2, -1]
    oris, dirs = _rays_for(poses_ori[:, :3, :4], target_height, target_width, focal)
    n = images.shape[0]
    oris, dirs = oris.reshape(n, -1, 3), dirs.reshape(n, -1, 3)
    near, far = float(np.min(bds) * 0.9), float(np.max(bds) * 1.0)
    i_train = [i for i in range(n) if i != i_test]
    pick = lambda x, idx: x[idx].reshape(-1, 3).contiguous()
    imgs_t = torch.as_tensor(images_r, dtype=torch.float32).reshape(n, -1, 3).to(du._dev())
    return ((pick(imgs_t, i_train), pick(oris, i_train), pick(dirs, i_train)),
            (pick(imgs_t, [i_test]), pick(oris, [i_test]), pick(dirs, [i_test])), (near, far), focal)


# --------------------------------------------------------------------------------------------------
# Blender / NeRF-synthetic (extension): transforms_<split>.json with camera_angle_x and per-frame transform_matrix
# --------------------------------------------------------------------------------------------------
def load_blender_data(basedir, split="train", skip=1, white_bkgd=True) -> Tuple[np.ndarray, np.ndarray, float]:
    """-> images (N,H,W,3) f32 (RGBA composited on white or black), poses (N,4,4) f32, focal."""
    meta_path = os.path.join(basedir, f"transforms_{split}.json")
    if not os.path.exists(meta_path):
        raise FileNotFoundError(f"{meta_path} not found")
    with open(meta_path) as f:
        meta = json.load(f)
    imgs, poses = [], []
    for frame in meta["frames"][::skip]:
        path = os.path.join(basedir, frame["file_path"])
        if not os.path.splitext(path)[1]:
            path += ".png"
        a = _read_image(path).astype(np.float32) / np.float32(255.0)
        if a.shape[-1] == 4:
            rgb, alpha = a[..., :3], a[..., 3:4]
            a = rgb * alpha + (1.0 - alpha) if white_bkgd else rgb * alpha
        imgs.append(a[..., :3])
        poses.append(np.asarray(frame["transform_matrix"], dtype=np.float32))
    imgs = np.stack(imgs, 0)
    focal = float(np.float32(0.5 * imgs.shape[2] / np.tan(0.5 * float(meta["camera_angle_x"]))))
    return imgs, np.stack(poses, 0), focal


def prepare_blender_data(target_height, target_width, basedir, skip=1, white_bkgd=True):
    """Same return structure as `prepare_lego_data`, from the `train` and `val` splits of a Blender scene."""
    out = []
    focal = None
    for split in ("train", "val"):
        imgs, poses, f = load_blender_data(basedir, split, skip, white_bkgd)
        focal = f * target_width / imgs.shape[2]
        o, d = _rays_for(poses, target_height, target_width, focal)
        out.append((_flat_images(resize_images(imgs, target_height, target_width)), o, d))
    return out[0], out[1], (2.0, 6.0), focal
