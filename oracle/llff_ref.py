"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy, float64 intermediate like the reference's NumPy code) of the host-side
pose pipeline of the reference's loaders, SURVEY.md section 8(f) rank 3.  PARITY UNPINNED: the reference has no tests or
fixtures for these functions and its module cannot be imported here (needs tensorflow / keras / imageio).

Follows, step for step:
  fern_data_utils.py:251-262   normalize, viewmatrix
  fern_data_utils.py:268-278   poses_avg
  fern_data_utils.py:282-292   render_path_spiral
  fern_data_utils.py:296-309   recenter_poses
  fern_data_utils.py:315-366   spherify_poses
  fern_data_utils.py:135-137,176-177   poses_bounds.npy unpacking and the hwf column rewrite in _load_data
  fern_data_utils.py:393-457   load_fern_data after the images are read (axis fix, bound rescale, recentre, spiral, hold-out)
  fern_data_utils.py:479-500   prepare_fern_data: focal, near/far, hold-out split indices
  lego_data_utils.py:26,48-49 + data_utils.py:100-117   80/20 split, near/far of the Lego loader
Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np


def normalize(x):
    return x / np.linalg.norm(x)


def viewmatrix(z, up, pos):
    """fern_data_utils.py:254-260: columns (x, y, z, position) of a camera-to-world matrix from a forward axis and an up hint."""
    zc = normalize(z)
    xc = normalize(np.cross(up, zc))
    yc = normalize(np.cross(zc, xc))
    return np.stack([xc, yc, zc, pos], 1)


def poses_avg(poses):
    """fern_data_utils.py:268-278: mean position, summed z and y axes -> (3, 5) pose carrying the first pose's hwf column."""
    hwf = poses[0, :3, -1:]
    center = poses[:, :3, 3].mean(0)
    zsum = normalize(poses[:, :3, 2].sum(0))
    ysum = poses[:, :3, 1].sum(0)
    return np.concatenate([viewmatrix(zsum, ysum, center), hwf], 1)


def recenter_poses(poses):
    """fern_data_utils.py:296-309: left-multiply every pose by the inverse of the average pose."""
    out = poses + 0
    last = np.reshape([0, 0, 0, 1.0], [1, 4])
    avg = np.concatenate([poses_avg(poses)[:3, :4], last], -2)
    full = np.concatenate([poses[:, :3, :4], np.tile(last[None], [poses.shape[0], 1, 1])], -2)
    full = np.linalg.inv(avg) @ full
    out[:, :3, :4] = full[:, :3, :4]
    return out


def render_path_spiral(c2w, up, rads, focal, zdelta, zrate, rots, N):
    """fern_data_utils.py:282-292."""
    out = []
    rads = np.array(list(rads) + [1.0])
    hwf = c2w[:, 4:5]
    for theta in np.linspace(0.0, 2.0 * np.pi * rots, N + 1)[:-1]:
        c = np.dot(c2w[:3, :4], np.array([np.cos(theta), -np.sin(theta), -np.sin(theta * zrate), 1.0]) * rads)
        z = normalize(c - np.dot(c2w[:3, :4], np.array([0, 0, -focal, 1.0])))
        out.append(np.concatenate([viewmatrix(z, up, c), hwf], 1))
    return out


def spherify_poses(poses, bds):
    """fern_data_utils.py:315-366 (bds is rescaled in place there; a copy is returned here)."""
    bds = bds.copy()
    to44 = lambda p: np.concatenate([p, np.tile(np.reshape(np.eye(4)[-1, :], [1, 1, 4]), [p.shape[0], 1, 1])], 1)
    rd = poses[:, :3, 2:3]
    ro = poses[:, :3, 3:4]
    A = np.eye(3) - rd * np.transpose(rd, [0, 2, 1])
    b = -A @ ro
    center = np.squeeze(-np.linalg.inv((np.transpose(A, [0, 2, 1]) @ A).mean(0)) @ b.mean(0))
    up = (poses[:, :3, 3] - center).mean(0)
    v0 = normalize(up)
    v1 = normalize(np.cross([0.1, 0.2, 0.3], v0))
    v2 = normalize(np.cross(v0, v1))
    c2w = np.stack([v1, v2, v0, center], 1)
    reset = np.linalg.inv(to44(c2w[None])) @ to44(poses[:, :3, :4])
    rad = np.sqrt(np.mean(np.sum(np.square(reset[:, :3, 3]), -1)))
    sc = 1.0 / rad
    reset[:, :3, 3] *= sc
    bds *= sc
    rad *= sc
    zh = np.mean(reset[:, :3, 3], 0)[2]
    radcircle = np.sqrt(rad ** 2 - zh ** 2)
    ring = []
    for th in np.linspace(0.0, 2.0 * np.pi, 120):
        origin = np.array([radcircle * np.cos(th), radcircle * np.sin(th), zh])
        upv = np.array([0, 0, -1.0])
        a2 = normalize(origin)
        a0 = normalize(np.cross(a2, upv))
        a1 = normalize(np.cross(a2, a0))
        ring.append(np.stack([a0, a1, a2, origin], 1))
    ring = np.stack(ring, 0)
    ring = np.concatenate([ring, np.broadcast_to(poses[0, :3, -1:], ring[:, :3, -1:].shape)], -1)
    reset = np.concatenate([reset[:, :3, :4], np.broadcast_to(poses[0, :3, -1:], reset[:, :3, -1:].shape)], -1)
    return reset, ring, bds


def unpack_poses_bounds(poses_arr, image_hw, factor):
    """fern_data_utils.py:135-137 + 176-177: (N, 17) -> poses (3, 5, N), bds (2, N); hwf column = image size, focal / factor."""
    poses = poses_arr[:, :-2].reshape([-1, 3, 5]).transpose([1, 2, 0]).copy()
    bds = poses_arr[:, -2:].transpose([1, 0]).copy()
    poses[:2, 4, :] = np.array(image_hw).reshape([2, 1])
    poses[2, 4, :] = poses[2, 4, :] * 1.0 / factor
    return poses, bds


def llff_poses(poses, bds, recenter=True, bd_factor=0.75, spherify=False, path_zflat=False):
    """fern_data_utils.py:393-457 given `_load_data`'s (3,5,N) poses and (2,N) bounds:
    -> poses (N,3,5) f32, bds (N,2) f32, render_poses (M,3,5) f32, i_test."""
    poses = np.concatenate([poses[:, 1:2, :], -poses[:, 0:1, :], poses[:, 2:, :]], 1)
    poses = np.moveaxis(poses, -1, 0).astype(np.float32)
    bds = np.moveaxis(bds, -1, 0).astype(np.float32)
    sc = 1.0 if bd_factor is None else 1.0 / (bds.min() * bd_factor)
    poses[:, :3, 3] *= sc
    bds *= sc
    if recenter:
        poses = recenter_poses(poses)
    if spherify:
        poses, render_poses, bds = spherify_poses(poses, bds)
    else:
        c2w = poses_avg(poses)
        up = normalize(poses[:, :3, 1].sum(0))
        close_depth, inf_depth = bds.min() * 0.9, bds.max() * 5.0
        dt = 0.75
        focal = 1.0 / (((1.0 - dt) / close_depth + dt / inf_depth))
        zdelta = close_depth * 0.2
        rads = np.percentile(np.abs(poses[:, :3, 3]), 90, 0)
        n_views, n_rots = 120, 2
        if path_zflat:
            zloc = -close_depth * 0.1
            c2w[:3, 3] = c2w[:3, 3] + zloc * c2w[:3, 2]
            rads[2] = 0.0
            n_rots = 1
            n_views = n_views // 2
        render_poses = render_path_spiral(c2w, up, rads, focal, zdelta, zrate=0.5, rots=n_rots, N=n_views)
    render_poses = np.array(render_poses).astype(np.float32)
    c2w = poses_avg(poses)
    i_test = int(np.argmin(np.sum(np.square(c2w[:3, 3] - poses[:, :3, 3]), -1)))
    return poses.astype(np.float32), bds, render_poses, i_test


def fern_split(poses, bds, i_test):
    """fern_data_utils.py:479-500: focal from the first pose's hwf column, near = 0.9 min(bds), far = max(bds), hold-out split."""
    focal = poses[0, 2, -1]
    near = np.min(bds) * 0.9
    far = np.max(bds) * 1.0
    i_train = np.array([i for i in range(len(poses)) if i != i_test])
    return focal, (near, far), i_train, np.array([i_test])


def lego_split(n_items, split_ratio=0.8):
    """data_utils.py:111-117 through lego_data_utils.py:26: first int(n * ratio) items train, the rest validate; near/far 2/6."""
    k = int(n_items * split_ratio)
    return np.arange(0, k), np.arange(k, n_items), (2.0, 6.0)
