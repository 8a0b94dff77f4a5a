"""Real-data loaders: host-side pose pipeline against the oracle restatement of
fern_data_utils.py / lego_data_utils.py, plus file-format handling on small generated scenes (no GPU needed)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import llff_ref as R
from nerf_keras_b200 import real_data as rd


def _fake_poses_bounds(n=12, seed=0, inward=False):
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(n):
        if inward:
            th = 2 * np.pi * i / n
            pos = np.array([3 * np.cos(th), 3 * np.sin(th), 0.5 + 0.2 * rng.standard_normal()])
            z = pos / np.linalg.norm(pos)                       # LLFF cameras look along -z
            x = np.cross([0, 0, 1.0], z); x /= np.linalg.norm(x)
            y = np.cross(z, x)
        else:
            yaw, pitch = rng.uniform(-0.15, 0.15, 2)
            z = np.array([np.sin(yaw), np.sin(pitch), np.cos(yaw) * np.cos(pitch)]); z /= np.linalg.norm(z)
            x = np.cross([0, 1.0, 0], z); x /= np.linalg.norm(x)
            y = np.cross(z, x)
            pos = np.array([rng.uniform(-1, 1), rng.uniform(-0.6, 0.6), rng.uniform(-0.1, 0.1)])
        Rm = np.stack([x, y, z], 1)
        # LLFF stores columns as [down, right, back]; the loader turns them into [right, up, back]
        llff = np.stack([-Rm[:, 1], Rm[:, 0], Rm[:, 2], pos, np.array([3024.0, 4032.0, 3260.5])], 1)
        rows.append(np.concatenate([llff.reshape(-1), [rng.uniform(1.0, 2.0), rng.uniform(20.0, 90.0)]]))
    return np.stack(rows, 0)


@pytest.mark.parametrize("recenter,spherify,zflat", [(True, False, False), (False, False, False), (True, False, True),
                                                     (True, True, False)])
def test_llff_pose_pipeline_matches_oracle(recenter, spherify, zflat):
    arr = _fake_poses_bounds(inward=spherify)
    hw, factor = (378, 504), 8
    p0, b0 = R.unpack_poses_bounds(arr.copy(), hw, factor)
    ref = R.llff_poses(p0, b0, recenter=recenter, bd_factor=0.75, spherify=spherify, path_zflat=zflat)
    got = rd.llff_pose_pipeline(arr.copy(), hw, factor, recenter, 0.75, spherify, zflat)
    for a, b, name in zip(got[:3], ref[:3], ("poses", "bds", "render_poses")):
        assert a.shape == b.shape and a.dtype == np.float32, name
        np.testing.assert_allclose(a, b, rtol=2e-5, atol=2e-5, err_msg=name)
    assert got[3] == ref[3]
    assert got[2].shape[0] == (120 if not zflat else 60)


def test_recentred_average_pose_is_identity_and_hwf_kept():
    arr = _fake_poses_bounds(seed=3)
    poses, bds, render, i_test = rd.llff_pose_pipeline(arr, (378, 504), 8)
    avg = rd.average_pose(poses.astype(np.float64))
    np.testing.assert_allclose(avg[:3, :3], np.eye(3), atol=1e-5)
    np.testing.assert_allclose(avg[:3, 3], 0.0, atol=1e-5)
    np.testing.assert_allclose(poses[:, :, 4], np.broadcast_to([378.0, 504.0, 3260.5 / 8], poses[:, :, 4].shape), rtol=1e-6)
    assert abs(float(bds.min()) - 1.0 / 0.75) < 1e-5          # nearest bound rescaled to 1 / bd_factor
    # spiral frames are orthonormal and right-handed
    Rm = render[:, :3, :3].astype(np.float64)
    np.testing.assert_allclose(np.einsum("nij,nik->njk", Rm, Rm), np.broadcast_to(np.eye(3), Rm.shape), atol=1e-5)
    assert np.all(np.linalg.det(Rm) > 0.99)
    assert 0 <= i_test < poses.shape[0]


def test_split_rules():
    tr, va, bounds = R.lego_split(106)
    assert len(tr) == 84 and len(va) == 22 and bounds == (2.0, 6.0)
    from nerf_keras_b200 import data_utils as du
    imgs, poses = np.arange(106)[:, None], np.arange(106)[:, None]
    a, b, c, d = du.split_data(imgs, poses, 0.8)
    assert len(a) == len(c) == 84 and len(b) == len(d) == 22 and a[-1, 0] == 83 and b[0, 0] == 84


def test_resize_images_is_tf_bilinear_half_pixel():
    x = np.array([[0.0, 1.0], [2.0, 3.0]], dtype=np.float32)[None, :, :, None]
    same = rd.resize_images(x, 2, 2)
    assert torch.equal(same, torch.from_numpy(x))
    up = rd.resize_images(x, 4, 4)[0, :, :, 0].numpy()
    # half-pixel centres: output pixel i samples input coordinate (i + 0.5) / 2 - 0.5, clamped at the border
    want_row = np.array([0.0, 0.25, 0.75, 1.0], dtype=np.float32)
    np.testing.assert_allclose(up[0], want_row, atol=1e-6)
    np.testing.assert_allclose(up[:, 0], 2 * want_row, atol=1e-6)
    down = rd.resize_images(np.arange(16, dtype=np.float32).reshape(1, 4, 4, 1), 2, 2)[0, :, :, 0].numpy()
    np.testing.assert_allclose(down, [[2.5, 4.5], [10.5, 12.5]], atol=1e-6)   # no antialiasing: 2x2 average at the centres


def _write_llff_scene(root, n=5, hw=(24, 32), with_small=True):
    from PIL import Image
    os.makedirs(os.path.join(root, "images"), exist_ok=True)
    rng = np.random.default_rng(1)
    arr = _fake_poses_bounds(n=n, seed=5)
    np.save(os.path.join(root, "poses_bounds.npy"), arr)
    for i in range(n):
        a = rng.integers(0, 255, (hw[0] * 8, hw[1] * 8, 3), dtype=np.uint8)
        Image.fromarray(a).save(os.path.join(root, "images", f"img_{i:03d}.png"))
        if with_small:
            os.makedirs(os.path.join(root, "images_8"), exist_ok=True)
            Image.fromarray(a[::8, ::8]).save(os.path.join(root, "images_8", f"img_{i:03d}.png"))
    return arr


@pytest.mark.parametrize("with_small", [True, False])
def test_load_fern_data_reads_scene(tmp_path, with_small):
    arr = _write_llff_scene(str(tmp_path), with_small=with_small)
    imgs, poses, bds, render, i_test = rd.load_fern_data(str(tmp_path), factor=8)
    assert imgs.shape == (5, 24, 32, 3) and imgs.dtype == np.float32 and 0.0 <= imgs.min() and imgs.max() <= 1.0
    assert poses.shape == (5, 3, 5) and bds.shape == (5, 2) and render.shape == (120, 3, 5)
    np.testing.assert_allclose(poses[0, :, 4], [24.0, 32.0, 3260.5 / 8], rtol=1e-6)
    p0, b0 = R.unpack_poses_bounds(arr.copy(), (24, 32), 8)
    ref = R.llff_poses(p0, b0)
    np.testing.assert_allclose(poses, ref[0], rtol=2e-5, atol=2e-5)
    assert i_test == ref[3]
    focal, (near, far), i_train, i_val = R.fern_split(ref[0], ref[1], ref[3])
    assert abs(focal - 3260.5 / 8) < 1e-3 and near < far and len(i_train) == 4 and i_val[0] == i_test


def test_missing_files_raise(tmp_path):
    with pytest.raises(FileNotFoundError):
        rd.load_fern_data(str(tmp_path))
    with pytest.raises(FileNotFoundError):
        rd.prepare_lego_data(8, 8, npz_path=str(tmp_path / "tiny_nerf_data.npz"))
    with pytest.raises(FileNotFoundError):
        rd.load_blender_data(str(tmp_path))


def test_load_blender_data(tmp_path):
    from PIL import Image
    frames = []
    os.makedirs(tmp_path / "train")
    for i in range(3):
        a = np.zeros((16, 16, 4), dtype=np.uint8)
        a[..., 0] = 255; a[4:12, 4:12, 3] = 255                 # opaque red square on transparent background
        Image.fromarray(a).save(tmp_path / "train" / f"r_{i}.png")
        m = np.eye(4); m[2, 3] = 4.0
        frames.append({"file_path": f"./train/r_{i}", "transform_matrix": m.tolist()})
    with open(tmp_path / "transforms_train.json", "w") as f:
        json.dump({"camera_angle_x": 0.6911112070083618, "frames": frames}, f)
    imgs, poses, focal = rd.load_blender_data(str(tmp_path), "train", white_bkgd=True)
    assert imgs.shape == (3, 16, 16, 3) and poses.shape == (3, 4, 4)
    np.testing.assert_allclose(imgs[0, 0, 0], [1, 1, 1]); np.testing.assert_allclose(imgs[0, 8, 8], [1, 0, 0])
    assert abs(focal - 0.5 * 16 / np.tan(0.5 * 0.6911112070083618)) < 1e-4
    black, _, _ = rd.load_blender_data(str(tmp_path), "train", white_bkgd=False)
    np.testing.assert_allclose(black[0, 0, 0], [0, 0, 0])
