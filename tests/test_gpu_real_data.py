"""Real-data loaders end to end on small generated scenes: rays from the CUDA `get_rays` kernel, return structure of
`prepare_lego_data` (lego_data_utils.py:8-51) and `prepare_fern_data` (fern_data_utils.py:462-520)."""
import json
import os

import numpy as np
import pytest
import torch

import oracle as O
from oracle import llff_ref as R
from tests.test_real_data_cpu import _write_llff_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nk():
    import nerf_keras_b200 as nk
    return nk


def test_prepare_fern_data(nk, tmp_path):
    from nerf_keras_b200 import real_data as rd
    arr = _write_llff_scene(str(tmp_path), n=5, hw=(24, 32))
    train, val, (near, far), focal = rd.prepare_fern_data(12, 16, datadir=str(tmp_path))
    p0, b0 = R.unpack_poses_bounds(arr.copy(), (24, 32), 8)
    poses, bds, _, i_test = R.llff_poses(p0, b0)
    f_ref, (near_ref, far_ref), i_train, _ = R.fern_split(poses, bds, i_test)
    assert abs(float(focal) - float(f_ref)) < 1e-3 and abs(near - near_ref) < 1e-5 and abs(far - far_ref) < 1e-4
    for part, n_views in ((train, 4), (val, 1)):
        assert all(t.is_cuda and t.dtype == torch.float32 and t.shape == (n_views * 12 * 16, 3) for t in part)
    # reference quirk kept: rays at the TARGET size with the UN-rescaled focal (fern_data_utils.py:485-489)
    o_ref, d_ref = O.get_rays(12, 16, float(f_ref), torch.from_numpy(poses[i_test, :3, :4]))
    np.testing.assert_allclose(val[1].cpu().numpy(), o_ref.reshape(-1, 3).numpy(), atol=1e-5)
    np.testing.assert_allclose(val[2].cpu().numpy(), d_ref.reshape(-1, 3).numpy(), atol=1e-5)
    assert 0.0 <= float(train[0].min()) and float(train[0].max()) <= 1.0


def test_prepare_lego_data(nk, tmp_path):
    from nerf_keras_b200 import real_data as rd
    rng = np.random.default_rng(0)
    n = 10
    poses = np.stack([np.asarray(O.pose_spherical(float(t), -30.0, 4.0)) for t in np.linspace(-180, 180, n, endpoint=False)])
    path = str(tmp_path / "tiny_nerf_data.npz")
    np.savez(path, images=rng.random((n, 20, 20, 3), dtype=np.float32), poses=poses.astype(np.float32),
             focal=np.array(27.7, dtype=np.float64))
    train, val, bounds, focal = rd.prepare_lego_data(10, 10, npz_path=path)
    assert bounds == (2.0, 6.0) and float(focal) == 27.7
    assert train[0].shape == (8 * 100, 3) and val[0].shape == (2 * 100, 3)
    assert train[1].shape == train[2].shape == (800, 3) and val[1].shape == (200, 3)
    o_ref, d_ref = O.get_rays(10, 10, 27.7, torch.from_numpy(poses[0].astype(np.float32)))
    assert np.array_equal(train[1][:100].cpu().numpy(), o_ref.reshape(-1, 3).numpy())
    assert np.array_equal(train[2][:100].cpu().numpy(), d_ref.reshape(-1, 3).numpy())


def test_prepare_blender_data_trains(nk, tmp_path):
    """A generated Blender-format scene goes through the loader, the on-device batch sampler and one train step."""
    from PIL import Image
    from nerf_keras_b200 import real_data as rd
    from nerf_keras_b200.synthetic import BatchedRayDataset
    for split, k in (("train", 4), ("val", 2)):
        os.makedirs(tmp_path / split)
        frames = []
        for i in range(k):
            a = np.full((16, 16, 4), 255, dtype=np.uint8)
            a[..., 1] = 40 * i
            Image.fromarray(a).save(tmp_path / split / f"r_{i}.png")
            frames.append({"file_path": f"./{split}/r_{i}",
                           "transform_matrix": np.asarray(O.pose_spherical(40.0 * i, -30.0, 4.0)).tolist()})
        with open(tmp_path / f"transforms_{split}.json", "w") as f:
            json.dump({"camera_angle_x": 0.6911112070083618, "frames": frames}, f)
    train, val, (near, far), focal = rd.prepare_blender_data(16, 16, str(tmp_path))
    assert train[0].shape == (4 * 256, 3) and val[0].shape == (2 * 256, 3)
    ds = BatchedRayDataset(*train, 64, 256, near, far, shuffle=True, steps_per_epoch=1)
    nk.set_random_seed(0)
    tr = nk.NeRFTrainer(nk.create_nerf_complete_model(8, 256, 4, 10, 4), nk.create_nerf_complete_model(8, 256, 4, 10, 4),
                        256, 64, 128, 10, 4)
    tr.compile(nk.Adam(5e-4), nk.MeanSquaredError())
    batch = next(iter(ds))
    out = tr.train_step(batch)
    assert np.isfinite(float(out["loss"])) and np.isfinite(float(out["psnr"]))
