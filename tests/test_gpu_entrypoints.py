"""GPU: the reference-named entry scripts run end to end on the B200 path (synthetic data)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable] + args, cwd=cwd, capture_output=True, text=True, timeout=280, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return r.stdout


def test_train_lego_and_inference_scripts(tmp_path):
    cfg = os.path.join(ROOT, "config", "lego_batch_debug.json")
    out = _run([os.path.join(ROOT, "train_lego.py"), "--config", cfg, "--steps-per-epoch", "30", "--views", "5"], str(tmp_path))
    lines = [l for l in out.splitlines() if l.startswith("Epoch")]
    assert len(lines) == 2
    loss = [float(l.split("loss: ")[1].split()[0]) for l in lines]
    assert loss[1] < loss[0]                      # the fine loss goes down on the procedural scene
    w = tmp_path / "models" / "nerf_lego_l8_d256_n48_lego_batch_debug.npz"
    assert w.exists()
    _run([os.path.join(ROOT, "inference.py"), "--config", cfg, "--weights", str(w), "--frames", "2", "--out", "f.npy"],
         str(tmp_path))
    frames = np.load(tmp_path / "f.npy")
    assert frames.shape == (2, 100, 100, 3) and frames.dtype == np.uint8 and frames.std() > 0


def test_train_fern_script(tmp_path):
    cfg = os.path.join(ROOT, "config", "fern_batch_debug.json")
    out = _run([os.path.join(ROOT, "train_fern.py"), "--config", cfg, "--steps-per-epoch", "10", "--views", "4"], str(tmp_path))
    assert out.count("Epoch") == 2 and "nan" not in out.lower()
