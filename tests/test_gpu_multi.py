"""GPU, 2+ devices: data-parallel gradients over NCCL equal the single-GPU gradients of the concatenated batch
(models.py:107 under train_tpu_lego.py:127).  Skipped on a single-GPU box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_gradients_equal_single_gpu_on_concatenated_batch():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "dp_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("DP_RESULT ")][-1]
    res = json.loads(line[len("DP_RESULT "):])
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r2_dp2_gradient_equality.json"), "w"), indent=1)
    except OSError:
        pass
    # identical per-sample arithmetic on every rank; only the order of the fp32 sums differs
    assert res["grad_rel_err"] <= 1e-3, res
    for tag in ("overlap_graph", "single_allreduce_eager"):
        r = res[tag]
        assert r["ranks_identical"] and r["steps"] == 4, (tag, r)
        assert r["update_cosine_vs_1gpu"] >= 0.98 and r["median_abs_diff_vs_1gpu"] <= 2e-5, (tag, r)
    assert res["overlap_graph"]["graphs"] == 1
    bn = res["batch_norm_dp"]
    assert bn["ranks_identical"] and np.isfinite(bn["last_loss"]) and bn["last_loss"] < bn["first_loss"], bn
