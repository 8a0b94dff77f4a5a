"""Experiment: training step with an opt-in variant of the forward kernel (nerf_debug_pair_mode: 49 / 17 CTA pair with
shared chunks, 64 / 320 weight-stationary MMAs) vs the default single-CTA kernel: per-kernel CUDA-event times (eager
steps).  usage: r2_pair_train.py [mode ...]"""
import ctypes as C, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nerf_keras_b200 as nk
from nerf_keras_b200 import _lib, models as nkm
import bench as B
conf = B.load_conf("config/lego_batch_h256.json")
dev = torch.device("cuda", 0)
L = _lib.lib()
tr = B.make_trainer(nk, conf, 4096, 64, 128, True, use_graph=False)
db, _ = B.make_batches(nk, torch, dev, B.scene_of("config/lego_batch_h256.json", conf, "pinhole"), conf, "pinhole", 4096, 64, 128, 4, 0, False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(mode, steps=10):
    L.nerf_debug_pair_mode(mode)
    fn = lambda i: tr.train_step((db[i % 4][0], db[i % 4][1:4]))
    for i in range(4): fn(i)
    nkm.set_kernel_timing(True)
    ms = B.timed_region(fn, steps, flush, 1, dev, torch) / steps
    nkm.set_kernel_timing(False)
    out = {"step_ms": ms, "loss": float(tr.loss_tracker.result())}
    for kind, nm in ((0, "fwd"), (1, "chain"), (2, "wgrad")):
        a, n = C.c_double(), C.c_int64()
        L.nerf_timing_read(kind, C.byref(a), C.byref(n))
        out[nm] = a.value / steps
    return out
for rep in range(2):
    for mode in ([int(x) for x in sys.argv[1:]] or (0, 49, 17)):
        print(mode, json.dumps(run(mode)))
L.nerf_debug_pair_mode(0)
