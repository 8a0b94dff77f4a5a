"""Standalone HBM-bound kernels against the measured copy bandwidth (MEASURED_PEAKS.json): ray generation,
t-values, compositing, resample+merge, Adam.  Inputs/outputs are sized beyond L2 (126 MB) and each kernel is
timed with CUDA events over `reps` launches on rotating buffers."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerf_keras_b200 as nk
from nerf_keras_b200 import _lib

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
L = _lib.lib()
st = lambda: torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


rows = []
def report(name, ms, alg_bytes):
    gbs = alg_bytes / ms / 1e6
    rows.append((name, ms, alg_bytes, gbs, gbs / peak))
    print(f"{name:46s} {ms*1e3:9.1f} us  {alg_bytes/1e6:9.1f} MB algorithmic  {gbs:8.1f} GB/s  {gbs/peak:5.2f} of measured HBM copy", flush=True)

# get_rays: 24 B/ray, 8 frames of 2000x2000 rotated (96 MB each)
H = W = 2000
pose = nk.pose_spherical(30.0, -30.0, 4.0)
import ctypes as C
arr = (C.c_float * 12)(*[float(v) for v in pose[:3, :4].reshape(-1)])
bufs = [(torch.empty(H * W * 3, device="cuda"), torch.empty(H * W * 3, device="cuda")) for _ in range(4)]
report("get_rays 2000x2000 (24 B/ray)", timeit(lambda i: L.nerf_get_rays(H, W, 1111.0, arr, bufs[i % 4][0].data_ptr(), bufs[i % 4][1].data_ptr(), st())), H * W * 24)
# t_vals: 4N B/ray
B, N = 1 << 20, 64
u = torch.rand(N, device="cuda")
tb = [torch.empty(B * N, device="cuda") for _ in range(3)]
report("generate_t_vals 1Mi x 64 (256 B/ray)", timeit(lambda i: L.nerf_generate_t_vals(2.0, 6.0, B, N, u.data_ptr(), 0, tb[i % 3].data_ptr(), st())), B * N * 4)
# volume_render: 24 B/sample + 20 B/ray
for (Br, Ns) in ((1 << 19, 64), (1 << 18, 192)):
    sets = []
    for _ in range(3):
        preds = torch.randn(Br, Ns, 4, device="cuda"); t = torch.rand(Br, Ns, device="cuda").sort(-1)[0] * 4 + 2
        sets.append((preds, t, torch.empty(Br, 3, device="cuda"), torch.empty(Br, device="cuda"), torch.empty(Br, Ns, device="cuda"), torch.empty(Br, device="cuda")))
    def vr(i):
        p, t, r, d, w, a = sets[i % 3]
        L.nerf_volume_render(p.data_ptr(), t.data_ptr(), Br, Ns, r.data_ptr(), d.data_ptr(), w.data_ptr(), a.data_ptr(), st())
    report(f"volume_render {Br} x {Ns}", timeit(vr), Br * (Ns * 24 + 20))
    dsets = [(torch.randn(Br, 3, device="cuda"), torch.empty(Br, Ns, 4, device="cuda")) for _ in range(3)]
    def vrb(i):
        p, t = sets[i % 3][:2]; dr, dp = dsets[i % 3]
        L.nerf_volume_render_bwd(p.data_ptr(), t.data_ptr(), dr.data_ptr(), 0, Br, Ns, dp.data_ptr(), 0, st())
    report(f"volume_render_bwd {Br} x {Ns} (36 B/sample)", timeit(vrb), Br * (Ns * 36 + 12))
    del sets, dsets
# resample + merge: 1792 B/ray @ 64/128
Br, Nc, Nf = 1 << 18, 64, 128
sets = [(torch.rand(Br, Nc, device="cuda").sort(-1)[0] * 4 + 2, torch.rand(Br, Nc, device="cuda"), torch.rand(Br, Nf, device="cuda"), torch.empty(Br, Nc + Nf, device="cuda")) for _ in range(3)]
def rm(i):
    t, w, uu, o = sets[i % 3]
    L.nerf_resample_merge(t.data_ptr(), w.data_ptr(), uu.data_ptr(), Br, Nc, Nf, o.data_ptr(), 0, st())
report("resample_merge 262144 x 64+128 (1792 B/ray)", timeit(rm), Br * 1792)
del sets
# Adam: 28 B/param, 64 Mi params (flat buffers >> L2)
n = 1 << 26
p, g, m, v = (torch.randn(n, device="cuda") for _ in range(4)); v.abs_()
step = [0]
def adam(i):
    step[0] += 1
    L.nerf_adam_flat(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, step[0], 5e-4, 1.0, st())
report("adam 64Mi params (28 B/param)", timeit(adam), n * 28)
out = os.path.join(ROOT, "gpurun_out", "bench_ops.json")
json.dump([dict(name=r[0], ms=r[1], alg_bytes=r[2], gbs=r[3], frac=r[4]) for r in rows], open(out, "w"), indent=1)
