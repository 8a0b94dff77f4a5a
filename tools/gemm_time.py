"""Times the split-bf16 tcgen05 GEMMs of the BATCH_NORM path (csrc/gemm_tc.cu) at the 4096 x 192 step's shapes and
compares them with torch.matmul in fp32 (cuBLAS sgemm, what the path used before)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_keras_b200 import _lib
from nerf_keras_b200.models import _ptr, _stream
L = _lib.lib()
torch.backends.cuda.matmul.allow_tf32 = False
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 192
def timeit(fn, n=5):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for name, ta, tb, N, K in (("rows  NN 256x256", 0, 0, 256, 256), ("rows  NT 256x256", 0, 1, 256, 256), ("rows  NN 256x63 ", 0, 0, 256, 63),
                           ("rows  NN 128x256", 0, 0, 128, 256), ("trans TN 256x256", 1, 0, 256, 256), ("trans TN 63x256 ", 1, 0, 256, 63),
                           ("head  NN 1x256  ", 0, 0, 1, 256), ("head  TN 256x1  ", 1, 0, 1, 256)):
    if ta:
        A = torch.randn(M, K, device="cuda"); B = torch.randn(M, N, device="cuda"); Cc = torch.zeros(K, N, device="cuda")
        ours = timeit(lambda: L.nerf_selftest_gemm_f32(1, 0, K, N, M, _ptr(A), K, _ptr(B), N, 1.0, _ptr(Cc), N, 0, _stream()))
        ours3 = None
        ref = timeit(lambda: torch.addmm(Cc, A.t(), B))
        flop = 2.0 * M * N * K
    else:
        A = torch.randn(M, K, device="cuda"); B = torch.randn((N, K) if tb else (K, N), device="cuda"); Cc = torch.empty(M, N, device="cuda")
        ours = timeit(lambda: L.nerf_selftest_gemm_f32(0, tb, M, N, K, _ptr(A), K, _ptr(B), B.shape[1], 0.0, _ptr(Cc), N, 0, _stream()))
        ours3 = timeit(lambda: L.nerf_selftest_gemm_f32(0, tb, M, N, K, _ptr(A), K, _ptr(B), B.shape[1], 0.0, _ptr(Cc), N, 1, _stream()))
        ref = timeit(lambda: torch.matmul(A, B.t() if tb else B, out=Cc))
        flop = 2.0 * M * N * K
    byts = 4.0 * (A.numel() + (B.numel() if ta else Cc.numel()))
    three = f", three-way split {ours3:.3f} ms" if ours3 is not None else ""
    print(f"{name} M={M}: tcgen05 split-bf16 {ours:.3f} ms ({flop / ours / 1e9:.1f} TFLOP/s fp32-equivalent, {byts / ours / 1e6:.0f} GB/s){three}   sgemm {ref:.3f} ms")
