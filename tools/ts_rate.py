"""MMA issue-rate probe: TMEM-A MMAs, single CTA / CTA pair, with and without a tcgen05.commit every 4 MMAs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_keras_b200 import _lib
L = _lib.lib()
cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for pair in (0, 1):
    M = 256 if pair else 128
    for N in (128, 256):
        for K in (64, 256):
            for mode in (0, 2):
                reps = 64 * 256 // K
                a = torch.randn(M, K).cuda(); b = torch.randn(N, K).cuda(); c = torch.zeros(M, N, device="cuda")
                _lib.check(L.nerf_selftest_gemm_ts(a.data_ptr(), b.data_ptr(), c.data_ptr(), N, K, pair, reps, mode, cyc.data_ptr(), st))
                torch.cuda.synchronize()
                print(f"pair={pair} N={N} K={K} commit_every_4={mode == 2}: {cyc.item() / (reps * K // 16):.1f} cycles/MMA", flush=True)
