"""Self-test of MMAs with the A operand in tensor memory (single CTA and CTA pair)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_keras_b200 import _lib
L = _lib.lib()
torch.manual_seed(0)
for pair in (0, 1):
    M = 256 if pair else 128
    for N in (128, 256):
        for K in (64, 128, 256):
            a = torch.randn(M, K); b = torch.randn(N, K)
            ad, bd = a.cuda(), b.cuda()
            c = torch.zeros(M, N, device="cuda")
            _lib.check(L.nerf_selftest_gemm_ts(ad.data_ptr(), bd.data_ptr(), c.data_ptr(), N, K, pair, 1, 0, None, torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            ref = a.bfloat16().float() @ b.bfloat16().float().T
            err = (c.cpu() - ref).abs().max().item()
            print(f"pair={pair} N={N} K={K}: max err {err:.3e} (ref max {ref.abs().max():.1f})", flush=True)

cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
for pair in (0, 1):
    M = 256 if pair else 128
    for N in (128, 256):
        for ss in (0, 1):
          for K in (64, 128, 256):
            reps = 64 * 256 // K
            a = torch.randn(M, K).cuda(); b = torch.randn(N, K).cuda(); c = torch.zeros(M, N, device="cuda")
            _lib.check(L.nerf_selftest_gemm_ts(a.data_ptr(), b.data_ptr(), c.data_ptr(), N, K, pair, reps, ss, cyc.data_ptr(), torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            n_mma = reps * K // 16
            print(f"rate pair={pair} N={N} K={K} A={'smem' if ss else 'tmem'}: {cyc.item() / n_mma:.1f} cycles per M={M} x N={N} x K=16 MMA", flush=True)
