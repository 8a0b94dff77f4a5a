import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_keras_b200 import _lib
L = _lib.lib()
for N, K in [(256, 64), (256, 256), (128, 128)]:
    g = torch.Generator().manual_seed(N + K)
    a = torch.randn(256, K, generator=g); b = torch.randn(N, K, generator=g)
    ref = a.bfloat16().float() @ b.bfloat16().float().T
    ad, bd = a.cuda(), b.cuda(); c = torch.zeros(256, N, device="cuda")
    _lib.check(L.nerf_selftest_gemm_2cta(ad.data_ptr(), bd.data_ptr(), c.data_ptr(), N, K, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    err = (c.cpu() - ref).abs()
    print(f"2cta N={N} K={K}: max err {err.max().item():.3e}; rows0-127 {err[:128].max().item():.3e} rows128-255 {err[128:].max().item():.3e}; "
          f"cols0-{N//2-1} {err[:, :N//2].max().item():.3e} cols{N//2}- {err[:, N//2:].max().item():.3e}", flush=True)
