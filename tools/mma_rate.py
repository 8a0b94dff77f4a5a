import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_keras_b200 import _lib
L = _lib.lib()
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for mode in (0, 1):
    for n in (64, 128, 256):
        for reps in (64, 1024):
            _lib.check(L.nerf_selftest_mma_rate(n, reps, mode, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            cyc = int(out.item())
            print(f"mode={mode} N={n} reps={reps}: {cyc} cycles -> {cyc / (reps * 4):.1f} cycles per 128xNx16 MMA, "
                  f"{128 * n * 16 * reps * 4 / cyc:.0f} MAC/cycle")
