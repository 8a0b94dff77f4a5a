"""tcgen05 collector probe: correctness and cycles of plain / .ws (B kept) / A-kept MMAs (nerf_selftest_collector)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_keras_b200 import _lib
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
names = {0: "plain, 4 accumulators", 1: ".ws B keep/reuse, 4 accumulators", 2: "A keep/reuse, 4 accumulators", 3: "plain, 1 accumulator",
         4: "plain, 2 accumulators alternating", 5: "plain, accumulator changes every 16 MMAs", 6: ".ws B keep/reuse, 2 accumulators"}
N, K = 256, 64
a1, a2, b = torch.randn(128, K, device="cuda"), torch.randn(128, K, device="cuda"), torch.randn(N, K, device="cuda")
r1 = a1.bfloat16().float() @ b.bfloat16().float().T
r2 = a2.bfloat16().float() @ b.bfloat16().float().T
for v in range(7):
    c1, c2 = torch.zeros(128, N, device="cuda"), torch.zeros(128, N, device="cuda")
    cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
    msg = ""
    if v <= 2:
        _lib.check(L.nerf_selftest_collector(a1.data_ptr(), a2.data_ptr(), b.data_ptr(), c1.data_ptr(), c2.data_ptr(), N, K, v, 1,
                                             cyc.data_ptr(), st), "probe")
        torch.cuda.synchronize()
        msg = f"max err {(c1 - r1).abs().max().item():.1e} {(c2 - r2).abs().max().item():.1e}; "
    reps = 1024
    for _ in range(2):
        _lib.check(L.nerf_selftest_collector(a1.data_ptr(), a2.data_ptr(), b.data_ptr(), c1.data_ptr(), c2.data_ptr(), N, K, v, reps,
                                             cyc.data_ptr(), st), "probe")
        torch.cuda.synchronize()
    print(f"variant {v} ({names[v]}): {msg}{cyc.item() / (reps * 16):.1f} cycles per 128x128x16 MMA", flush=True)
