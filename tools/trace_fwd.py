import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_keras_b200 as nk
from nerf_keras_b200 import _lib
L = _lib.lib()
B, Nc, Nf = 4096, 64, 128
nk.set_random_seed(42)
c = nk.create_nerf_complete_model(8, 256, 4, 10, 4); f = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
tr = nk.NeRFTrainer(c, f, B, Nc, Nf, 10, 4); tr.build()
o, d = nk.get_rays(64, 64, 88.0, nk.pose_spherical(20.0, -30.0, 4.0))
o, d = o.reshape(-1, 3).contiguous(), d.reshape(-1, 3).contiguous()
t = nk.generate_t_vals(2.0, 6.0, B, Nc, False)
L.nerf_debug_pair_mode(int(os.environ.get("PAIR", "0")))
SAVE = int(os.environ.get("SAVE", "0"))       # 1: trace the training variant (saves operand images)
if SAVE:
    tr.compile(nk.Adam(5e-4), nk.MeanSquaredError())
    dp = torch.randn(B, Nc, 4, device="cuda")
    run = lambda: tr.debug_mlp_grads("coarse", o, d, t, dp)
else:
    run = lambda: tr.mlp_forward_rays("coarse", o, d, t)
for _ in range(2): run()
buf = torch.zeros(4 * 3 * 16 * 4, dtype=torch.int64, device="cuda")
L.nerf_debug_trace(buf.data_ptr())
run()
torch.cuda.synchronize()
L.nerf_debug_trace(None)
a = buf.cpu().numpy().reshape(4, 3, 16, 4)
t0 = a[..., :3][a[..., :3] > 0].min()
names = ["issA", "issB", "wrkA", "wrkB"]
for tile in range(2):
    print(f"tile {tile}: per phase [wait_start, ready, done] relative cycles")
    for ph in range(11):
        row = []
        for who in range(4):
            e = a[who, tile, ph]
            row.append(f"{names[who]}: " + " ".join(f"{(x - t0) if x > 0 else -1:7d}" for x in e[:3]))
            if who < 2 and e[3] > 0: row[-1] += f" (wait full {int(e[3]) & 0xffffffff} pfull {int(e[3]) >> 32})"
        print(f"  ph{ph:2d} | " + " | ".join(row))
