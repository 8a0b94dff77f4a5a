"""Timeline of the TS forward kernel (block 0): issuer and elected worker, cycles relative to the first event."""
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
L.nerf_debug_pair_mode(4)
for _ in range(2): tr.mlp_forward_rays("coarse", o, d, t)
buf = torch.zeros(4 * 3 * 16 * 4, dtype=torch.int64, device="cuda")
L.nerf_debug_trace(buf.data_ptr())
tr.mlp_forward_rays("coarse", o, d, t)
torch.cuda.synchronize()
L.nerf_debug_trace(None)
a = buf.cpu().numpy().reshape(4, 3, 16, 4)
t0 = a[[0, 2]][a[[0, 2]] > 0].min()
r = lambda x: int(x - t0) if x > 0 else -1
for tile in range(3):
    print(f"tile {tile}: issuer [start, aready0, aready1, issued]   worker [wait h0, h0 ready, wait h1, done]")
    for ph in range(11):
        i, w = a[0, tile, ph], a[2, tile, ph]
        x = a[1, tile, ph]
        print(f"  ph{ph:2d} | iss {r(i[0]):7d} {r(i[1]):7d} {r(i[3]):7d} {r(i[2]):7d} | wrk {r(w[0]):7d} {r(w[1]):7d} {r(w[3]):7d} {r(w[2]):7d} | wait full {x[0]:6d} pfull {x[1]:6d} issue {x[2]:6d}")
