"""Round-2 probe: per-CTA timeline of the weight-gradient kernel (fine net) when it runs next to the dX chain.
usage: r2_wg_stats.py budget[:flags]"""
import os, sys, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_keras_b200 as nk
from nerf_keras_b200 import _lib
from nerf_keras_b200.models import _ptr, _stream
L = _lib.lib()
B, Nc, Nf = 4096, 64, 128
nk.set_random_seed(42)
c = nk.create_nerf_complete_model(8, 256, 4, 10, 4); f = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
tr = nk.NeRFTrainer(c, f, B, Nc, Nf, 10, 4); tr.compile(nk.Adam(5e-4), nk.MeanSquaredError())
o, d = nk.get_rays(64, 64, 88.0, nk.pose_spherical(20.0, -30.0, 4.0))
o, d = o.reshape(-1, 3).contiguous(), d.reshape(-1, 3).contiguous()
t = nk.generate_t_vals(2.0, 6.0, B, Nc, True)
u = torch.rand(B, Nf, device="cuda"); img = torch.rand(B, 3, device="cuda")
metrics = torch.empty(3, device="cuda")
def fb():
    _lib.check(L.nerf_train_forward_backward(tr._ctx.handle, _ptr(img), _ptr(o), _ptr(d), _ptr(t), _ptr(u), B, _ptr(metrics), _stream()), "fb")
for spec in sys.argv[1:] or ["60"]:
    budget, _, fl = spec.partition(":")
    L.nerf_debug_flags((int(budget) << 8) | int(fl or 0))
    for _ in range(3): fb()
    stats = torch.zeros(2 * 148 * 8, dtype=torch.int64, device="cuda")
    L.nerf_debug_wgrad_stats(_ptr(stats))
    fb(); torch.cuda.synchronize()
    L.nerf_debug_wgrad_stats(None)
    s = stats.cpu().view(2, 148, 8)[1]
    s = s[s[:, 6] > 0]
    t0 = int(s[:, 6].min())
    print(f"== budget {budget} flags {fl or 0}: {len(s)} weight-gradient CTAs (fine net); times in us after the first CTA start")
    print("cta job tiles start  loader_end  cta_end  wait_chain  wait_slot")
    for i, r in enumerate(s.tolist()):
        print(f"{i:3d} {r[0]:3d} {r[1]:5d} {(r[6]-t0)/1e3:7.0f} {(r[5]-t0)/1e3:9.0f} {(r[2]-t0)/1e3:8.0f} {r[3]/1e3:10.0f} {r[4]/1e3:9.0f}")
L.nerf_debug_flags(0)
