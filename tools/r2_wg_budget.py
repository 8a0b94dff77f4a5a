"""Round-2 probe of the backward (dX chain + weight gradient): usage r2_wg_budget.py budget[:flags] ...
budget = SMs given to the weight-gradient kernel (0: the context's default); flags (nerf_debug_flags): 64 = no overlap
(one kernel after the other, each on the whole GPU), 128 = budgeted weight gradient AFTER the chain (what bounds one of
its CTAs when HBM is not the limit), 8 / 2 / 4 = side jobs / MMAs / final reduction off."""
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
for spec in sys.argv[1:] or ["0", "148", "89", "60", "44", "0"]:
    budget, _, fl = spec.partition(":")
    budget, fl = int(budget), int(fl or 0)
    L.nerf_debug_flags((budget << 8) | fl)
    for _ in range(3): fb()
    L.nerf_timing_enable(1)
    for _ in range(5): fb()
    torch.cuda.synchronize(); L.nerf_timing_enable(0)
    ms, n = C.c_double(), C.c_int64()
    out = []
    for kind, name in ((1, "chain"), (2, "wgrad"), (3, "chain+wgrad")):
        L.nerf_timing_read(kind, C.byref(ms), C.byref(n))
        out.append(f"{name} {ms.value / 5:.3f}")
    print(f"budget {budget:4d} CTAs flags {fl:3d}: " + "  ".join(out) + " ms per step", flush=True)
L.nerf_debug_flags(0)
