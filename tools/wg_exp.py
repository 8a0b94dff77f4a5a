import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_keras_b200 as nk
from nerf_keras_b200 import _lib
L = _lib.lib()
B, Nc, Nf = 4096, 64, 128
nk.set_random_seed(42)
c = nk.create_nerf_complete_model(8, 256, 4, 10, 4); f = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
tr = nk.NeRFTrainer(c, f, B, Nc, Nf, 10, 4)
tr.compile(nk.Adam(learning_rate=5e-4), nk.MeanSquaredError())
o, d = nk.get_rays(64, 64, 88.0, nk.pose_spherical(20.0, -30.0, 4.0))
o, d = o.reshape(-1, 3).contiguous(), d.reshape(-1, 3).contiguous()
t = nk.generate_t_vals(2.0, 6.0, B, Nc, True, u=np.random.default_rng(3).random(Nc, dtype=np.float32))
u = torch.rand(B, Nf, device="cuda"); img = torch.rand(B, 3, device="cuda")
for flags in (0, 8, 16, 1):
    L.nerf_debug_flags(flags)
    for _ in range(2): tr.train_step((img, (o, d, t)), u_pdf=u)
    torch.cuda.synchronize()
    L.nerf_timing_enable(1)
    for _ in range(3): tr.train_step((img, (o, d, t)), u_pdf=u)
    torch.cuda.synchronize()
    L.nerf_timing_enable(0)
    out = {}
    for kind, nm in ((0, "fwd"), (1, "chain"), (2, "wgrad")):
        ms, n = C.c_double(), C.c_int64()
        L.nerf_timing_read(kind, C.byref(ms), C.byref(n))
        out[nm] = round(ms.value / 3, 3)
    print("flags", flags, out, flush=True)
L.nerf_debug_flags(0)
