import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_keras_b200 as nk
from nerf_keras_b200 import _lib
L = _lib.lib()
B, Nc, Nf = 4096 + 37, 64, 128
nk.set_random_seed(42)
c = nk.create_nerf_complete_model(8, 256, 4, 10, 4); f = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
w = c.get_flat_weights(); w[-4:] = [0.1, -0.2, 0.3, 0.05]
c.set_flat_weights(w + np.random.default_rng(0).normal(0, 0.01, w.shape).astype(np.float32))
tr = nk.NeRFTrainer(c, f, B, Nc, Nf, 10, 4); tr.build()
o, d = nk.get_rays(80, 80, 100.0, nk.pose_spherical(20.0, -30.0, 4.0))
o, d = o.reshape(-1, 3)[:B].contiguous(), d.reshape(-1, 3)[:B].contiguous()
t = nk.generate_t_vals(2.0, 6.0, B, Nc, True, u=np.random.default_rng(3).random(Nc, dtype=np.float32))
u = torch.rand(B, Nf, device="cuda")
L.nerf_debug_pair_mode(0)
ref = tr.mlp_forward_rays("coarse", o, d, t).clone()
torch.cuda.synchronize()
for pm in (1, 17, 49, 64, 320):
    L.nerf_debug_pair_mode(pm)
    for rep in range(2):
        got = tr.mlp_forward_rays("coarse", o, d, t).clone()
        torch.cuda.synchronize()
        bad = ((got - ref).abs() > 1e-3).any(-1).any(-1)
        print(f"pair mode {pm} rep {rep}: max abs diff {(got - ref).abs().max().item():.3e}; rays with diff {int(bad.sum())} of {B}; "
              f"first bad {bad.nonzero()[:8].flatten().tolist()}", flush=True)
for mode in (0, 1, 17, 49, 64, 320):
    L.nerf_debug_pair_mode(mode)
    for _ in range(3): tr.forward_pass(o, d, t, u_pdf=u, maps_only=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): tr.forward_pass(o, d, t, u_pdf=u, maps_only=True)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"mode {mode}: forward_pass {ms:.3f} ms  {B / ms * 1e3:.3e} rays/s  {B * 256 * 1186816 / ms / 1e9:.0f} TFLOP/s", flush=True)
L.nerf_debug_pair_mode(0)
