"""Step time of BATCH_NORM=true training (layer-by-layer fp32 path) at the reference's BN config shapes."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_keras_b200 as nk
for B, Nc, Nf in ((256, 16, 32), (512, 16, 32), (512, 64, 128), (4096, 64, 128)):
    nk.set_random_seed(0)
    mk = lambda: nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=True)
    tr = nk.NeRFTrainer(mk(), mk(), B, Nc, Nf, 10, 4); tr.compile(nk.Adam(5e-4), nk.MeanSquaredError())
    o, d = nk.get_rays(64, 64, 88.0, nk.pose_spherical(20.0, -30.0, 4.0))
    o, d = o.reshape(-1, 3)[:B].contiguous(), d.reshape(-1, 3)[:B].contiguous()
    t = nk.generate_t_vals(2.0, 6.0, B, Nc, True); img = torch.rand(B, 3, device="cuda")
    for _ in range(3): tr.train_step((img, (o, d, t)))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): tr.train_step((img, (o, d, t)))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"BN training  B={B:5d} samples {Nc}+{Nf}: {ms:8.2f} ms/step  {B / ms * 1e3:10.0f} rays/s", flush=True)
