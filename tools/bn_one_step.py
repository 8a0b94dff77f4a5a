import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import nerf_keras_b200 as nk
B, Nc, Nf = 512, 64, 128
nk.set_random_seed(0)
mk = lambda: nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=True)
tr = nk.NeRFTrainer(mk(), mk(), B, Nc, Nf, 10, 4); tr.compile(nk.Adam(5e-4), nk.MeanSquaredError())
o, d = nk.get_rays(64, 64, 88.0, nk.pose_spherical(20.0, -30.0, 4.0))
o, d = o.reshape(-1, 3)[:B].contiguous(), d.reshape(-1, 3)[:B].contiguous()
t = nk.generate_t_vals(2.0, 6.0, B, Nc, True); img = torch.rand(B, 3, device="cuda")
for _ in range(4): tr.train_step((img, (o, d, t)))
torch.cuda.synchronize()
