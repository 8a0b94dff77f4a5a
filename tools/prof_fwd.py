"""Small fixed workload for ncu: a few NeRFTrainer.forward_pass calls at the headline shape."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_keras_b200 as nk  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "render"
B, Nc, Nf = 4096, 64, 128
nk.set_random_seed(42)
c = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
f = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
tr = nk.NeRFTrainer(c, f, B, Nc, Nf, 10, 4)
if mode == "train":
    tr.compile(nk.Adam(learning_rate=5e-4), nk.MeanSquaredError())
else:
    tr.build()
o, d = nk.get_rays(64, 64, 88.0, nk.pose_spherical(20.0, -30.0, 4.0))
o, d = o.reshape(-1, 3).contiguous(), d.reshape(-1, 3).contiguous()
t = nk.generate_t_vals(2.0, 6.0, B, Nc, True, u=np.random.default_rng(3).random(Nc, dtype=np.float32))
u = torch.rand(B, Nf, device="cuda")
img = torch.rand(B, 3, device="cuda")
for i in range(3):
    if mode == "train":
        m = tr.train_step((img, (o, d, t)), u_pdf=u)
    else:
        out = tr.forward_pass(o, d, t, u_pdf=u)
torch.cuda.synchronize()
print("done", mode)
