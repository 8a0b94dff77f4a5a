"""Per-layer gradient error of the tcgen05 backward vs (a) the fp32 oracle and (b) a bf16-emulating
oracle (weights/activations rounded to bf16 at the points the kernels round them)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
from oracle.models_ref import _params
import nerf_keras_b200 as nk
from tests.util import golden_weights, load_golden

def rnd(x):
    return x + (x.bfloat16().float() - x).detach()

def mlp_bf16(w, enc_x, enc_d):
    x = rnd(enc_x)
    for i in range(8):
        p = w[f"d{i}"]
        x = torch.relu(x @ rnd(p["W"]) + p["b"])
        x = rnd(x)
        if i == 4:
            x = torch.cat([x, rnd(enc_x)], -1)
    sigma = x @ w["sigma"]["W"] + w["sigma"]["b"]
    feat = rnd(x @ rnd(w["feature"]["W"]) + w["feature"]["b"])
    Wd = w["ddir"]["W"]
    hd = torch.relu(feat @ rnd(Wd[:256]) + enc_d @ Wd[256:] + w["ddir"]["b"])
    rgb = hd @ w["rgb"]["W"] + w["rgb"]["b"]
    return torch.cat([rgb, sigma], -1)

name, net = sys.argv[1], sys.argv[2]
g = load_golden(name)
wc, wf = golden_weights(g)
w = wc if net == "coarse" else wf
t = g["t"] if net == "coarse" else g["t_all"]
d_preds = torch.randn(t.shape + (4,), generator=torch.Generator().manual_seed(5)) * 0.1
o, d, tt = map(torch.from_numpy, (g["o"], g["d"], t))
params = _params(w)
for p in params: p.requires_grad_(True)
rays, dirs = O.sample_rays(o, d, tt)
ex, ed = O.encode_position(rays, 10), O.encode_position(dirs, 4)
g32 = torch.autograd.grad((O.nerf_mlp(w, ex, ed) * d_preds).sum(), params)
g16 = torch.autograd.grad((mlp_bf16(w, ex, ed) * d_preds).sum(), params)
for p in params: p.requires_grad_(False)
mc = nk.create_nerf_complete_model(8, 256, 4, 10, 4); mf = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
mc.set_flat_weights(O.flatten_weights(wc)); mf.set_flat_weights(O.flatten_weights(wf))
tr = nk.NeRFTrainer(mc, mf, g["o"].shape[0], int(g["Nc"]), int(g["Nf"]), 10, 4)
tr.compile(nk.Adam(learning_rate=5e-4), nk.MeanSquaredError())
preds, grads = tr.debug_mlp_grads(net, g["o"], g["d"], t, d_preds.numpy())
got = grads.cpu().numpy(); off = 0
print(f"{name} {net}: layer | rel err vs fp32 | rel err vs bf16-emulation | emulation vs fp32 | norm")
for (role, fi, fo), i in zip(O.layer_shapes(), range(12)):
    for kind, n, ref32, ref16 in (("W", fi * fo, g32[2 * i], g16[2 * i]), ("b", fo, g32[2 * i + 1], g16[2 * i + 1])):
        a = got[off:off + n]; off += n
        r32, r16 = ref32.numpy().reshape(-1), ref16.numpy().reshape(-1)
        nn = np.linalg.norm(r32) + 1e-20
        print(f"  {role}/{kind}: {np.linalg.norm(a - r32) / nn:.4f}  {np.linalg.norm(a - r16) / (np.linalg.norm(r16) + 1e-20):.4f}  "
              f"{np.linalg.norm(r16 - r32) / nn:.4f}  {nn:.4e}")
if role == "ddir":
    pass
