"""Generates tests/golden/*.npz from the oracle (the reference itself cannot run here: TF/Keras are
not installable).  Run from the repo root:  python tests/golden/make_golden.py
The fixtures pin the oracle against regressions (CPU suite) and are what the CUDA path is compared
with on the GPU box, where /root/reference and a second oracle run are not needed."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from oracle.models_ref import compute_grads  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def case(name, H, W, focal, pose, near, far, B, Nc, Nf, seed):
    rng = np.random.default_rng(seed)
    o, d = O.get_rays(H, W, focal, pose)
    sel = rng.choice(H * W, size=B, replace=False)
    oo, dd = o.reshape(-1, 3)[sel].contiguous(), d.reshape(-1, 3)[sel].contiguous()
    u_t = rng.random(Nc, dtype=np.float32)
    u_pdf = rng.random((B, Nf), dtype=np.float32)
    img = rng.random((B, 3), dtype=np.float32)
    t = O.generate_t_vals(near, far, B, Nc, True, u=u_t)
    t_nojit = O.generate_t_vals(near, far, B, Nc, False)
    wc, wf = O.init_weights(seed + 100, 0.1), O.init_weights(seed + 101, 0.1)
    rays, dirs = O.sample_rays(oo, dd, t)
    enc_x, enc_d = O.encode_position(rays, 10), O.encode_position(dirs, 4)
    with torch.no_grad():
        rgbs, depths, ws, preds, t_all = O.forward_pass(wc, wf, oo, dd, t, 10, 4, Nf, torch.from_numpy(u_pdf))
        t_mid = 0.5 * (t[:, 1:] + t[:, :-1])
        t_fine = O.sample_pdf(t_mid, ws[0], Nf, u=torch.from_numpy(u_pdf))
    grads_stop, metrics = compute_grads(wc, wf, torch.from_numpy(img), oo, dd, t, 10, 4, Nf, torch.from_numpy(u_pdf),
                                        stop_grad_samples=True)
    grads_ref, _ = compute_grads(wc, wf, torch.from_numpy(img), oo, dd, t, 10, 4, Nf, torch.from_numpy(u_pdf),
                                 stop_grad_samples=False)
    flat = lambda gs: np.concatenate([g.numpy().reshape(-1) for g in gs])
    norms = lambda gs: np.array([float(np.linalg.norm(g.numpy().astype(np.float64))) for g in gs], dtype=np.float64)
    out = dict(H=H, W=W, focal=np.float32(focal), pose=pose, near=near, far=far, Nc=Nc, Nf=Nf, sel=sel,
               rays_o_full=o.numpy(), rays_d_full=d.numpy(), o=oo.numpy(), d=dd.numpy(), u_t=u_t, u_pdf=u_pdf, img=img,
               t=t.numpy(), t_nojit=t_nojit.numpy(), pts=rays.numpy(), enc_x=enc_x.numpy(), enc_d=enc_d.numpy(),
               w_seed_coarse=seed + 100, w_seed_fine=seed + 101, w_bias_range=0.1,
               w_coarse_sample=O.flatten_weights(wc)[::997], w_fine_sample=O.flatten_weights(wf)[::997],
               rgb_c=rgbs[0].numpy(), rgb_f=rgbs[1].numpy(), depth_c=depths[0].numpy(), depth_f=depths[1].numpy(),
               wt_c=ws[0].numpy(), wt_f=ws[1].numpy(), pred_c=preds[0].numpy(), pred_f=preds[1].numpy(),
               t_fine=t_fine.numpy(), t_all=t_all.numpy(), grads_stop_sample=flat(grads_stop)[::61], grads_ref_sample=flat(grads_ref)[::61],
               grads_stop_norms=norms(grads_stop), grads_ref_norms=norms(grads_ref),
               metrics=np.array([metrics["loss_coarse"], metrics["loss"], metrics["psnr"]], dtype=np.float32))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "rgb_f[0]", out["rgb_f"][0], "metrics", out["metrics"])


if __name__ == "__main__":
    torch.set_num_threads(8)
    # Lego-shaped: spherical pose, near/far 2/6 (lego_data_utils.py:48-49), debug sample counts
    case("lego_small", 20, 20, 27.7778, O.pose_spherical(37.0, -30.0, 4.0), 2.0, 6.0, 96, 16, 32, 7)
    # Fern-shaped (pinhole, near/far from bounds as fern_data_utils.py:495-496), full sample counts
    pose = np.eye(4, dtype=np.float32)
    pose[:3, 3] = [0.12, -0.2, 0.05]
    c, s = np.cos(np.float32(0.1)), np.sin(np.float32(0.1))
    pose[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float32)
    case("fern_small", 12, 16, 13.0, pose, 1.2, 12.0, 40, 64, 128, 11)
