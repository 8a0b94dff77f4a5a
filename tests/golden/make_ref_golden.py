"""Generates tests/golden/ref_source.npz by EXECUTING THE REFERENCE'S OWN SOURCE (/root/reference/data_utils.py and
models.py, unmodified) under the keras / tensorflow API stand-ins of oracle/refshim (torch-CPU fp32, one eager op per
call; TensorFlow and Keras themselves cannot be installed here -- see oracle/refshim/README.md for what this pins).

Run here (the container that has /root/reference):   python tests/golden/make_ref_golden.py
The GPU box has no /root/reference; tests read only the committed .npz.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("NERF_REFERENCE_ROOT", "/root/reference")


def import_reference():
    """(data_utils, models, refshim_core) of the reference, imported from REF with the stand-ins shadowing keras / tensorflow."""
    shim = os.path.join(ROOT, "oracle", "refshim")
    for p in (REF, shim):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import refshim_core
    import keras
    import tensorflow
    assert "refshim" in keras.__file__ and "refshim" in tensorflow.__file__, "a real keras / tensorflow shadows the stand-ins"
    import data_utils as RD
    import models as RM
    assert RD.__file__.startswith(REF) and RM.__file__.startswith(REF), "data_utils / models were not imported from the reference"
    return RD, RM, refshim_core


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def set_model_weights(model, w):
    """oracle role-keyed weights -> the reference model's Dense layers (creation order d0..d7, sigma, feature, ddir, rgb)."""
    dense = [l for l in model.layers if hasattr(l, "kernel")]
    assert len(dense) == len(w)
    for layer, role in zip(dense, w):
        assert tuple(layer.kernel.shape) == tuple(w[role]["W"].shape), role
        with torch.no_grad():
            layer.kernel.copy_(w[role]["W"])
            layer.bias.copy_(w[role]["b"])


def model_weights(model):
    return np.concatenate([np.concatenate([l.kernel.detach().numpy().reshape(-1), l.bias.detach().numpy().reshape(-1)])
                           for l in model.layers if hasattr(l, "kernel")])


def lego_pose(RD, th=37.0, ph=-30.0):
    return np.asarray(RD.pose_spherical(th, ph, 4.0), dtype=np.float32)


def fern_pose():
    pose = np.eye(4, dtype=np.float32)
    pose[:3, 3] = [0.12, -0.2, 0.05]
    c, s = np.cos(np.float32(0.1)), np.sin(np.float32(0.1))
    pose[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float32)
    return pose


def grab_gradients(opt):
    """Record the gradients the reference hands to optimizer.apply_gradients (models.py:107)."""
    grabbed, real = [], opt.apply_gradients

    def apply(gv):
        gv = list(gv)
        grabbed.append([g.detach().clone() for g, _ in gv])
        real(gv)

    opt.apply_gradients = apply
    return grabbed


def main():
    sys.path.insert(0, ROOT)
    import oracle as O          # only for the seeded weight initialisation (an INPUT of both sides)
    RD, RM, core = import_reference()
    import keras
    out = {}
    rng = np.random.default_rng(2024)

    # ---- get_rays: BASELINE shapes (hash + strided sample) and a small full case -----------------
    for name, H, W, focal, pose in (("lego800", 800, 800, np.float32(0.5 * 800 / np.tan(0.5 * 0.6911112)), lego_pose(RD)),
                                    ("fern378", 378, 504, np.float32(407.6), fern_pose()),
                                    ("odd", 7, 5, np.float32(3.3), lego_pose(RD, -101.0, -7.0))):
        o, d = RD.get_rays(H, W, focal, torch.from_numpy(pose))
        o, d = o.numpy(), d.numpy()
        out[f"rays_{name}_H"], out[f"rays_{name}_W"], out[f"rays_{name}_focal"], out[f"rays_{name}_pose"] = H, W, focal, pose
        out[f"rays_{name}_o_sha"], out[f"rays_{name}_d_sha"] = sha(o), sha(d)
        out[f"rays_{name}_d_sample"] = d.reshape(-1, 3)[::997].copy()
        if H * W < 100:
            out[f"rays_{name}_o"], out[f"rays_{name}_d"] = o, d
    out["pose_spherical_cases"] = np.array([[37.0, -30.0, 4.0], [-101.0, -7.0, 4.0], [180.0, -90.0, 2.5]], np.float32)
    out["pose_spherical"] = np.stack([np.asarray(RD.pose_spherical(*map(float, c)), np.float32)
                                      for c in out["pose_spherical_cases"]])

    # ---- generate_t_vals (shared jitter vector, Q1) --------------------------------------------------
    for name, near, far, N in (("lego", 2.0, 6.0, 64), ("fern", 1.2, 12.0, 64), ("dbg", 2.0, 6.0, 16)):
        u = rng.random(N, dtype=np.float32)
        core.push_draw(u)
        out[f"tv_{name}_args"] = np.array([near, far, N], np.float64)
        out[f"tv_{name}_u"] = u
        out[f"tv_{name}_jit"] = RD.generate_t_vals(near, far, 5, N, True).numpy()
        out[f"tv_{name}_nojit"] = RD.generate_t_vals(near, far, 5, N, False).numpy()

    # ---- sample_rays / encode_position / volume_render / sample_pdf ------------------------------------------
    B, Nc, Nf = 48, 64, 128
    o, d = RD.get_rays(20, 20, np.float32(27.7778), torch.from_numpy(lego_pose(RD)))
    sel = rng.choice(400, B, replace=False)
    o, d = o.reshape(-1, 3)[sel], d.reshape(-1, 3)[sel]
    core.push_draw(rng.random(Nc, dtype=np.float32))
    t = RD.generate_t_vals(2.0, 6.0, B, Nc, True)
    pts, dirs = RD.sample_rays(o, d, t)
    out.update(op_o=o.numpy(), op_d=d.numpy(), op_t=t.numpy(), op_pts=pts.numpy(), op_dirs=dirs.numpy(),
               op_enc_x=RD.encode_position(pts, 10).numpy()[:12], op_enc_d=RD.encode_position(dirs, 4).numpy()[:12])
    preds = (rng.standard_normal((B, Nc, 4)) * 2.0).astype(np.float32)
    preds[0] = 0.0                       # sigma = 0 everywhere -> zero weights
    preds[1, :, 3] = -1.0; preds[1, 17, 3] = 40.0   # one opaque sample
    preds[2, :, 3] = -1.0; preds[2, -1, 3] = 1e-3   # any positive sigma on the last sample is opaque (delta = 1e10)
    rgb, depth, w = RD.volume_render(torch.from_numpy(preds), t)
    out.update(vr_preds=preds, vr_rgb=rgb.numpy(), vr_depth=depth.numpy(), vr_w=w.numpy())
    u_pdf = rng.random((B, Nf), dtype=np.float32)
    u_pdf[3, :4] = [0.0, 1.0 - 2.0 ** -24, 0.5, 1e-7]
    t_mid = 0.5 * (t[..., 1:] + t[..., :-1])
    w_in = w.numpy().copy()
    w_in[4] = 0.0                        # empty ray: uniform pdf from the 1e-5 floor
    w_in[5] = 0.0; w_in[5, 30] = 1.0     # spike
    core.push_draw(u_pdf)
    w_arg = core.T(w_in.copy())
    samples = RD.sample_pdf(t_mid, w_arg, Nf)
    assert np.array_equal(w_arg.numpy(), w_in), "sample_pdf mutated its argument (TF tensors are immutable)"
    out.update(sp_w=w_in, sp_u=u_pdf, sp_t_mid=t_mid.numpy(), sp_samples=samples.numpy())

    # ---- models.py: functional model, forward_pass, train_step (2 steps), test_step -----------------------------------
    def trainer(bn, seeds, bias_range, B, Nc, Nf):
        mc = RM.create_nerf_complete_model(8, 256, 4, 10, 4, bn=bn)
        mf = RM.create_nerf_complete_model(8, 256, 4, 10, 4, bn=bn)
        wc, wf = O.init_weights(seeds[0], bias_range), O.init_weights(seeds[1], bias_range)
        set_model_weights(mc, wc)
        set_model_weights(mf, wf)
        tr = RM.NeRFTrainer(mc, mf, B, Nc, Nf, 10, 4)
        tr.compile(keras.optimizers.Adam(learning_rate=5e-4), keras.losses.MeanSquaredError())
        return tr, mc, mf

    def rays(B, Nc, H, W, focal, pose, near, far):
        o, d = RD.get_rays(H, W, focal, torch.from_numpy(pose))
        sel = rng.choice(H * W, B, replace=False)
        u_t = rng.random(Nc, dtype=np.float32)
        core.push_draw(u_t)
        return o.reshape(-1, 3)[sel], d.reshape(-1, 3)[sel], RD.generate_t_vals(near, far, B, Nc, True), u_t

    for tag, (B, Nc, Nf), cam in (("m_lego", (40, 64, 128), (30, 30, np.float32(41.6), lego_pose(RD, 12.0, -45.0), 2.0, 6.0)),
                                  ("m_fern", (24, 16, 32), (12, 16, np.float32(13.0), fern_pose(), 1.2, 12.0))):
        seeds = (311, 312) if tag == "m_lego" else (411, 412)
        tr, mc, mf = trainer(False, seeds, 0.1, B, Nc, Nf)
        o, d, t, u_t = rays(B, Nc, *cam)
        img = rng.random((B, 3), dtype=np.float32)
        u1, u2, u3 = (rng.random((B, Nf), dtype=np.float32) for _ in range(3))
        out.update({f"{tag}_dims": np.array([B, Nc, Nf]), f"{tag}_seeds": np.array(seeds), f"{tag}_o": o.numpy(),
                    f"{tag}_d": d.numpy(), f"{tag}_t": t.numpy(), f"{tag}_img": img, f"{tag}_u1": u1, f"{tag}_u2": u2,
                    f"{tag}_u3": u3})
        pts, dirs = RD.sample_rays(o, d, t)
        ex, ed = RD.encode_position(pts, 10), RD.encode_position(dirs, 4)
        with torch.no_grad():
            out[f"{tag}_mlp_c"] = mc([ex, ed], training=False).numpy()          # the Keras model call (models.py:24-62)
            core.push_draw(u1)
            rgbs, depths, ws, preds = tr.forward_pass(o, d, t, 10, 4, training=False)      # models.py:151-176
            for k, pair in (("rgb", rgbs), ("depth", depths), ("w", ws), ("pred", preds)):
                out[f"{tag}_{k}_c"], out[f"{tag}_{k}_f"] = pair[0].numpy(), pair[1].numpy()
            for s in range(0, B, 16):                                             # models.py:178-225, 16-ray tiles
                core.push_draw(u1[s:s + 16])
            mb = tr.forward_pass_with_minibatch(o, d, t, 10, 4, batch_size=16)
            out[f"{tag}_mb_rgb_f"] = mb[0][1].numpy()
            core.push_draw(u1)
            ts = tr.test_step((torch.from_numpy(img), (o, d, t)))                 # models.py:122-145
            out[f"{tag}_test_metrics"] = np.array([ts["loss_coarse"], ts["loss"], float(ts["psnr"])], np.float64)
        # train_step x2 (models.py:88-120): gradients of the literal graph (no stop_gradient anywhere), Adam on coarse+fine
        tr.compile(tr.optimizer, tr.loss_fn)          # fresh metric trackers
        grabbed = grab_gradients(tr.optimizer)
        logs = []
        for u in (u2, u3):
            core.push_draw(u)
            lg = tr.train_step((torch.from_numpy(img), (o, d, t)))
            logs.append([lg["loss_coarse"], lg["loss"], float(lg["psnr"])])
            if len(logs) == 1:
                out[f"{tag}_weights_after1"] = np.concatenate([model_weights(mc), model_weights(mf)])[::53].copy()
        out[f"{tag}_train_logs"] = np.array(logs, np.float64)       # running means, as Keras reports them
        g0 = np.concatenate([g.numpy().reshape(-1) for g in grabbed[0]])
        out[f"{tag}_grad_step1"] = g0[::7].copy() if tag == "m_fern" else g0[::61].copy()
        out[f"{tag}_grad_step1_norms"] = np.array([float(np.linalg.norm(g.numpy().astype(np.float64))) for g in grabbed[0]])
        wa = np.concatenate([model_weights(mc), model_weights(mf)])
        out[f"{tag}_weights_after2"] = wa[::53].copy()
        out[f"{tag}_weights_after2_sha"] = sha(wa)

    # ---- BATCH_NORM=true variant (models.py:30-33, 49-52): training-mode forward updates the moving statistics -------
    tr, mc, mf = trainer(True, (511, 512), 0.1, 24, 16, 32)
    o, d, t, u_t = rays(24, 16, 12, 16, np.float32(13.0), fern_pose(), 1.2, 12.0)
    ub = rng.random((24, 32), dtype=np.float32)
    out.update(bn_o=o.numpy(), bn_d=d.numpy(), bn_t=t.numpy(), bn_u=ub, bn_seeds=np.array([511, 512]))
    with torch.no_grad():
        core.push_draw(ub)
        rgbs, _, _, preds = tr.forward_pass(o, d, t, 10, 4, training=True)
        out["bn_train_rgb_f"], out["bn_train_pred_c"] = rgbs[1].numpy(), preds[0].numpy()
        bnl = [l for l in mc.layers if hasattr(l, "moving_mean")]
        out["bn_moving_mean_c"] = np.concatenate([l.moving_mean.numpy() for l in bnl])
        out["bn_moving_var_c"] = np.concatenate([l.moving_variance.numpy() for l in bnl])
        core.push_draw(ub)
        rgbs, _, _, _ = tr.forward_pass(o, d, t, 10, 4, training=False)
        out["bn_infer_rgb_f"] = rgbs[1].numpy()

    assert core.pending_draws() == 0
    path = os.path.join(HERE, "ref_source.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB,", len(out), "arrays")


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
