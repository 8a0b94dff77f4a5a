"""torchrun worker of tests/test_gpu_multi.py: data-parallel training step over NCCL (one process per GPU) checked against
the single-GPU step on the concatenated batch ("N-GPU gradients == 1-GPU gradients")."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import nerf_keras_b200 as nk
    from nerf_keras_b200 import _lib
    from nerf_keras_b200.dist import init_from_env, shard_range
    import oracle as O
    rank, local, world = init_from_env()
    dev = torch.device("cuda", local)
    Bg, Nc, Nf = 1024, 64, 128
    rng = np.random.default_rng(5)
    pose = O.pose_spherical(20.0, -35.0, 4.0)
    o, d = nk.get_rays(200, 200, 277.0, pose)
    sel = torch.from_numpy(rng.choice(40000, Bg, replace=False)).to(dev)
    o, d = o.reshape(-1, 3)[sel].contiguous(), d.reshape(-1, 3)[sel].contiguous()
    t = nk.generate_t_vals(2.0, 6.0, Bg, Nc, True, u=rng.random(Nc, dtype=np.float32))
    img = torch.from_numpy(rng.random((Bg, 3), dtype=np.float32)).to(dev)
    u = torch.from_numpy(rng.random((Bg, Nf), dtype=np.float32)).to(dev)
    wc, wf = O.init_weights(42), O.init_weights(43)

    def trainer(B, **kw):
        mc = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
        mf = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
        mc.set_flat_weights(O.flatten_weights(wc)); mf.set_flat_weights(O.flatten_weights(wf))
        tr = nk.NeRFTrainer(mc, mf, B, Nc, Nf, 10, 4, **kw)
        tr.compile(nk.Adam(5e-4), nk.MeanSquaredError())
        return tr

    lo, hi = shard_range(Bg, rank, world)
    res = {}
    L = _lib.lib()
    st = lambda: torch.cuda.current_stream().cuda_stream
    m = torch.empty(3, device=dev)
    # (1) gradient buffers: local shards, all-reduced and averaged, vs the full batch on one GPU
    full = trainer(Bg, use_cuda_graph=False)
    _lib.check(L.nerf_train_phases(full._ctx.handle, img.data_ptr(), o.data_ptr(), d.data_ptr(), t.data_ptr(), u.data_ptr(), Bg,
                                   m.data_ptr(), 3, st()), "full")
    g_full = full._ctx.grad_tensor().clone()
    part = trainer(hi - lo, use_cuda_graph=False)
    sl = [x[lo:hi].contiguous() for x in (img, o, d, t, u)]
    _lib.check(L.nerf_train_phases(part._ctx.handle, *[x.data_ptr() for x in sl], hi - lo, m.data_ptr(), 3, st()), "part")
    g_dp = part._ctx.grad_tensor().clone()
    torch.distributed.all_reduce(g_dp)
    g_dp /= world
    res["grad_rel_err"] = float((g_dp - g_full).norm() / g_full.norm())
    res["grad_max_abs_err"] = float((g_dp - g_full).abs().max())
    res["grad_norm"] = float(g_full.norm())
    # (2) whole steps through NeRFTrainer.train_step (overlapped all-reduce, CUDA graph) vs the single-GPU trainer
    keep = []
    for tag, kw in (("overlap_graph", dict(use_cuda_graph=True, overlap_allreduce=True)),
                    ("single_allreduce_eager", dict(use_cuda_graph=False, overlap_allreduce=False))):
        a, b = trainer(hi - lo, stop_grad_samples=True, **kw), trainer(Bg, stop_grad_samples=True, use_cuda_graph=False)
        # every rank must hold identical weights after identical updates
        for _ in range(4):
            a.train_step((sl[0], (sl[1], sl[2], sl[3])), u_pdf=sl[4])
            b.train_step((img, (o, d, t)), u_pdf=u)
        wa = torch.from_numpy(np.concatenate([a.coarse_model.get_flat_weights(), a.fine_model.get_flat_weights()])).to(dev)
        wb = torch.from_numpy(np.concatenate([b.coarse_model.get_flat_weights(), b.fine_model.get_flat_weights()])).to(dev)
        w0 = torch.from_numpy(np.concatenate([O.flatten_weights(wc), O.flatten_weights(wf)])).to(dev)
        ref = wa.clone()
        torch.distributed.broadcast(ref, 0)
        res[tag] = {"ranks_identical": bool(torch.equal(ref, wa)),
                    "update_cosine_vs_1gpu": float(torch.dot(wa - w0, wb - w0) / ((wa - w0).norm() * (wb - w0).norm())),
                    "median_abs_diff_vs_1gpu": float((wa - wb).abs().median()),
                    "graphs": len(a._graphs), "steps": a._ctx.optimizer_state()[2]}
        keep += [a, b]
    # (3) BATCH_NORM=true under data parallelism: per-replica batch statistics (as Keras), averaged gradients and moving
    # statistics -> every rank holds the same weights and BN parameters after every step, and the loss goes down
    import warnings
    warnings.simplefilter("ignore")
    mcb = nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=True)
    mfb = nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=True)
    mcb.set_flat_weights(O.flatten_weights(wc)); mfb.set_flat_weights(O.flatten_weights(wf))
    nb = 64
    trb = nk.NeRFTrainer(mcb, mfb, nb, Nc, Nf, 10, 4, stop_grad_samples=True)
    trb.compile(nk.Adam(5e-4), nk.MeanSquaredError())
    sb = [x[rank * nb:(rank + 1) * nb].contiguous() for x in (img, o, d, t, u)]
    losses = []
    for _ in range(6):
        trb.reset_metrics()
        losses.append(float(trb.train_step((sb[0], (sb[1], sb[2], sb[3])), u_pdf=sb[4])["loss"]))
    wbn = torch.from_numpy(np.concatenate([mcb.get_flat_weights(), mfb.get_flat_weights()] +
                                          [mcb.get_bn_params()[r][k] for r in mcb.bn for k in ("gamma", "beta", "mean", "var")])).to(dev)
    ref = wbn.clone()
    torch.distributed.broadcast(ref, 0)
    res["batch_norm_dp"] = {"ranks_identical": bool(torch.equal(ref, wbn)), "first_loss": losses[0], "last_loss": losses[-1]}
    if rank == 0:
        print("DP_RESULT " + json.dumps(res), flush=True)
    from nerf_keras_b200.dist import shutdown
    shutdown(*keep)


if __name__ == "__main__":
    main()
