"""BASELINE config 5: ray-batch scaling sweep (1Ki..1Mi rays/step x sample counts) for render and train throughput
on one GPU.  Prints one JSON object per point; run under torchrun for data parallel (each rank its own batch)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerf_keras_b200 as nk
from nerf_keras_b200.dist import init_from_env

rank, local, world = init_from_env()
dev = torch.device("cuda", local)
points = [(1 << 10, 64, 128), (1 << 12, 64, 128), (1 << 14, 64, 128), (1 << 16, 64, 128), (1 << 18, 64, 128),
          (1 << 12, 64, 64), (1 << 16, 64, 64)]
if "--big" in sys.argv:
    points.append((1 << 20, 64, 128))
for B, Nc, Nf in points:
    nk.set_random_seed(42)
    c = nk.create_nerf_complete_model(8, 256, 4, 10, 4); f = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    train = B <= (1 << 16)     # saved activations: 5 KB/sample -> 64Ki rays x 256 samples = 84 GB is the cap on one GPU
    tr = nk.NeRFTrainer(c, f, B, Nc, Nf, 10, 4)
    if train:
        tr.compile(nk.Adam(learning_rate=5e-4), nk.MeanSquaredError())
    else:
        tr.build()
    side = int(np.ceil(np.sqrt(B)))
    o, d = nk.get_rays(side, side, 1.6 * side, nk.pose_spherical(20.0 + rank, -30.0, 4.0))
    o, d = o.reshape(-1, 3)[:B].contiguous(), d.reshape(-1, 3)[:B].contiguous()
    t = nk.generate_t_vals(2.0, 6.0, B, Nc, True, u=np.random.default_rng(3).random(Nc, dtype=np.float32))
    u = torch.rand(B, Nf, device=dev); img = torch.rand(B, 3, device=dev)
    res = {"rays_per_step_per_gpu": B, "ns_coarse": Nc, "ns_fine": Nf, "n_gpus": world}
    for mode in (("render", "train") if train else ("render",)):
        fn = (lambda: tr.train_step((img, (o, d, t)), u_pdf=u)) if mode == "train" else (lambda: tr.forward_pass(o, d, t, u_pdf=u))
        reps = 5 if B >= (1 << 16) else 20
        for _ in range(3): fn()
        if world > 1: torch.distributed.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
        if world > 1: torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        res[mode + "_ms"] = float(ms)
        res[mode + "_rays_per_s"] = world * B / float(ms) * 1e3
    if rank == 0:
        print(json.dumps(res), flush=True)
    del tr, c, f
    torch.cuda.empty_cache()
if world > 1:
    torch.distributed.destroy_process_group()
