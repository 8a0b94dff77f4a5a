#!/usr/bin/env python
"""inference.py --config config/*.json [--weights models/x.npz] : novel-view sweep like the reference's
inference.py:229-265 -- `pose_spherical(theta, -30, 4)` poses, `get_rays`, deterministic t-values,
`forward_pass_with_minibatch`, uint8 frames.  Frames are sharded across GPUs under torchrun (one process per
GPU, no collective but the final gather); mp4 writing is out of scope, frames are saved as .npy."""
import argparse
import os
import time

import numpy as np
import torch

import nerf_keras_b200 as nk
from nerf_keras_b200.config import load_config, model_kwargs
from nerf_keras_b200.dist import gather_rows, init_from_env, shard_range


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=str, default="config/lego_batch_debug.json")
    ap.add_argument("--weights", type=str, default=None)
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--tile", type=int, default=16384,
                    help="rays per forward_pass call; 16384 keeps launch gaps and tail quantisation under 2 % (4096: 2.7 M rays/s, 16384: 3.3 M)")
    ap.add_argument("--out", type=str, default="frames.npy")
    args = ap.parse_args()
    conf = load_config(args.config)
    rank, local, world = init_from_env()
    H, W, Nc, Nf = conf["HEIGHT"], conf["WIDTH"], conf["NS_COARSE"], conf["NS_FINE"]
    nk.set_random_seed(42)
    coarse = nk.create_nerf_complete_model(**model_kwargs(conf))
    fine = nk.create_nerf_complete_model(**model_kwargs(conf))
    trainer = nk.NeRFTrainer(coarse, fine, args.tile, Nc, Nf, conf["L_XYZ"], conf["L_DIR"])
    trainer.build()
    if args.weights:
        trainer.load_weights(args.weights)
    focal = float(np.float32(0.5 * W / np.tan(0.5 * 0.6911112)))
    near, far = 2.0, 6.0
    thetas = np.linspace(-45.0, 45.0, args.frames, endpoint=False)
    lo, hi = shard_range(args.frames, rank, world)
    frames = []
    # warm-up (module load, weight packing) outside the timed sweep
    o_w, d_w = nk.get_rays(H, W, focal, nk.pose_spherical(0.0, -30.0, 4.0))
    trainer.forward_pass_with_minibatch(o_w.reshape(-1, 3), d_w.reshape(-1, 3),
                                        nk.generate_t_vals(near, far, H * W, Nc, rand_sampling=False),
                                        batch_size=args.tile, maps_only=True)
    if world > 1:
        # the gather's point-to-point channels are set up on first use: do that outside the timed sweep, like the kernels' warm-up
        gather_rows(torch.zeros((1, 8, 8, 3), dtype=torch.uint8, device="cuda"))
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.time()
    for theta in thetas[lo:hi]:
        c2w = nk.pose_spherical(float(theta), -30.0, 4.0)
        o, d = nk.get_rays(H, W, focal, c2w)
        o, d = o.reshape(-1, 3), d.reshape(-1, 3)
        t = nk.generate_t_vals(near, far, o.shape[0], Nc, rand_sampling=False)
        rgbs, _, _, _ = trainer.forward_pass_with_minibatch(o, d, t, conf["L_XYZ"], conf["L_DIR"], batch_size=args.tile,
                                                            maps_only=True)
        frames.append(torch.clamp(255.0 * rgbs[1], 0.0, 255.0).to(torch.uint8).reshape(H, W, 3))
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    dt_render = time.time() - t0    # wall time of the slowest rank's render loop
    local_frames = torch.stack(frames) if frames else torch.empty((0, H, W, 3), dtype=torch.uint8, device="cuda")
    if world > 1:
        per = -(-args.frames // world)
        pad = torch.zeros((per, H, W, 3), dtype=torch.uint8, device="cuda")
        pad[: local_frames.shape[0]] = local_frames
        allf = gather_rows(pad)
        if rank == 0:
            chunks = [allf[r * per: r * per + (shard_range(args.frames, r, world)[1] - shard_range(args.frames, r, world)[0])]
                      for r in range(world)]
            local_frames = torch.cat(chunks)
        torch.cuda.synchronize()
        torch.distributed.barrier()
    dt = time.time() - t0           # ... including the only collective of the sweep, the gather of the frames to rank 0
    if rank == 0:
        np.save(args.out, local_frames.cpu().numpy())
        n_rays = args.frames * H * W
        print(f"rendered {local_frames.shape[0]} frames {H}x{W} on {world} GPU(s) in {dt:.3f} s including the gather "
              f"({dt_render:.3f} s render only): {n_rays / dt:.3e} rays/s aggregate -> {args.out}")


if __name__ == "__main__":
    main()
