"""Synthetic Lego-/Fern-shaped data with the return structure of the reference loaders
(`prepare_lego_data`, lego_data_utils.py:8-51; `prepare_fern_data`, fern_data_utils.py:462-520) and an
on-device replacement for `create_batched_dataset_pipeline` (data_utils.py:140-170).

The real datasets need a network download / ImageMagick / GCS (out of scope);
shapes, bounds and camera models follow SURVEY.md section 8(d) and Appendix B."""
from __future__ import annotations

import math
from typing import Iterator, Tuple

import numpy as np
import torch

from . import data_utils as du


def _procedural_rgb(o: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    """A cheap smooth function of the ray (so that training has something learnable)."""
    dn = d / d.norm(dim=-1, keepdim=True)
    p = o + 4.0 * dn
    return torch.stack([0.5 + 0.5 * torch.sin(3.0 * p[..., 0]), 0.5 + 0.5 * torch.cos(2.0 * p[..., 1] + p[..., 2]),
                        0.5 + 0.5 * torch.sin(p[..., 0] * p[..., 2])], dim=-1).clamp(0, 1).contiguous()


def _views(H, W, focal, poses, ndc_near=None):
    """Flattened per-pixel (images, origins, directions) of the views.  ndc_near: rays are moved to the near plane and
    projected to normalised device coordinates (extension) AFTER the target colours were taken from the
    world-space rays."""
    imgs, oris, dirs = [], [], []
    for pose in poses:
        o, d = du.get_rays(H, W, focal, pose)
        o, d = o.reshape(-1, 3), d.reshape(-1, 3)
        imgs.append(_procedural_rgb(o, d))
        if ndc_near is not None:
            o, d = du.ndc_rays(H, W, focal, ndc_near, o, d)
        oris.append(o); dirs.append(d)
    return torch.cat(imgs), torch.cat(oris), torch.cat(dirs)


def prepare_lego_data(H: int, W: int, n_views: int = 20, seed: int = 0):
    """((imgs_s, oris_s, dirs_s) train, (...) val, (near, far), focal): flattened per-pixel arrays, 80/20 split,
    near/far = 2/6 (lego_data_utils.py:26,39-49)."""
    rng = np.random.default_rng(seed)
    focal = float(np.float32(0.5 * W / math.tan(0.5 * 0.6911112)))
    poses = [du.pose_spherical(float(t), float(p), 4.0)
             for t, p in zip(rng.uniform(-180, 180, n_views), rng.uniform(-90, 0, n_views))]
    k = int(n_views * 0.8)
    return _views(H, W, focal, poses[:k]), _views(H, W, focal, poses[k:]), (2.0, 6.0), focal


def prepare_fern_data(H: int, W: int, n_views: int = 20, seed: int = 2, ndc: bool = False):
    """Forward-facing poses (identity rotation + small yaw/pitch, xy translation), pinhole rays with near/far from
    the bounds as the reference does (fern_data_utils.py:489-496); one view held out (:499-500).
    ndc=True (extension, not in the reference): forward-facing NDC rays as in the original NeRF, sampled on t in [0, 1]."""
    rng = np.random.default_rng(seed)
    focal = float(np.float32(407.6 * W / 504.0))
    poses = []
    for _ in range(n_views):
        yaw, pitch = np.deg2rad(rng.uniform(-10, 10, 2))
        cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
        R = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
        pose = np.eye(4, dtype=np.float32)
        pose[:3, :3] = R.astype(np.float32)
        pose[:2, 3] = rng.uniform(-0.3, 0.3, 2).astype(np.float32)
        poses.append(pose)
    if ndc:
        return _views(H, W, focal, poses[1:], 1.0), _views(H, W, focal, poses[:1], 1.0), (0.0, 1.0), focal
    return _views(H, W, focal, poses[1:]), _views(H, W, focal, poses[:1]), (1.2, 12.0), focal


class BatchedRayDataset:
    """On-device stand-in for `create_batched_dataset_pipeline` (data_utils.py:140-170): one shared jitter vector
    drawn when the dataset is built (:156, quirk Q1), uniformly random ray batches (drop_remainder), yielding
    (images, (ray_origins, ray_directions, t_vals)).  rank/world shard each global batch for data parallelism."""

    def __init__(self, images_s, ray_oris_s, ray_dirs_s, num_samples, batch_size, near=2.0, far=6.0, shuffle=True,
                 rand_sampling=True, steps_per_epoch=None, rank=0, world=1, seed=0, order="random"):
        # order="random": uniformly random rays (with replacement) per batch.  order="windowed": the reference's epoch
        # semantics -- every ray exactly once per epoch, drop_remainder, shuffled like tf.data's shuffle buffer of
        # 5 * batch_size elements (data_utils.py:162-164), i.e. a LOCAL shuffle: sort by (position + window * u).
        self.order = order
        self.img, self.o, self.d = images_s, ray_oris_s, ray_dirs_s
        self.n = self.o.shape[0]
        self.batch = int(batch_size)
        self.local = self.batch // world
        self.rank, self.world, self.shuffle = rank, world, shuffle
        self.steps = steps_per_epoch or max(1, self.n // self.batch)
        u = torch.rand(num_samples, device=self.o.device) if rand_sampling else None
        self.t_row = du.generate_t_vals(near, far, 1, num_samples, rand_sampling, u=u)
        self.gen = torch.Generator(device=self.o.device)
        self.gen.manual_seed(seed)

    def __len__(self):
        return self.steps

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]]:
        perm = None
        if self.order == "windowed" and self.shuffle:
            window = float(5 * self.batch if self.batch else 1024)
            keys = torch.arange(self.n, device=self.o.device, dtype=torch.float32) + \
                window * torch.rand(self.n, device=self.o.device, generator=self.gen)
            perm = torch.argsort(keys)
        # two alternating sets of batch buffers with fixed addresses: the trainer replays its step from a CUDA graph
        # captured on the buffers it was first handed, and a batch stays valid until the next-but-one is produced
        n_loc = self.local if self.world > 1 else self.batch
        if getattr(self, "_slots", None) is None or self._slots[0][0].shape[0] != n_loc:
            mk = lambda src: torch.empty((n_loc,) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
            self._slots = [(mk(self.img), mk(self.o), mk(self.d), self.t_row.expand(n_loc, -1).contiguous())
                           for _ in range(2)]
        for s in range(self.steps):
            if perm is not None:
                idx = perm[s * self.batch:(s + 1) * self.batch]
            elif self.shuffle:
                idx = torch.randint(0, self.n, (self.batch,), device=self.o.device, generator=self.gen)
            else:
                idx = (torch.arange(self.batch, device=self.o.device) + s * self.batch) % self.n
            idx = idx[self.rank * self.local:(self.rank + 1) * self.local]
            img, o, d, t = self._slots[s & 1]
            if idx.shape[0] != img.shape[0]:                 # ragged shard: plain gathers
                yield self.img[idx], (self.o[idx], self.d[idx], self.t_row.expand(idx.shape[0], -1).contiguous())
                continue
            torch.index_select(self.img, 0, idx, out=img)
            torch.index_select(self.o, 0, idx, out=o)
            torch.index_select(self.d, 0, idx, out=d)
            yield img, (o, d, t)


class HostPrefetcher:
    """Feeds batches that live in (pinned) host memory to the device with the copy of batch i+1 overlapped with the work
    on batch i: a second CUDA stream and two device staging slots.  `host_batches` is any iterable of tuples of CPU
    tensors (pin them for truly asynchronous copies); the iterator yields tuples of device tensors that stay valid
    until the next-but-one `next()`.  This is what replaces tf.data's prefetch (data_utils.py:166-167) when the ray set
    does not fit in HBM."""

    def __init__(self, host_batches, device=None):
        self.src = host_batches
        self.device = torch.device(device) if device is not None else du._dev()
        self.stream = torch.cuda.Stream(device=self.device)
        self._slots = [None, None]      # the two device staging slots survive re-iteration: their addresses are what the
                                        # trainer's captured step graphs are keyed on

    def __iter__(self):
        it = iter(self.src)
        slots, ready, free = self._slots, [None, None], [None, None]
        compute = torch.cuda.current_stream(self.device)

        def enqueue(k, batch):
            if slots[k] is None or any(a.shape != b.shape or a.dtype != b.dtype for a, b in zip(slots[k], batch)):
                slots[k] = tuple(torch.empty(b.shape, dtype=b.dtype, device=self.device) for b in batch)
            if free[k] is not None:
                self.stream.wait_event(free[k])          # the consumer is done with this slot
            with torch.cuda.stream(self.stream):
                for dst, b in zip(slots[k], batch):
                    dst.copy_(b, non_blocking=True)
                ready[k] = torch.cuda.Event()
                ready[k].record(self.stream)

        nxt = next(it, None)
        if nxt is None:
            return
        enqueue(0, nxt)
        k = 0
        while True:
            cur = k
            nxt = next(it, None)
            if nxt is not None:
                # slot 1-k was handed out one iteration ago; everything enqueued on the compute stream so far used it
                free[1 - k] = torch.cuda.Event()
                free[1 - k].record(compute)
                enqueue(1 - k, nxt)
            compute.wait_event(ready[cur])
            yield slots[cur]
            if nxt is None:
                return
            k = 1 - k
