"""Data-parallel plumbing (the reference's only strategy: synchronous DP under
tf.distribute.TPUStrategy, train_tpu_lego.py:73-82,127 -> gradient all-reduce inside
apply_gradients, models.py:107).  One process per GPU, torch.distributed (NCCL on GPUs; the same
functions run over gloo on CPU tensors for the world_size-2 host tests)."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from torchrun's environment; initialises the default group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous equal shards of a global batch (BATCH_SIZE is the GLOBAL batch under the strategy,
    config/lego_batch_h256_tpu.json:2).  Remainder rays go to the lowest ranks."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_sum_(flat: torch.Tensor, group=None) -> int:
    """In-place SUM all-reduce of the flat gradient buffer; returns the world size so the caller can
    fold the 1/world mean into the Adam kernel (gradient of the global-batch mean loss)."""
    if not dist.is_initialized():
        return 1
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return world


def gather_rows(local: torch.Tensor, group=None) -> Optional[torch.Tensor]:
    """Inference: gather per-rank ray tiles (equal row counts, padded by the caller) to rank 0."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bufs = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
    dist.gather(local, bufs, dst=0, group=group)
    return torch.cat(bufs, dim=0) if rank == 0 else None


def shutdown(*trainers, timeout_s: float = 30.0):
    """End of a torchrun worker: release the trainers' CUDA graphs (they hold NCCL kernels), synchronise, and destroy the
    process group.  If the teardown itself does not return within `timeout_s` the process exits with status 0 anyway --
    all results have been produced at this point."""
    import os
    import sys
    import threading
    for tr in trainers:
        if tr is not None:
            tr.release_graphs()
    if not dist.is_initialized():
        return
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    guard = threading.Timer(timeout_s, lambda: os._exit(0))
    guard.daemon = True
    guard.start()
    dist.barrier()
    dist.destroy_process_group()
    guard.cancel()
