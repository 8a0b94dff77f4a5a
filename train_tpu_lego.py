#!/usr/bin/env python
"""train_tpu_lego.py of the reference (TPUStrategy data parallelism, train_tpu_lego.py:78,127) maps to one process per
GPU here:  torchrun --nproc-per-node N train_tpu_lego.py --config config/lego_batch_h256_tpu.json
Every rank trains BATCH_SIZE / N rays per step; gradients are summed with one NCCL all-reduce (nerf_keras_b200/dist.py)."""
from train_lego import main

if __name__ == "__main__":
    main("lego")
