#!/usr/bin/env python
"""train_tpu_fern.py of the reference -> torchrun --nproc-per-node N train_tpu_fern.py --config config/fern_batch_h256_tpu.json
(see train_tpu_lego.py)."""
from train_lego import main

if __name__ == "__main__":
    main("fern")
