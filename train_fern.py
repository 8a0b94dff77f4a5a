#!/usr/bin/env python
"""train_fern.py --config config/fern_*.json (reference CLI: train_fern.py:25-27) on synthetic Fern-shaped views."""
from train_lego import main

if __name__ == "__main__":
    main("fern")
