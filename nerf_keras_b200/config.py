"""JSON config schema of the reference (config/*.json read at train_lego.py:30-50), kept verbatim:
flat dict of UPPER_CASE keys, no defaults for the required keys (KeyError when one is missing, as
`conf["KEY"]` does in the reference).  The reference's own config files load unchanged."""
from __future__ import annotations

import json
from typing import Optional

REQUIRED_KEYS = ("BATCH_SIZE", "NS_COARSE", "NS_FINE", "HEIGHT", "WIDTH", "L_XYZ", "L_DIR", "NUM_LAYERS", "HIDDEN_DIM",
                 "SKIP_LAYER", "EPOCHS", "LEARNING_RATE", "BATCH_NORM", "WITH_GCS")
OPTIONAL_KEYS = ("TEST_BATCH_SIZE",)  # read only by the Fern scripts (train_fern.py:38)


def load_config(path: str, override: Optional[dict] = None) -> dict:
    with open(path) as f:
        conf = json.load(f)
    if override:
        override = dict(override)
        drop = override.pop("__drop__", None)
        conf.update(override)
        if drop:
            conf.pop(drop, None)
    for k in REQUIRED_KEYS:
        if k not in conf:
            raise KeyError(k)
    return conf


def model_kwargs(conf: dict) -> dict:
    return dict(num_layers=conf["NUM_LAYERS"], hidden_dim=conf["HIDDEN_DIM"], skip_layer=conf["SKIP_LAYER"],
                lxyz=conf["L_XYZ"], ldir=conf["L_DIR"], bn=conf["BATCH_NORM"])
