#!/usr/bin/env python
"""Weight bridge between the reference's Keras-3 `.weights.h5` checkpoints and this repo's role-keyed `.npz`
(SURVEY.md section 8(f) rank 1).

RUN THIS INSIDE A CHECKOUT OF THE REFERENCE (ghif/nerf-keras) with its own environment (tensorflow 2.16 / keras 3.10 /
h5py): it imports the reference's `models.py` to build `NeRFTrainer` exactly as `inference.py:120-152` does, lets
Keras read or write the `.weights.h5` file (`train_lego.py:205-213`, `inference.py:170`), and converts to / from

    {coarse|fine}/{d0..d7|sigma|feature|ddir|rgb}/{W|b}        W: Keras layout (in, out), b: (out,)

which is what `nerf_keras_b200.NeRFTrainer.save_weights / load_weights` use.  This environment has neither h5py nor
Keras, so the script is NOT exercised by this repo's tests; the role mapping it relies on is:

  * Keras names Dense layers `dense`, `dense_1`, ... in CREATION order, and `create_nerf_complete_model`
    (models.py:24-62) creates them as d0..d7, sigma, feature, ddir, rgb.  `model.layers` / `model.get_weights()` are
    ordered by graph depth instead (feature and ddir come before sigma there), so the mapping sorts by the name suffix
    and then checks every kernel shape against the expected (in, out).

    python keras_weights_bridge.py export --config config/lego_batch_h256.json --h5 model.weights.h5 --npz weights.npz
    python keras_weights_bridge.py import --config config/lego_batch_h256.json --npz weights.npz --h5 model.weights.h5
"""
import argparse
import json
import re

import numpy as np

ROLES = ["d%d" % i for i in range(8)] + ["sigma", "feature", "ddir", "rgb"]


def expected_shapes(conf):
    h, e_xyz, e_dir = conf["HIDDEN_DIM"], 6 * conf["L_XYZ"] + 3, 6 * conf["L_DIR"] + 3
    n, skip = conf["NUM_LAYERS"], conf["SKIP_LAYER"]
    roles = ["d%d" % i for i in range(n)] + ["sigma", "feature", "ddir", "rgb"]
    shapes, fan_in = [], e_xyz
    for i in range(n):
        shapes.append((fan_in, h))
        fan_in = h + e_xyz if (i % skip == 0 and i > 0) else h
    shapes += [(fan_in, 1), (fan_in, h), (h + e_dir, h // 2), (h // 2, 3)]
    return roles, shapes


def layers_in_creation_order(model, cls="Dense"):
    found = [l for l in model.layers if l.__class__.__name__ == cls]
    suffix = lambda l: int(m.group(1)) if (m := re.search(r"_(\d+)$", l.name)) else 0
    return sorted(found, key=suffix)


def dense_layers_in_creation_order(model):
    return layers_in_creation_order(model, "Dense")


def bn_roles(conf):
    """BATCH_NORM=true: one BatchNormalization after every trunk Dense and after the direction Dense (models.py:30-33,49-52)."""
    return ["d%d" % i for i in range(conf["NUM_LAYERS"])] + ["ddir"]


BN_KEYS = ("gamma", "beta", "mean", "var")          # Keras weight order: gamma, beta, moving_mean, moving_variance


def build_trainer(conf):
    from models import create_nerf_complete_model, NeRFTrainer          # the reference's own module
    mk = lambda: create_nerf_complete_model(conf["NUM_LAYERS"], conf["HIDDEN_DIM"], conf["SKIP_LAYER"], conf["L_XYZ"],
                                            conf["L_DIR"], bn=conf.get("BATCH_NORM", False))
    trainer = NeRFTrainer(coarse_model=mk(), fine_model=mk(), batch_size=conf["BATCH_SIZE"], ns_coarse=conf["NS_COARSE"],
                          ns_fine=conf["NS_FINE"], l_xyz=conf["L_XYZ"], l_dir=conf["L_DIR"])
    trainer.build(input_shape=((3,), ((3,), (3,), (conf["NS_COARSE"],))))  # as inference.py:145-152
    return trainer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("direction", choices=["export", "import"])
    ap.add_argument("--config", required=True)
    ap.add_argument("--h5", required=True)
    ap.add_argument("--npz", required=True)
    args = ap.parse_args()
    conf = json.load(open(args.config))
    use_bn = bool(conf.get("BATCH_NORM", False))     # the B200 path renders BN checkpoints on the fused kernels (folded) and trains them on its layer-by-layer path
    roles, shapes = expected_shapes(conf)
    trainer = build_trainer(conf)
    nets = (("coarse", trainer.coarse_model), ("fine", trainer.fine_model))
    if args.direction == "export":
        trainer.load_weights(args.h5)
        out = {}
        for name, model in nets:
            layers = dense_layers_in_creation_order(model)
            assert len(layers) == len(roles), (name, len(layers))
            for role, shape, layer in zip(roles, shapes, layers):
                W, b = layer.get_weights()
                assert W.shape == shape, (name, role, W.shape, shape)
                out[f"{name}/{role}/W"], out[f"{name}/{role}/b"] = W.astype(np.float32), b.astype(np.float32)
            if use_bn:
                bns = layers_in_creation_order(model, "BatchNormalization")
                assert len(bns) == len(bn_roles(conf)), (name, len(bns))
                for role, layer in zip(bn_roles(conf), bns):
                    for key, arr in zip(BN_KEYS, layer.get_weights()):
                        out[f"{name}/{role}/bn_{key}"] = arr.astype(np.float32)
        np.savez(args.npz, **out)
        print("wrote", args.npz, "with", len(out), "arrays")
    else:
        data = np.load(args.npz)
        for name, model in nets:
            for role, shape, layer in zip(roles, shapes, dense_layers_in_creation_order(model)):
                W, b = data[f"{name}/{role}/W"], data[f"{name}/{role}/b"]
                assert W.shape == shape, (name, role, W.shape, shape)
                layer.set_weights([W, b])
            if use_bn:
                for role, layer in zip(bn_roles(conf), layers_in_creation_order(model, "BatchNormalization")):
                    layer.set_weights([data[f"{name}/{role}/bn_{key}"] for key in BN_KEYS])
        trainer.save_weights(args.h5)
        print("wrote", args.h5)


if __name__ == "__main__":
    main()
