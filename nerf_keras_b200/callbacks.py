"""Per-epoch validation callback of the reference's training scripts (SURVEY.md section 8(f) rank 4;
`TrainCallback`, train_lego.py:165-264 / train_fern.py:169-268) without Keras, matplotlib or GCS:

  * accumulates `history = {"losses_coarse", "losses", "psnrs"}` from the epoch logs and writes it as JSON;
  * renders the first validation views (`2*H*W` rays, jittered t-values, fine rgb + depth) through
    `forward_pass_with_minibatch`;
  * saves the weights (role-keyed `.npz` instead of `.weights.h5`, see tools/keras_weights_bridge.py);
  * writes `images/<checkpoint_dir>/<epoch:03d>.png`: predicted image | depth map side by side, each min-max scaled to
    0..255 the way `keras.utils.array_to_img` does (the reference's third panel, the matplotlib loss plot, is the
    `losses` series of the JSON instead).
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import data_utils as du


def array_to_uint8(x: np.ndarray) -> np.ndarray:
    """Min-max scaling of `keras.utils.array_to_img(..., scale=True)`: (x - min) / (max - min) * 255."""
    x = np.asarray(x, dtype=np.float32)
    x = x - x.min()
    m = x.max()
    if m != 0:
        x = x / m
    return (x * 255.0).astype(np.uint8)


class TrainCallback:
    def __init__(self, trainer, val_ray_oris, val_ray_dirs, height, width, near, far, ns_coarse, checkpoint_dir,
                 weight_name="nerf.npz", history_name="history.json", image_dir=None, render_batch=4096):
        self.model = trainer
        self.o, self.d = val_ray_oris, val_ray_dirs
        self.H, self.W, self.near, self.far, self.ns_coarse = height, width, near, far, ns_coarse
        self.checkpoint_dir = checkpoint_dir
        self.weight_name, self.history_name = weight_name, history_name
        self.image_dir = image_dir if image_dir is not None else os.path.join("images", checkpoint_dir)
        self.render_batch = render_batch
        self.history = {"losses_coarse": [], "losses": [], "psnrs": []}

    def render_validation(self):
        """-> rgb (nb,H,W,3), depth (nb,H,W) float32 numpy: the first `2*H*W` validation rays (train_lego.py:183-198)."""
        n = min(2 * self.H * self.W, self.o.shape[0]) // (self.H * self.W) * (self.H * self.W)
        o, d = self.o[:n].contiguous(), self.d[:n].contiguous()
        t = du.generate_t_vals(self.near, self.far, n, self.ns_coarse, rand_sampling=True)
        rgbs, depths, _, _ = self.model.forward_pass_with_minibatch(o, d, t, batch_size=self.render_batch,
                                                                    training=False, maps_only=True)
        nb = n // (self.H * self.W)
        return (rgbs[1].reshape(nb, self.H, self.W, 3).cpu().numpy(), depths[1].reshape(nb, self.H, self.W).cpu().numpy())

    def on_epoch_end(self, epoch, logs=None):
        from PIL import Image
        logs = logs or {}
        for key, src in (("losses_coarse", "loss_coarse"), ("losses", "loss"), ("psnrs", "psnr")):
            v = logs.get(src)
            self.history[key].append(None if v is None else float(v))
        rgb, depth = self.render_validation()
        os.makedirs(self.checkpoint_dir, exist_ok=True)
        self.model.save_weights(os.path.join(self.checkpoint_dir, self.weight_name))
        os.makedirs(self.image_dir, exist_ok=True)
        panel = np.concatenate([array_to_uint8(rgb[0]), np.repeat(array_to_uint8(depth[0])[..., None], 3, -1)], axis=1)
        Image.fromarray(panel).save(os.path.join(self.image_dir, f"{epoch:03d}.png"))
        with open(os.path.join(self.checkpoint_dir, self.history_name), "w") as f:
            json.dump(self.history, f)
