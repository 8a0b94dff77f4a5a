"""`keras.ops` stand-in: the functions data_utils.py / models.py call, one eager torch-CPU fp32 op each."""
import math

import numpy as np
import torch

from refshim_core import T, Tensor, dt


def sin(x): return torch.sin(T(x))
def cos(x): return torch.cos(T(x))
def exp(x): return torch.exp(T(x))
def sigmoid(x): return torch.sigmoid(T(x))
def relu(x): return torch.relu(T(x))
def ones_like(x): return torch.ones_like(T(x))
def shape(x): return tuple(T(x).shape)


def concatenate(xs, axis=-1):
    return torch.cat([T(x) for x in xs], dim=axis)


def stack(xs, axis=0):
    return torch.stack([T(x) for x in xs], dim=axis)


def arange(start, stop=None, step=1, dtype=None):
    if stop is None:
        start, stop = 0, start
    return T(torch.arange(start, stop, step, dtype=dt(dtype)))


def meshgrid(*xs, indexing="xy"):
    return tuple(T(g) for g in torch.meshgrid(*[T(x) for x in xs], indexing=indexing))


def sum(x, axis=None, keepdims=False):
    x = T(x)
    if axis is None:
        return torch.sum(x)
    if x.shape[axis] <= 4:
        # a short inner reduce is a plain left-to-right scalar loop in Eigen; torch's vectorised reduce is not
        parts = torch.unbind(x, dim=axis)
        acc = parts[0]
        for p in parts[1:]:
            acc = acc + p
        return acc.unsqueeze(axis) if keepdims else acc
    return torch.sum(x, dim=axis, keepdim=keepdims)


def broadcast_to(x, shape):
    return torch.broadcast_to(T(x), tuple(int(s) for s in shape)).contiguous().as_subclass(Tensor)


def cumprod(x, axis=None):
    return torch.cumprod(T(x), dim=axis)


def roll(x, shift, axis=None):
    return torch.roll(T(x), shifts=shift, dims=axis)


def ones(shape, dtype="float32"):
    return T(torch.ones(tuple(int(s) for s in shape), dtype=dt(dtype)))


def linspace(start, stop, num, dtype=None):
    """tf.linspace as restated in oracle.data_utils_ref.tf_linspace_f32 (TF kernel numerics, unverifiable offline)."""
    s = torch.tensor(float(start), dtype=torch.float32)
    e = torch.tensor(float(stop), dtype=torch.float32)
    if num == 1:
        return T(s.view(1))
    delta = (e - s) / torch.tensor(float(num - 1), dtype=torch.float32)
    i = torch.arange(1, num - 1, dtype=torch.float32)
    return T(torch.cat([s.view(1), s + delta * i, e.view(1)]))


def sort(x, axis=-1):
    return torch.sort(T(x), dim=axis).values


def psnr(x1, x2, max_val):
    """keras.ops.psnr: 20 log10(max_val) - 10 log10(mean((x1 - x2)^2))."""
    mse = torch.mean((T(x1) - T(x2)) ** 2)
    return 20.0 * math.log10(max_val) - 10.0 * torch.log10(mse)


def convert_to_tensor(x, dtype=None):
    """Nested python lists may hold 0-d tensors (data_utils.py:236-255 builds rotation matrices that way)."""
    def conv(v):
        if isinstance(v, (list, tuple)):
            return [conv(e) for e in v]
        return float(v) if isinstance(v, torch.Tensor) and dt(dtype) == torch.float32 else v
    if isinstance(x, (list, tuple)):
        # float(tensor) of an fp32 value is exact, and the final cast back to fp32 restores it bit for bit
        return T(np.asarray(conv(x), dtype=np.float64)).to(dt(dtype) or torch.float32).as_subclass(Tensor)
    return T(x, dt(dtype))
