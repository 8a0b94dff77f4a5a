"""Shared pieces of the keras / tensorflow stand-ins (see README.md).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import collections

import numpy as np
import torch

F32 = torch.float32


class Tensor(torch.Tensor):
    """Eager fp32 tensor with the two behaviours of TF eager tensors the reference relies on: immutability
    (`weights += 1e-5` at data_utils.py:179 REBINDS the name, the caller's tensor is untouched) and `.numpy()`."""

    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __imul__(self, other):
        return self * other

    def numpy(self):
        return self.detach().as_subclass(torch.Tensor).numpy()

    def __rmatmul__(self, other):
        # `np.array(...) @ tensor` (data_utils.py:266): NumPy converts the eager tensor through __array__ and returns an ndarray
        return np.asarray(other) @ self.numpy()


def T(x, dtype=None) -> Tensor:
    if isinstance(x, torch.Tensor):
        t = x if dtype is None else x.to(dtype)
    else:
        a = np.asarray(x)
        if dtype is None and a.dtype.kind == "f":
            a = a.astype(np.float32)          # python floats / float64 arrays become float32 like keras.ops on "float32" backends
        t = torch.as_tensor(a) if dtype is None else torch.as_tensor(a).to(dtype)
    return t.as_subclass(Tensor)


_DTYPES = {"float32": torch.float32, "int32": torch.int32, "int64": torch.int64, None: None}


def dt(name):
    return _DTYPES[name] if isinstance(name, (str, type(None))) else name


# ---- explicit random draws ----------------------------------------------------------------------
_draws = collections.deque()


def push_draw(a):
    """Queue the result of the next `*.random.uniform(shape)` call."""
    _draws.append(np.asarray(a, dtype=np.float32))


def pop_draw(shape):
    if not _draws:
        raise RuntimeError("refshim: random.uniform called but no draw was queued (push_draw)")
    a = _draws.popleft()
    shape = tuple(int(s) for s in shape)
    if tuple(a.shape) != shape:
        raise RuntimeError(f"refshim: queued draw has shape {a.shape}, the reference asked for {shape}")
    return T(a)


def pending_draws():
    return len(_draws)
