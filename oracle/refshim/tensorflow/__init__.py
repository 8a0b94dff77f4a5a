"""Stand-in for the `tensorflow` package: the raw-TF calls of data_utils.py:172-223 and models.py (GradientTape, tf.data)."""
import torch

from refshim_core import T, Tensor, pop_draw


class GradientTape:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def gradient(self, loss, variables):
        return list(torch.autograd.grad(loss, list(variables), allow_unused=True))


def reduce_sum(x, axis=None, keepdims=False):
    return torch.sum(T(x), dim=axis, keepdim=keepdims)


def cumsum(x, axis=0):
    return torch.cumsum(T(x), dim=axis)


def concat(xs, axis):
    return torch.cat([T(x) for x in xs], dim=axis)


def zeros_like(x): return torch.zeros_like(T(x))
def ones_like(x): return torch.ones_like(T(x))


class random:
    @staticmethod
    def uniform(shape, minval=0, maxval=None, dtype=None, seed=None):
        return pop_draw(shape)


def searchsorted(sorted_sequence, values, side="left"):
    return T(torch.searchsorted(T(sorted_sequence).detach().contiguous(), T(values).detach().contiguous(),
                                right=(side == "right")).to(torch.int32))


def maximum(a, b): return torch.maximum(*_pair(a, b))
def minimum(a, b): return torch.minimum(*_pair(a, b))


def _pair(a, b):
    a = a if isinstance(a, torch.Tensor) else None if a is None else a
    ta = isinstance(a, torch.Tensor)
    tb = isinstance(b, torch.Tensor)
    ref = a if ta else b
    if not ta:
        a = torch.full_like(ref, a)
    if not tb:
        b = torch.full_like(ref, b)
    return a, b


def stack(xs, axis=0):
    return torch.stack(list(xs), dim=axis)


def gather(params, indices, axis=-1, batch_dims=0):
    """tf.gather(params (B,K), indices (B,...), axis=-1, batch_dims=1): out[b, ...] = params[b, indices[b, ...]]."""
    p, i = T(params), indices.to(torch.int64)
    assert p.dim() == 2 and batch_dims == 1 and axis in (-1, 1)
    return torch.gather(p, 1, i.reshape(i.shape[0], -1)).reshape(i.shape)


def where(cond, a, b):
    return torch.where(cond, a, b)


class _Dataset:
    def __init__(self, tensors, batch=None):
        self.tensors, self.bs = tensors, batch

    @staticmethod
    def from_tensor_slices(tensors):
        return _Dataset(tuple(T(t) for t in tensors))

    def batch(self, n, drop_remainder=False, num_parallel_calls=None):
        return _Dataset(self.tensors, int(n))

    def prefetch(self, n):
        return self

    def __iter__(self):
        n = self.tensors[0].shape[0]
        for s in range(0, n, self.bs):
            yield tuple(t[s:s + self.bs] for t in self.tensors)


class data:
    Dataset = _Dataset
    AUTOTUNE = -1
