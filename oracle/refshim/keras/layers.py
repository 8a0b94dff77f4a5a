"""`keras.layers` stand-in: functional-API Dense / BatchNormalization / ReLU / concatenate on torch-CPU fp32."""
import math

import numpy as np
import torch

from refshim_core import T

_rng = {"rng": np.random.default_rng(0)}
_registry = []            # layers in creation order since the last keras.Input() that started a model


class KerasTensor:
    """Symbolic tensor of the functional API: remembers the layer call that produces it."""

    def __init__(self, layer, inputs, last_dim):
        self.layer, self.inputs, self.last_dim = layer, inputs, last_dim


def Input(shape):
    return KerasTensor(None, [], int(shape[-1]))


class Layer:
    def __init__(self):
        _registry.append(self)

    def __call__(self, x, training=False):
        xs = x if isinstance(x, (list, tuple)) else [x]
        if any(isinstance(v, KerasTensor) for v in xs):
            return KerasTensor(self, list(xs), self.symbolic(xs))
        return self.call(x, training)

    trainable = ()


class Dense(Layer):
    """y = x @ kernel + bias, glorot-uniform kernel, zero bias (Keras defaults); activation None | "relu"."""

    def __init__(self, units, activation=None):
        super().__init__()
        assert activation in (None, "relu")
        self.units, self.activation, self.kernel, self.bias = int(units), activation, None, None

    def symbolic(self, xs):
        fan_in = xs[0].last_dim
        lim = math.sqrt(6.0 / (fan_in + self.units))
        k = _rng["rng"].uniform(-lim, lim, size=(fan_in, self.units)).astype(np.float32)
        self.kernel = torch.from_numpy(k).requires_grad_(True)
        self.bias = torch.zeros(self.units, dtype=torch.float32).requires_grad_(True)
        return self.units

    @property
    def trainable(self):
        return (self.kernel, self.bias)

    def call(self, x, training):
        y = torch.matmul(T(x), self.kernel) + self.bias
        return torch.relu(y) if self.activation == "relu" else y


class BatchNormalization(Layer):
    """Keras defaults: axis -1, momentum 0.99, epsilon 1e-3; batch statistics over all leading axes when training."""

    def __init__(self, momentum=0.99, epsilon=1e-3):
        super().__init__()
        self.momentum, self.epsilon = momentum, epsilon

    def symbolic(self, xs):
        n = xs[0].last_dim
        self.gamma = torch.ones(n).requires_grad_(True)
        self.beta = torch.zeros(n).requires_grad_(True)
        self.moving_mean, self.moving_variance = torch.zeros(n), torch.ones(n)
        return n

    @property
    def trainable(self):
        return (self.gamma, self.beta)

    def call(self, x, training):
        x = T(x)
        if training:
            red = tuple(range(x.dim() - 1))
            mean, var = x.mean(dim=red), x.var(dim=red, unbiased=False)
            with torch.no_grad():
                self.moving_mean.mul_(self.momentum).add_((1 - self.momentum) * mean)
                self.moving_variance.mul_(self.momentum).add_((1 - self.momentum) * var)
        else:
            mean, var = self.moving_mean, self.moving_variance
        return (x - mean) * torch.rsqrt(var + self.epsilon) * self.gamma + self.beta


class ReLU(Layer):
    def symbolic(self, xs):
        return xs[0].last_dim

    def call(self, x, training):
        return torch.relu(T(x))


class _Concat(Layer):
    def __init__(self, axis):
        super().__init__()
        self.axis = axis

    def symbolic(self, xs):
        return sum(v.last_dim for v in xs)

    def call(self, xs, training):
        return torch.cat([T(v) for v in xs], dim=self.axis)


def concatenate(xs, axis=-1):
    return _Concat(axis)(list(xs))
