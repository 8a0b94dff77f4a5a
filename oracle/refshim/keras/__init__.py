"""Stand-in for the `keras` package: only what /root/reference/{data_utils,models}.py import.  See ../README.md."""
import numpy as np
import torch

from . import layers, ops, random  # noqa: F401
from .layers import Input, KerasTensor
from refshim_core import T


class Model:
    """Functional `keras.Model(inputs=..., outputs=...)` and the base of subclassed models (NeRFTrainer)."""

    def __init__(self, inputs=None, outputs=None):
        self._inputs, self._outputs = inputs, outputs
        if outputs is not None:
            # layers of this graph in creation order (= the order Keras builds `trainable_variables` for this topology
            # is by graph depth; the reference only zips gradients with the same list, so any fixed order is equivalent)
            seen, order = set(), []

            def walk(t):
                if id(t) in seen or t.layer is None:
                    return
                seen.add(id(t))
                for i in t.inputs:
                    walk(i)
                if t.layer not in order:
                    order.append(t.layer)

            walk(outputs)
            self.layers = sorted(order, key=layers._registry.index)

    def compile(self, *a, **k):
        pass

    @property
    def trainable_variables(self):
        return [v for layer in self.layers for v in layer.trainable]

    def __call__(self, inputs, training=False):
        ins = self._inputs if isinstance(self._inputs, (list, tuple)) else [self._inputs]
        vals = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        memo = {id(s): T(v) for s, v in zip(ins, vals)}

        def ev(t):
            if id(t) not in memo:
                args = [ev(i) for i in t.inputs]
                memo[id(t)] = t.layer.call(args if isinstance(t.layer, layers._Concat) else args[0], training)
            return memo[id(t)]

        return ev(self._outputs)

    def predict(self, inputs, batch_size=None):
        with torch.no_grad():
            return self(inputs, training=False)


class _Mean:
    def __init__(self, name=None):
        self.name, self.total, self.count = name, 0.0, 0

    def update_state(self, v):
        self.total += float(v)
        self.count += 1

    def result(self):
        return self.total / max(self.count, 1)

    def reset_state(self):
        self.total, self.count = 0.0, 0


class metrics:
    Mean = _Mean


class _MSE:
    """keras.losses.MeanSquaredError: mean over the last axis, then sum_over_batch_size."""

    def __call__(self, y_true, y_pred):
        per = torch.mean((T(y_pred) - T(y_true)) ** 2, dim=-1)
        return torch.sum(per) / per.numel()


class losses:
    MeanSquaredError = _MSE


class _Adam:
    """keras.optimizers.Adam update (Keras-3 `update_step`, restated as in oracle.models_ref.KerasAdam)."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps, self.t, self.state = learning_rate, beta_1, beta_2, epsilon, 0, {}

    @torch.no_grad()
    def apply_gradients(self, grads_and_vars):
        import math
        self.t += 1
        alpha = np.float32(self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t))
        for g, p in grads_and_vars:
            m, v = self.state.setdefault(id(p), (torch.zeros_like(p), torch.zeros_like(p)))
            m.add_((g - m) * np.float32(1.0 - self.b1))
            v.add_((g * g - v) * np.float32(1.0 - self.b2))
            p.sub_(float(alpha) * m / (torch.sqrt(v) + np.float32(self.eps)))


class optimizers:
    Adam = _Adam


class utils:
    @staticmethod
    def set_random_seed(seed):
        layers._rng["rng"] = np.random.default_rng(seed)
