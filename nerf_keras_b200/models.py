"""Drop-in for the reference's `models.py`: `create_nerf_complete_model` and `NeRFTrainer` with the
reference's signatures, running on the sm_100a kernels of libnerf_b200.so (no Keras, no TF).

Differences that are forced by the platform and documented in DESIGN.md:
  * random draws (`u_pdf` for sample_pdf) may be passed explicitly; otherwise they are generated inside the resampling
    kernel (Philox keyed by `set_random_seed`, the optimiser step / call counter, the ray and the draw index);
  * BATCH_NORM=true models render on the fused kernels (inference affine folded into the Dense weights) and train on a
    separate layer-by-layer fp32 path (csrc/bn_train.cu);
  * weights are saved as `.npz` keyed by layer role (h5py is not available);
  * the training step is replayed from a CUDA graph once its input buffers have been seen twice (no per-kernel launch
    cost, no host work between kernels).
"""
from __future__ import annotations

import ctypes as C
import math
import weakref
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .data_utils import _dev, _f32, _ptr, _stream

PRECISION_BF16_TC = 0
PRECISION_FP32 = 1

_seed_state = {"rng": np.random.default_rng(42), "seed": 42}


def set_random_seed(seed: int):
    """keras.utils.set_random_seed (train_lego.py:22)."""
    _seed_state["rng"] = np.random.default_rng(seed)
    _seed_state["seed"] = int(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def layer_roles(num_layers: int) -> List[str]:
    return [f"d{i}" for i in range(num_layers)] + ["sigma", "feature", "ddir", "rgb"]


def layer_shapes(num_layers, hidden_dim, skip_layer, lxyz, ldir) -> List[Tuple[str, int, int]]:
    """(role, fan_in, fan_out) in creation order -- models.py:24-62."""
    exyz, edir = 3 + 6 * lxyz, 3 + 6 * ldir
    out, fan_in = [], exyz
    for i in range(num_layers):
        out.append((f"d{i}", fan_in, hidden_dim))
        fan_in = hidden_dim
        if i % skip_layer == 0 and i > 0:
            fan_in = hidden_dim + exyz
    out += [("sigma", fan_in, 1), ("feature", fan_in, hidden_dim),
            ("ddir", hidden_dim + edir, hidden_dim // 2), ("rgb", hidden_dim // 2, 3)]
    return out


class Adam:
    """keras.optimizers.Adam(learning_rate=...) (train_lego.py:149-151); Keras defaults otherwise."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        if (beta_1, beta_2, epsilon) != (0.9, 0.999, 1e-7):
            raise ValueError("the fused Adam kernel implements the Keras defaults beta_1=0.9, beta_2=0.999, epsilon=1e-7")
        self._lr = float(learning_rate)
        self._trainer = None

    @property
    def learning_rate(self) -> float:
        return self._lr

    @learning_rate.setter
    def learning_rate(self, value):
        self._lr = float(value)
        tr = self._trainer() if self._trainer is not None else None
        if tr is not None:
            tr._learning_rate_changed()


class MeanSquaredError:
    """keras.losses.MeanSquaredError (train_lego.py:154): mean over every element."""

    def __call__(self, y_true, y_pred):
        return torch.mean((_f32(y_pred) - _f32(y_true)) ** 2)


class _Mean:
    """keras.metrics.Mean stand-in.  With `bind(sums, i)` the sum and the count live in the context's device buffer
    (float[4]: three sums + step count, updated by the metrics kernel itself), so a step enqueues no extra kernels and
    `result()` costs one 16-byte read when somebody asks for the number."""

    def __init__(self, name):
        self.name, self.total, self.count, self._sums, self._i = name, 0.0, 0, None, 0

    def bind(self, sums: Optional[torch.Tensor], i: int):
        self._sums, self._i = sums, i

    def update_state(self, v):
        self.total = self.total + v
        self.count += 1

    def result(self):
        if self._sums is not None:
            h = self._sums.tolist()
            return h[self._i] / h[3] if h[3] else 0.0
        return self.total / self.count if self.count else 0.0

    def reset_state(self):
        self.total, self.count = 0.0, 0
        if self._sums is not None:
            self._sums.zero_()


class _Lazy:
    """A metric value that is read back from the device only when it is converted to a number."""

    def __init__(self, mean: _Mean):
        self._m = mean

    def __float__(self):
        return float(self._m.result())

    def item(self):
        return float(self)

    def __format__(self, spec):
        return format(float(self), spec)

    def __repr__(self):
        return f"{float(self):.6g}"

    # plain-number behaviour for callers that do arithmetic on the logs
    def __add__(self, o): return float(self) + o
    def __radd__(self, o): return o + float(self)
    def __sub__(self, o): return float(self) - o
    def __rsub__(self, o): return o - float(self)
    def __mul__(self, o): return float(self) * o
    def __rmul__(self, o): return o * float(self)
    def __truediv__(self, o): return float(self) / o
    def __neg__(self): return -float(self)
    def __abs__(self): return abs(float(self))
    def __lt__(self, o): return float(self) < o
    def __le__(self, o): return float(self) <= o
    def __gt__(self, o): return float(self) > o
    def __ge__(self, o): return float(self) >= o


def _timing_on() -> bool:
    """bench.py's per-kernel CUDA-event timers bracket individual launches: such steps run eagerly."""
    return bool(_timing_state["on"])


_timing_state = {"on": False}


def set_kernel_timing(on: bool):
    _timing_state["on"] = bool(on)
    _lib.check(_lib.lib().nerf_timing_enable(1 if on else 0), "nerf_timing_enable")


class _Ctx:
    """Owns one nerf_ctx (C side)."""

    def __init__(self, arch: dict, ns_coarse, ns_fine, max_rays, training, learning_rate, stop_grad_samples=True):
        _dev()
        cfg = _lib.NerfConfig(arch["num_layers"], arch["hidden_dim"], arch["skip_layer"], arch["lxyz"], arch["ldir"],
                              int(ns_coarse), int(ns_fine), int(max_rays), 0, int(bool(training)),   # BN is folded on the host
                              float(learning_rate), int(bool(stop_grad_samples)))
        self.cfg = cfg
        self.handle = C.c_void_p()
        _lib.check(_lib.lib().nerf_create(C.byref(cfg), C.byref(self.handle)), "nerf_create")
        self.n_params = int(_lib.lib().nerf_param_count(C.byref(cfg)))
        self.max_rays = int(max_rays)

    def set_weights(self, net: int, blob: torch.Tensor):
        blob = _f32(blob).reshape(-1)
        _lib.check(_lib.lib().nerf_set_weights(self.handle, net, _ptr(blob), blob.numel(), _stream()), "nerf_set_weights")
        torch.cuda.current_stream().synchronize()  # blob may be a temporary

    def get_weights(self, net: int) -> torch.Tensor:
        out = torch.empty((self.n_params,), device=_dev(), dtype=torch.float32)
        _lib.check(_lib.lib().nerf_get_weights(self.handle, net, _ptr(out), out.numel(), _stream()), "nerf_get_weights")
        return out

    @staticmethod
    def _view(ptr: int, n: int) -> torch.Tensor:
        class _Raw:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}

        return torch.as_tensor(_Raw(), device=_dev())

    def grad_tensor(self) -> torch.Tensor:
        """Zero-copy torch view of the ctx-owned flat gradient buffer [coarse | fine]."""
        if getattr(self, "_grad_view", None) is None:
            p, n = C.c_void_p(), C.c_int64()
            _lib.check(_lib.lib().nerf_grad_buffer(self.handle, C.byref(p), C.byref(n)), "nerf_grad_buffer")
            self._grad_view = self._view(p.value, n.value)
        return self._grad_view

    def metric_sums(self) -> torch.Tensor:
        """Zero-copy view of the running metric sums (loss_coarse, loss, psnr, count)."""
        if getattr(self, "_sums_view", None) is None:
            p = C.c_void_p()
            _lib.check(_lib.lib().nerf_metric_sums(self.handle, C.byref(p)), "nerf_metric_sums")
            self._sums_view = self._view(p.value, 4)
        return self._sums_view

    def optimizer_state(self):
        m = torch.empty((2 * self.n_params,), device=_dev(), dtype=torch.float32)
        v = torch.empty_like(m)
        step = C.c_int64()
        _lib.check(_lib.lib().nerf_get_optimizer_state(self.handle, _ptr(m), _ptr(v), C.byref(step), _stream()), "optimizer state")
        return m, v, int(step.value)

    def set_optimizer_state(self, m, v, step):
        _lib.check(_lib.lib().nerf_set_optimizer_state(self.handle, _ptr(m), _ptr(v), int(step), _stream()), "optimizer state")

    def close(self):
        if self.handle:
            self._grad_view = self._sums_view = None
            _lib.lib().nerf_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NerfModel:
    """What `create_nerf_complete_model` returns: the 8x256 skip MLP of models.py:24-62.

    Callable like the Keras model: `model([rays_enc, dirs_enc])` -> (..., 4) = [r,g,b,sigma] raw
    (fp32 kernels).  Weights follow the Keras Dense convention y = x @ W + b, W (in, out)."""

    BN_EPS = 1e-3   # keras.layers.BatchNormalization default epsilon (momentum 0.99 is only used in training)

    def __init__(self, num_layers, hidden_dim, skip_layer, lxyz, ldir, bn=False):
        self.arch = dict(num_layers=int(num_layers), hidden_dim=int(hidden_dim), skip_layer=int(skip_layer),
                         lxyz=int(lxyz), ldir=int(ldir), bn=bool(bn))
        self.shapes = layer_shapes(num_layers, hidden_dim, skip_layer, lxyz, ldir)
        rng = _seed_state["rng"]
        parts = []
        for _, fi, fo in self.shapes:  # Keras defaults: glorot_uniform kernel, zero bias
            lim = math.sqrt(6.0 / (fi + fo))
            parts.append(rng.uniform(-lim, lim, size=(fi * fo,)).astype(np.float32))
            parts.append(np.zeros((fo,), dtype=np.float32))
        self._host_blob = np.concatenate(parts)
        self._owner: Optional[Tuple[_Ctx, int]] = None
        self._own_ctx: Optional[_Ctx] = None
        self._bn_trainer = None      # weakref to the NeRFTrainer that holds this model's BATCH_NORM training state
        # BATCH_NORM=true (models.py:30-33, 49-52): one BatchNormalization after every trunk Dense and after the
        # direction Dense.  For rendering the moving statistics are folded into the Dense weights on the host,
        #   y = ((xW + b) - mean) * gamma / sqrt(var + eps) + beta  =  x (W s) + ((b - mean) s + beta),  s = gamma / sqrt(var + eps)
        # so the fused kernels are unchanged; training with batch statistics runs in NeRFTrainer (csrc/bn_train.cu).
        self.bn: Optional[Dict[str, Dict[str, np.ndarray]]] = None
        if bn:
            self.bn = {role: {"gamma": np.ones(fo, np.float32), "beta": np.zeros(fo, np.float32),
                              "mean": np.zeros(fo, np.float32), "var": np.ones(fo, np.float32)}
                       for role, _, fo in self.shapes if role.startswith("d")}      # d0..d7 and ddir

    # -- weights ----------------------------------------------------------------------------------
    def count_params(self) -> int:
        return int(self._host_blob.size)

    def _bn_owner(self):
        return self._bn_trainer() if self._bn_trainer is not None else None

    def get_flat_weights(self) -> np.ndarray:
        """The Dense kernels and biases (un-folded when BATCH_NORM=true)."""
        tr = self._bn_owner()
        if tr is not None:
            tr._bn_sync_models()                             # device training state -> host copies
        if self._owner is not None and self.bn is None:      # a BN model's device blob is FOLDED: its host copy is the truth
            ctx, net = self._owner
            self._host_blob = ctx.get_weights(net).cpu().numpy()
        return self._host_blob.copy()

    def device_blob(self) -> np.ndarray:
        """What the kernels see: the Dense weights, with the BatchNormalization inference affine folded in."""
        if self.bn is None:
            return self._host_blob
        parts, off = [], 0
        for role, fi, fo in self.shapes:
            W = self._host_blob[off:off + fi * fo].reshape(fi, fo); off += fi * fo
            b = self._host_blob[off:off + fo]; off += fo
            if role in self.bn:
                st = self.bn[role]
                sc = (st["gamma"].astype(np.float64) / np.sqrt(st["var"].astype(np.float64) + self.BN_EPS))
                W = (W.astype(np.float64) * sc[None, :]).astype(np.float32)
                b = ((b.astype(np.float64) - st["mean"]) * sc + st["beta"]).astype(np.float32)
            parts += [W.reshape(-1), b]
        return np.concatenate(parts)

    def _push(self):
        blob = torch.from_numpy(np.ascontiguousarray(self.device_blob()))
        if self._owner is not None:
            ctx, net = self._owner
            ctx.set_weights(net, blob)
        if self._own_ctx is not None:
            self._own_ctx.set_weights(0, blob)

    def set_flat_weights(self, blob):
        blob = np.asarray(blob, dtype=np.float32).reshape(-1)
        if blob.size != self._host_blob.size:
            raise ValueError(f"expected {self._host_blob.size} floats, got {blob.size}")
        self._host_blob = blob.copy()
        self._push()
        tr = self._bn_owner()
        if tr is not None:
            tr._bn_mark_stale()                              # the trainer re-uploads its training copy before the next step

    def get_bn_params(self) -> Optional[Dict[str, Dict[str, np.ndarray]]]:
        tr = self._bn_owner()
        if tr is not None:
            tr._bn_sync_models()
        return None if self.bn is None else {r: {k: v.copy() for k, v in st.items()} for r, st in self.bn.items()}

    def set_bn_params(self, params: Dict[str, Dict[str, np.ndarray]]):
        """role -> {gamma, beta, mean, var} (Keras order: gamma, beta, moving_mean, moving_variance)."""
        if self.bn is None:
            raise ValueError("the model was created with bn=False")
        for role, st in self.bn.items():
            for k in ("gamma", "beta", "mean", "var"):
                v = np.asarray(params[role][k], dtype=np.float32)
                if v.shape != st[k].shape:
                    raise ValueError(f"bad shape for {role}/{k}")
                st[k] = v.copy()
        self._push()
        tr = self._bn_owner()
        if tr is not None:
            tr._bn_mark_stale()

    def get_weights(self) -> Dict[str, Dict[str, np.ndarray]]:
        blob, out, off = self.get_flat_weights(), {}, 0
        for role, fi, fo in self.shapes:
            W = blob[off:off + fi * fo].reshape(fi, fo); off += fi * fo
            b = blob[off:off + fo]; off += fo
            out[role] = {"W": W.copy(), "b": b.copy()}
        return out

    def set_weights(self, weights: Dict[str, Dict[str, np.ndarray]]):
        parts = []
        for role, fi, fo in self.shapes:
            W = np.asarray(weights[role]["W"], dtype=np.float32)
            b = np.asarray(weights[role]["b"], dtype=np.float32)
            if W.shape != (fi, fo) or b.shape != (fo,):
                raise ValueError(f"bad shape for layer {role}")
            parts += [W.reshape(-1), b]
        self.set_flat_weights(np.concatenate(parts))

    # -- call -------------------------------------------------------------------------------------
    def __call__(self, inputs, training=False):
        if training and self.bn is not None:
            raise NotImplementedError("training-mode calls of a BATCH_NORM=true model go through NeRFTrainer.train_step "
                                      "(batch statistics are computed by the trainer's layer-by-layer path)")
        tr = self._bn_owner()
        if tr is not None:
            tr._bn_sync_models()
        rays_enc, dirs_enc = inputs
        x, dd = _f32(rays_enc), _f32(dirs_enc)
        ex, ed = 3 + 6 * self.arch["lxyz"], 3 + 6 * self.arch["ldir"]
        if x.shape[-1] != ex or dd.shape[-1] != ed or x.shape[:-1] != dd.shape[:-1]:
            raise ValueError(f"expected inputs [(..., {ex}), (..., {ed})]")
        if self._owner is not None:
            ctx, net = self._owner
        else:
            if self._own_ctx is None:
                self._own_ctx = _Ctx(self.arch, 2, 1, 1, False, 0.0)
                self._own_ctx.set_weights(0, torch.from_numpy(np.ascontiguousarray(self.device_blob())))
            ctx, net = self._own_ctx, 0
        n = x.numel() // ex
        out = torch.empty(x.shape[:-1] + (4,), device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().nerf_mlp_forward_encoded(ctx.handle, net, _ptr(x), _ptr(dd), n, _ptr(out), _stream()),
                   "model call")
        return out

    predict = __call__


def create_nerf_complete_model(num_layers, hidden_dim, skip_layer, lxyz, ldir, bn=False) -> NerfModel:
    """models.py:24-62."""
    return NerfModel(num_layers, hidden_dim, skip_layer, lxyz, ldir, bn)


class NeRFTrainer:
    """models.py:64-225 -- coarse->fine forward pass, train/test steps, minibatched rendering."""

    MAX_GRAPHS = 16          # captured training-step graphs kept per trainer (one per distinct set of input buffers)

    def __init__(self, coarse_model, fine_model, batch_size, ns_coarse, ns_fine, l_xyz, l_dir,
                 precision=PRECISION_BF16_TC, stop_grad_samples=False, process_group=None, use_cuda_graph=True,
                 overlap_allreduce=True, exact_far_sigma=False, backward_overlap_sms=0):
        if not isinstance(coarse_model, NerfModel):
            raise TypeError("coarse_model must be a NerfModel (create_nerf_complete_model) instance")
        if not isinstance(fine_model, NerfModel):
            raise TypeError("fine_model must be a NerfModel (create_nerf_complete_model) instance")
        if coarse_model.arch != fine_model.arch:
            raise ValueError("coarse and fine models must share one architecture")
        self.coarse_model, self.fine_model = coarse_model, fine_model
        self.batch_size, self.ns_coarse, self.ns_fine = int(batch_size), int(ns_coarse), int(ns_fine)
        self.l_xyz, self.l_dir = int(l_xyz), int(l_dir)
        self.precision = precision
        self.stop_grad_samples = bool(stop_grad_samples)
        self.process_group = process_group
        self.use_cuda_graph = bool(use_cuda_graph)
        self.overlap_allreduce = bool(overlap_allreduce)
        # rendering option: the last sample of every ray (delta = 1e10, data_utils.py:82: colour is discontinuous in its raw
        # sigma at 0) takes its sigma from the fp32 path, so bf16 rounding cannot flip that decision
        self.exact_far_sigma = bool(exact_far_sigma)
        # training option (default off, measured slower -- DESIGN.md 4.3): the weight-gradient kernel on this many SMs
        # NEXT TO the dX chain, consuming every tile's dZ images as the chain publishes them (nerf_set_backward_overlap)
        self.backward_overlap_sms = int(backward_overlap_sms)
        self.optimizer = None
        self.loss_fn = None
        self._ctx: Optional[_Ctx] = None
        self._bn_state = None
        self._graphs: Dict[tuple, tuple] = {}
        self._seen: Dict[tuple, int] = {}
        self._metrics_buf: Optional[torch.Tensor] = None
        self._seed = _seed_state["seed"]
        self.loss_coarse_tracker = _Mean("loss_coarse")
        self.loss_tracker = _Mean("loss")
        self.psnr_tracker = _Mean("psnr")

    # -- setup ------------------------------------------------------------------------------------
    def compile(self, optimizer, loss_fn):
        """models.py:80-86."""
        if not isinstance(optimizer, Adam):
            raise TypeError("optimizer must be nerf_keras_b200.models.Adam")
        if (self.coarse_model.bn is None) != (self.fine_model.bn is None):
            raise ValueError("coarse and fine model must both be created with the same bn flag")
        if not isinstance(loss_fn, MeanSquaredError):
            raise TypeError("loss_fn must be nerf_keras_b200.models.MeanSquaredError (the kernels compute the reference's "
                            "MSE(images, rgb_coarse) + MSE(images, rgb_fine), models.py:98-102)")
        self.optimizer, self.loss_fn = optimizer, loss_fn
        optimizer._trainer = weakref.ref(self)
        self._bn_state = None
        if self.coarse_model.bn is not None:
            if self.coarse_model.arch["hidden_dim"] % 4:
                raise ValueError("BATCH_NORM training needs HIDDEN_DIM % 4 == 0 (16-byte aligned parameter blocks)")
            if not self.stop_grad_samples:
                import warnings
                warnings.warn("BATCH_NORM=true training stops the gradient at the fine sample positions: the reference's "
                              "un-stopped term (models.py:166-175) is not carried through this path (DESIGN.md)")
            # BATCH_NORM=true: batch statistics couple all samples of a batch between consecutive layers, so training
            # runs on the layer-by-layer fp32 path (csrc/bn_train.cu); rendering keeps the fused kernels (folded BN).
            self.optimizer = optimizer
            self._rebuild_ctx_inference_only()
            self._bn_init_state()
            for m in (self.coarse_model, self.fine_model):
                m._bn_trainer = weakref.ref(self)
            return
        self._rebuild_ctx()

    def build(self, input_shape=None):
        if self._ctx is None:
            self._rebuild_ctx()

    # -- BATCH_NORM=true training state (csrc/bn_train.cu) -----------------------------------------------------------
    def _rebuild_ctx_inference_only(self):
        opt, self.optimizer = self.optimizer, None           # the ctx of a BN trainer is a render-only ctx
        try:
            self._rebuild_ctx()
        finally:
            self.optimizer = opt

    def _bn_init_state(self):
        dev = _dev()
        models = (self.coarse_model, self.fine_model)
        n = self.coarse_model.count_params()
        roles = list(self.coarse_model.bn.keys())
        pack = lambda m, k: np.concatenate([m.bn[r][k] for r in roles]).astype(np.float32)
        params = torch.from_numpy(np.concatenate([m._host_blob for m in models])).to(dev)
        bn = torch.from_numpy(np.concatenate([np.concatenate([pack(m, k) for k in ("gamma", "beta", "mean", "var")])
                                              for m in models])).to(dev)
        nbn = bn.numel() // 8
        z = lambda k: torch.zeros(k, device=dev, dtype=torch.float32)
        self._bn_state = dict(params=params, bn=bn, n=n, nbn=nbn, roles=roles, grads=z(2 * n), bn_grads=z(4 * nbn),
                              m=z(2 * n), v=z(2 * n), bm=z(8 * nbn), bv=z(8 * nbn), step=0, ws=None, ws_rays=0, dirty=False,
                              stale=False, syncing=False)

    def _bn_mark_stale(self):
        """A model's weights / BN parameters were set from outside (load_weights, set_weights ...): the device training
        copy is re-uploaded from the models before the next step (the Adam moments are kept)."""
        st = getattr(self, "_bn_state", None)
        if st and not st.get("syncing"):
            st["stale"] = True

    def _bn_refresh_state(self):
        st = self._bn_state
        dev = st["params"].device
        models = (self.coarse_model, self.fine_model)
        pack = lambda m, k: np.concatenate([m.bn[r][k] for r in st["roles"]]).astype(np.float32)
        st["params"].copy_(torch.from_numpy(np.concatenate([m._host_blob for m in models])).to(dev))
        st["bn"].copy_(torch.from_numpy(np.concatenate([np.concatenate([pack(m, k) for k in ("gamma", "beta", "mean", "var")])
                                                        for m in models])).to(dev))
        st["stale"] = False

    def _bn_sync_models(self):
        """Device training state -> the two models (un-folded weights + BN parameters) -> folded weights of the render ctx."""
        st = getattr(self, "_bn_state", None)
        if not st or not st["dirty"] or st.get("syncing"):
            return
        st["dirty"] = False
        st["syncing"] = True
        try:
            self._bn_sync_models_locked(st)
        finally:
            st["syncing"] = False

    def _bn_sync_models_locked(self, st):
        params, bn = st["params"].cpu().numpy(), st["bn"].cpu().numpy()
        n, nbn = st["n"], st["nbn"]
        for i, m in enumerate((self.coarse_model, self.fine_model)):
            blk = bn[i * 4 * nbn:(i + 1) * 4 * nbn]
            off, out = 0, {}
            for r in st["roles"]:
                c = m.bn[r]["gamma"].shape[0]
                out[r] = {k: blk[j * nbn + off: j * nbn + off + c].copy() for j, k in enumerate(("gamma", "beta", "mean", "var"))}
                off += c
            m.bn = out
            m.set_flat_weights(params[i * n:(i + 1) * n])      # pushes the folded blob to the ctx

    def _bn_train_step(self, images, o, d, t, u):
        st = self._bn_state
        if st.get("stale"):
            self._bn_refresh_state()
        B = o.shape[0]
        L = _lib.lib()
        cfg = self._ctx.cfg
        if st["ws"] is None or st["ws_rays"] < B:
            st["ws"] = torch.empty(int(L.nerf_bn_workspace_bytes(C.byref(cfg), B)), dtype=torch.uint8, device=o.device)
            st["ws_rays"] = B
        metrics = torch.empty((3,), device=o.device, dtype=torch.float32)
        _lib.check(L.nerf_bn_forward_backward(C.byref(cfg), _ptr(st["params"]), _ptr(st["bn"]), _ptr(images), _ptr(o), _ptr(d),
                                              _ptr(t), _ptr(u), B, _ptr(st["grads"]), _ptr(st["bn_grads"]), _ptr(metrics),
                                              _ptr(st["ws"]), st["ws"].numel(), _stream()), "bn train_step")
        st["step"] += 1
        lr, n, nbn = float(self.optimizer.learning_rate), st["n"], st["nbn"]
        # data parallel (train_tpu_*.py): gradients of the Dense and BatchNormalization parameters are averaged over the
        # replicas; batch statistics stay per replica (Keras' BatchNormalization is not synchronised across replicas) and the
        # moving statistics are averaged, which is what a MEAN-aggregated replica variable ends up holding
        world, scale = self._world(), 1.0
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(st["grads"], group=self.process_group)
            dist.all_reduce(st["bn_grads"], group=self.process_group)
            for i in range(2):
                mv = st["bn"][4 * i * nbn + 2 * nbn:4 * (i + 1) * nbn]
                dist.all_reduce(mv, group=self.process_group)
                mv.mul_(1.0 / world)
            scale = 1.0 / world
        adam = lambda p, g, m, v, k: _lib.check(L.nerf_adam_flat(p, g, m, v, k, st["step"], lr, scale, _stream()), "adam")
        adam(_ptr(st["params"]), _ptr(st["grads"]), _ptr(st["m"]), _ptr(st["v"]), 2 * n)
        for i in range(2):      # gamma | beta of each net are the first 2 nbn floats of its [gamma | beta | mean | var] block
            e = 4 * i * nbn * 4
            adam(_ptr(st["bn"]) + e, _ptr(st["bn_grads"]) + 2 * i * nbn * 4, _ptr(st["bm"]) + e, _ptr(st["bv"]) + e, 2 * nbn)
        st["dirty"] = True
        return self._update_metrics(metrics)

    def _rebuild_ctx(self, max_rays=None):
        for m in (self.coarse_model, self.fine_model):
            m.get_flat_weights()                              # refresh the host copies from the old ctx
        blobs = [np.ascontiguousarray(self.coarse_model.device_blob()), np.ascontiguousarray(self.fine_model.device_blob())]
        training = self.optimizer is not None
        opt_state = None
        if self._ctx is not None:
            if training and self._ctx.cfg.training:
                opt_state = self._ctx.optimizer_state()       # Adam moments and step survive a workspace resize
                torch.cuda.current_stream().synchronize()
            self._graphs.clear()
            self._seen.clear()
            self._ctx.close()
        lr = self.optimizer.learning_rate if training else 0.0
        self._ctx = _Ctx(self.coarse_model.arch, self.ns_coarse, self.ns_fine, max_rays or self.batch_size, training, lr,
                         self.stop_grad_samples)
        _lib.check(_lib.lib().nerf_set_seed(self._ctx.handle, int(self._seed) & 0xFFFFFFFFFFFFFFFF), "nerf_set_seed")
        _lib.check(_lib.lib().nerf_set_exact_far_sigma(self._ctx.handle, int(self.exact_far_sigma)), "exact_far_sigma")
        if training and self.backward_overlap_sms:
            _lib.check(_lib.lib().nerf_set_backward_overlap(self._ctx.handle, self.backward_overlap_sms), "backward_overlap")
        for net, (m, blob) in enumerate(zip((self.coarse_model, self.fine_model), blobs)):
            m._owner = (self._ctx, net)
            self._ctx.set_weights(net, torch.from_numpy(blob))
        if opt_state is not None:
            self._ctx.set_optimizer_state(*opt_state)
        # BATCH_NORM trainers keep host-side trackers (their training step does not go through this context)
        sums = self._ctx.metric_sums() if self.coarse_model.bn is None else None
        for i, tr in enumerate((self.loss_coarse_tracker, self.loss_tracker, self.psnr_tracker)):
            tr.bind(sums, i)

    def _learning_rate_changed(self):
        if self._ctx is not None and self._ctx.cfg.training:
            _lib.check(_lib.lib().nerf_set_learning_rate(self._ctx.handle, float(self.optimizer.learning_rate), _stream()),
                       "set_learning_rate")

    @property
    def metrics(self):
        return [self.loss_tracker, self.psnr_tracker]  # models.py:147-149 (loss_coarse is not listed, Q15)

    def reset_metrics(self):
        for m in (self.loss_coarse_tracker, self.loss_tracker, self.psnr_tracker):
            m.reset_state()

    # -- forward ----------------------------------------------------------------------------------
    def _forward_tile(self, o, d, t, u_pdf, precision, maps_only=False):
        B = o.shape[0]
        Nc, Na = self.ns_coarse, self.ns_coarse + self.ns_fine
        dev = o.device
        e = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
        out = dict(rgb_c=e(B, 3), rgb_f=e(B, 3), depth_c=e(B), depth_f=e(B))
        if not maps_only:   # per-sample outputs (weights, raw predictions, sample positions): 4.6 KB per ray
            out.update(w_c=e(B, Nc), w_f=e(B, Na), pred_c=e(B, Nc, 4), pred_f=e(B, Na, 4), t_all=e(B, Na))
        if self.ns_fine == 0:       # single-net shape: the "fine" outputs do not exist
            for k in ("rgb_f", "depth_f", "w_f", "pred_f", "t_all"):
                out.pop(k, None)
        fo = _lib.ForwardOut(*[_ptr(out.get(k)) for k, _ in _lib.ForwardOut._fields_])
        _lib.check(_lib.lib().nerf_forward_pass(self._ctx.handle, _ptr(o), _ptr(d), _ptr(t), _ptr(u_pdf), B, precision,
                                                C.byref(fo), _stream()), "forward_pass")
        return out

    def forward_pass(self, ray_origins, ray_directions, t_vals, l_xyz=None, l_dir=None, training=False,
                     batch_size=None, u_pdf=None, precision=None, return_t_all=False, maps_only=False):
        """models.py:151-176.  Returns ((rgb_c,rgb_f),(depth_c,depth_f),(w_c,w_f),(pred_c,pred_f)).
        maps_only=True skips the per-sample outputs (their pairs are (None, None)): what a renderer needs."""
        if (l_xyz not in (None, self.l_xyz)) or (l_dir not in (None, self.l_dir)):
            raise ValueError("l_xyz / l_dir differ from the trainer's configuration")
        if training and (self.coarse_model.bn is not None or self.fine_model.bn is not None):
            raise NotImplementedError("BATCH_NORM=true is inference-only on the B200 path")
        if self._ctx is None:
            self._rebuild_ctx()
        self._bn_sync_models()
        o, d, t = _f32(ray_origins), _f32(ray_directions), _f32(t_vals)
        if t.shape != (o.shape[0], self.ns_coarse):
            raise ValueError(f"t_vals must have shape (n_rays, {self.ns_coarse})")
        B = o.shape[0]
        u = None if u_pdf is None else _f32(u_pdf)      # None: drawn inside the resampling kernel (data_utils.py:196)
        if u is not None and u.shape != (B, self.ns_fine):
            raise ValueError(f"u_pdf must have shape (n_rays, {self.ns_fine})")
        precision = self.precision if precision is None else precision
        tile = self._ctx.max_rays
        outs = [self._forward_tile(o[s:s + tile], d[s:s + tile], t[s:s + tile], None if u is None else u[s:s + tile],
                                   precision, maps_only)
                for s in range(0, B, tile)]
        cat = (lambda k: None if k not in outs[0] else
               (outs[0][k] if len(outs) == 1 else torch.cat([x[k] for x in outs], dim=0)))
        if maps_only:
            return ((cat("rgb_c"), cat("rgb_f")), (cat("depth_c"), cat("depth_f")), (None, None), (None, None))
        res = ((cat("rgb_c"), cat("rgb_f")), (cat("depth_c"), cat("depth_f")), (cat("w_c"), cat("w_f")),
               (cat("pred_c"), cat("pred_f")))
        return res + (cat("t_all"),) if return_t_all else res

    def mlp_forward_rays(self, net, ray_origins, ray_directions, t_vals, precision=None):
        """Fused sample_rays + encode_position x2 + model call for one net (models.py:152-157 /
        169-173): returns the raw predictions (B, N, 4).  net: "coarse" | "fine"."""
        if self._ctx is None:
            self._rebuild_ctx()
        o, d, t = _f32(ray_origins), _f32(ray_directions), _f32(t_vals)
        B, N = t.shape
        idx = {"coarse": 0, "fine": 1}[net]
        precision = self.precision if precision is None else precision
        out = torch.empty((B, N, 4), device=o.device, dtype=torch.float32)
        tile = self._ctx.max_rays                       # larger inputs are tiled: the context (and Adam's state) stays
        for s0 in range(0, B, tile):
            n = min(tile, B - s0)
            _lib.check(_lib.lib().nerf_mlp_forward_rays(self._ctx.handle, idx, _ptr(o[s0:s0 + n]), _ptr(d[s0:s0 + n]),
                                                        _ptr(t[s0:s0 + n]), n, N, precision, _ptr(out[s0:s0 + n]),
                                                        _stream()), "mlp_forward_rays")
        return out

    def debug_mlp_grads(self, net, ray_origins, ray_directions, t_vals, d_preds, return_input_grad=False):
        """Diagnostics: (preds, d(sum(preds*d_preds))/d(weights of `net`)) through the tcgen05 forward
        (saved activations) and backward kernels.  Needs compile() (training workspace)."""
        o, d, t, dp = _f32(ray_origins), _f32(ray_directions), _f32(t_vals), _f32(d_preds)
        B, N = t.shape
        idx = {"coarse": 0, "fine": 1}[net]
        preds = torch.empty((B, N, 4), device=o.device, dtype=torch.float32)
        _lib.check(_lib.lib().nerf_debug_mlp_grads(self._ctx.handle, idx, _ptr(o), _ptr(d), _ptr(t), B, N, _ptr(dp),
                                                   _ptr(preds), _stream()), "debug_mlp_grads")
        g = self._ctx.grad_tensor()
        n = self._ctx.n_params
        grads = g[idx * n:(idx + 1) * n].clone()
        if not return_input_grad:
            return preds, grads
        dtp = torch.empty((B, N), device=o.device, dtype=torch.float32)
        if return_input_grad == "fused":      # by-product of the weight-gradient kernel (what train_step uses)
            _lib.check(_lib.lib().nerf_debug_fused_input_grad(self._ctx.handle, B, N, _ptr(dtp), _stream()), "fused_input_grad")
        else:                                 # the stand-alone input-gradient kernel
            _lib.check(_lib.lib().nerf_debug_input_grad(self._ctx.handle, idx, _ptr(o), _ptr(d), _ptr(t), B, N, _ptr(dtp),
                                                        _stream()), "debug_input_grad")
        return preds, grads, dtp

    def forward_pass_with_minibatch(self, ray_origins, ray_directions, t_vals, l_xyz=None, l_dir=None, batch_size=512,
                                    training=False, u_pdf=None, precision=None, maps_only=False):
        """models.py:178-225 -- ray-tile loop; tiles are `batch_size` rays (capped by the workspace)."""
        o, d, t = _f32(ray_origins), _f32(ray_directions), _f32(t_vals)
        B = o.shape[0]
        u = None if u_pdf is None else _f32(u_pdf)
        outs = [self.forward_pass(o[s:s + batch_size], d[s:s + batch_size], t[s:s + batch_size], l_xyz, l_dir,
                                  training=training, u_pdf=None if u is None else u[s:s + batch_size], precision=precision,
                                  maps_only=maps_only)
                for s in range(0, B, batch_size)]
        cat = lambda i, j: None if outs[0][i][j] is None else torch.cat([x[i][j] for x in outs], dim=0)
        return tuple((cat(i, 0), cat(i, 1)) for i in range(4))

    # -- steps ------------------------------------------------------------------------------------
    def _world(self) -> int:
        if self.process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return torch.distributed.get_world_size(self.process_group)
        return 1

    def _step_body(self, images, o, d, t, u, B, accumulate=False, grad_scale=1.0, apply=True):
        """One training step enqueued on the current stream: forward, loss, backward, gradient all-reduce, Adam.
        No host synchronisation and no allocation: this is what the CUDA graph captures."""
        L, h = _lib.lib(), self._ctx.handle
        args = (h, _ptr(images), _ptr(o), _ptr(d), _ptr(t), _ptr(u), B, _ptr(self._metrics_buf))
        acc = 4 if accumulate else 0
        world = self._world()
        if world == 1 or not apply:
            _lib.check(L.nerf_train_phases(*args, 3 | acc, _stream()), "train_step")
        else:
            import torch.distributed as dist
            g, n = self._ctx.grad_tensor(), self._ctx.n_params
            if self.overlap_allreduce and self.ns_fine > 0:
                # the fine net's half of the flat gradient buffer is complete after phase 0: reduce it over NVLink while
                # the coarse net's backward still runs; the collective runs on NCCL's own stream
                _lib.check(L.nerf_train_phases(*args, 1 | acc, _stream()), "train_step (forward + fine backward)")
                w_fine = dist.all_reduce(g[n:], op=dist.ReduceOp.SUM, group=self.process_group, async_op=True)
                _lib.check(L.nerf_train_phases(*args, 2 | acc, _stream()), "train_step (coarse backward)")
                dist.all_reduce(g[:n], op=dist.ReduceOp.SUM, group=self.process_group)
                w_fine.wait()
            else:
                _lib.check(L.nerf_train_phases(*args, 3 | acc, _stream()), "train_step")
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.process_group)
            grad_scale = grad_scale / world
        if apply:
            _lib.check(L.nerf_adam_step(h, float(grad_scale), _stream()), "adam_step")

    def train_step(self, inputs, u_pdf=None):
        """models.py:88-120.  inputs = (images (B,3), (ray_origins, ray_directions, t_vals)).
        u_pdf=None (production): the uniform draws of sample_pdf are generated inside the kernels."""
        if self.optimizer is None:
            raise RuntimeError("call compile(optimizer, loss_fn) before train_step")
        images, (o, d, t) = inputs
        images, o, d, t = _f32(images), _f32(o), _f32(d), _f32(t)
        B = o.shape[0]
        if getattr(self, "_bn_state", None) is not None:
            u = torch.rand((B, self.ns_fine), device=o.device, dtype=torch.float32) if u_pdf is None else _f32(u_pdf)
            return self._bn_train_step(images, o, d, t, u)
        if t.shape != (B, self.ns_coarse) or images.shape != (B, 3):
            raise ValueError(f"expected images (n_rays, 3) and t_vals (n_rays, {self.ns_coarse})")
        u = None if u_pdf is None else _f32(u_pdf)
        if self._metrics_buf is None or self._metrics_buf.device != o.device:
            self._metrics_buf = torch.empty((3,), device=o.device, dtype=torch.float32)
        cap = self._ctx.max_rays
        if B > cap:
            # a batch larger than the workspace: equal micro-batches, gradients accumulated in the context, one update.
            # mean over the batch = mean of the micro-batch means (the metrics are averaged the same way)
            if B % cap:
                raise ValueError(f"a batch of {B} rays must be a multiple of the context's {cap}-ray workspace")
            k = B // cap
            for i in range(k):
                sl = slice(i * cap, (i + 1) * cap)
                self._step_body(images[sl], o[sl], d[sl], t[sl], None if u is None else u[sl], cap, accumulate=i > 0,
                                grad_scale=1.0 / k, apply=(i == k - 1))
            return self._logs()
        key = (images.data_ptr(), o.data_ptr(), d.data_ptr(), t.data_ptr(), 0 if u is None else u.data_ptr(), B)
        if self.use_cuda_graph and not _timing_on():
            entry = self._graphs.get(key)
            if entry is None and self._seen.get(key, 0) >= 1 and len(self._graphs) < self.MAX_GRAPHS:
                entry = self._capture(key, images, o, d, t, u, B)
            if entry is not None:
                entry[0].replay()
                return self._logs()
            self._seen[key] = self._seen.get(key, 0) + 1
            if len(self._seen) > 4096:
                self._seen.clear()
        self._step_body(images, o, d, t, u, B)
        return self._logs()     # stays on the device: no host synchronisation inside the step

    def release_graphs(self):
        """Destroy the captured step graphs.  Call this before torch.distributed.destroy_process_group(): NCCL does not
        tear a communicator down while a CUDA graph that captured its collectives is still alive."""
        self._graphs.clear()
        self._seen.clear()
        import gc
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def _capture(self, key, images, o, d, t, u, B):
        """Capture one training step on these input buffers.  The buffers are kept alive with the graph; the step count,
        the learning rate and the random draws are read from device memory, so every replay is a fresh step."""
        g = torch.cuda.CUDAGraph()
        torch.cuda.current_stream().synchronize()
        try:
            with torch.cuda.graph(g):
                self._step_body(images, o, d, t, u, B)
        except Exception:
            # a failed capture leaves no partial update behind (nothing was executed); fall back to eager for this key
            self._seen[key] = -(1 << 30)
            torch.cuda.synchronize()
            return None
        # capturing does not execute: the host-side mirror of the step count was advanced once by nerf_adam_step
        self._graphs[key] = (g, (images, o, d, t, u))
        return self._graphs[key]

    def test_step(self, inputs, u_pdf=None):
        """models.py:122-145."""
        images, (o, d, t) = inputs
        images = _f32(images)
        rgbs, _, _, _ = self.forward_pass(o, d, t, u_pdf=u_pdf, maps_only=True)
        metrics = torch.empty((3,), device=images.device, dtype=torch.float32)
        rgb_f = rgbs[1] if rgbs[1] is not None else rgbs[0]
        if self.coarse_model.bn is not None:
            _lib.check(_lib.lib().nerf_metrics(_ptr(images), _ptr(rgbs[0]), _ptr(rgb_f), images.shape[0], _ptr(metrics),
                                               _stream()), "test_step")
            return self._update_metrics(metrics)
        _lib.check(_lib.lib().nerf_metrics_accumulate(self._ctx.handle, _ptr(images), _ptr(rgbs[0]), _ptr(rgb_f),
                                                      images.shape[0], _ptr(metrics), _stream()), "test_step")
        self.last_metrics = metrics
        return self._logs()

    def _logs(self):
        """Running means as Keras reports them (models.py:116-120); read back lazily."""
        return {"loss_coarse": _Lazy(self.loss_coarse_tracker), "loss": _Lazy(self.loss_tracker),
                "psnr": _Lazy(self.psnr_tracker)}

    def _update_metrics(self, m):
        self.loss_coarse_tracker.update_state(m[0])
        self.loss_tracker.update_state(m[1])  # `loss` is the FINE loss only (models.py:114)
        self.psnr_tracker.update_state(m[2])
        return {"loss_coarse": self.loss_coarse_tracker.result(), "loss": self.loss_tracker.result(),
                "psnr": self.psnr_tracker.result()}

    def fit(self, train_ds: Iterable, validation_data: Optional[Iterable] = None, epochs=1, callbacks=None, verbose=1):
        """Minimal stand-in for keras Model.fit as used at train_lego.py:279-284."""
        history: Dict[str, list] = {"loss": [], "psnr": [], "loss_coarse": [], "val_loss": [], "val_psnr": []}
        for epoch in range(epochs):
            self.reset_metrics()
            logs = {}
            for batch in train_ds:
                logs = self.train_step(batch)
            logs = {k: float(v) for k, v in logs.items()}     # read the running means back before they are reset
            for k in ("loss", "psnr", "loss_coarse"):
                history[k].append(logs[k] if k in logs else None)
            if validation_data is not None:
                self.reset_metrics()
                vlogs = {}
                for batch in validation_data:
                    vlogs = self.test_step(batch)
                vlogs = {k: float(v) for k, v in vlogs.items()}
                history["val_loss"].append(float(vlogs["loss"]) if "loss" in vlogs else None)
                history["val_psnr"].append(float(vlogs["psnr"]) if "psnr" in vlogs else None)
                logs = dict(logs, val_loss=vlogs.get("loss"), val_psnr=vlogs.get("psnr"))
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs} " + " ".join(f"{k}: {float(v):.5f}" for k, v in logs.items() if v is not None))
            for cb in callbacks or []:
                cb.on_epoch_end(epoch, logs)
        return history

    # -- weights I/O (train_lego.py:199-213, inference.py:170) ------------------------------------
    def save_weights(self, path: str):
        self._bn_sync_models()
        data = {}
        for name, m in (("coarse", self.coarse_model), ("fine", self.fine_model)):
            for role, wb in m.get_weights().items():
                data[f"{name}/{role}/W"] = wb["W"]
                data[f"{name}/{role}/b"] = wb["b"]
            for role, st in (m.get_bn_params() or {}).items():
                for k, v in st.items():
                    data[f"{name}/{role}/bn_{k}"] = v
        np.savez(path, **data)

    def load_weights(self, path: str):
        data = np.load(path if path.endswith(".npz") else path + ".npz")
        for name, m in (("coarse", self.coarse_model), ("fine", self.fine_model)):
            m.set_weights({role: {"W": data[f"{name}/{role}/W"], "b": data[f"{name}/{role}/b"]}
                           for role, _, _ in m.shapes})
            if m.bn is not None:
                m.set_bn_params({role: {k: data[f"{name}/{role}/bn_{k}"] for k in ("gamma", "beta", "mean", "var")}
                                 for role in m.bn})
