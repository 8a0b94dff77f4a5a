"""ctypes loader for libnerf_b200.so (the C-ABI declared in include/nerf_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  PyTorch is used only for device memory, streams and torch.distributed."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnerf_b200.so")

_lib = None


class NerfConfig(C.Structure):
    _fields_ = [
        ("num_layers", C.c_int32), ("hidden_dim", C.c_int32), ("skip_layer", C.c_int32),
        ("l_xyz", C.c_int32), ("l_dir", C.c_int32), ("ns_coarse", C.c_int32), ("ns_fine", C.c_int32),
        ("max_rays", C.c_int32), ("batch_norm", C.c_int32), ("training", C.c_int32),
        ("learning_rate", C.c_float), ("stop_grad_samples", C.c_int32),
    ]


class ForwardOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("rgb_c", "rgb_f", "depth_c", "depth_f", "w_c", "w_f", "pred_c", "pred_f", "t_all", "acc_c", "acc_f")]


# name -> (restype, argtypes); must list every symbol include/nerf_b200.h declares
_P, _I, _L, _F, _D = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
SIGNATURES = {
    "nerf_last_error": (C.c_char_p, []),
    "nerf_version": (_I, []),
    "nerf_param_count": (_L, [C.POINTER(NerfConfig)]),
    "nerf_create": (_I, [C.POINTER(NerfConfig), C.POINTER(_P)]),
    "nerf_destroy": (_I, [_P]),
    "nerf_set_weights": (_I, [_P, _I, _P, _L, _P]),
    "nerf_get_weights": (_I, [_P, _I, _P, _L, _P]),
    "nerf_grad_buffer": (_I, [_P, C.POINTER(_P), C.POINTER(_L)]),
    "nerf_get_rays": (_I, [_I, _I, _F, C.POINTER(C.c_float), _P, _P, _P]),
    "nerf_ndc_rays": (_I, [_I, _I, _F, _F, _P, _P, _P, _P, _L, _P]),
    "nerf_generate_t_vals": (_I, [_D, _D, _L, _I, _P, _I, _P, _P]),
    "nerf_sample_rays": (_I, [_P, _P, _P, _L, _I, _P, _P, _P]),
    "nerf_encode_position": (_I, [_P, _L, _I, _P, _P]),
    "nerf_volume_render": (_I, [_P, _P, _L, _I, _P, _P, _P, _P, _P]),
    "nerf_sample_pdf": (_I, [_P, _P, _P, _L, _I, _I, _P, _P]),
    "nerf_resample_merge": (_I, [_P, _P, _P, _L, _I, _I, _P, _P, _P]),
    "nerf_mlp_forward_encoded": (_I, [_P, _I, _P, _P, _L, _P, _P]),
    "nerf_mlp_forward_rays": (_I, [_P, _I, _P, _P, _P, _L, _I, _I, _P, _P]),
    "nerf_forward_pass": (_I, [_P, _P, _P, _P, _P, _L, _I, C.POINTER(ForwardOut), _P]),
    "nerf_train_forward_backward": (_I, [_P, _P, _P, _P, _P, _P, _L, _P, _P]),
    "nerf_train_phases": (_I, [_P, _P, _P, _P, _P, _P, _L, _P, _I, _P]),
    "nerf_adam_step": (_I, [_P, _F, _P]),
    "nerf_set_seed": (_I, [_P, C.c_uint64]),
    "nerf_set_exact_far_sigma": (_I, [_P, _I]),
    "nerf_set_backward_overlap": (_I, [_P, _I]),
    "nerf_set_learning_rate": (_I, [_P, _F, _P]),
    "nerf_get_optimizer_state": (_I, [_P, _P, _P, C.POINTER(_L), _P]),
    "nerf_set_optimizer_state": (_I, [_P, _P, _P, _L, _P]),
    "nerf_metric_sums": (_I, [_P, C.POINTER(_P)]),
    "nerf_metrics_accumulate": (_I, [_P, _P, _P, _P, _L, _P, _P]),
    "nerf_metrics": (_I, [_P, _P, _P, _L, _P, _P]),
    "nerf_launch_count": (_L, []),
    "nerf_timing_enable": (_I, [_I]),
    "nerf_timing_read": (_I, [_I, C.POINTER(C.c_double), C.POINTER(_L)]),
    "nerf_debug_mlp_grads": (_I, [_P, _I, _P, _P, _P, _L, _I, _P, _P, _P]),
    "nerf_selftest_mma_rate": (_I, [_I, _I, _I, _P, _P]),
    "nerf_debug_input_grad": (_I, [_P, _I, _P, _P, _P, _L, _I, _P, _P]),
    "nerf_debug_fused_input_grad": (_I, [_P, _L, _I, _P, _P]),
    "nerf_sample_pdf_bwd": (_I, [_P, _P, _P, _P, _P, _P, _L, _I, _I, _P, _P]),
    "nerf_debug_pdf_draws": (_I, [C.c_uint64, C.c_uint64, _L, _I, _P, _P]),
    "nerf_debug_flags": (_I, [_I]),
    "nerf_debug_pair_mode": (_I, [_I]),
    "nerf_debug_trace": (_I, [_P]),
    "nerf_debug_wgrad_stats": (_I, [_P]),
    "nerf_selftest_gemm_f32": (_I, [_I, _I, _L, _I, _I, _P, _L, _P, _L, C.c_float, _P, _L, _I, _P]),
    "nerf_selftest_gemm_2cta": (_I, [_P, _P, _P, _I, _I, _P]),
    "nerf_selftest_gemm_ts": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "nerf_selftest_gemm": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "nerf_selftest_collector": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "nerf_bn_param_count": (_L, [_P]),
    "nerf_bn_workspace_bytes": (_L, [_P, _L]),
    "nerf_bn_forward_backward": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _L, _P, _P, _P, _P, _L, _P]),
    "nerf_adam_flat": (_I, [_P, _P, _P, _P, _L, _L, _F, _F, _P]),
    # internal building blocks exported for unit tests (not part of the public header)
    "nerf_volume_render_bwd": (_I, [_P, _P, _P, _P, _L, _I, _P, _P, _P]),
}


def lib():
    """Load the shared library once; fail loudly when it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nerf_keras_b200 has no CPU or PyTorch fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().nerf_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what}: {msg} (code {rc})")


def launch_count() -> int:
    return int(lib().nerf_launch_count())
