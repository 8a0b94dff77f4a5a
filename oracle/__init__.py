"""CPU restatement of the ghif/nerf-keras hot path.  TEST INFRASTRUCTURE ONLY.

This package is the parity oracle: a torch-CPU fp32, op-for-op restatement of the
reference's `data_utils.py` + `models.py` (TensorFlow 2.16.2 / Keras 3.10.0 are not
installable in this image, so the reference itself cannot be run).

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 and TF/Keras cannot be imported here, so this restatement is
pinned only by the analytic known-answer tests in tests/test_oracle_known_answers.py
and by the committed fixtures under tests/golden/ that it generated itself.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import this package -- as the checker or the timed CPU
baseline, never as the product path.  The product (`nerf_keras_b200`) never imports it.
"""
from .data_utils_ref import (  # noqa: F401
    encode_position, get_rays, sample_rays, volume_render, generate_t_vals,
    sample_pdf, pose_spherical, get_translation_t, get_rotation_phi,
    get_rotation_theta, split_data, ndc_rays,
)
from .models_ref import (  # noqa: F401
    LAYER_ROLES, layer_shapes, param_count, init_weights, flatten_weights,
    unflatten_weights, nerf_mlp, forward_pass, forward_pass_with_minibatch,
    train_step, test_step, KerasAdam, mse, psnr,
)
