"""nerf_keras_b200 -- B200-native (sm_100a) drop-in for the ghif/nerf-keras ray-render / train path.

Host-side mirror of the reference's `data_utils.py` and `models.py`; all arithmetic runs in
hand-written CUDA kernels behind the C-ABI of libnerf_b200.so (include/nerf_b200.h)."""
from . import data_utils, models, dist, real_data  # noqa: F401
from .data_utils import (encode_position, get_rays, sample_rays, volume_render, generate_t_vals,  # noqa: F401
                         sample_pdf, pose_spherical, split_data, ndc_rays, resample_merge,
                         create_batched_dataset_pipeline)
from .models import (create_nerf_complete_model, NeRFTrainer, NerfModel, Adam, MeanSquaredError,  # noqa: F401
                     set_random_seed, PRECISION_BF16_TC, PRECISION_FP32)

__version__ = "0.1.0"
