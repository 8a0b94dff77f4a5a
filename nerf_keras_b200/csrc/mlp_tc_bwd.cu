// Backward of the fused tcgen05 MLP (placeholder until the backward kernels land in this file).
#include "common.cuh"
#include "ctx.cuh"

namespace nerf {
int64_t tc_save_bytes_per_tile();

int tc_train_alloc(nerf_ctx* ctx) {
    const nerf_config& c = ctx->cfg;
    const int64_t tiles_c = ceil_div((int64_t)c.max_rays * c.ns_coarse, 256) * 2;
    const int64_t tiles_f = ceil_div((int64_t)c.max_rays * (c.ns_coarse + c.ns_fine), 256) * 2;
    NERF_CUDA(cudaMalloc((void**)&ctx->act_save[0], (size_t)(tiles_c * tc_save_bytes_per_tile())));
    NERF_CUDA(cudaMalloc((void**)&ctx->act_save[1], (size_t)(tiles_f * tc_save_bytes_per_tile())));
    return NERF_OK;
}

int tc_backward(nerf_ctx*, int, const float*, const float*, const float*, int64_t, int, const float*, cudaStream_t) {
    return fail(NERF_ERR_STATE, "tc_backward: not built yet");
}
}  // namespace nerf
