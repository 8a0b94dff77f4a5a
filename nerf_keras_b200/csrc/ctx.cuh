// nerf_ctx: per-GPU state of the B200 NeRF path.  The library owns weights (fp32 master + packed bf16
// operand images), gradients, Adam moments and all workspaces; callers own every I/O buffer.
#pragma once
#include <cuda_bf16.h>

#include <vector>

#include "common.cuh"
#include "rng.cuh"

namespace nerf {

struct LayerInfo {
    int fan_in, fan_out;
    int64_t w_off, b_off;  // float offsets inside one net's flat blob
};

}  // namespace nerf

// Step state that kernels read from device memory, so that a captured CUDA graph of the training step stays valid
// from one replay to the next (kernel arguments are frozen at capture time).
struct nerf_dev_state {
    unsigned long long step;   // optimiser updates applied so far (Adam's t - 1); also keys the in-kernel uniform draws
    unsigned long long seed;   // nerf_set_seed
    float lr;                  // LEARNING_RATE
    float pad;
};

struct nerf_ctx {
    nerf_config cfg;
    int device = 0;
    std::vector<nerf::LayerInfo> layers;  // d0..d{L-1}, sigma, feature, ddir, rgb
    int64_t n_params = 0;                 // per net

    // fp32 master state, [coarse | fine]
    float* params = nullptr;
    float* grads = nullptr;
    float* adam_m = nullptr;
    float* adam_v = nullptr;
    int64_t adam_step = 0;                // host mirror of dev_state->step
    nerf_dev_state* dev_state = nullptr;
    uint64_t seed = 0;
    bool exact_far_sigma = false;         // rendering: last sample's sigma from the fp32 path (nerf_set_exact_far_sigma)
    float *far_t = nullptr, *far_pred = nullptr;   // (max_rays), (max_rays, 4): workspace of that option
    uint64_t render_draws = 0;            // counter keying the in-kernel draws of inference forward passes
    float* metric_sums = nullptr;         // device float[4]: running sums of loss_coarse, loss, psnr + step count
    bool weights_set[2] = {false, false};

    // tcgen05 operand images (bf16, 128B-swizzled 16 KB chunks) + fp32 side tables, per net
    __nv_bfloat16* w_fwd[2] = {nullptr, nullptr};  // forward chunk stream  (72 chunks)
    __nv_bfloat16* w_bwd[2] = {nullptr, nullptr};  // backward (transposed) chunk stream
    float* side[2] = {nullptr, nullptr};           // biases + fp32 head weights (see mlp_tc.cu)
    bool packed_valid[2] = {false, false};
    bool bwd_packed_valid = false;                 // w_bwd[0], w_bwd[1] and w_ig match the current weights

    // fp32 path workspace (allocated on first use)
    int64_t fp32_chunk = 0;
    float *ws_a = nullptr, *ws_b = nullptr, *ws_c = nullptr, *ws_encx = nullptr, *ws_encd = nullptr;

    // forward_pass / train workspace, sized for cfg.max_rays
    float *fw_pred_c = nullptr, *fw_pred_f = nullptr, *fw_w_c = nullptr, *fw_w_f = nullptr, *fw_t_all = nullptr;
    float *fw_rgb_c = nullptr, *fw_rgb_f = nullptr;
    float* fw_dirbias[2] = {nullptr, nullptr};     // per-ray ddir bias of each net
    int32_t* fw_src_idx = nullptr;

    // training workspace (cfg.training): saved activation images and gradient images, per net
    __nv_bfloat16* act_save[2] = {nullptr, nullptr};
    __nv_bfloat16* dz_save[2] = {nullptr, nullptr};
    uint32_t* mask_save[2] = {nullptr, nullptr};
    float *tr_dpred_c = nullptr, *tr_dpred_f = nullptr, *tr_drgb_c = nullptr, *tr_drgb_f = nullptr;
    float* tr_ddirsum[2] = {nullptr, nullptr};     // (max_rays, 128) per net: per-ray sums of dZ of the ddir layer (chain kernel)
    // un-stopped gradient through the fine sample positions (stop_grad_samples = 0)
    __nv_bfloat16* w_ig = nullptr;                 // W0^T / W5b^T operand image of the input-gradient kernel
    float *tr_ddelta_f = nullptr, *tr_dtp_f = nullptr, *tr_dw_extra = nullptr;
    // backward overlap: the weight-gradient kernel runs on `wgrad_ctas` SMs NEXT TO the dX chain (side stream, forked and
    // joined by events so that it is captured into the step graph), consuming each tile's dZ images as the chain
    // publishes them (chain_progress, one counter per tile).  0 = one kernel after the other.
    int wgrad_ctas = 0;
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    uint32_t* chain_progress[2] = {nullptr, nullptr};
    int64_t progress_tiles[2] = {0, 0};
};

namespace nerf {

int ensure_fp32_workspace(nerf_ctx* ctx);
int mlp_fp32_forward_encoded(nerf_ctx* ctx, int net, const float* enc_x, const float* enc_d, int64_t n, float* preds,
                             cudaStream_t st);
int mlp_fp32_forward_rays(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t, int64_t B, int N,
                          float* preds, cudaStream_t st);

// mlp_tc.cu
int tc_supported(const nerf_config& cfg, std::string* why);
int tc_alloc(nerf_ctx* ctx);
void tc_free(nerf_ctx* ctx);
int tc_pack_weights(nerf_ctx* ctx, int net, cudaStream_t st);
int tc_pack_all(nerf_ctx* ctx, bool tick_step, cudaStream_t st);     // forward images of both nets (+ step counter tick)
int tc_pack_backward(nerf_ctx* ctx, cudaStream_t st);                // transposed images of both nets + input-gradient image
int tc_dirbias(nerf_ctx* ctx, const float* d, int64_t B, int nets_mask, cudaStream_t st);
bool stream_is_capturing(cudaStream_t st);
int tc_input_grad(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t, int64_t B, int N, float* dtp,
                  cudaStream_t st);
int sample_pdf_backward(const float* t, const float* weights, const PdfDraws& dr, const int32_t* src_idx, const float* dtp,
                        const float* d_delta, int64_t B, int nc, int nf, float* d_w, cudaStream_t st);
int tc_forward_rays(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t, int64_t B, int N,
                    float* preds, bool save_acts, cudaStream_t st, bool dirbias_ready = false);
int resample_merge(const float* t, const float* weights, PdfDraws dr, int64_t B, int nc, int nf, float* t_all,
                   int32_t* src_idx, cudaStream_t st);

}  // namespace nerf
