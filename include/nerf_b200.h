/*
 * nerf_b200.h -- C-ABI of libnerf_b200.so: the B200 (sm_100a) NeRF ray-render / train hot path.
 *
 * Drop-in boundary for the Python call surface of ghif/nerf-keras (the reference has no FFI layer;
 * its boundary is `data_utils.py` + `models.py`).  Each entry point cites the reference interface
 * it replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes binding a
 * reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; nerf_last_error() gives the message
 *     (thread-local).  The library never throws, aborts or silently falls back to the CPU.
 *   - all tensor pointers are DEVICE pointers to contiguous row-major float32 unless noted; the
 *     caller owns them (borrowed for the call, never retained).  `stream` is a cudaStream_t passed
 *     as void*; all work is enqueued on it and no call synchronises the device unless it says so.
 *   - one nerf_ctx per GPU/trainer; a ctx is not thread-safe, distinct ctxs are independent.
 */
#ifndef NERF_B200_H
#define NERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NERF_OK 0
#define NERF_ERR_INVALID (-1)   /* bad argument / unsupported configuration */
#define NERF_ERR_CUDA (-2)      /* CUDA runtime error (message has the cudaError string) */
#define NERF_ERR_STATE (-3)     /* call made in the wrong state (e.g. no weights set) */

#define NERF_NET_COARSE 0
#define NERF_NET_FINE 1

#define NERF_PRECISION_BF16_TC 0 /* tcgen05 bf16-in / fp32-accumulate fused MLP (product path)     */
#define NERF_PRECISION_FP32 1    /* fp32 CUDA-core layer-by-layer MLP (tight-parity / debug path)  */

/* Mirrors the JSON config schema read at train_lego.py:30-50 (BATCH_SIZE, NS_COARSE, NS_FINE,
 * L_XYZ, L_DIR, NUM_LAYERS, HIDDEN_DIM, SKIP_LAYER, LEARNING_RATE, BATCH_NORM). */
typedef struct nerf_config {
    int32_t num_layers;  /* NUM_LAYERS  (8)   */
    int32_t hidden_dim;  /* HIDDEN_DIM  (256) */
    int32_t skip_layer;  /* SKIP_LAYER  (4)   */
    int32_t l_xyz;       /* L_XYZ       (10)  */
    int32_t l_dir;       /* L_DIR       (4)   */
    int32_t ns_coarse;   /* NS_COARSE         */
    int32_t ns_fine;     /* NS_FINE; 0 = single-net benchmark shape (coarse net only, SURVEY 8(d) sweep) */
    int32_t max_rays;    /* largest ray batch any call will pass (workspace is sized for it) */
    int32_t batch_norm;  /* BATCH_NORM: must be 0 here; the host mirror folds BN inference into W, b   */
    int32_t training;    /* 1: allocate gradient / Adam / saved-activation storage            */
    float learning_rate; /* LEARNING_RATE (Adam, Keras defaults b1 .9 b2 .999 eps 1e-7)       */
    int32_t stop_grad_samples; /* 1: no gradient through the fine sample positions; 0: reference semantics */
} nerf_config;

typedef struct nerf_ctx nerf_ctx;

const char* nerf_last_error(void);
int nerf_version(void);
/* floats in one net's flat weight blob: roles d0..d{L-1}, sigma, feature, ddir, rgb; per role the
 * Keras Dense kernel W (in,out) row-major, then the bias (models.py:24-62).  595844 for 8x256. */
int64_t nerf_param_count(const nerf_config* cfg);

/* ---- context ---------------------------------------------------------------------------------- */
int nerf_create(const nerf_config* cfg, nerf_ctx** out);
int nerf_destroy(nerf_ctx* ctx);
/* replaces keras Model.set_weights / load_weights (inference.py:170); blob is a device pointer */
int nerf_set_weights(nerf_ctx* ctx, int net, const float* blob, int64_t n, void* stream);
int nerf_get_weights(nerf_ctx* ctx, int net, float* blob, int64_t n, void* stream);
/* flat gradient buffer [coarse | fine] (2 * param_count floats), device pointer owned by the ctx;
 * the data-parallel trainer all-reduces it in place (models.py:107 under train_tpu_lego.py:127). */
int nerf_grad_buffer(nerf_ctx* ctx, float** grads, int64_t* n);

/* ---- data_utils.py ------------------------------------------------------------------------------ */
/* get_rays(height,width,focal,pose) -> (ray_origins, ray_directions)   data_utils.py:23-52.
 * pose: 12 HOST floats = rows of pose[:3,:4].  o, d: (H,W,3).  Bit-exact op order. */
int nerf_get_rays(int height, int width, float focal, const float* pose3x4_host, float* o, float* d, void* stream);
/* EXTENSION: original-NeRF NDC transform of n rays, in place allowed. */
int nerf_ndc_rays(int height, int width, float focal, float near_plane, const float* o_in, const float* d_in,
                  float* o_out, float* d_out, int64_t n, void* stream);
/* generate_t_vals(near,far,batch_size,num_samples,rand_sampling)      data_utils.py:119-138.
 * u: NULL (rand_sampling=False) or device uniforms; u_per_ray=0: shape (N,) shared by all rays
 * (the reference, Q1); 1: shape (B,N).  t: (B,N). */
int nerf_generate_t_vals(double near_plane, double far_plane, int64_t batch, int num_samples, const float* u,
                         int u_per_ray, float* t, void* stream);
/* sample_rays(o,d,t) -> (rays (B,N,3), dirs (B,N,3))                   data_utils.py:55-73 */
int nerf_sample_rays(const float* o, const float* d, const float* t, int64_t batch, int num_samples, float* rays,
                     float* dirs, void* stream);
/* encode_position(x, L): x (n,3) -> (n, 3+6L)                          data_utils.py:7-21 */
int nerf_encode_position(const float* x, int64_t n, int L, float* out, void* stream);
/* volume_render(preds (B,N,4), t (B,N)) -> rgb (B,3), depth (B), weights (B,N); acc (B) optional
 * extension (sum of weights), may be NULL.                             data_utils.py:75-98 */
int nerf_volume_render(const float* preds, const float* t, int64_t batch, int num_samples, float* rgb, float* depth,
                       float* weights, float* acc, void* stream);
/* sample_pdf(t_vals_mid (B,Nc-1), weights (B,Nc), ns_fine) with explicit uniforms u (B,ns_fine)
 * -> samples (B,ns_fine)                                               data_utils.py:172-223 */
int nerf_sample_pdf(const float* t_mid, const float* weights, const float* u, int64_t batch, int nc, int ns_fine,
                    float* samples, void* stream);
/* fused t_mid + sample_pdf + sort(concat([t, t_fine]))                 models.py:165-167.
 * t (B,Nc), weights (B,Nc), u (B,Nf) -> t_all (B,Nc+Nf) ascending.  src_idx (B,Nc+Nf) int32
 * optional: for each sorted slot, the index into concat([t, t_fine]) it came from. */
int nerf_resample_merge(const float* t, const float* weights, const float* u, int64_t batch, int nc, int nf,
                        float* t_all, int32_t* src_idx, void* stream);

/* ---- models.py ---------------------------------------------------------------------------------- */
/* The Keras model call `model([rays_enc, dirs_enc])` (models.py:24-62,157,173):
 * rays_enc (n,3+6*l_xyz), dirs_enc (n,3+6*l_dir) -> preds (n,4) = [r,g,b,sigma] raw. fp32 path. */
int nerf_mlp_forward_encoded(nerf_ctx* ctx, int net, const float* rays_enc, const float* dirs_enc, int64_t n,
                             float* preds, void* stream);
/* Fused sample_rays + encode_position x2 + model (models.py:152-157 / 169-173) for one net:
 * o,d (B,3), t (B,N) -> preds (B,N,4).  precision selects the tcgen05 or the fp32 path. */
int nerf_mlp_forward_rays(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t, int64_t batch,
                          int num_samples, int precision, float* preds, void* stream);
/* NeRFTrainer.forward_pass (models.py:151-176), inference mode.  u_pdf (B,Nf) explicit uniforms, or NULL: the
 * draws of tf.random.uniform (data_utils.py:196) are generated inside the resampling kernel (Philox4x32-10 keyed by
 * nerf_set_seed, a per-call counter, the ray and the draw index -- no (B,Nf) buffer in HBM).
 * Outputs (any may be NULL): rgb_c/f (B,3), depth_c/f (B), w_c (B,Nc), w_f (B,Nc+Nf),
 * pred_c (B,Nc,4), pred_f (B,Nc+Nf,4), t_all (B,Nc+Nf). */
typedef struct nerf_forward_out {
    float *rgb_c, *rgb_f, *depth_c, *depth_f, *w_c, *w_f, *pred_c, *pred_f, *t_all, *acc_c, *acc_f;
} nerf_forward_out;
int nerf_forward_pass(nerf_ctx* ctx, const float* o, const float* d, const float* t, const float* u_pdf,
                      int64_t batch, int precision, const nerf_forward_out* out, void* stream);
/* NeRFTrainer.train_step forward+backward (models.py:88-106): fills the ctx gradient buffer with
 * d(MSE(rgb_c)+MSE(rgb_f))/d(weights) for the LOCAL batch and writes metrics[3] (device floats):
 * loss_coarse, loss (fine), psnr (models.py:110-120).  images (B,3). */
int nerf_train_forward_backward(nerf_ctx* ctx, const float* images, const float* o, const float* d, const float* t,
                                const float* u_pdf, int64_t batch, float* metrics_dev, void* stream);
/* The same step in two halves so that a data-parallel caller can all-reduce the fine net's half of the gradient buffer
 * while the coarse net's backward still runs: phases bit 0 = forward of both nets, loss, metrics and the
 * fine net's backward (fills grads[n_params, 2 n_params)); bit 1 = the coarse net's backward, including the term through
 * the fine sample positions (fills grads[0, n_params)).  phases = 3 is nerf_train_forward_backward.  u_pdf may be NULL
 * (in-kernel Philox draws keyed by the optimiser step).  No host synchronisation; safe to capture in a CUDA graph. */
#define NERF_PHASE_FORWARD_FINE 1
#define NERF_PHASE_COARSE 2
int nerf_train_phases(nerf_ctx* ctx, const float* images, const float* o, const float* d, const float* t,
                      const float* u_pdf, int64_t batch, float* metrics_dev, int phases, void* stream);
/* Rendering option (default off).  The reference sets delta = 1e10 on a ray's LAST sample (data_utils.py:82), so the ray's
 * colour is discontinuous in that sample's raw sigma at 0 (alpha jumps 0 -> 1).  With this option the bf16 tensor-core
 * forward re-evaluates the last sample of every ray with the fp32 CUDA-core MLP and uses its sigma, so the sign decision
 * is the fp32 one (one extra fp32 MLP evaluation per ray and net).  Inference passes only. */
int nerf_set_exact_far_sigma(nerf_ctx* ctx, int on);
/* Training option (default 0 = off).  wgrad_sms > 0: during the backward of a net the weight-gradient kernel runs on
 * `wgrad_sms` SMs NEXT TO the dX-chain kernel (which keeps the others) and consumes every tile's dZ images as the chain
 * publishes them, instead of after it on all SMs.  Same results (different summation order of the split-K reduction);
 * measured slower on B200 at every split, kept as the built form of "dW accumulated where dZ is produced".
 * Accepts 13 .. SMs/2; applies to nets with at least one tile pair per SM. */
int nerf_set_backward_overlap(nerf_ctx* ctx, int wgrad_sms);
/* Seed of the in-kernel uniform draws (keras.utils.set_random_seed, train_lego.py:22). */
int nerf_set_seed(nerf_ctx* ctx, uint64_t seed);
/* LEARNING_RATE lives in device memory (the step may be replayed from a CUDA graph): change it between steps. */
int nerf_set_learning_rate(nerf_ctx* ctx, float learning_rate, void* stream);
/* Optimiser state [coarse | fine] (2 * param_count floats each) and the number of applied updates: read / restore it when a
 * trainer moves to a context with a larger workspace.  Device pointers; `step` is a host value. */
int nerf_get_optimizer_state(nerf_ctx* ctx, float* m, float* v, int64_t* step, void* stream);
int nerf_set_optimizer_state(nerf_ctx* ctx, const float* m, const float* v, int64_t step, void* stream);
/* Running sums kept by the context for keras.metrics.Mean (models.py:113-120): device float[4] = sum of loss_coarse,
 * sum of loss (fine), sum of psnr, number of steps; every train phase-0 call and nerf_metrics_accumulate add to it. */
int nerf_metric_sums(nerf_ctx* ctx, float** sums_dev);
int nerf_metrics_accumulate(nerf_ctx* ctx, const float* images, const float* rgb_c, const float* rgb_f, int64_t batch,
                            float* metrics_dev, void* stream);
/* keras.optimizers.Adam.apply_gradients (train_lego.py:149-151, models.py:107) on the ctx gradient
 * buffer scaled by grad_scale (1/world_size after a sum all-reduce); bumps the step count. */
int nerf_adam_step(nerf_ctx* ctx, float grad_scale, void* stream);
/* ---- BATCH_NORM=true training (models.py:30-33, 49-52 with training=True) --------------------------------------------
 * Batch statistics couple all samples of a batch between consecutive layers, so this is a separate, layer-by-layer path on
 * fp32 activations: tcgen05 GEMMs on split bf16 operands (hi + residual, fp32 accumulation: results to ~1e-5 of the
 * products' scale; csrc/gemm_tc.cu, no library GEMM) between hand-written statistics / normalise / backward kernels.  It
 * serves the reference's small-batch BN configs and does not use a nerf_ctx: every buffer is the caller's.
 *   params   [coarse | fine], nerf_param_count() floats each: Dense kernels and biases, un-folded
 *   bn       [coarse | fine] x [gamma | beta | moving_mean | moving_variance], nerf_bn_param_count() floats each, layers in
 *            the order d0..d(L-1), ddir; the moving statistics are updated in place (momentum 0.99, epsilon 1e-3)
 *   grads    like params, overwritten;  bn_grads [coarse | fine] x [dgamma | dbeta], overwritten
 *   metrics  3 device floats: loss_coarse, loss (fine), psnr
 * Gradient semantics: stop-gradient on the fine sample positions. */
int64_t nerf_bn_param_count(const nerf_config* cfg);
int64_t nerf_bn_workspace_bytes(const nerf_config* cfg, int64_t batch);
int nerf_bn_forward_backward(const nerf_config* cfg, const float* params, float* bn, const float* images, const float* o,
                             const float* d, const float* t, const float* u_pdf, int64_t batch, float* grads, float* bn_grads,
                             float* metrics, void* workspace, int64_t workspace_bytes, void* stream);
/* The Adam update of nerf_adam_step on caller-owned flat buffers (step counts from 1). */
int nerf_adam_flat(float* params, const float* grads, float* m, float* v, int64_t n, int64_t step, float learning_rate,
                   float grad_scale, void* stream);
/* NeRFTrainer.test_step metrics (models.py:122-145): mse_c, mse_f, psnr from rgb_c, rgb_f, images. */
int nerf_metrics(const float* images, const float* rgb_c, const float* rgb_f, int64_t batch, float* metrics_dev,
                 void* stream);

/* ---- diagnostics -------------------------------------------------------------------------------- */
/* Number of kernel launches this library has issued since load (bench.py's gpu_launches). */
int64_t nerf_launch_count(void);
/* Per-kernel device timing for bench.py's roofline leg: when enabled, every launch of kernel class
 * `kind` (0 = fused MLP forward, 1 = fused MLP backward chain, 2 = weight-gradient GEMM) is bracketed
 * by CUDA events on its stream.  nerf_timing_read sums and clears them (synchronises on the events). */
int nerf_timing_enable(int on);
int nerf_timing_read(int kind, double* total_ms, int64_t* launches);
/* Gradient / draw diagnostics used by the tests, the tcgen05 building-block self-tests, MMA issue-rate probes and the selector
 * of the experimental forward-kernel variants are declared in nerf_b200_debug.h (not part of the drop-in surface). */

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_H */
