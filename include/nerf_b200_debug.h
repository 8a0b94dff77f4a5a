/*
 * nerf_b200_debug.h -- experimental kernel variants, tcgen05 self-tests and probes of libnerf_b200.so.
 * NOT part of the drop-in boundary (include/nerf_b200.h); used by tests/ and tools/ only.
 */
#ifndef NERF_B200_DEBUG_H
#define NERF_B200_DEBUG_H

#include "nerf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Diagnostics: d(sum(preds * d_preds))/d(weights of `net`) for given rays and t-values (forward with
 * saved activations + the tcgen05 backward); preds (B,N,4) is also returned.  Gradients land in the
 * ctx gradient buffer (nerf_grad_buffer), the other net's half is zero. */
int nerf_debug_mlp_grads(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t, int64_t batch,
                         int num_samples, const float* d_preds, float* preds, void* stream);
/* Diagnostics for the un-stopped sample-position gradient (quirk Q5): after nerf_debug_mlp_grads(net, ...),
 * dtp (B,N) = < d_ray, d(sum(preds*d_preds))/d pts >; and the backward of sort(concat([t, sample_pdf])) alone:
 * d_w (B,nc) from dL/dt_all given as a direct part dtp (B,Na) and/or dL/d(delta) of the fine compositing. */
int nerf_debug_input_grad(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t, int64_t batch,
                          int num_samples, float* dtp, void* stream);
/* The same quantity as left behind by the fused path of the weight-gradient kernel (what the training step uses)
 * during the last nerf_debug_mlp_grads(NERF_NET_FINE, ...): device-to-device copy of batch * num_samples floats. */
int nerf_debug_fused_input_grad(nerf_ctx* ctx, int64_t batch, int num_samples, float* dtp, void* stream);
int nerf_sample_pdf_bwd(const float* t, const float* weights, const float* u, const int32_t* src_idx,
                        const float* dtp, const float* d_delta, int64_t batch, int nc, int nf, float* d_w, void* stream);
/* Test hook: out (batch, nf) = the uniforms the in-kernel generator gives sample_pdf for (seed, counter); a training step
 * uses counter = number of optimiser updates applied so far, an inference pass 2^62 + its call index since nerf_set_seed. */
int nerf_debug_pdf_draws(uint64_t seed, uint64_t counter, int64_t batch, int nf, float* out, void* stream);
/* Experiments on the backward.  Timing only (results wrong by construction): bit0 / bit3 skip the CUDA-core side
 * jobs, bit1 skip the MMAs, bit2 skip the final reduction of the weight-gradient kernel.  Scheduling (results
 * unchanged): bit6 no overlap (weight gradient after the dX chain, each on all SMs), bit7 weight gradient after the
 * chain but on the overlap budget of SMs, bits 8.. = that budget (0: the context's default). */
int nerf_debug_flags(int flags);
/* Test hook: the fp32 GEMM of the BATCH_NORM training path (csrc/gemm_tc.cu: tcgen05 MMAs on split bf16 operands).
 * Row-major C (M x N) = op(A) op(B) + beta C on device buffers; ta: A stored (K x M), result ADDED to C (beta = 1);
 * precise: three-way operand split (fp32-grade; what the forward GEMMs use) instead of two-way. */
int nerf_selftest_gemm_f32(int ta, int tb, int64_t M, int N, int K, const float* A, int64_t lda, const float* B,
                           int64_t ldb, float beta, float* C, int64_t ldc, int precise, void* stream);
/* Diagnostics: per weight-gradient CTA eight int64 {job, tiles, end of CTA (globaltimer ns), ns its loader waited for
 * the dX chain's progress counters, ns it waited for free ring slots, end of its loader (ns), start of the CTA (ns), 0}; device buffer of
 * 2 nets x 148 x 8 int64, NULL disables. */
int nerf_debug_wgrad_stats(long long* dev_buf);
/* Diagnostics: timeline trace of CTA 0 of the fused forward kernel (device buffer of 768 int64 clock stamps,
 * NULL disables). */
int nerf_debug_trace(long long* dev_buf);
/* Self-test of the CTA-pair (cta_group::2) path: C (256,n) = A (256,k) x B (n,k)^T. */
int nerf_selftest_gemm_2cta(const float* a, const float* b, float* c, int n, int k, void* stream);
/* C = A * B^T with the A operand resident in tensor memory (tcgen05.st + TMEM-A MMAs); pair = 1: M = 256 over a CTA pair.
   reps > 1 repeats the MMA sequence (issue-rate probe; C is then reps x the product), probe = 2 adds a tcgen05.commit
   per four MMAs, cycles_dev (optional) receives the SM cycles from first issue to completion */
int nerf_selftest_gemm_ts(const float* a, const float* b, float* c, int n, int k, int pair, int reps, int probe,
                          long long* cycles_dev, void* stream);
/* Selects an experimental variant of the fused forward kernel (same results within bf16 rounding; slower than the default,
   kept for the measurements in DESIGN.md): 0 = default single-CTA kernel, 1 = CTA pair (cta_group::2, M = 256 MMAs over an
   SM pair), 17 = CTA pair with every weight chunk staged once for both sub-tiles, 49 = 17 with a CTA-scope proxy fence,
   4 = CTA pair with the activations resident in tensor memory (TMEM-A "TS" MMAs), 64 = single CTA with weight-stationary
   MMA pairs (tcgen05.mma.ws, B kept in the collector) from one issuer warp, 192 = the same issue order with plain MMAs,
   320 = weight-stationary pairs from two issuer warps (one per N-half, collector buffers b0 / b1). */
int nerf_debug_pair_mode(int on);
/* MMA issue-rate probe (cycles for `reps` x 4 back-to-back 128 x n x 16 MMAs). */
int nerf_selftest_mma_rate(int n, int reps, int mode, long long* cycles_dev, void* stream);
/* Self-test of the tcgen05 building block: C (M,N) fp32 = A (M,K) bf16-rounded x B^T, B (N,K). */
int nerf_selftest_gemm(const float* a, const float* b, float* c, int m, int n, int k, int mode, void* stream);

/* Probe of the tcgen05 collector variants: two A tiles against one B tile.  variant 0 plain, 1 tcgen05.mma.ws with B kept
   in the collector for the second MMA, 2 A kept across the two N-halves (n = 256).  C1 = A1 B^T, C2 = A2 B^T, (128, n). */
int nerf_selftest_collector(const float* a1, const float* a2, const float* b, float* c1, float* c2, int n, int k, int variant,
                            int reps, long long* cycles_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_DEBUG_H */
