// Definitions shared by the product forward kernel (mlp_tc.cu) and the experimental forward variants
// (mlp_tc_experimental.cu): the forward chunk program, the kernel parameter block, the ReLU sign-mask helpers.
#pragma once
#include "mlp_tc_common.cuh"

namespace tcmlp {

struct FwdParams {
    const float* o;
    const float* d;
    const float* t;
    int N;                 // samples per ray
    int64_t M;             // total samples = rays * N
    int64_t n_pairs;       // ceil(M / 256)
    const __nv_bfloat16* w_chunks;
    const float* side;
    const float* dirbias;  // (rays, 128)
    float4* preds;         // (M) [r,g,b,sigma] raw
    uint8_t* act_save;     // optional saved-activation images (training)
    uint32_t* mask_save;   // optional ReLU masks (training)
    long long* trace;      // optional timeline trace buffer (diagnostics)
    int dbg;               // experiments: bit0 = no K-half early start in pair mode
};

}  // namespace tcmlp

namespace {

using namespace tcmlp;

// forward program: L0, L1, L2, L3, L4, L5a (h part), L5b (skip part), L6, L7, feature, ddir.
// L0 / L5b consume the encoding as TWO k-blocks: [bf16(enc) | bf16 residual of the raw xyz channels]
// (the raw coordinates reach |x| ~ 6, where a single bf16 loses ~1e-2; the hi+lo split restores them).
constexpr int N_PHASES = 11;
constexpr int N_CHUNKS = 76;
__constant__ Program c_fwd_prog = {
    N_PHASES, N_CHUNKS,
    {4, 8, 8, 8, 8, 8, 4, 8, 8, 8, 4, 0},
    {2, 4, 4, 4, 4, 4, 2, 4, 4, 4, 4, 0},
    {PH_ENC, 0, 0, 0, 0, 0, PH_ENC | PH_ACC, 0, 0, 0, 0, 0}};

// ReLU sign masks for the backward pass, one funnel shift per element: `neg` collects the fp32 sign bits of the
// pre-activations in arrival order (element 0 ends up in bit 31); the mask is the bit-reversed complement, i.e.
// bit q = (x_q >= +0).  (x == +0.0 exactly counts as active; TF's ReLU'(0) = 0 differs only on that measure-zero set.)
__device__ __forceinline__ uint32_t push_signs(uint32_t neg, float x0, float x1, float x2, float x3) {
    neg = __funnelshift_l(__float_as_uint(x0), neg, 1);
    neg = __funnelshift_l(__float_as_uint(x1), neg, 1);
    neg = __funnelshift_l(__float_as_uint(x2), neg, 1);
    neg = __funnelshift_l(__float_as_uint(x3), neg, 1);
    return neg;
}
__device__ __forceinline__ uint32_t signs_to_mask(uint32_t neg) { return ~__brev(neg); }

}  // namespace
