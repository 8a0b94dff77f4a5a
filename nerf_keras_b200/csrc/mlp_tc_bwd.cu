// Backward of the fused tcgen05 MLP (models.py:94-106 via tf.GradientTape in the reference; SURVEY
// Appendix A is the hand-derived target).  Three kinds of kernels, all sm_100a:
//
//  1. nerf_mlp_bwd_tc_kernel -- the dX chain.  Same skeleton as the forward kernel (two 128-row
//     sub-tiles per CTA, shared weight ring, TMEM accumulators), streaming the TRANSPOSED weights:
//     rgb head and ddir ReLU on CUDA cores -> dZ_ddir; then dZ_ddir x Wddir^T, x Wfeat^T (+ sigma
//     head), x W7^T ... x W1^T, each epilogue masking with the forward ReLU sign bits and writing the
//     bf16 dZ tile both as the next A operand and (bulk store) as an image for the weight-gradient
//     kernel.
//  2. nerf_wgrad_tc_kernel -- dW = X^T dZ per layer as a split-K tcgen05 GEMM over the saved activation
//     and dZ images.  The images are exactly the 128B-swizzled operand tiles, so a [samples x features]
//     tile is consumed as an MN-major operand (features along M/N, samples along K) without any
//     transpose.  Accumulators stay in TMEM for a CTA's whole slab of samples; bias gradients (column
//     sums of dZ) and the tiny sigma / rgb head gradients are CUDA-core side jobs on the same tiles.
//     The sigma / rgb heads and the direction rows of Wddir are three more jobs of the same kernel: the
//     chain kernel emits a [d rgb, d sigma] "head" dZ image, and
//  3. dir_image_kernel writes the per-sample direction-encoding images (the forward kernel hoists that
//     product into a per-ray bias, so the operand has to be materialised for the backward).
#include "common.cuh"
#include "ctx.cuh"
#include "mlp_tc_common.cuh"

using namespace nerf;
using namespace tc5;
using namespace tcmlp;

namespace {

// ------------------------------------------------------------------------------------------------
// 1. dX chain
// ------------------------------------------------------------------------------------------------
// phases: P0 dZ_ddir x Wddir[0:256]^T (K=128) ; P1 dFeat x Wfeat^T ; P2..P8 dZ_l x W_l^T for l = 7..1
constexpr int B_PHASES = 9;
constexpr int B_CHUNKS = 68;
__constant__ Program c_bwd_prog = {
    B_PHASES, B_CHUNKS,
    {4, 8, 8, 8, 8, 8, 8, 8, 8, 0, 0, 0},
    {2, 4, 4, 4, 4, 4, 4, 4, 4, 0, 0, 0},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}};

// B[n][k] = W_phase[n][k]  (weights in their natural (in, out) layout are K-major for dX = dZ x W^T)
__device__ __forceinline__ float bwd_weight_at(const float* __restrict__ blob, const BlobOffsets& off, int phase,
                                               int n, int k) {
    if (phase == 0) return blob[off.w[10] + (int64_t)n * (H / 2) + k];   // Wddir rows 0..255, 128 cols
    if (phase == 1) return blob[off.w[9] + (int64_t)n * H + k];          // Wfeature
    const int l = 9 - phase;                                             // 7..1
    return blob[off.w[l] + (int64_t)n * H + k];                          // rows 0..255 (h part for l = 5)
}

struct PackBwdArgs {
    const float* blob[2];
    __nv_bfloat16* chunks[2];
    const float* ig_blob;            // weights whose W0^T / W5b^T image is packed by grid row 2 (the fine net), or null
    __nv_bfloat16* ig_img;
};
__device__ void pack_ig_rows(const float* __restrict__ blob, const BlobOffsets& off, __nv_bfloat16* __restrict__ img);

__global__ void __launch_bounds__(256) pack_bwd_kernel(const PackBwdArgs A, BlobOffsets off) {
    if (blockIdx.y == 2) {
        if (A.ig_blob) pack_ig_rows(A.ig_blob, off, A.ig_img);
        return;
    }
    const float* __restrict__ blob = A.blob[blockIdx.y];
    __nv_bfloat16* __restrict__ chunks = A.chunks[blockIdx.y];
    if (!blob) return;
    const int total = B_CHUNKS * 128 * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int g = i / (128 * 8);
        int rem = i - g * 128 * 8;
        int n = rem >> 3, kg = rem & 7;
        int phase, j;
        if (g < 4) { phase = 0; j = g; } else { phase = 1 + (g - 4) / 8; j = (g - 4) % 8; }
        int kbs = c_bwd_prog.kb[phase];
        int h = j / kbs, kb = j - h * kbs;
        uint32_t packed[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int k0 = kb * 64 + kg * 8 + 2 * q;
            packed[q] = pack_bf16x2(bwd_weight_at(blob, off, phase, h * 128 + n, k0),
                                    bwd_weight_at(blob, off, phase, h * 128 + n, k0 + 1));
        }
        uint8_t* dst = reinterpret_cast<uint8_t*>(chunks) + (size_t)g * CHUNK_BYTES + sw128_offset(n, kg * 8);
        *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
}

struct BwdParams {
    const float4* dpreds;   // (M) dL/d[r,g,b,sigma] raw
    int64_t M;
    int64_t n_pairs;
    const __nv_bfloat16* w_chunks;
    const float* side;
    const uint32_t* mask_save;
    uint8_t* dz_save;
};

// one 32-column group of a chain epilogue: optional sigma-head term, ReLU mask, bf16; stored to the A tile or held
template <bool SIGMA, bool MASK, int CG, bool STORE>
__device__ __forceinline__ void chain_group(const uint32_t (&v)[32], float dsig, const float* wsig, uint32_t mk,
                                            const RowStore& rs, uint32_t* held) {
    uint32_t pk[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        float x0 = __uint_as_float(v[2 * q]), x1 = __uint_as_float(v[2 * q + 1]);
        if (SIGMA) {
            x0 = fmaf(dsig, wsig[CG * 32 + 2 * q], x0);
            x1 = fmaf(dsig, wsig[CG * 32 + 2 * q + 1], x1);
        }
        if (MASK) {
            x0 = ((mk >> (2 * q)) & 1u) ? x0 : 0.f;
            x1 = ((mk >> (2 * q + 1)) & 1u) ? x1 : 0.f;
        }
        pk[q] = cvt_bf16x2<false>(x0, x1);
    }
    if (STORE) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
            rs.store<CG / 2>((CG & 1) * 4 + c, pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
    } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) held[CG * 16 + q] = pk[q];
    }
}

// columns 0..127 while the second N-half's MMAs are in flight: results held in registers
template <bool SIGMA, bool MASK>
__device__ __forceinline__ void chain_part1(uint32_t t_lane, float dsig, const float* wsig, const uint32_t (&mask)[8],
                                            const RowStore& rs, uint32_t (&held)[64]) {
    uint32_t v[32];
    tmem_ld32(t_lane, v); tmem_ld_wait();
    chain_group<SIGMA, MASK, 0, false>(v, dsig, wsig, mask[0], rs, held);
    tmem_ld32(t_lane + 32, v); tmem_ld_wait();
    chain_group<SIGMA, MASK, 1, false>(v, dsig, wsig, mask[1], rs, held);
    tmem_ld32(t_lane + 64, v); tmem_ld_wait();
    chain_group<SIGMA, MASK, 2, false>(v, dsig, wsig, mask[2], rs, held);
    tmem_ld32(t_lane + 96, v); tmem_ld_wait();
    chain_group<SIGMA, MASK, 3, false>(v, dsig, wsig, mask[3], rs, held);
}
__device__ __forceinline__ void store_held(const RowStore& rs, const uint32_t (&held)[64]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) rs.store<0>(c, held[4 * c], held[4 * c + 1], held[4 * c + 2], held[4 * c + 3]);
#pragma unroll
    for (int c = 0; c < 8; ++c) rs.store<1>(c, held[32 + 4 * c], held[32 + 4 * c + 1], held[32 + 4 * c + 2], held[32 + 4 * c + 3]);
}
// columns 128..255 straight into K-blocks 2,3
template <bool SIGMA, bool MASK>
__device__ __forceinline__ void chain_part2(uint32_t t_lane, float dsig, const float* wsig, const uint32_t (&mask)[8],
                                            const RowStore& rs) {
    uint32_t va[32], vb[32];
    tmem_ld32(t_lane + 128, va);
    tmem_ld_wait();
    tmem_ld32(t_lane + 160, vb);
    chain_group<SIGMA, MASK, 4, true>(va, dsig, wsig, mask[4], rs, nullptr);
    tmem_ld_wait();
    tmem_ld32(t_lane + 192, va);
    chain_group<SIGMA, MASK, 5, true>(vb, dsig, wsig, mask[5], rs, nullptr);
    tmem_ld_wait();
    tmem_ld32(t_lane + 224, vb);
    chain_group<SIGMA, MASK, 6, true>(va, dsig, wsig, mask[6], rs, nullptr);
    tmem_ld_wait();
    chain_group<SIGMA, MASK, 7, true>(vb, dsig, wsig, mask[7], rs, nullptr);
}

__global__ void __launch_bounds__(NUM_THREADS, 1) nerf_mlp_bwd_tc_kernel(const BwdParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    Barriers B;
    init_barriers(base, B);
    float* side = reinterpret_cast<float*>(smem + SM_SIDE);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM);
    if (warp == 9) tmem_alloc(base + SM_TMEM, 512);
    for (int i = threadIdx.x; i < SIDE_FLOATS; i += NUM_THREADS) side[i] = P.side[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int my_pairs = (P.n_pairs > blockIdx.x) ? (int)((P.n_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    const int total_chunks = my_pairs * B_CHUNKS;

    if (warp == 8) {
        if (lane == 0) producer_loop(base, B, P.w_chunks, B_CHUNKS, total_chunks);
    } else if (warp >= 9) {
        if (lane == 0) issuer_loop(base, B, tmem_base, c_bwd_prog, warp - 9, my_pairs);
    } else {
        const int s = warp >> 2;
        const int row = threadIdx.x - s * TILE_M;
        const uint32_t act_base = base + SM_ACT + s * 65536;
        const uint32_t t_lane = tmem_base + (uint32_t(32 * (warp & 3)) << 16) + s * 256;
        const bool elected = (row == 0);
        RowStore rs;
        rs.init(act_base, row);
        const uint32_t bar_lo = B.actr + 16 * s, bar_hi = bar_lo + 8;    // A tile hand-off, K-halves
        const uint32_t bar_h0 = B.accf + 16 * s, bar_h1 = bar_h0 + 8;    // accumulator hand-off, N-halves
        uint32_t accf_par = 0;

        for (int it = 0; it < my_pairs; ++it) {
            const int64_t pair = blockIdx.x + (int64_t)it * gridDim.x;
            const int64_t tile = pair * 2 + s;
            const int64_t g_row = tile * TILE_M + row;
            const bool valid = g_row < P.M;
            const uint32_t* mask_tile = P.mask_save + tile * (MASK_TILE_BYTES / 4);
            uint8_t* dz_tile = P.dz_save + tile * DZ_TILE_BYTES;
            const float4 dp = valid ? P.dpreds[g_row] : make_float4(0.f, 0.f, 0.f, 0.f);

            // ---- prologue: rgb head backward + ddir ReLU mask -> dZ_ddir (128 wide) ----
            if (elected) bulk_wait_read0();
            named_bar_sync(1 + s, TILE_M);
            {
                const uint4 mk4 = *reinterpret_cast<const uint4*>(mask_tile + ((size_t)8 * 128 + row) * 8);
                const uint32_t mk[4] = {mk4.x, mk4.y, mk4.z, mk4.w};
#define NERF_DDIR_BWD_GROUP(CG)                                                                                       \
    {                                                                                                                 \
        uint32_t pk[16];                                                                                              \
        _Pragma("unroll") for (int q = 0; q < 16; ++q) {                                                              \
            const int c0 = CG * 32 + 2 * q;                                                                           \
            float x0 = dp.x * side[SIDE_WRGB + c0] + dp.y * side[SIDE_WRGB + 128 + c0] + dp.z * side[SIDE_WRGB + 256 + c0]; \
            float x1 = dp.x * side[SIDE_WRGB + c0 + 1] + dp.y * side[SIDE_WRGB + 128 + c0 + 1] +                      \
                       dp.z * side[SIDE_WRGB + 256 + c0 + 1];                                                         \
            x0 = ((mk[CG] >> (2 * q)) & 1u) ? x0 : 0.f;                                                               \
            x1 = ((mk[CG] >> (2 * q + 1)) & 1u) ? x1 : 0.f;                                                           \
            pk[q] = cvt_bf16x2<false>(x0, x1);                                                                        \
        }                                                                                                             \
        _Pragma("unroll") for (int c = 0; c < 4; ++c)                                                                 \
            rs.store<CG / 2>((CG & 1) * 4 + c, pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);               \
    }
                NERF_DDIR_BWD_GROUP(0)
                NERF_DDIR_BWD_GROUP(1)
                NERF_DDIR_BWD_GROUP(2)
                NERF_DDIR_BWD_GROUP(3)
#undef NERF_DDIR_BWD_GROUP
            }
            // head image (K-block 2 is free until the first epilogue): [d r, d g, d b, d sigma, 0 ...]
            rs.store<2>(0, pack_bf16x2(dp.x, dp.y), pack_bf16x2(dp.z, dp.w), 0u, 0u);
#pragma unroll
            for (int c = 1; c < 8; ++c) rs.store<2>(c, 0u, 0u, 0u, 0u);
            tc_fence_before();
            fence_proxy_async_smem();
            named_bar_sync(1 + s, TILE_M);
            if (elected) {
                bulk_s2g(dz_tile + DZ_DDIR, act_base, 32768);
                bulk_s2g(dz_tile + DZ_HEAD, act_base + 2 * 16384, 16384);
                bulk_commit();
            }
            mbar_arrive(bar_lo);
            mbar_arrive(bar_hi);

            for (int ph = 0; ph < B_PHASES; ++ph) {
                int64_t dst;
                uint32_t mask[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                int ml = -1;
                if (ph == 0) dst = DZ_FEAT;                          // feature layer is linear: no mask
                else {
                    ml = (ph == 1) ? 7 : (8 - ph);                   // ReLU output feeding this gradient: h8 .. h1
                    const uint4* mp = reinterpret_cast<const uint4*>(mask_tile + ((size_t)ml * 128 + row) * 8);
                    const uint4 m0 = mp[0], m1 = mp[1];
                    mask[0] = m0.x; mask[1] = m0.y; mask[2] = m0.z; mask[3] = m0.w;
                    mask[4] = m1.x; mask[5] = m1.y; mask[6] = m1.z; mask[7] = m1.w;
                    dst = DZ_Z + (int64_t)65536 * ml;
                }
                const float* wsig = side + SIDE_WSIG;
                uint32_t held[64];
                mbar_wait(bar_h0, accf_par, 2);
                tc_fence_after();
                if (ph == 0) chain_part1<false, false>(t_lane, 0.f, wsig, mask, rs, held);
                else if (ph == 1) chain_part1<true, true>(t_lane, dp.w, wsig, mask, rs, held);
                else chain_part1<false, true>(t_lane, 0.f, wsig, mask, rs, held);
                mbar_wait(bar_h1, accf_par, 6);                      // every MMA of the phase is complete
                accf_par ^= 1;
                tc_fence_after();
                if (elected) bulk_wait_read0();                      // previous image store has finished reading the tile
                named_bar_sync(1 + s, TILE_M);
                store_held(rs, held);
                tc_fence_before();
                fence_proxy_async_smem();
                if (ph < B_PHASES - 1) mbar_arrive(bar_lo);
                // first half of the dZ image (K-blocks 0,1) leaves now, under the second half of the epilogue: one 64 KB
                // burst of TMA shared-memory reads at the start of the next MMA phase costs more than two 32 KB ones
                named_bar_sync(1 + s, TILE_M);
                if (elected) { bulk_s2g(dz_tile + dst, act_base, 32768); bulk_commit(); }
                if (ph == 0) chain_part2<false, false>(t_lane, 0.f, wsig, mask, rs);
                else if (ph == 1) chain_part2<true, true>(t_lane, dp.w, wsig, mask, rs);
                else chain_part2<false, true>(t_lane, 0.f, wsig, mask, rs);
                tc_fence_before();
                fence_proxy_async_smem();
                named_bar_sync(1 + s, TILE_M);
                if (elected) { bulk_s2g(dz_tile + dst + 32768, act_base + 32768, 32768); bulk_commit(); }
                if (ph < B_PHASES - 1) mbar_arrive(bar_hi);
            }
        }
        if (elected) bulk_wait_all0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// 2. weight gradients: dW[k_in][n_out] = sum over samples  X[m][k_in] * dZ[m][n_out]
// ------------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 320;            // warps 0-3: side jobs + final reduction, 4: loader, 5: MMA issuer, 6-9: side jobs
constexpr int WG_SIDE_WARPS = 8;
constexpr int WG_STAGES = 3;
constexpr int WG_SLOT = 65536;             // A half-blocks (4 x 8 KB) + B half-blocks (4 x 8 KB)
constexpr int WG_SM_BAR = WG_STAGES * WG_SLOT;
constexpr int WG_SMEM = WG_SM_BAR + 128 + 1024;
constexpr int WG_NJOBS = 13;

struct WgJob {
    int64_t a_off;       // byte offset of the X image inside a saved-activation tile
    int64_t b_off;       // byte offset of the dZ image inside a dZ tile
    int n_a;             // 64-feature blocks of X: 1, 2 or 4
    int n_b;             // 64-feature blocks of dZ: 2 or 4
    int64_t w_dst;       // float offset (in the grads blob of this net) of dW row 0
    int ld;              // fan_out
    int rows;            // valid rows of dW produced by this job
    int col_lo, col_hi;  // valid columns [col_lo, col_hi) of the accumulator that are written to dW
    int64_t bias_dst;    // float offset of the bias gradient (column sums of dZ), or -1
    int64_t sig_dst;     // float offset of dW_sigma (256,1): column sums of the X image weighted by d sigma per sample, or -1
};
struct WgParams {
    int debug;              // bit0: skip side jobs, bit1: skip MMAs, bit2: skip final reduction (timing experiments)
    WgJob jobs[WG_NJOBS];
    int cta_first[WG_NJOBS + 1];   // 1-D grid: job j owns CTAs [cta_first[j], cta_first[j+1]) and splits its samples among them
    const uint8_t* act_save;
    const uint8_t* dz_save;
    const float4* dpreds;
    int64_t M;
    int64_t n_half_tiles;   // 2 * tiles
    float* grads;           // this net's gradient blob
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1) nerf_wgrad_tc_kernel(const WgParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int job_id = 0;
    while (job_id + 1 < WG_NJOBS && (int)blockIdx.x >= P.cta_first[job_id + 1]) ++job_id;
    const WgJob& J = P.jobs[job_id];
    const int job_ctas = P.cta_first[job_id + 1] - P.cta_first[job_id], job_cta = (int)blockIdx.x - P.cta_first[job_id];
    const uint32_t bar_full = base + WG_SM_BAR, bar_empty = bar_full + 8 * WG_STAGES, bar_done = bar_empty + 8 * WG_STAGES;
    const int STG = (P.debug & 16) ? 2 : WG_STAGES;     // experiment: ring depth 2 (16 instead of 24 bulk copies in flight)
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + WG_SM_BAR + 8 * (2 * WG_STAGES + 1));

    if (threadIdx.x == 0) {
        for (int i = 0; i < WG_STAGES; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1 + 32 * WG_SIDE_WARPS);   // tcgen05.commit + the side-job threads
        }
        mbar_init(bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(base + WG_SM_BAR + 8 * (2 * WG_STAGES + 1), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this CTA's slab of half-tiles (64 samples each)
    const int64_t per = (P.n_half_tiles + job_ctas - 1) / job_ctas;
    const int64_t ht0 = (int64_t)job_cta * per;
    const int64_t ht1 = (ht0 + per < P.n_half_tiles) ? ht0 + per : P.n_half_tiles;
    const int n_ht = (ht1 > ht0) ? (int)(ht1 - ht0) : 0;
    const int n_a_load = (J.n_a == 1) ? 2 : J.n_a;   // a single 64-feature block is loaded twice (M = 128 MMA)
    const int n_mh = (J.n_a == 4) ? 2 : 1;

    if (warp == 4) {
        // ===================== loader =====================
        if (lane == 0) {
            int slot = 0;
            uint32_t par = 1;
            for (int i = 0; i < n_ht; ++i) {
                const int64_t ht = ht0 + i;
                const int64_t tile = ht >> 1;
                const int half = (int)(ht & 1);
                mbar_wait(bar_empty + 8 * slot, par, 11);
                const uint32_t dst = base + slot * WG_SLOT;
                mbar_arrive_expect_tx(bar_full + 8 * slot, (uint32_t)(n_a_load + J.n_b) * 8192u);
                const uint8_t* a_src = P.act_save + tile * SAVE_TILE_BYTES + J.a_off + half * 8192;
                const uint8_t* b_src = P.dz_save + tile * DZ_TILE_BYTES + J.b_off + half * 8192;
                if (P.debug & 32) {
                    // TIMING EXPERIMENT ONLY (wrong operands): the same bytes as two large copies per stage
                    bulk_g2s(dst, a_src - half * 8192 + half * (n_a_load * 8192), n_a_load * 8192, bar_full + 8 * slot);
                    bulk_g2s(dst + 32768, b_src - half * 8192 + half * (J.n_b * 8192), J.n_b * 8192, bar_full + 8 * slot);
                } else {
                for (int b = 0; b < n_a_load; ++b)
                    bulk_g2s(dst + b * 8192, a_src + (J.n_a == 1 ? 0 : b) * 16384, 8192, bar_full + 8 * slot);
                for (int b = 0; b < J.n_b; ++b)
                    bulk_g2s(dst + 32768 + b * 8192, b_src + b * 16384, 8192, bar_full + 8 * slot);
                }
                if (++slot == STG) { slot = 0; par ^= 1; }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, 64 * J.n_b, 1, 1);
            int slot = 0;
            uint32_t par = 0;
            for (int i = 0; i < n_ht; ++i) {
                mbar_wait(bar_full + 8 * slot, par, 12);
                tc_fence_after();
                const uint32_t a0 = base + slot * WG_SLOT, b0 = a0 + 32768;
                for (int mh = 0; mh < ((P.debug & 2) ? 0 : n_mh); ++mh) {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const uint64_t ad = make_sdesc_sw128(a0 + mh * 16384 + k4 * 2048, 8192, 1024);
                        const uint64_t bd = make_sdesc_sw128(b0 + k4 * 2048, 8192, 1024);
                        mma_bf16_ss(tmem_base + mh * 256, ad, bd, idesc, (i > 0 || k4 > 0) ? 1u : 0u);
                    }
                }
                mma_commit(bar_empty + 8 * slot);
                if (++slot == STG) { slot = 0; par ^= 1; }
            }
            mma_commit(bar_done);
        }
    } else {
        // ===================== side jobs + final reduction =====================
        // Eight side warps (one single warp per scheduler is latency-bound: a dependent-issue chain of ~1.1 K instructions
        // per half-tile made the job's CTAs the stragglers); side warp sw takes rows [8 sw, 8 sw + 8) of each 64-sample
        // half-tile, lane <-> (64-column block, 16-byte chunk) = 8 columns.
        const int sw = (warp < 4) ? warp : warp - 2;
        const int cb = lane >> 3, ch = lane & 7;
        float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // bias: partial column sums (8 columns per lane)
        float sg[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // sigma head: column sums of X weighted by d sigma
        auto accum = [&](uint32_t img, float wgt, float (&acc)[8], int rr) {
            const uint32_t addr = img + cb * 8192 + sw * 1024 + rr * 128 + ((ch ^ rr) << 4);
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(addr));
            acc[0] = fmaf(wgt, __uint_as_float(w0 << 16), acc[0]); acc[1] = fmaf(wgt, __uint_as_float(w0 & 0xFFFF0000u), acc[1]);
            acc[2] = fmaf(wgt, __uint_as_float(w1 << 16), acc[2]); acc[3] = fmaf(wgt, __uint_as_float(w1 & 0xFFFF0000u), acc[3]);
            acc[4] = fmaf(wgt, __uint_as_float(w2 << 16), acc[4]); acc[5] = fmaf(wgt, __uint_as_float(w2 & 0xFFFF0000u), acc[5]);
            acc[6] = fmaf(wgt, __uint_as_float(w3 << 16), acc[6]); acc[7] = fmaf(wgt, __uint_as_float(w3 & 0xFFFF0000u), acc[7]);
        };
        int slot = 0;
        uint32_t par = 0;
        for (int i = 0; i < n_ht; ++i) {
            // d sigma of this warp's 8 samples: one load per lane, issued before the wait so that its latency is hidden
            float my_ds = 0.f;
            if (J.sig_dst >= 0 && lane < 8) my_ds = __ldg(&P.dpreds[(ht0 + i) * 64 + sw * 8 + lane].w);
            mbar_wait(bar_full + 8 * slot, par, 13);
            const uint32_t a0 = base + slot * WG_SLOT, b0 = a0 + 32768;
            if (J.sig_dst >= 0 && !(P.debug & (1 | 8))) {
                // dW_sigma[f] += sum over the half-tile's samples of h8[sample][f] * d sigma[sample]  (models.py:42): the
                // X image of this job IS h8, so the sigma head costs no extra HBM traffic (it used to be a job of its own)
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) accum(a0, __shfl_sync(0xffffffffu, my_ds, rr), sg, rr);
            }
            if (J.bias_dst >= 0 && !(P.debug & (1 | 8)) && cb < J.n_b) {
                // bias gradient: column sums of the dZ half-tile
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) accum(b0, 1.0f, cs, rr);
            }
            mbar_arrive(bar_empty + 8 * slot);
            if (++slot == STG) { slot = 0; par ^= 1; }
        }
        if (n_ht > 0) {
            if (J.bias_dst >= 0 && (lane >> 3) < J.n_b) {
#pragma unroll
                for (int q = 0; q < 8; ++q) atomicAdd(P.grads + J.bias_dst + (lane >> 3) * 64 + (lane & 7) * 8 + q, cs[q]);
            }
            if (J.sig_dst >= 0) {
#pragma unroll
                for (int q = 0; q < 8; ++q) atomicAdd(P.grads + J.sig_dst + (lane >> 3) * 64 + (lane & 7) * 8 + q, sg[q]);
            }
            if (n_mh > 0 && !(P.debug & 4) && warp < 4) {     // TMEM lane quarters belong to warps 0-3
                mbar_wait(bar_done, 0, 14);
                tc_fence_after();
                const int N = 64 * J.n_b;
                for (int mh = 0; mh < n_mh; ++mh) {
                    const int r = mh * 128 + 32 * warp + lane;   // dW row = input feature
                    for (int cg = 0; cg < N / 32; ++cg) {
                        uint32_t v[32];
                        tmem_ld32(tmem_base + (uint32_t(32 * warp) << 16) + mh * 256 + cg * 32, v);
                        tmem_ld_wait();
                        if (r < J.rows) {
                            float* dst = P.grads + J.w_dst + (int64_t)r * J.ld + cg * 32;
                            const bool full = (J.col_lo <= cg * 32) && (J.col_hi >= cg * 32 + 32);
                            if (full && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                                for (int q = 0; q < 8; ++q)
                                    red_add_v4(dst + 4 * q, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                               __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                            } else {   // partial column window, or a layer that is not 16-byte aligned in the blob
#pragma unroll
                                for (int q = 0; q < 32; ++q)
                                    if (cg * 32 + q >= J.col_lo && cg * 32 + q < J.col_hi) atomicAdd(dst + q, __uint_as_float(v[q]));
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// 3. per-sample direction-encoding images for the direction rows of Wddir (models.py:48-54): the forward
// kernel hoists that product into a per-ray bias, the backward wants it as a GEMM operand again.
// One warp per ray writes the ray's N rows ([27 values | zeros], 128 bytes each) into the tiles' images.
// ------------------------------------------------------------------------------------------------
struct DirImageArgs {
    int N[2];
    uint8_t* act_save[2];
};
__global__ void __launch_bounds__(128) dir_image_kernel(const float* __restrict__ d, int64_t rays, const DirImageArgs A) {
    const int N = A.N[blockIdx.y];
    uint8_t* __restrict__ act_save = A.act_save[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < rays; ray += warps_total) {
        // lane <-> channel pair (2 lane, 2 lane + 1) of the 64-wide row
        float v[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int c = 2 * lane + q;
            float x = 0.f;
            if (c < 3) x = d[ray * 3 + c];
            else if (c < ENC_D) {
                int qq = c - 3, oct = qq / 6, r6 = qq - oct * 6, comp = r6 % 3;
                float arg = __fmul_rn(exp2f((float)oct), d[ray * 3 + comp]);
                x = (r6 >= 3) ? cosf(arg) : sinf(arg);
            }
            v[q] = x;
        }
        const uint32_t w = pack_bf16x2(v[0], v[1]);
        for (int n = 0; n < N; ++n) {
            const int64_t m = ray * N + n;
            const int64_t tile = m >> 7;
            const int r = (int)(m & 127);
            uint8_t* p = act_save + tile * SAVE_TILE_BYTES + SAVE_DIR + (r >> 3) * 1024 + (r & 7) * 128 +
                         ((((lane >> 2) ^ (r & 7))) << 4) + (lane & 3) * 4;
            *reinterpret_cast<uint32_t*>(p) = w;
        }
    }
}

// bias gradients of the two heads: b_rgb = sum d rgb_raw, b_sigma = sum d sigma_raw  (float4 grid-stride reduce)
__global__ void __launch_bounds__(256) head_bias_kernel(const float4* __restrict__ dpreds, int64_t M,
                                                        float* __restrict__ g_brgb, float* __restrict__ g_bsig) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = dpreds[i];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o); a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
        a.z += __shfl_xor_sync(0xffffffffu, a.z, o); a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(g_brgb, a.x); atomicAdd(g_brgb + 1, a.y); atomicAdd(g_brgb + 2, a.z); atomicAdd(g_bsig, a.w);
    }
}

// ------------------------------------------------------------------------------------------------
// 4. Input gradient of the (fine) MLP for the reference's un-stopped gradient through the fine sample positions
// (models.py:166-175 has no stop_gradient):  d_enc = dZ0 W0^T + dZ5 W5[256:319]^T  (tcgen05, M=128, N=64),
// positional-encoding backward per row, and  dtp[m] = < d_ray, dL/dpts[m] >  (data_utils.py:17-21,68-70).
// ------------------------------------------------------------------------------------------------
constexpr int IG_THREADS = 192;
constexpr int IG_W_BYTES = 2 * 4 * 8192;               // W0^T and W5b^T: 4 K-blocks of [64 n x 64 k] each
constexpr int IG_SM_W = 0;
constexpr int IG_SM_A = IG_W_BYTES;                    // 2 stages x 64 KB (one dZ image each)
constexpr int IG_SM_BAR = IG_SM_A + 2 * 65536;
constexpr int IG_SMEM = IG_SM_BAR + 128 + 1024;

__device__ void pack_ig_rows(const float* __restrict__ blob, const BlobOffsets& off, __nv_bfloat16* __restrict__ img) {
    // img[w][kb][n (64) x k (64)]: B[n][k] = W0[n][kb*64+k] (w=0) or W5[256+n][kb*64+k] (w=1); rows n >= 63 are zero
    const int total = 2 * 4 * 64 * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int w = i / (4 * 64 * 8), rem = i % (4 * 64 * 8);
        const int kb = rem / (64 * 8), n = (rem / 8) % 64, kg = rem % 8;
        uint32_t packed[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k0 = kb * 64 + kg * 8 + 2 * q;
            float a = 0.f, b = 0.f;
            if (n < ENC_X) {
                const int64_t base = (w == 0) ? off.w[0] + (int64_t)n * H : off.w[5] + (int64_t)(H + n) * H;
                a = blob[base + k0]; b = blob[base + k0 + 1];
            }
            packed[q] = pack_bf16x2(a, b);
        }
        uint8_t* dst = reinterpret_cast<uint8_t*>(img) + (size_t)(w * 4 + kb) * 8192 + sw128_offset(n, kg * 8);
        *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
}

struct IgParams {
    const uint8_t* dz_save;
    const __nv_bfloat16* w_img;
    const float* o; const float* d; const float* t;
    int N; int64_t M; int64_t n_tiles;
    float* dtp;                      // (M) < d, dL/dpts >
};

__global__ void __launch_bounds__(IG_THREADS, 1) nerf_input_grad_tc_kernel(const IgParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // barriers: wfull, full[2], empty[2], accf[2], acce[2]
    const uint32_t bar_w = base + IG_SM_BAR, bar_full = bar_w + 8, bar_empty = bar_full + 16, bar_accf = bar_empty + 16,
                   bar_acce = bar_accf + 16;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + IG_SM_BAR + 96);
    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
            mbar_init(bar_accf + 8 * i, 1);
            mbar_init(bar_acce + 8 * i, 128);
        }
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(base + IG_SM_BAR + 96, 128);     // two 64-column accumulators
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_tiles = (P.n_tiles > blockIdx.x) ? (int)((P.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    if (warp == 4) {
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_w, IG_W_BYTES);
            bulk_g2s(base + IG_SM_W, P.w_img, IG_W_BYTES, bar_w);
            int slot = 0; uint32_t par = 1;
            for (int i = 0; i < my_tiles; ++i) {
                const int64_t tile = blockIdx.x + (int64_t)i * gridDim.x;
                const uint8_t* tb = P.dz_save + tile * DZ_TILE_BYTES;
                for (int img = 0; img < 2; ++img) {            // dZ0 then dZ5
                    mbar_wait(bar_empty + 8 * slot, par, 21);
                    mbar_arrive_expect_tx(bar_full + 8 * slot, 65536);
                    bulk_g2s(base + IG_SM_A + slot * 65536, tb + DZ_Z + (img == 0 ? 0 : 5 * 65536), 65536, bar_full + 8 * slot);
                    if (++slot == 2) { slot = 0; par ^= 1; }
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
            mbar_wait(bar_w, 0, 22);
            int slot = 0; uint32_t par = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int ab = i & 1;                           // accumulator buffer
                const uint32_t ap = (uint32_t)((i >> 1) & 1);
                if (i >= 2) mbar_wait(bar_acce + 8 * ab, ap ^ 1, 23);   // epilogue of tile i-2 has drained this buffer
                tc_fence_after();
                for (int img = 0; img < 2; ++img) {
                    mbar_wait(bar_full + 8 * slot, par, 24);
                    tc_fence_after();
                    const uint32_t a0 = base + IG_SM_A + slot * 65536, b0 = base + IG_SM_W + img * 32768;
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t ad = make_sdesc_sw128(a0 + kb * 16384 + k * 32, 16, 1024);
                            const uint64_t bd = make_sdesc_sw128(b0 + kb * 8192 + k * 32, 16, 1024);
                            mma_bf16_ss(tmem_base + ab * 64, ad, bd, idesc, (img > 0 || kb > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    mma_commit(bar_empty + 8 * slot);
                    if (++slot == 2) { slot = 0; par ^= 1; }
                }
                mma_commit(bar_accf + 8 * ab);
            }
        }
    } else {
        const int row = threadIdx.x;
        for (int i = 0; i < my_tiles; ++i) {
            const int64_t tile = blockIdx.x + (int64_t)i * gridDim.x;
            const int64_t m = tile * TILE_M + row;
            const bool valid = m < P.M;
            const int64_t mm = valid ? m : (P.M - 1);
            const int64_t ray = mm / P.N;
            const float tv = P.t[mm];
            float dr[3], p[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                dr[c] = P.d[ray * 3 + c];
                p[c] = __fadd_rn(P.o[ray * 3 + c], __fmul_rn(dr[c], tv));
            }
            const int ab = i & 1;
            mbar_wait(bar_accf + 8 * ab, (uint32_t)((i >> 1) & 1), 25);
            tc_fence_after();
            uint32_t v0[32], v1[32];
            const uint32_t ta = tmem_base + (uint32_t(32 * warp) << 16) + ab * 64;
            tmem_ld32(ta, v0);
            tmem_ld32(ta + 32, v1);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(bar_acce + 8 * ab);
            float de[64];
#pragma unroll
            for (int q = 0; q < 32; ++q) { de[q] = __uint_as_float(v0[q]); de[32 + q] = __uint_as_float(v1[q]); }
            // encoding backward: e = [p, sin(2^i p), cos(2^i p)]_i  ->  dp = de_p + sum_i 2^i (cos * de_sin - sin * de_cos)
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float dp = de[c];
                float sv, cv;
#pragma unroll
                for (int oct = 0; oct < 10; ++oct) {
                    if (oct == 0 || oct == 5) sincosf((float)(1 << oct) * p[c], &sv, &cv);
                    else { const float s2 = 2.f * sv * cv, c2 = fmaf(-2.f * sv, sv, 1.f); sv = s2; cv = c2; }
                    dp = fmaf((float)(1 << oct), cv * de[3 + 6 * oct + c] - sv * de[3 + 6 * oct + 3 + c], dp);
                }
                acc = fmaf(dr[c], dp, acc);
            }
            if (valid) P.dtp[m] = acc;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------
// 5. Backward of sort(concat([t, sample_pdf(...)])) + sample_pdf into the coarse compositing weights
// (models.py:165-167, data_utils.py:179-220).  One warp per ray.
//   d_t_all[pos] = dtp[pos] + d_delta[pos-1] - d_delta[pos]      (delta_n = t[n+1] - t[n] in the fine compositing)
//   only sorted slots that came from sample_pdf (src_idx >= Nc) carry gradient; t, t_mid, u are constants.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(128) sample_pdf_bwd_kernel(const float* __restrict__ t, const float* __restrict__ weights,
                                                             const PdfDraws dr, const int32_t* __restrict__ src_idx,
                                                             const float* __restrict__ dtp, const float* __restrict__ d_delta,
                                                             int64_t B, int nc, int nf, float* __restrict__ d_w) {
    extern __shared__ float smem_pb[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int per_warp = 3 * (nc + 1) + nc;
    float* cdf = smem_pb + (size_t)wib * per_warp;      // nc + 1
    float* tm = cdf + nc + 1;                           // nc - 1 (+pad)
    float* dcdf = tm + nc + 1;                          // nc + 1
    float* pdf = dcdf + nc + 1;                         // nc
    const int na = nc + nf;
    const unsigned long long dstep = pdf_step(dr);
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < B; ray += warps_total) {
        const float* w = weights + ray * nc;
        const float* tr = t + ray * nc;
        // forward quantities (same arithmetic as build_cdf / invert_cdf in ops.cu)
        float ssum = 0.f;
        for (int n = lane; n < nc; n += 32) ssum += (w[n] + 1e-5f);
        ssum = warp_sum_f(ssum);
        float carry = 0.f;
        if (lane == 0) cdf[0] = 0.f;
        for (int b0 = 0; b0 < nc; b0 += 32) {
            const int n = b0 + lane;
            const float pv = (n < nc) ? __fdiv_rn(w[n] + 1e-5f, ssum) : 0.f;
            float incl = pv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const float x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
            if (n < nc) { cdf[n + 1] = carry + incl; pdf[n] = pv; }
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        for (int n = lane; n < nc - 1; n += 32) tm[n] = __fmul_rn(0.5f, __fadd_rn(tr[n + 1], tr[n]));
        for (int n = lane; n <= nc; n += 32) dcdf[n] = 0.f;
        __syncwarp();
        for (int pos = lane; pos < na; pos += 32) {
            const int k = src_idx[ray * na + pos];
            if (k < nc) continue;
            const int j = k - nc;
            float g = dtp[ray * na + pos];
            if (pos > 0) g += d_delta[ray * na + pos - 1];
            if (pos < na - 1) g -= d_delta[ray * na + pos];
            const float uu = pdf_draw(dr, dstep, ray, nf, j);
            int lo = 0, hi = nc + 1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] > uu) hi = mid; else lo = mid + 1; }
            const int below = max(0, lo - 1), above = min(nc, lo);
            const float cb = cdf[below], ca = cdf[above];
            const float tb = tm[min(nc - 2, below)], ta = tm[min(nc - 2, above)];
            const float den = ca - cb;
            const float span = ta - tb;
            if (den < 1e-5f) {
                atomicAdd(&dcdf[below], -g * span);             // samples = tb + (u - cdf_b) * span
            } else {
                const float f = (uu - cb) / den;
                atomicAdd(&dcdf[below], g * span * (f - 1.0f) / den);
                atomicAdd(&dcdf[above], -g * span * f / den);
            }
        }
        __syncwarp();
        // d_pdf[i] = sum_{k > i} d_cdf[k]  (cdf = [0, cumsum(pdf)]); reverse scan in chunks of 32 from the top
        float tail = 0.f, dot = 0.f;
        for (int b0 = ((nc - 1) / 32) * 32; b0 >= 0; b0 -= 32) {
            const int n = b0 + lane;
            float v = (n < nc) ? dcdf[n + 1] : 0.f;
            float incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const float x = __shfl_down_sync(0xffffffffu, incl, o); if (lane + o < 32) incl += x; }
            const float dp = incl + tail;
            if (n < nc) { dcdf[n + 1] = dp; dot += dp * pdf[n]; }
            tail += __shfl_sync(0xffffffffu, incl, 0);
        }
        dot = warp_sum_f(dot);
        __syncwarp();
        for (int n = lane; n < nc; n += 32) d_w[ray * nc + n] = (dcdf[n + 1] - dot) / ssum;
        __syncwarp();
    }
}

}  // namespace

namespace nerf {
int64_t tc_save_bytes_per_tile();
int g_wg_debug = 0;

int tc_train_alloc(nerf_ctx* ctx) {
    const nerf_config& c = ctx->cfg;
    // the CTA-pair forward kernel works on 512-row quads: size the per-tile storage for whole quads
    const int64_t tiles[2] = {ceil_div((int64_t)c.max_rays * c.ns_coarse, 512) * 4,
                              c.ns_fine > 0 ? ceil_div((int64_t)c.max_rays * (c.ns_coarse + c.ns_fine), 512) * 4 : 0};
    for (int net = 0; net < (c.ns_fine > 0 ? 2 : 1); ++net) {
        NERF_CUDA(cudaMalloc((void**)&ctx->act_save[net], (size_t)(tiles[net] * SAVE_TILE_BYTES)));
        // rows of a ragged last tile are read by the weight-gradient GEMMs (times zero gradients): keep them finite
        NERF_CUDA(cudaMemset(ctx->act_save[net], 0, (size_t)(tiles[net] * SAVE_TILE_BYTES)));
        NERF_CUDA(cudaMalloc((void**)&ctx->dz_save[net], (size_t)(tiles[net] * DZ_TILE_BYTES)));
        NERF_CUDA(cudaMalloc((void**)&ctx->mask_save[net], (size_t)(tiles[net] * MASK_TILE_BYTES)));
        NERF_CUDA(cudaMalloc((void**)&ctx->w_bwd[net], (size_t)B_CHUNKS * CHUNK_BYTES));
    }
    NERF_CUDA(cudaMalloc((void**)&ctx->w_ig, IG_W_BYTES));
    NERF_CUDA(cudaFuncSetAttribute(nerf_input_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IG_SMEM));
    NERF_CUDA(cudaFuncSetAttribute(nerf_mlp_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    NERF_CUDA(cudaFuncSetAttribute(nerf_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    return NERF_OK;
}

// transposed weight images of both nets and the input-gradient image of the fine net, one launch
int tc_pack_backward(nerf_ctx* ctx, cudaStream_t st) {
    BlobOffsets off = make_offsets(ctx);
    PackBwdArgs A = {};
    const bool single = ctx->cfg.ns_fine == 0;
    for (int net = 0; net < (single ? 1 : 2); ++net) {
        A.blob[net] = ctx->params + (int64_t)net * ctx->n_params;
        A.chunks[net] = ctx->w_bwd[net];
    }
    if (!single) {      // 64 KB: also packed for stop_grad_samples contexts (diagnostic entry points use it)
        A.ig_blob = ctx->params + ctx->n_params;
        A.ig_img = ctx->w_ig;
    }
    pack_bwd_kernel<<<dim3(num_sms() / 3, 3), 256, 0, st>>>(A, off);
    NERF_LAUNCHED();
    ctx->bwd_packed_valid = true;
    return NERF_OK;
}

// direction-encoding operand images of both nets (rows of Wddir below the feature part), one launch
int tc_dir_images(nerf_ctx* ctx, const float* d, int64_t B, int nc, int na, cudaStream_t st) {
    DirImageArgs A = {};
    A.N[0] = nc; A.act_save[0] = reinterpret_cast<uint8_t*>(ctx->act_save[0]);
    A.N[1] = na; A.act_save[1] = reinterpret_cast<uint8_t*>(ctx->act_save[1]);
    const int nets = (na > 0) ? 2 : 1;
    dir_image_kernel<<<dim3(stream_grid(B * 32, 128, 4), nets), 128, 0, st>>>(d, B, A);
    NERF_LAUNCHED();
    return NERF_OK;
}

// gradients of one net given dL/dpreds; the forward must have run with save_acts on the same batch.
// flags bit 0: the head bias gradients were already accumulated by volume_render_bwd; bit 1: the direction images of
// this net are already in place (tc_dir_images)
int tc_backward(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t, int64_t B, int N,
                const float* d_preds, cudaStream_t st, int flags) {
    (void)o; (void)t;
    const int64_t M = B * (int64_t)N;
    const int64_t n_pairs = ceil_div(M, 2 * TILE_M);
    // d_preds is the ctx-owned, tile-padded buffer: rows [M, padded) must carry zero gradient
    if (n_pairs * 2 * TILE_M > M)
        NERF_CUDA(cudaMemsetAsync(const_cast<float*>(d_preds) + M * 4, 0, (size_t)(n_pairs * 2 * TILE_M - M) * 16, st));
    BlobOffsets off = make_offsets(ctx);
    float* grads = ctx->grads + (int64_t)net * ctx->n_params;

    if (!ctx->bwd_packed_valid) {
        int rc = tc_pack_backward(ctx, st);
        if (rc) return rc;
    }

    BwdParams P;
    P.dpreds = reinterpret_cast<const float4*>(d_preds);
    P.M = M;
    P.n_pairs = n_pairs;
    P.w_chunks = ctx->w_bwd[net];
    P.side = ctx->side[net];
    P.mask_save = ctx->mask_save[net];
    P.dz_save = reinterpret_cast<uint8_t*>(ctx->dz_save[net]);
    int grid = (int)(n_pairs < num_sms() ? n_pairs : num_sms());
    timing_begin(1, st);
    nerf_mlp_bwd_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(P);
    timing_end(1, st);
    NERF_LAUNCHED();

    WgParams W;
    W.debug = g_wg_debug;
    W.act_save = reinterpret_cast<const uint8_t*>(ctx->act_save[net]);
    W.dz_save = reinterpret_cast<const uint8_t*>(ctx->dz_save[net]);
    W.dpreds = reinterpret_cast<const float4*>(d_preds);
    W.M = M;
    W.n_half_tiles = n_pairs * 4;
    W.grads = grads;
    auto job = [&](int i, int64_t a_off, int n_a, int64_t b_off, int n_b, int64_t w_dst, int ld, int rows, int col_lo,
                   int col_hi, int64_t bias, int64_t sig = -1) {
        W.jobs[i] = WgJob{a_off, b_off, n_a, n_b, w_dst, ld, rows, col_lo, col_hi, bias, sig};
    };
    job(0, SAVE_ENC, 1, DZ_Z, 4, off.w[0], H, ENC_X, 0, H, off.b[0]);
    for (int l = 1; l <= 4; ++l)
        job(l, SAVE_H + 65536 * (l - 1), 4, DZ_Z + 65536 * l, 4, off.w[l], H, H, 0, H, off.b[l]);
    job(5, SAVE_H + 65536 * 4, 4, DZ_Z + 65536 * 5, 4, off.w[5], H, H, 0, H, off.b[5]);
    job(6, SAVE_ENC, 1, DZ_Z + 65536 * 5, 4, off.w[5] + (int64_t)H * H, H, ENC_X, 0, H, -1);
    job(7, SAVE_H + 65536 * 5, 4, DZ_Z + 65536 * 6, 4, off.w[6], H, H, 0, H, off.b[6]);
    job(8, SAVE_H + 65536 * 6, 4, DZ_Z + 65536 * 7, 4, off.w[7], H, H, 0, H, off.b[7]);
    // feature layer; its X image is h8, so the same job also produces dW_sigma = h8^T d sigma as a side job
    job(9, SAVE_H + 65536 * 7, 4, DZ_FEAT, 4, off.w[9], H, H, 0, H, off.b[9], off.w[8]);
    job(10, SAVE_FEAT, 4, DZ_DDIR, 2, off.w[10], H / 2, H, 0, H / 2, off.b[10]);
    // rgb head: hd^T [d rgb, ..] -> columns 0..2 are dW_rgb (128,3)
    job(11, SAVE_HD, 2, DZ_HEAD, 1, off.w[11], 3, H / 2, 0, 3, -1);
    // direction rows of Wddir: dirimg^T dZ_ddir -> rows 256..282 of dW_ddir
    job(12, SAVE_DIR, 1, DZ_DDIR, 2, off.w[10] + (int64_t)H * (H / 2), H / 2, ENC_D, 0, H / 2, -1);
    if (!(flags & 1)) {
        head_bias_kernel<<<num_sms(), 256, 0, st>>>(reinterpret_cast<const float4*>(d_preds), M, grads + off.b[11],
                                                    grads + off.b[8]);
        NERF_LAUNCHED();
    }
    if (!(flags & 2)) {
        DirImageArgs A = {};
        A.N[0] = N; A.act_save[0] = reinterpret_cast<uint8_t*>(ctx->act_save[net]);
        dir_image_kernel<<<dim3(stream_grid(B * 32, 128), 1), 128, 0, st>>>(d, B, A);
        NERF_LAUNCHED();
    }
    int slabs = num_sms() / WG_NJOBS;
    if (slabs < 1) slabs = 1;
    if (slabs > W.n_half_tiles) slabs = (int)W.n_half_tiles;
    // CTAs per job: equal shares, the spare SMs go to the feature job (its side warps also compute the sigma head)
    int spare = num_sms() - slabs * WG_NJOBS;
    if (slabs >= W.n_half_tiles) spare = 0;
    W.cta_first[0] = 0;
    for (int j = 0; j < WG_NJOBS; ++j) W.cta_first[j + 1] = W.cta_first[j] + slabs + ((j == 9 && spare > 0) ? spare : 0);
    timing_begin(2, st);
    nerf_wgrad_tc_kernel<<<W.cta_first[WG_NJOBS], WG_THREADS, WG_SMEM, st>>>(W);
    timing_end(2, st);
    NERF_LAUNCHED();
    return NERF_OK;
}
}  // namespace nerf

extern "C" int nerf_debug_flags(int flags) { nerf::g_wg_debug = flags; return NERF_OK; }

namespace nerf {
// dtp[m] = < d_ray, dL/dpts[m] > of one net from its saved dZ0 / dZ5 images (tc_backward must have run)
int tc_input_grad(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t, int64_t B, int N, float* dtp,
                  cudaStream_t st) {
    const int64_t M = B * (int64_t)N;
    if (net != 1) return fail(NERF_ERR_INVALID, "tc_input_grad: the input-gradient image is packed for the fine net only");
    if (!ctx->bwd_packed_valid) {
        int rc = tc_pack_backward(ctx, st);
        if (rc) return rc;
    }
    IgParams P;
    P.dz_save = reinterpret_cast<const uint8_t*>(ctx->dz_save[net]);
    P.w_img = ctx->w_ig;
    P.o = o; P.d = d; P.t = t; P.N = N; P.M = M;
    P.n_tiles = ceil_div(M, TILE_M);
    P.dtp = dtp;
    const int grid = (int)(P.n_tiles < num_sms() ? P.n_tiles : num_sms());
    nerf_input_grad_tc_kernel<<<grid, IG_THREADS, IG_SMEM, st>>>(P);
    NERF_LAUNCHED();
    return NERF_OK;
}

int sample_pdf_backward(const float* t, const float* weights, const PdfDraws& dr, const int32_t* src_idx, const float* dtp,
                        const float* d_delta, int64_t B, int nc, int nf, float* d_w, cudaStream_t st) {
    const int threads = 128;
    const size_t smem = (size_t)(threads / 32) * (3 * (nc + 1) + nc) * sizeof(float);
    if (smem > 48 * 1024) return fail(NERF_ERR_INVALID, "sample_pdf_backward: nc too large");
    sample_pdf_bwd_kernel<<<stream_grid(B * 32, threads), threads, smem, st>>>(t, weights, dr, src_idx, dtp, d_delta, B, nc, nf, d_w);
    NERF_LAUNCHED();
    return NERF_OK;
}
}  // namespace nerf

// test hook: backward of sort(concat([t, sample_pdf])) given dL/dt_all contributions: dtp (direct) and, optionally,
// d_delta (dL/d(delta_n) of the fine compositing, delta_n = t_all[n+1] - t_all[n]); either may be NULL (= zeros)
extern "C" int nerf_sample_pdf_bwd(const float* t, const float* weights, const float* u, const int32_t* src_idx,
                                   const float* dtp, const float* d_delta, int64_t batch, int nc, int nf, float* d_w,
                                   void* stream) {
    NERF_CHECK_ARG(t && weights && u && src_idx && d_w && batch >= 1 && nc >= 2 && nf >= 1, "bad arguments");
    float* zeros = nullptr;
    NERF_CUDA(cudaMalloc(&zeros, (size_t)batch * (nc + nf) * 4));
    NERF_CUDA(cudaMemsetAsync(zeros, 0, (size_t)batch * (nc + nf) * 4, (cudaStream_t)stream));
    int rc = nerf::sample_pdf_backward(t, weights, nerf::PdfDraws{u, 0ull, 0ull, nullptr}, src_idx, dtp ? dtp : zeros,
                                       d_delta ? d_delta : zeros, batch, nc, nf, d_w, (cudaStream_t)stream);
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(zeros);
    return rc;
}
