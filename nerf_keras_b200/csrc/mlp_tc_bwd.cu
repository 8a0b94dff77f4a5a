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
//     The rgb head is one more job of the same kernel (the chain kernel emits a [d rgb, d sigma] "head" dZ image),
//     the sigma head a side job of the feature layer's job.
//  3. dirw_grad_kernel -- the direction rows of dW_ddir.  The forward kernel hoists that product into a per-ray bias;
//     its backward is enc_d(ray)^T times the per-ray sums of dZ_ddir, which the chain kernel takes while the tile is
//     in shared memory.
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
    uint32_t* progress;     // per tile: number of dZ images whose bulk stores are complete (null: nobody is listening)
    int N;                  // samples per ray
    float* ddir_sum;        // (rays, 128): per-ray column sums of dZ of the ddir layer, accumulated here (zeroed by the caller)
};

// The weight-gradient kernel may run NEXT TO this kernel on other SMs and consume a tile's dZ images while they are
// still in L2.  The thread that issues a sub-tile's bulk stores publishes how many images of the tile are complete:
// 1 = [ddir + head], 2 = feature, 3 .. 10 = Z7 .. Z0.  cp.async.bulk.wait_group N (not .read) returns once all but the N
// most recent groups have been written; the counter is then released at gpu scope.
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void publish_progress(uint32_t* p, uint32_t v) {
    asm volatile("fence.proxy.async.global;" ::: "memory");
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

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
        int64_t prev_tile = -1;                                          // elected thread: tile whose last stores are in flight

        for (int it = 0; it < my_pairs; ++it) {
            const int64_t pair = blockIdx.x + (int64_t)it * gridDim.x;
            const int64_t tile = pair * 2 + s;
            const int64_t g_row = tile * TILE_M + row;
            const bool valid = g_row < P.M;
            const uint32_t* mask_tile = P.mask_save + tile * (MASK_TILE_BYTES / 4);
            uint8_t* dz_tile = P.dz_save + tile * DZ_TILE_BYTES;
            const float4 dp = valid ? P.dpreds[g_row] : make_float4(0.f, 0.f, 0.f, 0.f);

            // ---- prologue: rgb head backward + ddir ReLU mask -> dZ_ddir (128 wide) ----
            if (elected) {
                bulk_wait_read0();
                if (P.progress && prev_tile >= 0) {      // all but the previous tile's very last store (Z0, second half)
                    bulk_wait_group<1>();
                    publish_progress(P.progress + prev_tile, 9);
                }
            }
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
                bulk_s2g_stream(dz_tile + DZ_DDIR, act_base, 32768);
                bulk_s2g_stream(dz_tile + DZ_HEAD, act_base + 2 * 16384, 16384);
                bulk_commit();
            }
            mbar_arrive(bar_lo);
            mbar_arrive(bar_hi);
            if (P.ddir_sum) {
                // Direction rows of dW_ddir (models.py:48-54): the direction encoding is constant along a ray, so
                //   sum_samples enc_d[k] dZ[sample][j] = sum_rays enc_d(ray)[k] (sum of the ray's dZ rows)[j].
                // Thread <-> column j of the 128-wide dZ_ddir tile just written (bf16, as the GEMMs see it): per-ray sums go
                // to ddir_sum, dirw_grad_kernel finishes the product.  Runs under the first MMA phase; K-blocks 0,1 are not
                // overwritten before the barrier in front of store_held.  (This replaced a per-sample direction image,
                // 16 KB per tile written by a kernel of its own and read back by a weight-gradient job.)
                // Thread <-> (8-column group cg, 16-row block rb): sixteen 16-byte loads, eight independent sums, flushed
                // with vector reductions at ray boundaries; four rows in flight at a time (register pressure).  A/B on one box:
                // +0.013 ms on the chain kernels of a step, -0.05 ms on the weight gradient, -0.03 ms of image kernel.
                const int cg = row & 15, rb = row >> 4;
                const uint32_t blk = act_base + (cg >> 3) * (TILE_M * 128);
                int64_t g = tile * TILE_M + rb * 16;
                int64_t ray = g / P.N;
                int rem = (int)(g - ray * P.N), cnt = 0;
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                auto flush = [&]() {
                    float* dstp = P.ddir_sum + ray * 128 + cg * 8;
                    red_add_v4(dstp, acc[0], acc[1], acc[2], acc[3]);
                    red_add_v4(dstp + 4, acc[4], acc[5], acc[6], acc[7]);
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
                };
#pragma unroll 1
                for (int q4 = 0; q4 < 4; ++q4) {                      // four rows at a time: 16 registers of loads in flight
                    uint4 v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = rb * 16 + q4 * 4 + i;
                        const uint32_t addr = blk + (r >> 3) * 1024 + (r & 7) * 128 + ((uint32_t)((cg & 7) ^ (r & 7)) << 4);
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w) : "r"(addr));
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i, ++g) {
                        if (g < P.M) {
                            acc[0] += __uint_as_float(v[i].x << 16); acc[1] += __uint_as_float(v[i].x & 0xFFFF0000u);
                            acc[2] += __uint_as_float(v[i].y << 16); acc[3] += __uint_as_float(v[i].y & 0xFFFF0000u);
                            acc[4] += __uint_as_float(v[i].z << 16); acc[5] += __uint_as_float(v[i].z & 0xFFFF0000u);
                            acc[6] += __uint_as_float(v[i].w << 16); acc[7] += __uint_as_float(v[i].w & 0xFFFF0000u);
                            ++cnt;
                            if (++rem == P.N) { flush(); rem = 0; cnt = 0; ++ray; }
                        }
                    }
                }
                if (cnt > 0) flush();
            }

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
                if (elected) {
                    bulk_wait_read0();                               // previous image store has finished reading the tile
                    if (P.progress && ph >= 1) {
                        // groups so far: [ddir + head], then two per phase; all but the last two (= phase ph-1) are written
                        bulk_wait_group<2>();
                        if (ph == 1 && prev_tile >= 0) publish_progress(P.progress + prev_tile, 10);
                        publish_progress(P.progress + tile, (uint32_t)ph);
                    }
                }
                named_bar_sync(1 + s, TILE_M);
                store_held(rs, held);
                tc_fence_before();
                fence_proxy_async_smem();
                if (ph < B_PHASES - 1) mbar_arrive(bar_lo);
                // first half of the dZ image (K-blocks 0,1) leaves now, under the second half of the epilogue: one 64 KB
                // burst of TMA shared-memory reads at the start of the next MMA phase costs more than two 32 KB ones
                named_bar_sync(1 + s, TILE_M);
                if (elected) { bulk_s2g_stream(dz_tile + dst, act_base, 32768); bulk_commit(); }
                if (ph == 0) chain_part2<false, false>(t_lane, 0.f, wsig, mask, rs);
                else if (ph == 1) chain_part2<true, true>(t_lane, dp.w, wsig, mask, rs);
                else chain_part2<false, true>(t_lane, 0.f, wsig, mask, rs);
                tc_fence_before();
                fence_proxy_async_smem();
                named_bar_sync(1 + s, TILE_M);
                if (elected) { bulk_s2g_stream(dz_tile + dst + 32768, act_base + 32768, 32768); bulk_commit(); }
                if (ph < B_PHASES - 1) mbar_arrive(bar_hi);
            }
            prev_tile = tile;
        }
        if (elected) {
            bulk_wait_all0();
            if (P.progress && prev_tile >= 0) publish_progress(P.progress + prev_tile, 10);
        }
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
// Two rings of WHOLE 128-sample tiles.  X ring: 3 stages x 32 KB (two 64-feature K-blocks = one M=128 block of the
// operand, contiguous in the saved image); dZ ring: 128 KB, stages of n_b x 16 KB (the whole dZ image, contiguous).
// Every stage is filled by ONE bulk copy of 16 .. 64 KB: the copy engine of an SM retires one cp.async.bulk per ~130 ns
// whatever its size (tools/micro/bulk_bw.cu: 63 GB/s per SM with 8 KB copies, 125 / 159 / 198 GB/s with 16 / 32 / 64 KB),
// so a CTA fed by 8 KB copies cannot outrun ~50 GB/s -- invisible while 148 CTAs share HBM, decisive when the kernel runs
// on a fraction of the SMs next to the dX chain (tc_backward).
constexpr int WG_A_STAGES = 3;
constexpr int WG_A_SLOT = 32768;
constexpr int WG_B_MAX_STAGES = 4;
constexpr int WG_B_RING = 131072;
constexpr int WG_SM_A = 0;
constexpr int WG_SM_B = WG_A_STAGES * WG_A_SLOT;
constexpr int WG_SM_BAR = WG_SM_B + WG_B_RING;
constexpr int WG_SMEM = WG_SM_BAR + 256 + 1024;
constexpr int WG_NJOBS = 12;

struct WgJob {
    int64_t a_off;       // byte offset of the X image inside a saved-activation tile
    int64_t b_off;       // byte offset of the dZ image inside a dZ tile
    int n_a;             // 64-feature blocks of X: 1, 2 or 4
    int n_b;             // 64-feature blocks of dZ: 1, 2 or 4
    int64_t w_dst;       // float offset (in the grads blob of this net) of dW row 0
    int ld;              // fan_out
    int rows;            // valid rows of dW produced by this job
    int col_lo, col_hi;  // valid columns [col_lo, col_hi) of the accumulator that are written to dW
    int64_t bias_dst;    // float offset of the bias gradient (column sums of dZ), or -1
    int64_t sig_dst;     // float offset of dW_sigma (256,1): column sums of the X image weighted by d sigma per sample, or -1
    int need;            // progress count of the dX chain (images stored so far for a tile) at which this job's dZ image exists
    int ig;              // -1, or which block of the input-gradient weight image (0: W0^T, 1: W5[256:]^T) this job's dZ image meets
};
struct WgParams {
    int debug;              // bit0 / bit3: skip side jobs, bit1: skip MMAs, bit2: skip final reduction (timing experiments)
    WgJob jobs[WG_NJOBS];
    int cta_first[WG_NJOBS + 1];   // 1-D grid: job j owns CTAs [cta_first[j], cta_first[j+1]); its tiles are dealt round-robin
    const uint8_t* act_save;
    const uint8_t* dz_save;
    const float4* dpreds;
    int64_t M;
    int64_t n_tiles;
    const uint32_t* progress;   // per tile: images the concurrently running dX chain has completed (null: chain already done)
    long long* stats;           // diagnostics (nerf_debug_wgrad_stats): per CTA {job, tiles, end ns, ns waiting for the chain, ns waiting for slots}
    float* grads;           // this net's gradient blob
    // input gradient of the net (reference semantics, Q5), fused into the two jobs that stream dZ0 / dZ5 anyway:
    //   dtp[m] += < d_ray, J_enc(p_m)^T (dZ_l[m] W_l^T) >;  null: not wanted
    const __nv_bfloat16* w_ig;
    const float *ig_o, *ig_d, *ig_t;
    int ig_N;
    float* dtp;
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Bounded wait for the chain's progress counter of one tile (same policy as mbar_wait: trap instead of hanging)
__device__ __forceinline__ void wait_progress(const uint32_t* p, uint32_t need, int code) {
    if (ld_acquire_gpu(p) >= need) return;
    long long t0 = 0;
    uint32_t n = 0;
    while (ld_acquire_gpu(p) < need) {
        __nanosleep(200);
        if ((++n & 1023u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > TC5_TIMEOUT_CYCLES) {
                printf("wgrad: chain progress wait timeout code=%d block=%d need=%u have=%u\n", code, blockIdx.x, need,
                       ld_acquire_gpu(p));
                __trap();
            }
        }
    }
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
    const uint32_t bar_afull = base + WG_SM_BAR, bar_aempty = bar_afull + 8 * WG_A_STAGES,
                   bar_bfull = bar_aempty + 8 * WG_A_STAGES, bar_bempty = bar_bfull + 8 * WG_B_MAX_STAGES,
                   bar_done = bar_bempty + 8 * WG_B_MAX_STAGES, bar_d2full = bar_done + 8, bar_d2free = bar_d2full + 16,
                   bar_w = bar_d2free + 16;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + WG_SM_BAR + 200);
    const int n_mh = (J.n_a == 4) ? 2 : 1;                     // M = 128 blocks of the operand (X stages per tile)
    const uint32_t b_bytes = (uint32_t)J.n_b * 16384u;
    int b_stages = WG_B_RING / (int)b_bytes;
    if (b_stages > WG_B_MAX_STAGES) b_stages = WG_B_MAX_STAGES;
    const bool side_off = (P.debug & (1 | 8)) != 0;
    const bool do_sig = J.sig_dst >= 0 && !side_off && !(P.debug & 32), do_bias = J.bias_dst >= 0 && !side_off;
    // fused input gradient: these jobs have a 16 KB X image (two X stages are plenty); the third X slot holds the 32 KB
    // weight image, tensor-memory columns 256..383 two 64-column accumulators
    const bool do_ig = J.ig >= 0 && P.dtp != nullptr;
    const int a_stages = do_ig ? 2 : WG_A_STAGES;
    const uint32_t w_img = base + WG_SM_A + 2 * WG_A_SLOT;

    if (threadIdx.x == 0) {
        for (int i = 0; i < WG_A_STAGES; ++i) {
            mbar_init(bar_afull + 8 * i, 1);
            mbar_init(bar_aempty + 8 * i, 1 + (do_sig ? 32 * WG_SIDE_WARPS : 0));   // tcgen05.commit (+ the side-job threads)
        }
        for (int i = 0; i < WG_B_MAX_STAGES; ++i) {
            mbar_init(bar_bfull + 8 * i, 1);
            mbar_init(bar_bempty + 8 * i, 1 + (do_bias ? 32 * WG_SIDE_WARPS : 0));
        }
        mbar_init(bar_done, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(bar_d2full + 8 * i, 1); mbar_init(bar_d2free + 8 * i, 128); }
        mbar_init(bar_w, 1);
        fence_barrier_init();
    }
    if (P.stats && threadIdx.x == 0) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.stats[(size_t)blockIdx.x * 8 + 6] = t;
    }
    if (warp == 5) tmem_alloc(base + WG_SM_BAR + 200, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this CTA's tiles: job_cta, job_cta + job_ctas, ... (the order in which the dX chain produces them)
    // next to the chain (P.progress): job_cta, job_cta + job_ctas, .. -- the order in which the chain produces tiles;
    // alone: one contiguous slab of tiles per CTA (same speed as the interleaved order: 1.69 ms per step, HBM-bound)
    const bool interleave = P.progress != nullptr;
    const int64_t slab = (P.n_tiles + job_ctas - 1) / job_ctas;
    const int64_t t_first = interleave ? job_cta : (int64_t)job_cta * slab, t_step = interleave ? job_ctas : 1;
    int n_my;
    if (interleave) n_my = (P.n_tiles > job_cta) ? (int)((P.n_tiles - job_cta + job_ctas - 1) / job_ctas) : 0;
    else n_my = (int)((t_first + slab <= P.n_tiles) ? slab : (P.n_tiles > t_first ? P.n_tiles - t_first : 0));

    if (warp == 4) {
        // ===================== loader =====================
        if (lane == 0) {
            int as = 0, bs = 0;
            uint32_t apar = 1, bpar = 1;
            long long t_flag = 0, t_slot = 0;
            auto now_ns = []() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
            if (do_ig) {
                mbar_arrive_expect_tx(bar_w, 32768);
                bulk_g2s(w_img, reinterpret_cast<const uint8_t*>(P.w_ig) + (size_t)J.ig * 32768, 32768, bar_w);
            }
            for (int i = 0; i < n_my; ++i) {
                const int64_t tile = t_first + (int64_t)i * t_step;
                long long t0 = P.stats ? now_ns() : 0;
                mbar_wait(bar_bempty + 8 * bs, bpar, 11);
                if (P.stats) { const long long t1 = now_ns(); t_slot += t1 - t0; t0 = t1; }
                if (P.progress) {
                    // the chain kernel runs next to this one: its bulk stores of this image have completed (and were
                    // released at gpu scope) once the tile's counter has reached `need`; order the acquire before the
                    // async-proxy read below
                    wait_progress(P.progress + tile, (uint32_t)J.need, 15);
                    fence_proxy_async_all();
                    if (P.stats) t_flag += now_ns() - t0;
                }
                mbar_arrive_expect_tx(bar_bfull + 8 * bs, b_bytes);
                bulk_g2s(base + WG_SM_B + bs * b_bytes, P.dz_save + tile * DZ_TILE_BYTES + J.b_off, b_bytes, bar_bfull + 8 * bs);
                if (++bs == b_stages) { bs = 0; bpar ^= 1; }
                const uint8_t* a_src = P.act_save + tile * SAVE_TILE_BYTES + J.a_off;
                for (int mh = 0; mh < n_mh; ++mh) {
                    t0 = P.stats ? now_ns() : 0;
                    mbar_wait(bar_aempty + 8 * as, apar, 12);
                    if (P.stats) t_slot += now_ns() - t0;
                    const uint32_t dst = base + WG_SM_A + as * WG_A_SLOT;
                    mbar_arrive_expect_tx(bar_afull + 8 * as, WG_A_SLOT);
                    if (J.n_a == 1) {        // a single 64-feature block is loaded twice (M = 128 MMA)
                        bulk_g2s(dst, a_src, 16384, bar_afull + 8 * as);
                        bulk_g2s(dst + 16384, a_src, 16384, bar_afull + 8 * as);
                    } else {
                        bulk_g2s(dst, a_src + mh * 32768, 32768, bar_afull + 8 * as);
                    }
                    if (++as == a_stages) { as = 0; apar ^= 1; }
                }
            }
            if (P.stats) {
                long long* o = P.stats + (size_t)blockIdx.x * 8;
                o[0] = job_id; o[1] = n_my; o[3] = t_flag; o[4] = t_slot; o[5] = now_ns();
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, 64 * J.n_b, 1, 1);
            const uint32_t idesc_ig = make_idesc_bf16(128, 64, 0, 0);      // [128 samples x 256] K-major times [64 x 256] K-major
            int as = 0, bs = 0;
            uint32_t apar = 0, bpar = 0;
            if (do_ig) mbar_wait(bar_w, 0, 18);
            for (int i = 0; i < n_my; ++i) {
                mbar_wait(bar_bfull + 8 * bs, bpar, 13);
                const uint32_t b0 = base + WG_SM_B + bs * b_bytes;
                for (int mh = 0; mh < n_mh; ++mh) {
                    mbar_wait(bar_afull + 8 * as, apar, 14);
                    tc_fence_after();
                    const uint32_t a0 = base + WG_SM_A + as * WG_A_SLOT;
                    if (!(P.debug & 2)) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {      // 128 samples = 8 x K16; 16 sample rows = 2048 bytes
                            const uint64_t ad = make_sdesc_sw128(a0 + k * 2048, 16384, 1024);
                            const uint64_t bd = make_sdesc_sw128(b0 + k * 2048, 16384, 1024);
                            mma_bf16_ss(tmem_base + mh * 256, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    mma_commit(bar_aempty + 8 * as);
                    if (++as == a_stages) { as = 0; apar ^= 1; }
                }
                if (do_ig) {
                    // the dZ image of this stage once more, now as the K-major A operand: d_enc = dZ_l W_l^T  (M = 128, N = 64)
                    const int buf = i & 1;
                    mbar_wait(bar_d2free + 8 * buf, (uint32_t)(((i >> 1) & 1) ^ 1), 19);
                    tc_fence_after();
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            mma_bf16_ss(tmem_base + 256 + buf * 64, make_sdesc_sw128(b0 + kb * 16384 + k * 32, 16, 1024),
                                        make_sdesc_sw128(w_img + kb * 8192 + k * 32, 16, 1024), idesc_ig, (kb > 0 || k > 0) ? 1u : 0u);
                    mma_commit(bar_d2full + 8 * buf);
                }
                mma_commit(bar_bempty + 8 * bs);
                if (++bs == b_stages) { bs = 0; bpar ^= 1; }
            }
            mma_commit(bar_done);
        }
    } else {
        // ===================== side jobs + final reduction =====================
        // Eight side warps; side warp sw takes rows [16 sw, 16 sw + 16) of each 128-sample tile.
        //   bias gradient (column sums of the dZ image): lane <-> (64-column block cb = lane / 8, 16-byte chunk ch) = 8 columns
        //   sigma head (X = h8, two stages of 128 features): lane <-> (block cb = (lane / 8) & 1, row half (lane / 16), chunk ch)
        const int sw = (warp < 4) ? warp : warp - 2;
        const int sub = lane >> 3, ch = lane & 7;
        float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // bias: partial column sums (8 columns per lane)
        float sg[2][8] = {{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}};
        // eight rows at a time: all loads first, then the FMAs (two warps per scheduler: the side jobs live on instruction-
        // level parallelism; one load -> eight dependent FMAs per row ran at ~80 cycles per row)
        auto load8 = [&](uint32_t blk, int row0, uint4 (&v)[8]) {
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                const int row = row0 + rr;
                const uint32_t addr = blk + row * 128 + ((ch ^ (row & 7)) << 4);
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[rr].x), "=r"(v[rr].y), "=r"(v[rr].z), "=r"(v[rr].w) : "r"(addr));
            }
        };
        auto fma8 = [&](const uint4& w, float wgt, float (&acc)[8]) {
            acc[0] = fmaf(wgt, __uint_as_float(w.x << 16), acc[0]); acc[1] = fmaf(wgt, __uint_as_float(w.x & 0xFFFF0000u), acc[1]);
            acc[2] = fmaf(wgt, __uint_as_float(w.y << 16), acc[2]); acc[3] = fmaf(wgt, __uint_as_float(w.y & 0xFFFF0000u), acc[3]);
            acc[4] = fmaf(wgt, __uint_as_float(w.z << 16), acc[4]); acc[5] = fmaf(wgt, __uint_as_float(w.z & 0xFFFF0000u), acc[5]);
            acc[6] = fmaf(wgt, __uint_as_float(w.w << 16), acc[6]); acc[7] = fmaf(wgt, __uint_as_float(w.w & 0xFFFF0000u), acc[7]);
        };
        if (do_sig || do_bias || do_ig) {
            int as = 0, bs = 0;
            uint32_t apar = 0, bpar = 0;
            // d sigma of this warp's 16 samples of a tile: one load per lane, fetched ONE TILE AHEAD
            float next_ds = 0.f;
            if (do_sig && lane < 16 && n_my > 0) next_ds = __ldg(&P.dpreds[t_first * TILE_M + sw * 16 + lane].w);
            for (int i = 0; i < n_my; ++i) {
                const int64_t tile = t_first + (int64_t)i * t_step;
                const float my_ds = next_ds;
                if (do_sig && lane < 16 && i + 1 < n_my) next_ds = __ldg(&P.dpreds[(tile + t_step) * TILE_M + sw * 16 + lane].w);
                if (do_bias) {
                    mbar_wait(bar_bfull + 8 * bs, bpar, 16);
                    if (sub < J.n_b && !(P.debug & 16)) {
                        const uint32_t blk = base + WG_SM_B + bs * b_bytes + sub * 16384;
                        uint4 v[8];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            load8(blk, sw * 16 + 8 * h, v);
#pragma unroll
                            for (int rr = 0; rr < 8; ++rr) fma8(v[rr], 1.0f, cs);
                        }
                    }
                    mbar_arrive(bar_bempty + 8 * bs);
                    if (++bs == b_stages) { bs = 0; bpar ^= 1; }
                }
                if (do_sig) {
                    // dW_sigma[f] += sum over the tile's samples of h8[sample][f] * d sigma[sample]  (models.py:42): the X
                    // image of this job IS h8, so the sigma head costs no extra HBM traffic
                    float wrow[8];
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) wrow[rr] = __shfl_sync(0xffffffffu, my_ds, (sub >> 1) * 8 + rr);
                    for (int mh = 0; mh < n_mh; ++mh) {
                        mbar_wait(bar_afull + 8 * as, apar, 17);
                        const uint32_t blk = base + WG_SM_A + as * WG_A_SLOT + (sub & 1) * 16384;
                        uint4 v[8];
                        load8(blk, sw * 16 + (sub >> 1) * 8, v);
                        float part[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) fma8(v[rr], wrow[rr], part);
                        mbar_arrive(bar_aempty + 8 * as);
                        if (++as == a_stages) { as = 0; apar ^= 1; }
#pragma unroll
                        for (int q = 0; q < 8; ++q) { if (mh == 0) sg[0][q] += part[q]; else sg[1][q] += part[q]; }
                    }
                }
                if (do_ig && warp < 4) {
                    // positional-encoding backward of this tile's d_enc rows (thread <-> sample), data_utils.py:17-21,68-70
                    const int buf = i & 1;
                    const int64_t m = tile * TILE_M + 32 * warp + lane;
                    const bool valid = m < P.M;
                    const int64_t mm = valid ? m : (P.M - 1);
                    const int64_t ray = mm / P.ig_N;
                    const float tv = P.ig_t[mm];
                    float dr[3], pt[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        dr[c] = P.ig_d[ray * 3 + c];
                        pt[c] = __fadd_rn(P.ig_o[ray * 3 + c], __fmul_rn(dr[c], tv));
                    }
                    mbar_wait(bar_d2full + 8 * buf, (uint32_t)((i >> 1) & 1), 20);
                    tc_fence_after();
                    uint32_t v0[32], v1[32];
                    const uint32_t ta = tmem_base + (uint32_t(32 * warp) << 16) + 256 + buf * 64;
                    tmem_ld32(ta, v0);
                    tmem_ld32(ta + 32, v1);
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(bar_d2free + 8 * buf);
                    float de[64];
#pragma unroll
                    for (int q = 0; q < 32; ++q) { de[q] = __uint_as_float(v0[q]); de[32 + q] = __uint_as_float(v1[q]); }
                    // e = [p, sin(2^i p), cos(2^i p)]_i  ->  dp = de_p + sum_i 2^i (cos * de_sin - sin * de_cos)
                    float accp = 0.f;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float dp = de[c];
                        float sv, cv;
#pragma unroll
                        for (int oct = 0; oct < 10; ++oct) {
                            if (oct == 0 || oct == 5) sincosf((float)(1 << oct) * pt[c], &sv, &cv);
                            else { const float s2 = 2.f * sv * cv, c2 = fmaf(-2.f * sv, sv, 1.f); sv = s2; cv = c2; }
                            dp = fmaf((float)(1 << oct), cv * de[3 + 6 * oct + c] - sv * de[3 + 6 * oct + 3 + c], dp);
                        }
                        accp = fmaf(dr[c], dp, accp);
                    }
                    if (valid) atomicAdd(P.dtp + m, accp);
                }
            }
        }
        if (n_my > 0) {
            if (do_bias && sub < J.n_b) {
#pragma unroll
                for (int q = 0; q < 8; ++q) atomicAdd(P.grads + J.bias_dst + sub * 64 + ch * 8 + q, cs[q]);
            }
            if (do_sig) {
#pragma unroll
                for (int mh = 0; mh < 2; ++mh)
#pragma unroll
                    for (int q = 0; q < 8; ++q) atomicAdd(P.grads + J.sig_dst + mh * 128 + (sub & 1) * 64 + ch * 8 + q, sg[mh][q]);
            }
            if (!(P.debug & 4) && warp < 4) {     // TMEM lane quarters belong to warps 0-3
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
    if (P.stats && threadIdx.x == 0) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.stats[(size_t)blockIdx.x * 8 + 2] = t;
    }
    if (warp == 5) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// 3. direction rows of dW_ddir (rows 256..282, models.py:48-54) from the per-ray sums of dZ_ddir the chain kernel took:
//    dW[256 + k][j] += sum_rays enc_d(ray)[k] * S[ray][j].  The forward hoists this product into a per-ray bias
//    (dirbias_kernel); this is its backward.  Thread <-> column j, 27 accumulators, one atomic per (k, j) and block.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) dirw_grad_kernel(const float* __restrict__ d, int64_t rays, const float* __restrict__ S,
                                                        float* __restrict__ g_wdir /* (27, 128) */) {
    __shared__ float enc[4][32];
    const int j = threadIdx.x;
    float acc[ENC_D];
#pragma unroll
    for (int k = 0; k < ENC_D; ++k) acc[k] = 0.f;
    // four rays per round: their four S loads are in flight together (one dependent L2 / HBM round trip per ray made this
    // 30 us); threads 0..107 compute the 4 x 27 encoding values
    for (int64_t ray0 = (int64_t)blockIdx.x * 4; ray0 < rays; ray0 += (int64_t)gridDim.x * 4) {
        float sv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) sv[q] = (ray0 + q < rays) ? __ldg(S + (ray0 + q) * 128 + j) : 0.f;
        if (j < 4 * ENC_D) {
            const int q = j / ENC_D, c = j - q * ENC_D;
            const int64_t ray = ray0 + q;
            float v = 0.f;
            if (ray < rays) {
                if (c < 3) v = d[ray * 3 + c];
                else {
                    const int qq = c - 3, i = qq / 6, r = qq - i * 6, comp = r % 3;
                    const float arg = __fmul_rn(exp2f((float)i), d[ray * 3 + comp]);
                    v = (r >= 3) ? cosf(arg) : sinf(arg);
                }
            }
            enc[q][c] = v;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int k = 0; k < ENC_D; ++k) acc[k] = fmaf(enc[q][k], sv[q], acc[k]);
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < ENC_D; ++k) atomicAdd(g_wdir + k * (H / 2) + j, acc[k]);
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
long long* g_wg_stats = nullptr;
constexpr int WG_OVERLAP_CTAS = 0;       // default: no overlap (measured slower); nerf_set_backward_overlap opts in

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
        NERF_CUDA(cudaMalloc((void**)&ctx->chain_progress[net], (size_t)tiles[net] * 4));
        NERF_CUDA(cudaMalloc((void**)&ctx->tr_ddirsum[net], (size_t)c.max_rays * (H / 2) * 4));
        ctx->progress_tiles[net] = tiles[net];
    }
    NERF_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
    NERF_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    NERF_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    ctx->wgrad_ctas = WG_OVERLAP_CTAS;
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

// gradients of one net given dL/dpreds; the forward must have run with save_acts on the same batch.
// flags bit 0: the head bias gradients were already accumulated by volume_render_bwd (bit 1: unused); bit 2: also
// leave the net's input gradient dtp[m] = < d_ray, dL/dpts[m] > in ctx->tr_dtp_f (fine net; see WgParams::dtp)
int tc_backward(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t, int64_t B, int N,
                const float* d_preds, cudaStream_t st, int flags) {
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
    P.N = N;
    P.ddir_sum = ctx->tr_ddirsum[net];
    NERF_CUDA(cudaMemsetAsync(ctx->tr_ddirsum[net], 0, (size_t)B * (H / 2) * 4, st));
    // flags bit 2: the input gradient of this net (dtp = ctx->tr_dtp_f) as a by-product of the jobs that stream dZ0 / dZ5.
    // Zeroed HERE: in front of the fork, the weight-gradient kernel may run on the side stream next to the chain
    const bool fuse_ig = (flags & 4) != 0 && ctx->tr_dtp_f != nullptr && ctx->w_ig != nullptr;
    if (fuse_ig) {
        if (net != 1) return fail(NERF_ERR_INVALID, "tc_backward: the input-gradient image is packed for the fine net only");
        NERF_CUDA(cudaMemsetAsync(ctx->tr_dtp_f, 0, (size_t)M * 4, st));
    }
    // Overlap: with at least one full wave of tile pairs the chain gives up `wg_ctas` SMs to the weight-gradient kernel,
    // which follows it tile by tile (see publish_progress / wait_progress).  Both kernels need a whole SM per CTA
    // (shared memory), so chain CTAs + weight-gradient CTAs <= SMs keeps every CTA resident: the chain never waits for
    // the consumer, the consumer only for counters the chain is certain to publish -- no deadlock.
    int wg_ctas = (g_wg_debug >> 8) ? (g_wg_debug >> 8) : ctx->wgrad_ctas;
    if (g_wg_debug & 64) wg_ctas = 0;
    const bool budgeted = wg_ctas >= WG_NJOBS && wg_ctas <= num_sms() / 2 && n_pairs >= num_sms();
    const bool overlap = budgeted && !(g_wg_debug & 128) && ctx->side_stream && n_pairs * 2 <= ctx->progress_tiles[net];
    const int chain_sms = overlap ? num_sms() - wg_ctas : num_sms();
    P.progress = overlap ? ctx->chain_progress[net] : nullptr;
    int grid = (int)(n_pairs < chain_sms ? n_pairs : chain_sms);
    cudaStream_t wst = st;
    timing_begin(3, st);
    if (overlap) {
        NERF_CUDA(cudaMemsetAsync(ctx->chain_progress[net], 0, (size_t)n_pairs * 2 * 4, st));
        NERF_CUDA(cudaEventRecord(ctx->ev_fork, st));
        NERF_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
        wst = ctx->side_stream;
    }
    timing_begin(1, st);
    nerf_mlp_bwd_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(P);
    timing_end(1, st);
    NERF_LAUNCHED();
    // direction rows of dW_ddir from the per-ray sums the chain just took
    dirw_grad_kernel<<<(int)(ceil_div(B, 4) < 2 * num_sms() ? ceil_div(B, 4) : 2 * num_sms()), 128, 0, st>>>(
        d, B, ctx->tr_ddirsum[net], grads + off.w[10] + (int64_t)H * (H / 2));
    NERF_LAUNCHED();

    WgParams W;
    W.debug = g_wg_debug;
    W.act_save = reinterpret_cast<const uint8_t*>(ctx->act_save[net]);
    W.dz_save = reinterpret_cast<const uint8_t*>(ctx->dz_save[net]);
    W.dpreds = reinterpret_cast<const float4*>(d_preds);
    W.M = M;
    W.n_tiles = n_pairs * 2;
    W.progress = P.progress;
    W.stats = g_wg_stats ? g_wg_stats + (size_t)net * 148 * 8 : nullptr;
    W.w_ig = ctx->w_ig; W.ig_o = o; W.ig_d = d; W.ig_t = t; W.ig_N = N;
    W.dtp = fuse_ig ? ctx->tr_dtp_f : nullptr;
    W.grads = grads;
    auto job = [&](int i, int64_t a_off, int n_a, int64_t b_off, int n_b, int64_t w_dst, int ld, int rows, int col_lo,
                   int col_hi, int64_t bias, int64_t sig = -1) {
        // images of a tile leave the chain kernel in the order [ddir + head], feature, Z7, Z6, .. Z0 (progress 1 .. 10)
        const int need = (b_off >= DZ_DDIR) ? 1 : (b_off == DZ_FEAT) ? 2 : 10 - (int)((b_off - DZ_Z) / 65536);
        W.jobs[i] = WgJob{a_off, b_off, n_a, n_b, w_dst, ld, rows, col_lo, col_hi, bias, sig, need, -1};
    };
    job(0, SAVE_ENC, 1, DZ_Z, 4, off.w[0], H, ENC_X, 0, H, off.b[0]);
    for (int l = 1; l <= 4; ++l)
        job(l, SAVE_H + 65536 * (l - 1), 4, DZ_Z + 65536 * l, 4, off.w[l], H, H, 0, H, off.b[l]);
    job(5, SAVE_H + 65536 * 4, 4, DZ_Z + 65536 * 5, 4, off.w[5], H, H, 0, H, off.b[5]);
    job(6, SAVE_ENC, 1, DZ_Z + 65536 * 5, 4, off.w[5] + (int64_t)H * H, H, ENC_X, 0, H, -1);
    W.jobs[0].ig = 0;      // dZ0 W0^T
    W.jobs[6].ig = 1;      // dZ5 W5[256:319]^T
    job(7, SAVE_H + 65536 * 5, 4, DZ_Z + 65536 * 6, 4, off.w[6], H, H, 0, H, off.b[6]);
    job(8, SAVE_H + 65536 * 6, 4, DZ_Z + 65536 * 7, 4, off.w[7], H, H, 0, H, off.b[7]);
    // feature layer; its X image is h8, so the same job also produces dW_sigma = h8^T d sigma as a side job
    job(9, SAVE_H + 65536 * 7, 4, DZ_FEAT, 4, off.w[9], H, H, 0, H, off.b[9], off.w[8]);
    job(10, SAVE_FEAT, 4, DZ_DDIR, 2, off.w[10], H / 2, H, 0, H / 2, off.b[10]);
    // rgb head: hd^T [d rgb, ..] -> columns 0..2 are dW_rgb (128,3)
    job(11, SAVE_HD, 2, DZ_HEAD, 1, off.w[11], 3, H / 2, 0, 3, -1);
    if (!(flags & 1)) {
        head_bias_kernel<<<num_sms(), 256, 0, st>>>(reinterpret_cast<const float4*>(d_preds), M, grads + off.b[11],
                                                    grads + off.b[8]);
        NERF_LAUNCHED();
    }
    W.cta_first[0] = 0;
    if (!budgeted) {
        // the kernel owns the GPU: equal shares, the spare SMs go to the feature job (its side warps also compute the sigma head)
        int slabs = num_sms() / WG_NJOBS;
        if (slabs < 1) slabs = 1;
        if (slabs > W.n_tiles) slabs = (int)W.n_tiles;
        int spare = num_sms() - slabs * WG_NJOBS;
        if (slabs >= W.n_tiles) spare = 0;
        for (int j = 0; j < WG_NJOBS; ++j) W.cta_first[j + 1] = W.cta_first[j] + slabs + ((j == 9 && spare > 0) ? spare : 0);
    } else {
        // a budget of SMs next to the chain: shares in proportion to a job's time per tile = bytes streamed (units of
        // 16 KB at the ~60 GB/s one CTA sustains) + ~0.5 us of fixed cost per tile (2 units; measured on the rgb-head job,
        // tools/r2_wg_stats.py: with a share by bytes alone its two CTAs finished 1 ms after everybody else)
        static const int units[WG_NJOBS] = {8, 10, 10, 10, 10, 10, 8, 10, 10, 10, 8, 5};
        int cnt[WG_NJOBS];
        for (int j = 0; j < WG_NJOBS; ++j) cnt[j] = 1;
        for (int sum = WG_NJOBS; sum < wg_ctas; ++sum) {       // next CTA to the job with the most bytes per CTA
            int best = 0;
            for (int j = 1; j < WG_NJOBS; ++j) if (units[j] * cnt[best] > units[best] * cnt[j]) best = j;
            ++cnt[best];
        }
        for (int j = 0; j < WG_NJOBS; ++j) W.cta_first[j + 1] = W.cta_first[j] + cnt[j];
    }
    timing_begin(2, wst);
    nerf_wgrad_tc_kernel<<<W.cta_first[WG_NJOBS], WG_THREADS, WG_SMEM, wst>>>(W);
    timing_end(2, wst);
    NERF_LAUNCHED();
    if (overlap) {
        NERF_CUDA(cudaEventRecord(ctx->ev_join, ctx->side_stream));
        NERF_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    }
    timing_end(3, st);
    return NERF_OK;
}
}  // namespace nerf

extern "C" int nerf_debug_flags(int flags) { nerf::g_wg_debug = flags; return NERF_OK; }
extern "C" int nerf_debug_wgrad_stats(long long* dev_buf) { nerf::g_wg_stats = dev_buf; return NERF_OK; }

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
