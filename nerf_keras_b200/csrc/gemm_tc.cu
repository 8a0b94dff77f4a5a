// fp32-in / fp32-out GEMMs on the Blackwell tensor cores for the BATCH_NORM training path (bn_train.cu), sm_100a only.
//
// BATCH_NORM=true (models.py:30-33,49-52) needs the statistics of the WHOLE batch between two Dense layers, so that path
// runs layer by layer and its contractions are plain GEMMs over row-major fp32 activations.  They used to go to cuBLAS
// sgemm; they now run here as tcgen05 MMAs with split operands ("bf16x3"): every fp32 value x is written to shared
// memory as hi = bf16(x) and lo = bf16(x - hi), and  A B ~= Ahi Bhi + Alo Bhi + Ahi Blo  accumulates in fp32 in tensor
// memory (the dropped Alo Blo term is 2^-16 relative) -- fp32-grade results at a third of the bf16 MMA rate, which is
// still ~20x the fp32 FMA rate.  The forward GEMMs use a three-way split (hi + mid + lo = x exactly, six products):
// a pre-activation that is off by 1e-5 flips ReLU decisions, and every flip is a visible difference in the gradients.
// Two kernels:
//   gemm_rows_kernel   C (M x N) = A (M x K) op(B) [+ C]      M = samples (large), K, N <= 256
//       a pipeline over the 64-wide K-blocks of 128-row tiles: 8 warps convert A (K-major, 128B-swizzled part tiles) into
//       a 3-stage ring, a producer thread streams the pre-split B chunks ([128 n x 64 k] per part) through a 4-slot ring,
//       one thread issues the MMAs into one of two 256-column accumulators, 4 warps write the other one out.
//   gemm_tn_kernel     C (Mo x N) += A^T B,  A (Ms x Mo), B (Ms x N): the contraction runs over the SAMPLES
//       a [64 samples x 64 features] tile stored K-major IS an MN-major operand of the transposed product (the trick of
//       nerf_wgrad_tc_kernel), so the same conversion feeds it; each CTA owns a slab of 64-sample half-tiles (two stages:
//       one is converted while the other is multiplied) and one 128-column half of N, keeps its (<= 256 x 128) accumulator
//       in tensor memory and adds it to C with atomics at the end.
// Skinny shapes (N <= 4 or K <= 4: the sigma / rgb heads) are CUDA-core kernels at the end of this file.
#include "common.cuh"
#include "tc5.cuh"

using namespace nerf;
using namespace tc5;

namespace {

constexpr int GT_THREADS = 320;      // transposed GEMM: warps 0-7 operand conversion (0-3 also the final write-out), 9: MMA issuer
constexpr int GT_WORKERS = 256;

__device__ __forceinline__ void split_store(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, float v) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    *reinterpret_cast<__nv_bfloat16*>(hi_tile + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(lo_tile + off) = l;
}

__device__ __forceinline__ void split_store4(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, const float4& v) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z),
                        h3 = __float2bfloat16_rn(v.w);
    uint2 hi, lo;
    hi.x = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
    hi.y = pack_bf16x2(__bfloat162float(h2), __bfloat162float(h3));
    lo.x = pack_bf16x2(v.x - __bfloat162float(h0), v.y - __bfloat162float(h1));
    lo.y = pack_bf16x2(v.z - __bfloat162float(h2), v.w - __bfloat162float(h3));
    *reinterpret_cast<uint2*>(hi_tile + off) = hi;
    *reinterpret_cast<uint2*>(lo_tile + off) = lo;
}

// rows [row0, row0 + 64) x columns [col0, col0 + 64 * blocks) of a row-major fp32 matrix -> hi / lo bf16 half-tiles,
// [block][64 rows][64] K-major 128B-swizzled (8 KB per block); rows >= rows_end and columns >= cols_end read as zero.
// `worker` = 0 .. 255 (8 warps): warp w takes rows w, w + 8, ..; lanes run along the columns (coalesced).  All loads of
// a batch of rows are issued before the first conversion.
__device__ __forceinline__ void convert_half_tile(const float* __restrict__ src, int64_t ld, int64_t row0, int64_t rows_end,
                                                  int col0, int cols_end, int blocks, uint8_t* hi_tile, uint8_t* lo_tile,
                                                  int worker) {
    const int w = worker >> 5, lane = worker & 31;
    const int ncols = blocks * 64;
    const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(src + col0) & 15) == 0) &&
                     ((cols_end - col0) % 4 == 0 || cols_end - col0 >= ncols);
    if (vec) {
        // lane <-> four consecutive columns; up to two float4 per row and lane (256 columns), four rows in flight
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float4 v[4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t g = row0 + w + 8 * (half * 4 + i);
                const float* __restrict__ p = src + g * ld + col0;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c = 4 * lane + 128 * j;
                    v[i][j] = (g < rows_end && c < ncols && col0 + c < cols_end) ? __ldg(reinterpret_cast<const float4*>(p + c))
                                                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c = 4 * lane + 128 * j, r = w + 8 * (half * 4 + i);
                    if (c < ncols) split_store4(hi_tile, lo_tile, (uint32_t)(c >> 6) * 8192u + sw128_offset(r, c & 63), v[i][j]);
                }
        }
    } else {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
            float v[2][8];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int64_t g = row0 + w + 8 * (q4 * 2 + i);
                const float* __restrict__ p = src + g * ld + col0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = lane + 32 * j;
                    v[i][j] = (g < rows_end && c < ncols && col0 + c < cols_end) ? __ldg(p + c) : 0.f;
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = lane + 32 * j, r = w + 8 * (q4 * 2 + i);
                    if (c < ncols) split_store(hi_tile, lo_tile, (uint32_t)(c >> 6) * 8192u + sw128_offset(r, c & 63), v[i][j]);
                }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Row GEMM:  C (M x N) = A (M x K) op(B) [+ C].  PARTS = 2: x ~ hi + lo (three products, ~1e-5 of the products' scale);
// PARTS = 3: x = hi + mid + lo exactly (six products, fp32-grade) -- the FORWARD GEMMs of the BATCH_NORM path use it,
// because a pre-activation that moves by 1e-5 flips ReLU decisions and every flip is a visible gradient difference.
// ------------------------------------------------------------------------------------------------
constexpr int GR_THREADS = 448;      // warps 0-7: A conversion, 8-11: epilogue, 12: B producer, 13: MMA issuer
constexpr int GR_CONV = 256;
constexpr int GR_A_STAGES = 3;       // stage = one 64-wide K-block of a 128-row tile, PARTS x 16 KB
constexpr int GR_B_STAGES = 4;       // 16 KB chunks [128 n x 64 k] of one part of B
constexpr int GR_SM_A = 0;
constexpr int GR_SM_B = GR_A_STAGES * 3 * 16384;
constexpr int GR_SM_OUT = GR_SM_B + GR_B_STAGES * 16384;      // 4 epilogue warps x [16 rows][36 floats]: transpose staging
constexpr int GR_OUT_WARP = 16 * 36 * 4;
constexpr int GR_SM_BAR = GR_SM_OUT + 4 * GR_OUT_WARP;
constexpr int GR_SMEM = GR_SM_BAR + 256 + 1024;
static_assert(GR_SMEM <= 232448, "shared memory of gemm_rows_kernel");

template <int PARTS>
__device__ __forceinline__ void split_parts(float v, __nv_bfloat16 (&o)[3]) {
    o[0] = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(o[0]);
    o[1] = __float2bfloat16_rn(r1);
    o[2] = (PARTS == 3) ? __float2bfloat16_rn(r1 - __bfloat162float(o[1])) : __float2bfloat16_rn(0.f);
}

// B split and packed once per call: chunk ((kb * NH + h) * PARTS + part) = [128 n x 64 k], K-major, 128B-swizzled
//   tb = 0: B is (K x N) row-major -> B[n][k] = B[k * ldb + n];  tb = 1: B is (N x K) row-major
template <int PARTS>
__global__ void __launch_bounds__(256) pack_b_split_kernel(const float* __restrict__ B, int64_t ldb, int tb, int N, int K,
                                                           int NH, int KB, uint8_t* __restrict__ out) {
    const int total = NH * KB * 128 * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int k = i & 63, n = (i >> 6) & 127;
        const int h = (i >> 13) % NH, kb = (i >> 13) / NH;
        if (!tb) {                       // run along n for coalesced reads of a (K x N) matrix
            n = i & 127; k = (i >> 7) & 63;
        }
        const int gn = h * 128 + n, gk = kb * 64 + k;
        float v = 0.f;
        if (gn < N && gk < K) v = tb ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn];
        __nv_bfloat16 parts[3];
        split_parts<PARTS>(v, parts);
        uint8_t* chunk = out + (size_t)((kb * NH + h) * PARTS) * 16384 + sw128_offset(n, k);
#pragma unroll
        for (int q = 0; q < PARTS; ++q) *reinterpret_cast<__nv_bfloat16*>(chunk + q * 16384) = parts[q];
    }
}

// one K-block (columns [col0, col0 + 64)) of rows [row0, row0 + 128) -> PARTS tiles of [128][64] at stage + part * 16 KB.
// Aligned case: lane <-> four consecutive columns, 16 lanes per row, two rows per warp instruction, eight float4 per thread.
__device__ __forceinline__ bool block_is_vec(const float* src, int64_t ld, int cols_end) {
    return (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((cols_end & 3) == 0);
}
__device__ __forceinline__ void load_block_vec(const float* __restrict__ src, int64_t ld, int64_t row0, int64_t rows_end, int col0,
                                               int cols_end, int worker, float4 (&v)[8]) {
    const int w = worker >> 5, lane = worker & 31;
    const int c = 4 * (lane & 15), rr = 2 * w + (lane >> 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t g = row0 + rr + 16 * i;
        v[i] = (g < rows_end && col0 + c < cols_end) ? __ldg(reinterpret_cast<const float4*>(src + g * ld + col0 + c))
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
template <int PARTS>
__device__ __forceinline__ void store_block_vec(const float4 (&v)[8], uint8_t* stage, int worker) {
    const int w = worker >> 5, lane = worker & 31;
    const int c = 4 * (lane & 15), rr = 2 * w + (lane >> 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t off = sw128_offset(rr + 16 * i, c);
        __nv_bfloat16 a[3], b[3], cc[3], d[3];
        split_parts<PARTS>(v[i].x, a); split_parts<PARTS>(v[i].y, b); split_parts<PARTS>(v[i].z, cc); split_parts<PARTS>(v[i].w, d);
#pragma unroll
        for (int q = 0; q < PARTS; ++q) {
            uint2 o;
            o.x = (uint32_t)__bfloat16_as_ushort(a[q]) | ((uint32_t)__bfloat16_as_ushort(b[q]) << 16);
            o.y = (uint32_t)__bfloat16_as_ushort(cc[q]) | ((uint32_t)__bfloat16_as_ushort(d[q]) << 16);
            *reinterpret_cast<uint2*>(stage + q * 16384 + off) = o;
        }
    }
}
// unaligned / ragged case (K = 63, 27, odd leading dimensions): lanes along the columns, scalar loads, all loads of eight rows
// in flight before the first conversion
template <int PARTS>
__device__ __forceinline__ void convert_block_scalar(const float* __restrict__ src, int64_t ld, int64_t row0, int64_t rows_end,
                                                     int col0, int cols_end, uint8_t* stage, int worker) {
    const int w = worker >> 5, lane = worker & 31;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float v[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t g = row0 + w + 8 * (half * 8 + i);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = lane + 32 * j;
                v[i][j] = (g < rows_end && col0 + c < cols_end) ? __ldg(src + g * ld + col0 + c) : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t off = sw128_offset(w + 8 * (half * 8 + i), lane + 32 * j);
                __nv_bfloat16 parts[3];
                split_parts<PARTS>(v[i][j], parts);
#pragma unroll
                for (int q = 0; q < PARTS; ++q) *reinterpret_cast<__nv_bfloat16*>(stage + q * 16384 + off) = parts[q];
            }
    }
}

struct RowsParams {
    const float* A; int64_t lda;
    const uint8_t* b_chunks;
    float* C; int64_t ldc;
    int64_t M; int N; int K;
    int NH, KB;
    float beta;
};

template <int PARTS>
__global__ void __launch_bounds__(GR_THREADS, 1) gemm_rows_kernel(const RowsParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t A_STAGE = PARTS * 16384;
    const uint32_t bar_afull = base + GR_SM_BAR, bar_aempty = bar_afull + 8 * GR_A_STAGES, bar_bfull = bar_aempty + 8 * GR_A_STAGES,
                   bar_bempty = bar_bfull + 8 * GR_B_STAGES, bar_accfull = bar_bempty + 8 * GR_B_STAGES, bar_accfree = bar_accfull + 16;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + GR_SM_BAR + 192);
    if (threadIdx.x == 0) {
        for (int i = 0; i < GR_A_STAGES; ++i) { mbar_init(bar_afull + 8 * i, GR_CONV); mbar_init(bar_aempty + 8 * i, 1); }
        for (int i = 0; i < GR_B_STAGES; ++i) { mbar_init(bar_bfull + 8 * i, 1); mbar_init(bar_bempty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_accfull + 8 * i, 1); mbar_init(bar_accfree + 8 * i, 128); }
        fence_barrier_init();
    }
    if (warp == 13) tmem_alloc(base + GR_SM_BAR + 192, 512);           // two 256-column accumulators
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t n_tiles = (P.M + 127) / 128;
    const int my_tiles = (n_tiles > blockIdx.x) ? (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    if (warp < 8) {
        // ===================== A conversion: one K-block per stage =====================
        int s = 0; uint32_t par = 1;
        if (block_is_vec(P.A, P.lda, P.K)) {
            // the loads of block n + 1 are issued before block n is converted and stored: two blocks (64 KB per SM) in flight
            const int n_blocks = my_tiles * P.KB;
            float4 cur[8], nxt[8];
            if (n_blocks > 0) load_block_vec(P.A, P.lda, (int64_t)blockIdx.x * 128, P.M, 0, P.K, threadIdx.x, cur);
            int it = 0, kb = 0;
            for (int n = 0; n < n_blocks; ++n) {
                int it2 = it, kb2 = kb + 1;
                if (kb2 == P.KB) { kb2 = 0; ++it2; }
                if (n + 1 < n_blocks)
                    load_block_vec(P.A, P.lda, (blockIdx.x + (int64_t)it2 * gridDim.x) * 128, P.M, kb2 * 64, P.K, threadIdx.x, nxt);
                mbar_wait(bar_aempty + 8 * s, par, 31);
                store_block_vec<PARTS>(cur, smem + GR_SM_A + s * A_STAGE, threadIdx.x);
                fence_proxy_async_smem();
                mbar_arrive(bar_afull + 8 * s);
                if (++s == GR_A_STAGES) { s = 0; par ^= 1; }
#pragma unroll
                for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
                it = it2; kb = kb2;
            }
        } else {
            for (int it = 0; it < my_tiles; ++it) {
                const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
                for (int kb = 0; kb < P.KB; ++kb) {
                    mbar_wait(bar_aempty + 8 * s, par, 31);
                    convert_block_scalar<PARTS>(P.A, P.lda, tile * 128, P.M, kb * 64, P.K, smem + GR_SM_A + s * A_STAGE, threadIdx.x);
                    fence_proxy_async_smem();
                    mbar_arrive(bar_afull + 8 * s);
                    if (++s == GR_A_STAGES) { s = 0; par ^= 1; }
                }
            }
        }
    } else if (warp < 12) {
        // ===================== epilogue: accumulator -> C =====================
        const int q = warp - 8;                                        // TMEM lane quarter (warp % 4)
        const bool vec = (P.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0);
        for (int it = 0; it < my_tiles; ++it) {
            const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
            const int buf = it & 1;
            mbar_wait(bar_accfull + 8 * buf, (uint32_t)((it >> 1) & 1), 32);
            tc_fence_after();
            const int64_t row = tile * 128 + 32 * q + lane;
            for (int cg = 0; cg < P.NH * 4; ++cg) {
                if (cg * 32 >= P.N) break;
                uint32_t v[32];
                tmem_ld32(tmem_base + (uint32_t(32 * q) << 16) + buf * 256 + cg * 32, v);
                tmem_ld_wait();
                if (vec && cg * 32 + 32 <= P.N) {
                    // through shared memory: thread <-> row on the way in, 8 lanes <-> one row's 128 bytes on the way out, so
                    // that a warp store writes four complete 128-byte lines instead of 16 bytes of 32 different ones
                    float* stg = reinterpret_cast<float*>(smem + GR_SM_OUT + q * GR_OUT_WARP);
                    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
                    for (int p = 0; p < 2; ++p) {                      // sixteen rows at a time (staging fits next to the rings)
                        if ((lane >> 4) == p) {
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                *reinterpret_cast<float4*>(stg + (lane & 15) * 36 + 4 * e) =
                                    make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                                                __uint_as_float(v[4 * e + 3]));
                        }
                        __syncwarp();
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int rl = 4 * e + rsub;
                            const int64_t grow = tile * 128 + 32 * q + 16 * p + rl;
                            float4 o = *reinterpret_cast<const float4*>(stg + rl * 36 + c4);
                            if (grow < P.M) {
                                float* dst = P.C + grow * P.ldc + cg * 32 + c4;
                                if (P.beta != 0.f) {
                                    const float4 c = *reinterpret_cast<const float4*>(dst);
                                    o.x = fmaf(P.beta, c.x, o.x); o.y = fmaf(P.beta, c.y, o.y);
                                    o.z = fmaf(P.beta, c.z, o.z); o.w = fmaf(P.beta, c.w, o.w);
                                }
                                *reinterpret_cast<float4*>(dst) = o;
                            }
                        }
                        __syncwarp();
                    }
                } else if (row < P.M) {
                    float* dst = P.C + row * P.ldc + cg * 32;
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (cg * 32 + e < P.N) {
                            float o = __uint_as_float(v[e]);
                            if (P.beta != 0.f) o = fmaf(P.beta, dst[e], o);
                            dst[e] = o;
                        }
                }
            }
            tc_fence_before();
            mbar_arrive(bar_accfree + 8 * buf);
        }
    } else if (warp == 12) {
        // ===================== B producer =====================
        if (lane == 0) {
            const int n_chunks = P.KB * P.NH * PARTS;
            int slot = 0; uint32_t par = 1;
            for (int it = 0; it < my_tiles; ++it)
                for (int c = 0; c < n_chunks; ++c) {
                    mbar_wait(bar_bempty + 8 * slot, par, 33);
                    mbar_arrive_expect_tx(bar_bfull + 8 * slot, 16384);
                    bulk_g2s(base + GR_SM_B + slot * 16384, P.b_chunks + (size_t)c * 16384, 16384, bar_bfull + 8 * slot);
                    if (++slot == GR_B_STAGES) { slot = 0; par ^= 1; }
                }
        }
    } else {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
            int as = 0, bslot = 0; uint32_t apar = 0, bpar = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int buf = it & 1;
                mbar_wait(bar_accfree + 8 * buf, (uint32_t)(((it >> 1) & 1) ^ 1), 34);    // epilogue of tile it - 2 has drained it
                tc_fence_after();
                for (int kb = 0; kb < P.KB; ++kb) {
                    mbar_wait(bar_afull + 8 * as, apar, 35);
                    tc_fence_after();
                    const uint32_t a0 = base + GR_SM_A + as * A_STAGE;
                    for (int h = 0; h < P.NH; ++h) {
                        const uint32_t d = tmem_base + buf * 256 + h * 128;
#pragma unroll
                        for (int j = 0; j < PARTS; ++j) {                 // part j of B against parts 0 .. PARTS-1-j of A
                            mbar_wait(bar_bfull + 8 * bslot, bpar, 36);
                            tc_fence_after();
                            const uint32_t b0 = base + GR_SM_B + bslot * 16384;
#pragma unroll
                            for (int i = 0; i < PARTS - j; ++i)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    mma_bf16_ss(d, make_sdesc_sw128(a0 + i * 16384 + k * 32, 16, 1024),
                                                make_sdesc_sw128(b0 + k * 32, 16, 1024), idesc, (kb > 0 || j > 0 || i > 0 || k > 0) ? 1u : 0u);
                            mma_commit(bar_bempty + 8 * bslot);
                            if (++bslot == GR_B_STAGES) { bslot = 0; bpar ^= 1; }
                        }
                    }
                    mma_commit(bar_aempty + 8 * as);
                    if (++as == GR_A_STAGES) { as = 0; apar ^= 1; }
                }
                mma_commit(bar_accfull + 8 * buf);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 13) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// C (Mo x N) += A^T B over the samples:  A (Ms x Mo), B (Ms x N) row-major fp32.  grid = (slabs, N halves)
// ------------------------------------------------------------------------------------------------
// Two stages of 64-sample half-tiles: A [4 blocks of 64 features][64 samples][64] hi + lo (64 KB), B [2 blocks] hi + lo (32 KB)
constexpr int TN_STAGE = 98304;
constexpr int TN_A_HI = 0, TN_A_LO = 32768, TN_B_HI = 65536, TN_B_LO = 65536 + 16384;
constexpr int TN_SM_BAR = 2 * TN_STAGE;
constexpr int TN_SMEM = TN_SM_BAR + 128 + 1024;

struct TnParams {
    const float* A; int64_t lda;
    const float* B; int64_t ldb;
    float* C; int64_t ldc;
    int64_t Ms; int Mo; int N;
};

__global__ void __launch_bounds__(GT_THREADS, 1) gemm_tn_kernel(const TnParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = base + TN_SM_BAR, bar_empty = bar_full + 16, bar_done = bar_empty + 16;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + TN_SM_BAR + 64);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(bar_full + 8 * i, GT_WORKERS); mbar_init(bar_empty + 8 * i, 1); }
        mbar_init(bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(base + TN_SM_BAR + 64, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nh = blockIdx.y;
    const int MB = (P.Mo + 127) / 128;                 // M = 128 blocks of the transposed operand
    const int64_t n_half = (P.Ms + 63) / 64;           // 64-sample half-tiles, dealt round-robin to the slabs
    const int my_half = (n_half > blockIdx.x) ? (int)((n_half - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    if (warp == 9) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, 128, 1, 1);      // both operands MN-major
            int s = 0; uint32_t par = 0;
            for (int it = 0; it < my_half; ++it) {
                mbar_wait(bar_full + 8 * s, par, 36);
                tc_fence_after();
                const uint32_t st = base + s * TN_STAGE;
                for (int mb = 0; mb < MB; ++mb) {
                    const uint32_t ahi = st + TN_A_HI + mb * 16384, alo = st + TN_A_LO + mb * 16384;
                    const uint32_t d = tmem_base + mb * 128;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                       // 16 samples = 16 rows = 2048 bytes; 64-feature blocks 8 KB apart
                        const uint64_t dah = make_sdesc_sw128(ahi + k * 2048, 8192, 1024), dal = make_sdesc_sw128(alo + k * 2048, 8192, 1024);
                        const uint64_t dbh = make_sdesc_sw128(st + TN_B_HI + k * 2048, 8192, 1024), dbl = make_sdesc_sw128(st + TN_B_LO + k * 2048, 8192, 1024);
                        mma_bf16_ss(d, dah, dbh, idesc, (it > 0 || k > 0) ? 1u : 0u);
                        mma_bf16_ss(d, dal, dbh, idesc, 1u);
                        mma_bf16_ss(d, dah, dbl, idesc, 1u);
                    }
                }
                mma_commit(bar_empty + 8 * s);
                if (++s == 2) { s = 0; par ^= 1; }
            }
            mma_commit(bar_done);
        }
    } else if (warp < 8) {
        const int worker = threadIdx.x;
        int s = 0; uint32_t par = 1;
        for (int it = 0; it < my_half; ++it) {
            const int64_t ht = blockIdx.x + (int64_t)it * gridDim.x;
            mbar_wait(bar_empty + 8 * s, par, 37);                      // the MMAs that read this stage have completed
            uint8_t* st = smem + s * TN_STAGE;
            convert_half_tile(P.A, P.lda, ht * 64, P.Ms, 0, P.Mo, MB * 2, st + TN_A_HI, st + TN_A_LO, worker);
            convert_half_tile(P.B, P.ldb, ht * 64, P.Ms, nh * 128, P.N, 2, st + TN_B_HI, st + TN_B_LO, worker);
            fence_proxy_async_smem();
            mbar_arrive(bar_full + 8 * s);
            if (++s == 2) { s = 0; par ^= 1; }
        }
        if (my_half > 0 && warp < 4) {
            mbar_wait(bar_done, 0, 38);
            tc_fence_after();
            for (int mb = 0; mb < MB; ++mb) {
                const int r = mb * 128 + 32 * warp + lane;              // row of C = feature of A
                for (int cg = 0; cg < 4; ++cg) {
                    const int c0 = nh * 128 + cg * 32;
                    if (c0 >= P.N) break;
                    uint32_t v[32];
                    tmem_ld32(tmem_base + (uint32_t(32 * warp) << 16) + mb * 128 + cg * 32, v);
                    tmem_ld_wait();
                    if (r < P.Mo) {
                        float* dst = P.C + (int64_t)r * P.ldc + c0;
#pragma unroll
                        for (int q = 0; q < 32; ++q)
                            if (c0 + q < P.N) atomicAdd(dst + q, __uint_as_float(v[q]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------
// skinny shapes on CUDA cores
// ------------------------------------------------------------------------------------------------
// C (M x N) = A (M x K) op(B) [+ C], N <= 4: one warp per row, lanes along K
__global__ void __launch_bounds__(256) small_rows_dot_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                             int64_t ldb, int tb, float* __restrict__ C, int64_t ldc, int64_t M,
                                                             int N, int K, float beta) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = lane; k < K; k += 32) {
            const float a = A[m * lda + k];
#pragma unroll
            for (int n = 0; n < 4; ++n)
                if (n < N) acc[n] = fmaf(a, tb ? B[(int64_t)n * ldb + k] : B[(int64_t)k * ldb + n], acc[n]);
        }
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
        if (lane < N) {
            float v = acc[0];
            if (lane == 1) v = acc[1]; else if (lane == 2) v = acc[2]; else if (lane == 3) v = acc[3];
            float* dst = C + m * ldc + lane;
            *dst = (beta != 0.f) ? fmaf(beta, *dst, v) : v;
        }
    }
}
// C (M x N) = A (M x K) op(B) [+ C], K <= 4: one thread per output element
__global__ void __launch_bounds__(256) small_rows_outer_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                               int64_t ldb, int tb, float* __restrict__ C, int64_t ldc, int64_t M,
                                                               int N, int K, float beta) {
    const int64_t total = M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / N;
        const int n = (int)(i - m * N);
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(A[m * lda + k], tb ? B[(int64_t)n * ldb + k] : B[(int64_t)k * ldb + n], acc);
        float* dst = C + m * ldc + n;
        *dst = (beta != 0.f) ? fmaf(beta, *dst, acc) : acc;
    }
}
// C (Mo x N) += A^T B, N <= 4, Mo <= 256: thread j owns column j of A over a slab of samples
__global__ void __launch_bounds__(256) small_tn_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                       int64_t ldb, float* __restrict__ C, int64_t ldc, int64_t Ms, int Mo, int N) {
    const int j = threadIdx.x;
    const int64_t per = (Ms + gridDim.x - 1) / gridDim.x;
    const int64_t m0 = (int64_t)blockIdx.x * per, m1 = (m0 + per < Ms) ? m0 + per : Ms;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (j < Mo) {
        for (int64_t m = m0; m < m1; m += 8) {          // eight rows in flight
            float av[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) av[q] = (m + q < m1) ? A[(m + q) * lda + j] : 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (m + q < m1) {
#pragma unroll
                    for (int n = 0; n < 4; ++n)
                        if (n < N) acc[n] = fmaf(av[q], __ldg(B + (m + q) * ldb + n), acc[n]);
                }
            }
        }
#pragma unroll
        for (int n = 0; n < 4; ++n)
            if (n < N) atomicAdd(C + (int64_t)j * ldc + n, acc[n]);
    }
}

uint8_t* g_b_scratch = nullptr;      // split B chunks of the GEMM in flight (one stream at a time: the BN path is serial)
constexpr size_t B_SCRATCH_BYTES = 2 * 4 * 3 * 16384;      // N halves x K-blocks x parts
bool g_attr_set = false;

}  // namespace

namespace nerf {

// row-major C (M x N) = op(A) op(B) + beta C, fp32 in and out.  ta: A is stored (K x M) and the contraction runs over
// its rows (the sample dimension; beta must be 1: the result is ADDED to C); tb: B is stored (N x K).
// precise: three-way operand split (fp32-grade; row GEMMs only) instead of two-way (~1e-5 of the products' scale).
int tc_gemm_f32(cudaStream_t st, bool ta, bool tb, int64_t M, int N, int K, const float* A, int64_t lda, const float* B,
                int64_t ldb, float beta, float* C, int64_t ldc, bool precise) {
    if (M <= 0 || N <= 0 || K <= 0) return NERF_OK;
    if (!g_attr_set) {
        NERF_CUDA(cudaFuncSetAttribute(gemm_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_SMEM));
        NERF_CUDA(cudaFuncSetAttribute(gemm_rows_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_SMEM));
        NERF_CUDA(cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM));
        NERF_CUDA(cudaMalloc((void**)&g_b_scratch, B_SCRATCH_BYTES));
        g_attr_set = true;
    }
    if (ta) {
        // M = rows of C = features of A (<= 256), K = samples
        if (tb) return fail(NERF_ERR_INVALID, "tc_gemm_f32: A^T B^T is not used by this library");
        if (beta != 1.f) return fail(NERF_ERR_INVALID, "tc_gemm_f32: the transposed product accumulates into C (beta = 1)");
        if (M > 256 || N > 256) return fail(NERF_ERR_INVALID, "tc_gemm_f32: transposed product limited to 256 x 256 outputs");
        if (N <= 4) {
            const int grid = (int)(ceil_div(K, 256) < 16 * num_sms() ? ceil_div(K, 256) : 16 * num_sms());     // ~256 samples per block
            small_tn_kernel<<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, (int64_t)K, (int)M, N);
            NERF_LAUNCHED();
            return NERF_OK;
        }
        TnParams P = {A, lda, B, ldb, C, ldc, (int64_t)K, (int)M, N};
        const int64_t tiles = ceil_div(K, 64);           // 64-sample half-tiles
        const int NH = (N + 127) / 128;
        int slabs = num_sms() / NH;
        if (slabs > tiles) slabs = (int)tiles;
        gemm_tn_kernel<<<dim3(slabs, NH), GT_THREADS, TN_SMEM, st>>>(P);
        NERF_LAUNCHED();
        return NERF_OK;
    }
    if (N <= 4) {
        small_rows_dot_kernel<<<stream_grid(M * 32, 256), 256, 0, st>>>(A, lda, B, ldb, tb ? 1 : 0, C, ldc, M, N, K, beta);
        NERF_LAUNCHED();
        return NERF_OK;
    }
    if (K <= 4) {
        small_rows_outer_kernel<<<stream_grid(M * N, 256), 256, 0, st>>>(A, lda, B, ldb, tb ? 1 : 0, C, ldc, M, N, K, beta);
        NERF_LAUNCHED();
        return NERF_OK;
    }
    if (N > 256 || K > 256) return fail(NERF_ERR_INVALID, "tc_gemm_f32: N and K are limited to 256");
    const int NH = (N + 127) / 128, KB = (K + 63) / 64;
    RowsParams P = {A, lda, g_b_scratch, C, ldc, M, N, K, NH, KB, beta};
    const int64_t tiles = ceil_div(M, 128);
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    if (precise) {
        pack_b_split_kernel<3><<<NH * KB * 4, 256, 0, st>>>(B, ldb, tb ? 1 : 0, N, K, NH, KB, g_b_scratch);
        NERF_LAUNCHED();
        gemm_rows_kernel<3><<<grid, GR_THREADS, GR_SMEM, st>>>(P);
    } else {
        pack_b_split_kernel<2><<<NH * KB * 4, 256, 0, st>>>(B, ldb, tb ? 1 : 0, N, K, NH, KB, g_b_scratch);
        NERF_LAUNCHED();
        gemm_rows_kernel<2><<<grid, GR_THREADS, GR_SMEM, st>>>(P);
    }
    NERF_LAUNCHED();
    return NERF_OK;
}

}  // namespace nerf

// test hook (nerf_b200_debug.h): the GEMM above on caller buffers
extern "C" int nerf_selftest_gemm_f32(int ta, int tb, int64_t M, int N, int K, const float* A, int64_t lda, const float* B,
                                      int64_t ldb, float beta, float* C, int64_t ldc, int precise, void* stream) {
    NERF_CHECK_ARG(A && B && C, "null pointer");
    return nerf::tc_gemm_f32((cudaStream_t)stream, ta != 0, tb != 0, M, N, K, A, lda, B, ldb, beta, C, ldc, precise != 0);
}
