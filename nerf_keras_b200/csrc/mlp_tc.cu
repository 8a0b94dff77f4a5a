// Fused NeRF MLP forward on Blackwell tensor cores (tcgen05 + TMEM), sm_100a only.
//
// One persistent CTA per SM processes pairs of 128-sample sub-tiles (A, B).  Per sub-tile the whole
// network of models.py:24-62 runs without leaving the SM:
//   positional encoding (data_utils.py:7-21, fused sample_rays :68-70) is computed in registers and
//   written as the bf16 A-operand tile of the first GEMM; every Dense layer is a chain of
//   tcgen05.mma (M=128, N=128, K=16, bf16 in / fp32 accumulate in TMEM) over weight chunks that a
//   producer warp streams from L2 with bulk async copies (TMA engine) through a 5-stage ring; the
//   epilogue warps read the accumulator with tcgen05.ld, add bias, apply ReLU, round to bf16 and
//   write the next layer's A-operand tile straight back into shared memory (128B-swizzled,
//   K-major).  The two sub-tiles share every weight chunk (halving L2->SMEM traffic) and are kept
//   a few chunks apart so that one sub-tile's epilogue overlaps the other's MMAs.
//   sigma (256->1) and rgb (128->3) heads are fp32 dot products in the epilogues of the layers that
//   feed them; the direction encoding is constant along a ray, so its contribution to the `ddir`
//   layer (models.py:48-54) is hoisted into a per-ray fp32 bias computed by dirbias_kernel.
//
// Shared memory (225.4 KB): 2 x 64 KB activation tiles, 5 x 16 KB weight ring, 12 KB fp32 side
// table (biases + head weights), mbarriers.  TMEM: 2 x 256 fp32 columns (all 512).
#include "common.cuh"
#include "ctx.cuh"
#include "mlp_tc_common.cuh"
#include "mlp_tc_fwd.cuh"

using namespace nerf;
using namespace tc5;
using namespace tcmlp;

namespace {

__constant__ int c_fwd_first[N_PHASES] = {0, 4, 12, 20, 28, 36, 44, 48, 56, 64, 72};
__constant__ int c_fwd_chunk_phase[N_CHUNKS] = {
    0, 0, 0, 0,
    1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4,
    5, 5, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6,
    7, 7, 7, 7, 7, 7, 7, 7, 8, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 9, 9, 10, 10, 10, 10};


// ------------------------------------------------------------------------------------------------
// weight packing: fp32 master blob -> bf16 chunk stream (K-major B operand, B[n][k] = W[k][n])
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fwd_weight_at(const float* __restrict__ blob, const BlobOffsets& off, int phase,
                                               int n, int k) {
    // returns W_phase[k][n] with zero padding; k in [64,128) of the encoding phases is the residual block
    switch (phase) {
        case 0:
            if (k >= 64) k -= 64, k = (k < 3) ? k : 1 << 20;
            return (k < ENC_X) ? blob[off.w[0] + (int64_t)k * H + n] : 0.f;
        case 1: case 2: case 3: case 4: return blob[off.w[phase] + (int64_t)k * H + n];
        case 5: return blob[off.w[5] + (int64_t)k * H + n];
        case 6:
            if (k >= 64) k -= 64, k = (k < 3) ? k : 1 << 20;
            return (k < ENC_X) ? blob[off.w[5] + (int64_t)(H + k) * H + n] : 0.f;
        case 7: return blob[off.w[6] + (int64_t)k * H + n];
        case 8: return blob[off.w[7] + (int64_t)k * H + n];
        case 9: return blob[off.w[9] + (int64_t)k * H + n];                 // feature
        default: return blob[off.w[10] + (int64_t)k * (H / 2) + n];         // ddir rows 0..255, n < 128
    }
}

struct PackFwdArgs {
    const float* blob[2];
    __nv_bfloat16* chunks[2];
    float* side[2];
    nerf_dev_state* tick;      // optional: the optimiser step counter is bumped here (this launch follows the Adam kernel)
};
__global__ void __launch_bounds__(256) pack_fwd_kernel(const PackFwdArgs A, BlobOffsets off) {
    const float* __restrict__ blob = A.blob[blockIdx.y];
    __nv_bfloat16* __restrict__ chunks = A.chunks[blockIdx.y];
    float* __restrict__ side = A.side[blockIdx.y];
    if (A.tick && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) A.tick->step += 1ull;
    // one thread per 16-byte group: chunk g, row n (0..127), k-group kg (0..7)
    const int total = N_CHUNKS * 128 * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int g = i / (128 * 8);
        int rem = i - g * 128 * 8;
        int n = rem >> 3, kg = rem & 7;
        int phase = c_fwd_chunk_phase[g];
        int j = g - c_fwd_first[phase];
        int kbs = c_fwd_prog.kb[phase];
        int h = j / kbs, kb = j - h * kbs;
        uint32_t packed[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int k0 = kb * 64 + kg * 8 + 2 * q;
            float a = fwd_weight_at(blob, off, phase, h * 128 + n, k0);
            float b = fwd_weight_at(blob, off, phase, h * 128 + n, k0 + 1);
            packed[q] = pack_bf16x2(a, b);
        }
        uint8_t* dst = reinterpret_cast<uint8_t*>(chunks) + (size_t)g * CHUNK_BYTES + sw128_offset(n, kg * 8);
        *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
    // side table
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < SIDE_FLOATS; i += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (i < SIDE_BFEAT) v = blob[off.b[i >> 8] + (i & 255)];
        else if (i < SIDE_BDDIR) v = blob[off.b[9] + (i - SIDE_BFEAT)];
        else if (i < SIDE_WSIG) v = blob[off.b[10] + (i - SIDE_BDDIR)];
        else if (i < SIDE_WRGB) v = blob[off.w[8] + (i - SIDE_WSIG)];                    // sigma W (256,1)
        else if (i < SIDE_BSIG) { int q = i - SIDE_WRGB; int ch = q >> 7, k = q & 127; v = blob[off.w[11] + k * 3 + ch]; }
        else if (i == SIDE_BSIG) v = blob[off.b[8]];
        else if (i < SIDE_BRGB + 3) v = blob[off.b[11] + (i - SIDE_BRGB)];
        side[i] = v;
    }
}

// per-ray fp32 bias of the ddir layer: dirbias[ray][j] = sum_k enc_dir(d_ray)[k] * Wddir[256+k][j]
struct DirbiasArgs {
    const float* wddir[2];   // (283,128)
    const float* bddir[2];   // (128)
    float* out[2];           // (rays,128)
};
__global__ void __launch_bounds__(128) dirbias_kernel(const float* __restrict__ d, int64_t rays, const DirbiasArgs A) {
    const float* __restrict__ wddir = A.wddir[blockIdx.y];
    const float* __restrict__ bddir = A.bddir[blockIdx.y];
    float* __restrict__ dirbias = A.out[blockIdx.y];
    __shared__ float enc[ENC_D];
    for (int64_t ray = blockIdx.x; ray < rays; ray += gridDim.x) {
        if (threadIdx.x < ENC_D) {
            int c = threadIdx.x;
            float v;
            if (c < 3) v = d[ray * 3 + c];
            else {
                int q = c - 3, i = q / 6, r = q - i * 6, comp = r % 3;
                float arg = __fmul_rn(exp2f((float)i), d[ray * 3 + comp]);
                v = (r >= 3) ? cosf(arg) : sinf(arg);
            }
            enc[c] = v;
        }
        __syncthreads();
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < ENC_D; ++k) acc = fmaf(enc[k], wddir[(int64_t)(H + k) * (H / 2) + threadIdx.x], acc);
        dirbias[ray * (H / 2) + threadIdx.x] = acc + bddir[threadIdx.x];
        __syncthreads();
    }
}


// Training variant: how the saved operand images reach HBM.  false (default): the tile is bulk-stored (TMA engine)
// from shared memory after each epilogue; that costs another 64 KB of shared-memory reads per sub-tile and phase plus a
// wait + barrier before the tile may be overwritten.  true: every epilogue thread writes its own row pieces straight
// from registers with 256-bit global stores (two swizzled 16-byte chunks = one 32-byte sector).  Measured on B200,
// forward of both nets per 4096-ray step: bulk stores 1.85 ms, 256-bit register stores 2.06 ms, 128-bit register
// stores 2.44 ms -- the LSU path loses to the TMA engine even though it spares the shared-memory reads.
#ifdef NERF_EXP_COALESCED_SAVE
// TIMING EXPERIMENT (round 2; the weight-gradient kernel does not read this layout): images as [16-byte column group][row],
// so that a warp's 128-bit register stores cover 512 contiguous bytes instead of 32 different 128-byte lines.  Measured
// 1.91 ms per step (forward of both nets) against 1.57-1.75 ms with bulk stores: coalescing is not what the LSU path lacks.
constexpr bool kDirectSave = true;
constexpr bool kCoalescedSave = true;
#else
constexpr bool kDirectSave = false;
constexpr bool kCoalescedSave = false;
#endif
// The three named barriers of a saving trunk epilogue (previous store has read the tile / tile halves complete before the
// elected thread issues their bulk stores) cost 0.035 ms per step of forward (timing experiment without them, round 2):
// a dedicated store-issuing thread fed by mbarriers could win back at most that much.
constexpr bool kSplitImageStore = true;
constexpr bool kHybridSave = false;       // K-blocks 0,1 by 256-bit register stores during the first epilogue half, K-blocks 2,3 by
                                          // one bulk store: measured 1.90 ms per step vs 1.65 with two bulk stores -- off   // bulk-store K-blocks 0,1 as soon as they are written (two 32 KB stores per phase)
__device__ __forceinline__ void st_global_v4(uint64_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Two adjacent 16-byte chunks (2j, 2j+1) of a swizzled row share one 32-byte sector (the XOR only swaps them on odd
// rows): one 256-bit store writes the whole sector.  pk8 = chunk 2j then chunk 2j+1; off_c0 = address of chunk 2j.
__device__ __forceinline__ void st_global_pair(uint64_t gimg, uint32_t off_c0, bool odd_row, const uint32_t* pk8) {
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        lo[i] = odd_row ? pk8[4 + i] : pk8[i];
        hi[i] = odd_row ? pk8[i] : pk8[4 + i];
    }
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(gimg + (off_c0 & ~16u)), "r"(lo[0]),
                 "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3])
                 : "memory");
}

// ---- epilogue building blocks ---------------------------------------------------------------------
// one 32-column group of a trunk / feature layer: acc + bias (packed fp32x2 adds), fused ReLU + bf16
// convert, optional sigma head accumulation and ReLU mask, then four 16-byte swizzled stores.
template <bool RELU, bool SIGMA, bool SAVE, int CG, bool STORE>
__device__ __forceinline__ void trunk_group(const uint32_t (&v)[32], const float* bias, const float* wsig,
                                            uint64_t& sig2, uint32_t& mk, const RowStore& rs, uint32_t* held,
                                            uint64_t gimg /* image base minus the tile's shared address */) {
    const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(bias + CG * 32);
    const ulonglong2* s2 = reinterpret_cast<const ulonglong2*>(wsig + CG * 32);
    uint32_t pk[16];
    uint32_t neg = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const ulonglong2 bb = b2[q];
        uint64_t x01 = f2_add(f2_pack(v[4 * q], v[4 * q + 1]), bb.x);
        uint64_t x23 = f2_add(f2_pack(v[4 * q + 2], v[4 * q + 3]), bb.y);
        float x0, x1, x2, x3;
        f2_unpack(x01, x0, x1);
        f2_unpack(x23, x2, x3);
        if (SAVE && RELU) neg = push_signs(neg, x0, x1, x2, x3);
        if (SIGMA) {
            x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f);
            const ulonglong2 ws = s2[q];
            sig2 = f2_fma(f2_pack(__float_as_uint(x0), __float_as_uint(x1)), ws.x, sig2);
            sig2 = f2_fma(f2_pack(__float_as_uint(x2), __float_as_uint(x3)), ws.y, sig2);
        }
        pk[2 * q] = cvt_bf16x2<RELU>(x0, x1);
        pk[2 * q + 1] = cvt_bf16x2<RELU>(x2, x3);
    }
    mk = signs_to_mask(neg);
    if (STORE) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
            rs.store<CG / 2>((CG & 1) * 4 + c, pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
    } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) held[CG * 16 + q] = pk[q];
    }
    if (SAVE && kCoalescedSave) {
        const uint32_t row = threadIdx.x & 127;
        const uint64_t b = gimg + (rs.off[0] & ~127u) - ((row >> 3) * 1024 + (row & 7) * 128) + (CG / 2) * (TILE_M * 128) + row * 16;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            st_global_v4(b + ((CG & 1) * 4 + c) * 2048, pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
    } else if (SAVE && (kDirectSave || (kHybridSave && !STORE))) {
        const bool odd = (rs.off[0] >> 7) & 1;                 // row parity (bit 7 of the row's shared-memory address)
#pragma unroll
        for (int j = 0; j < 2; ++j)
            st_global_pair(gimg + (CG / 2) * (TILE_M * 128), rs.off[(CG & 1) * 4 + 2 * j], odd, pk + 8 * j);
    }
}

// First half of a trunk / feature epilogue: columns 0..127 (runs while the MMAs of columns 128..255 are still in
// flight).  The converted bf16 values are HELD in registers: K-blocks 0,1 of the A tile are still being read.
template <bool RELU, bool SIGMA, bool SAVE>
__device__ __forceinline__ void trunk_part1(uint32_t t_lane, const float* bias, const float* wsig, uint64_t& sig2,
                                            uint32_t (&mask)[8], const RowStore& rs, uint32_t (&held)[64], uint64_t gimg) {
    uint32_t v[32];
    tmem_ld32(t_lane, v); tmem_ld_wait();
    trunk_group<RELU, SIGMA, SAVE, 0, false>(v, bias, wsig, sig2, mask[0], rs, held, gimg);
    tmem_ld32(t_lane + 32, v); tmem_ld_wait();
    trunk_group<RELU, SIGMA, SAVE, 1, false>(v, bias, wsig, sig2, mask[1], rs, held, gimg);
    tmem_ld32(t_lane + 64, v); tmem_ld_wait();
    trunk_group<RELU, SIGMA, SAVE, 2, false>(v, bias, wsig, sig2, mask[2], rs, held, gimg);
    tmem_ld32(t_lane + 96, v); tmem_ld_wait();
    trunk_group<RELU, SIGMA, SAVE, 3, false>(v, bias, wsig, sig2, mask[3], rs, held, gimg);
}
// all MMAs of the phase are complete: K-blocks 0,1 may be overwritten with the held first half
__device__ __forceinline__ void store_held(const RowStore& rs, const uint32_t (&held)[64]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) rs.store<0>(c, held[4 * c], held[4 * c + 1], held[4 * c + 2], held[4 * c + 3]);
#pragma unroll
    for (int c = 0; c < 8; ++c) rs.store<1>(c, held[32 + 4 * c], held[32 + 4 * c + 1], held[32 + 4 * c + 2], held[32 + 4 * c + 3]);
}
// second half: columns 128..255 straight into K-blocks 2,3 (TMEM loads pipelined one group ahead)
template <bool RELU, bool SIGMA, bool SAVE>
__device__ __forceinline__ void trunk_part2(uint32_t t_lane, const float* bias, const float* wsig, uint64_t& sig2,
                                            uint32_t (&mask)[8], const RowStore& rs, uint64_t gimg) {
    uint32_t va[32], vb[32];
    tmem_ld32(t_lane + 128, va);
    tmem_ld_wait();
    tmem_ld32(t_lane + 160, vb);
    trunk_group<RELU, SIGMA, SAVE, 4, true>(va, bias, wsig, sig2, mask[4], rs, nullptr, gimg);
    tmem_ld_wait();
    tmem_ld32(t_lane + 192, va);
    trunk_group<RELU, SIGMA, SAVE, 5, true>(vb, bias, wsig, sig2, mask[5], rs, nullptr, gimg);
    tmem_ld_wait();
    tmem_ld32(t_lane + 224, vb);
    trunk_group<RELU, SIGMA, SAVE, 6, true>(va, bias, wsig, sig2, mask[6], rs, nullptr, gimg);
    tmem_ld_wait();
    trunk_group<RELU, SIGMA, SAVE, 7, true>(vb, bias, wsig, sig2, mask[7], rs, nullptr, gimg);
}

// whole accumulator (256 columns) in one pass, TMEM loads pipelined one group ahead, direct stores: for the CTA-pair
// kernel, where the accumulator is handed over complete (no MMA of the phase still reads the A tile)
template <bool RELU, bool SIGMA, bool SAVE>
__device__ __forceinline__ void trunk_full(uint32_t t_lane, const float* bias, const float* wsig, uint64_t& sig2,
                                           uint32_t (&mask)[8], const RowStore& rs, uint64_t gimg) {
    uint32_t va[32], vb[32];
    tmem_ld32(t_lane, va);
    tmem_ld_wait();
    tmem_ld32(t_lane + 32, vb);
    trunk_group<RELU, SIGMA, SAVE, 0, true>(va, bias, wsig, sig2, mask[0], rs, nullptr, gimg);
    tmem_ld_wait();
    tmem_ld32(t_lane + 64, va);
    trunk_group<RELU, SIGMA, SAVE, 1, true>(vb, bias, wsig, sig2, mask[1], rs, nullptr, gimg);
    tmem_ld_wait();
    tmem_ld32(t_lane + 96, vb);
    trunk_group<RELU, SIGMA, SAVE, 2, true>(va, bias, wsig, sig2, mask[2], rs, nullptr, gimg);
    tmem_ld_wait();
    tmem_ld32(t_lane + 128, va);
    trunk_group<RELU, SIGMA, SAVE, 3, true>(vb, bias, wsig, sig2, mask[3], rs, nullptr, gimg);
    tmem_ld_wait();
    tmem_ld32(t_lane + 160, vb);
    trunk_group<RELU, SIGMA, SAVE, 4, true>(va, bias, wsig, sig2, mask[4], rs, nullptr, gimg);
    tmem_ld_wait();
    tmem_ld32(t_lane + 192, va);
    trunk_group<RELU, SIGMA, SAVE, 5, true>(vb, bias, wsig, sig2, mask[5], rs, nullptr, gimg);
    tmem_ld_wait();
    tmem_ld32(t_lane + 224, vb);
    trunk_group<RELU, SIGMA, SAVE, 6, true>(va, bias, wsig, sig2, mask[6], rs, nullptr, gimg);
    tmem_ld_wait();
    trunk_group<RELU, SIGMA, SAVE, 7, true>(vb, bias, wsig, sig2, mask[7], rs, nullptr, gimg);
}

// one 32-column group of the ddir epilogue: + (bias + per-ray direction bias), ReLU, rgb head dot products
template <bool SAVE, int CG>
__device__ __forceinline__ void ddir_group(const uint32_t (&v)[32], const float* dbias /* this ray, 128 floats */,
                                           const float* side, uint64_t& r2, uint64_t& g2, uint64_t& b2acc, uint32_t& mk,
                                           const RowStore& rs, uint64_t gimg) {
    const ulonglong2* d2 = reinterpret_cast<const ulonglong2*>(dbias + CG * 32);
    const ulonglong2* wr = reinterpret_cast<const ulonglong2*>(side + SIDE_WRGB + CG * 32);
    const ulonglong2* wg = reinterpret_cast<const ulonglong2*>(side + SIDE_WRGB + 128 + CG * 32);
    const ulonglong2* wb = reinterpret_cast<const ulonglong2*>(side + SIDE_WRGB + 256 + CG * 32);
    uint32_t pk[16];
    uint32_t neg = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const ulonglong2 dd = d2[q];
        float x0, x1, x2, x3;
        f2_unpack(f2_add(f2_pack(v[4 * q], v[4 * q + 1]), dd.x), x0, x1);
        f2_unpack(f2_add(f2_pack(v[4 * q + 2], v[4 * q + 3]), dd.y), x2, x3);
        if (SAVE) neg = push_signs(neg, x0, x1, x2, x3);
        x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f);
        const uint64_t x01 = f2_pack(__float_as_uint(x0), __float_as_uint(x1));
        const uint64_t x23 = f2_pack(__float_as_uint(x2), __float_as_uint(x3));
        const ulonglong2 a = wr[q], b = wg[q], c = wb[q];
        r2 = f2_fma(x01, a.x, r2); r2 = f2_fma(x23, a.y, r2);
        g2 = f2_fma(x01, b.x, g2); g2 = f2_fma(x23, b.y, g2);
        b2acc = f2_fma(x01, c.x, b2acc); b2acc = f2_fma(x23, c.y, b2acc);
        if (SAVE) {
            pk[2 * q] = cvt_bf16x2<false>(x0, x1);
            pk[2 * q + 1] = cvt_bf16x2<false>(x2, x3);
        }
    }
    mk = signs_to_mask(neg);
    if (SAVE) {
#pragma unroll
        if (kDirectSave) {
            const bool odd = (rs.off[0] >> 7) & 1;
#pragma unroll
            for (int j = 0; j < 2; ++j)
                st_global_pair(gimg + (CG / 2) * (TILE_M * 128), rs.off[(CG & 1) * 4 + 2 * j], odd, pk + 8 * j);
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                rs.store<CG / 2>((CG & 1) * 4 + c, pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
    }
}

// positional encoding of one point: accurate sincosf at octaves 0 and 5, exact angle doubling in between
// (sin 2a = 2 sin a cos a, cos 2a = 1 - 2 sin^2 a; at most 4 doublings -> error <= ~2e-6, far below a bf16 ulp)
__device__ __forceinline__ void encode_xyz(const float (&p)[3], float (&e)[64]) {
    e[0] = p[0]; e[1] = p[1]; e[2] = p[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float sv, cv;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            if (i == 0 || i == 5) sincosf((float)(1 << i) * p[c], &sv, &cv);
            else {
                const float s2 = 2.f * sv * cv;
                const float c2 = fmaf(-2.f * sv, sv, 1.f);
                sv = s2; cv = c2;
            }
            e[3 + 6 * i + c] = sv;
            e[3 + 6 * i + 3 + c] = cv;
        }
    }
    e[63] = 0.f;
}

// ------------------------------------------------------------------------------------------------
// the fused forward kernel
// ------------------------------------------------------------------------------------------------
// PAIR = false: one CTA per SM, independent.  PAIR = true: clusters of two CTAs (an SM pair) sharing M = 256 MMAs
// (cta_group::2): grid = 2 x clusters, launched with cudaLaunchAttributeClusterDimension {2,1,1}.
template <bool SAVE, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1) nerf_mlp_fwd_tc_kernel(const FwdParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int rank = PAIR ? (int)cluster_ctarank() : 0;              // 0 = leader of the CTA pair
    // work units: 256-row tile pairs (one per CTA) or 512-row quads (one per cluster)
    const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_units_grid = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int64_t n_units = PAIR ? (P.n_pairs + 1) / 2 : P.n_pairs;
    Barriers B;
    const bool shared_chunks = PAIR && (P.dbg & 2);     // pair mode: one ring fill per chunk, consumed by both sub-tiles
    const bool ws_issue = !PAIR && (P.dbg & 8);         // weight-stationary MMAs: one issuer, B read once for both sub-tiles
    init_barriers(base, B, PAIR, shared_chunks, ws_issue);
    float* side = reinterpret_cast<float*>(smem + SM_SIDE);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM);
    if (warp == 9) {
        if (PAIR) tmem_alloc_2cta(base + SM_TMEM, 512); else tmem_alloc(base + SM_TMEM, 512);
    }
    for (int i = threadIdx.x; i < SIDE_FLOATS; i += NUM_THREADS) side[i] = P.side[i];
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync();                                         // peer barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int my_pairs = (n_units > unit) ? (int)((n_units - unit + n_units_grid - 1) / n_units_grid) : 0;
    const int total_chunks = my_pairs * N_CHUNKS;
    int steps_per_tile = 0;
    for (int ph = 0; ph < N_PHASES; ++ph) steps_per_tile += c_fwd_prog.kb[ph];

    if (warp == 8) {
        if (lane == 0) {
            if (PAIR) producer_loop_pair(base, B, P.w_chunks, c_fwd_prog, rank, my_pairs, shared_chunks);
            else producer_loop(base, B, P.w_chunks, N_CHUNKS, total_chunks);
        }
    } else if (warp >= 9 && !PAIR && ws_issue) {
        // weight-stationary issue: the whole warp runs the loop (uniform datapath), one lane's MMAs are predicated on
        if (P.dbg & 32) {          // two issuer warps, one per N-half
            if (warp == 9) issuer_loop_ws2<0>(base, B, tmem_base, c_fwd_prog, my_pairs, P.trace);
            else issuer_loop_ws2<1>(base, B, tmem_base, c_fwd_prog, my_pairs, P.trace);
        } else if (warp == 9) issuer_loop_ws(base, B, tmem_base, c_fwd_prog, my_pairs, P.trace, (P.dbg & 16) != 0);
    } else if (warp >= 9) {
        if (lane == 0) {
            if (!PAIR) issuer_loop(base, B, tmem_base, c_fwd_prog, warp - 9, my_pairs, P.trace);
            else if (rank == 0 && shared_chunks) issuer_loop_pair_shared(base, B, tmem_base, c_fwd_prog, warp - 9, my_pairs, P.trace);
            else if (rank == 0) issuer_loop_pair(base, B, tmem_base, c_fwd_prog, warp - 9, my_pairs, P.trace, P.dbg, false);
            else if (warp == 9) forwarder_loop_pair(B, my_pairs * steps_per_tile * (shared_chunks ? 1 : 2));
        }
    } else {
        // ===================== workers: PE prologue + epilogues =====================
        const int s = warp >> 2;
        const int row = threadIdx.x - s * TILE_M;
        const uint32_t act_base = base + SM_ACT + s * 65536;
        const uint32_t t_lane = tmem_base + (uint32_t(32 * (warp & 3)) << 16) + s * 256;
        const bool elected = (row == 0);
        float* dbs = reinterpret_cast<float*>(smem + SM_DIRB) + s * (DIRB_ROWS * 128);   // staged per-ray ddir biases
        RowStore rs;
        rs.init(act_base, row);
        const uint32_t bar_lo_l = B.actr + 16 * s;                       // A tile hand-off, K-halves (leader's barriers)
        const uint32_t bar_lo = (PAIR && rank != 0) ? map_to_cta(bar_lo_l, 0) : bar_lo_l;
        const uint32_t bar_hi = (PAIR && rank != 0) ? map_to_cta(bar_lo_l + 8, 0) : bar_lo_l + 8;
        const uint32_t bar_h0 = B.accf + 16 * s, bar_h1 = bar_h0 + 8;    // accumulator hand-off, N-halves
        uint32_t accf_par = 0;
        uint32_t E[32];     // bf16(enc), 64 channels packed
        uint32_t Elo[2];    // bf16 residuals of the raw x, y, z channels
        // pair mode: the writes are read by the async proxy of THIS SM (each SM's tensor core reads its own A rows); the
        // cross-CTA ordering is carried by the cluster-scope release / acquire of the barrier (P.dbg bit 2: cheaper fence)
        const bool cheap_fence = PAIR && (P.dbg & 4);
        auto fence_async = [&]() { if (PAIR && !cheap_fence) fence_proxy_async_all(); else fence_proxy_async_smem(); };
        auto arrive = [&](uint32_t bar) {                                // the peer's workers arrive on the leader's barrier
            if (PAIR && rank != 0) mbar_arrive_cluster(bar); else mbar_arrive(bar);
        };

        for (int it = 0; it < my_pairs; ++it) {
            const int64_t wu = unit + (int64_t)it * n_units_grid;
            const int64_t tile = PAIR ? (wu * 4 + s * 2 + rank) : (wu * 2 + s);
            const int64_t g_row = tile * TILE_M + row;
            const bool valid = g_row < P.M;
            const int64_t gr = valid ? g_row : (P.M - 1);
            const int64_t ray = gr / P.N;
            uint8_t* save_tile = SAVE ? P.act_save + tile * SAVE_TILE_BYTES : nullptr;
            uint32_t* mask_tile = SAVE ? P.mask_save + tile * (MASK_TILE_BYTES / 4) : nullptr;

            // ---- stage the ddir biases (bias + direction term) of the rays this sub-tile touches ----
            const int64_t m_first = tile * TILE_M;
            const int64_t ray0 = ((m_first < P.M) ? m_first : (P.M - 1)) / P.N;
            const int64_t m_last = (m_first + TILE_M - 1 < P.M) ? (m_first + TILE_M - 1) : (P.M - 1);
            const int n_rays = (int)(m_last / P.N - ray0) + 1;
            const bool staged = n_rays <= DIRB_ROWS && n_rays * 32 <= TILE_M;   // one float4 per thread
            float4 db_pref = make_float4(0.f, 0.f, 0.f, 0.f);
            const int db_idx = row;                          // thread i prefetches float4 i of the staged block
            if (staged && db_idx < n_rays * 32)
                db_pref = __ldg(reinterpret_cast<const float4*>(P.dirbias + ray0 * 128) + db_idx);

            // ---- positional encoding of this row's sample point (fp32) ----
            {
                const float tv = P.t[gr];
                float p[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) p[c] = __fadd_rn(P.o[ray * 3 + c], __fmul_rn(P.d[ray * 3 + c], tv));
                float e[64];
                encode_xyz(p, e);
#pragma unroll
                for (int q = 0; q < 32; ++q) E[q] = pack_bf16x2(e[2 * q], e[2 * q + 1]);
                float lo[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) lo[c] = p[c] - __bfloat162float(__float2bfloat16_rn(p[c]));
                Elo[0] = pack_bf16x2(lo[0], lo[1]);
                Elo[1] = pack_bf16x2(lo[2], 0.f);
            }
            // every thread of the sub-tile is past the previous tile's ddir epilogue (last reader of the staged
            // biases), and in training mode the previous tile's last bulk store has finished reading the tile
            if (SAVE && !kDirectSave && elected) bulk_wait_read0();
            named_bar_sync(1 + s, TILE_M);
#pragma unroll
            for (int c = 0; c < 8; ++c) rs.store<0>(c, E[4 * c], E[4 * c + 1], E[4 * c + 2], E[4 * c + 3]);
            rs.store<1>(0, Elo[0], Elo[1], 0u, 0u);
            rs.store<1>(1, 0u, 0u, 0u, 0u);
            if (staged && db_idx < n_rays * 32) reinterpret_cast<float4*>(dbs)[db_idx] = db_pref;
            tc_fence_before();
            fence_async();
            if (SAVE && kDirectSave) {
                const uint64_t genc = reinterpret_cast<uint64_t>(save_tile + SAVE_ENC) - act_base;
#pragma unroll
                for (int j = 0; j < 4; ++j) st_global_pair(genc, rs.off[2 * j], (rs.off[0] >> 7) & 1, E + 8 * j);
            } else if (SAVE) {
                named_bar_sync(1 + s, TILE_M);
                if (elected) { bulk_s2g_stream(save_tile + SAVE_ENC, act_base, 16384); bulk_commit(); }
            }
            arrive(bar_lo);
            arrive(bar_hi);

            float sig = 0.f;
            for (int ph = 0; ph < N_PHASES; ++ph) {
                if (elected) trace_ev(P.trace, 2 + s, it, ph, 0);      // worker: starts waiting for the accumulator
                mbar_wait(bar_h0, accf_par, 2);
                tc_fence_after();
                if (elected) trace_ev(P.trace, 2 + s, it, ph, 1);      // worker: first accumulator half ready
                if (ph == 5) {
                    // L5a done: stage the skip-connection encoding as K-blocks 0 (hi) and 1 (residual) for L5b
                    mbar_wait(bar_h1, accf_par, 6);
                    accf_par ^= 1;
                    tc_fence_after();
                    if (SAVE && !kDirectSave) {
                        if (elected) bulk_wait_read0();
                        named_bar_sync(1 + s, TILE_M);
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) rs.store<0>(c, E[4 * c], E[4 * c + 1], E[4 * c + 2], E[4 * c + 3]);
                    rs.store<1>(0, Elo[0], Elo[1], 0u, 0u);
                    rs.store<1>(1, 0u, 0u, 0u, 0u);
                    tc_fence_before();
                    fence_async();
                    arrive(bar_lo);
                    arrive(bar_hi);
                    continue;
                }
                if (ph < 10) {
                    // trunk layer / feature epilogue: bias (+ReLU) -> bf16 -> next A tile, in two column halves
                    const int layer = (ph <= 4) ? ph : (ph == 6 ? 5 : (ph == 7 ? 6 : (ph == 8 ? 7 : 8)));
                    const float* bias = side + (layer < 8 ? SIDE_BIAS + layer * H : SIDE_BFEAT);
                    const bool relu = layer < 8;
                    const float* wsig = side + SIDE_WSIG;
                    uint32_t mask[8];
                    uint32_t held[64];
                    uint64_t sig2 = 0ull;
                    const uint64_t gimg = SAVE ? reinterpret_cast<uint64_t>(save_tile + ((layer < 8) ? SAVE_H + 65536 * layer
                                                                                                    : SAVE_FEAT)) - act_base
                                               : 0ull;
                    if (PAIR && (P.dbg & 2)) {
                        // pair kernel with shared chunks: the accumulator arrives complete -> one pass, one fence, one hand-off
                        mbar_wait(bar_h1, accf_par, 6);
                        accf_par ^= 1;
                        tc_fence_after();
                        if (SAVE && !kDirectSave) {
                            if (elected) bulk_wait_read0();
                            named_bar_sync(1 + s, TILE_M);
                        }
                        if (ph == 8) trunk_full<true, true, SAVE>(t_lane, bias, wsig, sig2, mask, rs, gimg);
                        else if (relu) trunk_full<true, false, SAVE>(t_lane, bias, wsig, sig2, mask, rs, gimg);
                        else trunk_full<false, false, SAVE>(t_lane, bias, wsig, sig2, mask, rs, gimg);
                        if (ph == 8) {
                            float a, b;
                            f2_unpack(sig2, a, b);
                            sig = a + b;
                        }
                        tc_fence_before();
                        fence_async();
                        if (SAVE) {
                            if (relu) {
                                uint4* mp = reinterpret_cast<uint4*>(mask_tile + ((size_t)layer * 128 + row) * 8);
                                mp[0] = make_uint4(mask[0], mask[1], mask[2], mask[3]);
                                mp[1] = make_uint4(mask[4], mask[5], mask[6], mask[7]);
                            }
                            if (!kDirectSave) {
                                named_bar_sync(1 + s, TILE_M);
                                if (elected) {
                                    int64_t off = (layer < 8) ? SAVE_H + 65536 * layer : SAVE_FEAT;
                                    bulk_s2g_stream(save_tile + off, act_base, 32768);
                                    bulk_s2g_stream(save_tile + off + 32768, act_base + 32768, 32768);
                                    bulk_commit();
                                }
                            }
                        }
                        if (elected) trace_ev(P.trace, 2 + s, it, ph, 2);
                        arrive(bar_lo);
                        arrive(bar_hi);
                        continue;
                    }
                    if (ph == 8) trunk_part1<true, true, SAVE>(t_lane, bias, wsig, sig2, mask, rs, held, gimg);
                    else if (relu) trunk_part1<true, false, SAVE>(t_lane, bias, wsig, sig2, mask, rs, held, gimg);
                    else trunk_part1<false, false, SAVE>(t_lane, bias, wsig, sig2, mask, rs, held, gimg);
                    mbar_wait(bar_h1, accf_par, 6);                    // every MMA of the phase is complete
                    accf_par ^= 1;
                    tc_fence_after();
                    if (SAVE && !kDirectSave) {
                        if (elected) bulk_wait_read0();
                        named_bar_sync(1 + s, TILE_M);
                    }
                    store_held(rs, held);
                    tc_fence_before();
                    fence_async();
                    arrive(bar_lo);                                    // next phase may start on K-blocks 0,1
                    if (SAVE && !kDirectSave && kSplitImageStore && !kHybridSave) {
                        // first half of the image (K-blocks 0,1) leaves now, under the second half of the epilogue
                        named_bar_sync(1 + s, TILE_M);
                        if (elected) {
                            int64_t off = (layer < 8) ? SAVE_H + 65536 * layer : SAVE_FEAT;
                            bulk_s2g_stream(save_tile + off, act_base, 32768);
                            bulk_commit();
                        }
                    }
                    if (ph == 8) trunk_part2<true, true, SAVE>(t_lane, bias, wsig, sig2, mask, rs, gimg);
                    else if (relu) trunk_part2<true, false, SAVE>(t_lane, bias, wsig, sig2, mask, rs, gimg);
                    else trunk_part2<false, false, SAVE>(t_lane, bias, wsig, sig2, mask, rs, gimg);
                    if (ph == 8) {
                        float a, b;
                        f2_unpack(sig2, a, b);
                        sig = a + b;
                    }
                    tc_fence_before();
                    fence_async();
                    if (SAVE) {
                        if (relu) {
                            uint4* mp = reinterpret_cast<uint4*>(mask_tile + ((size_t)layer * 128 + row) * 8);
                            mp[0] = make_uint4(mask[0], mask[1], mask[2], mask[3]);
                            mp[1] = make_uint4(mask[4], mask[5], mask[6], mask[7]);
                        }
                        if (!kDirectSave) {
                            named_bar_sync(1 + s, TILE_M);
                            if (elected) {
                                int64_t off = (layer < 8) ? SAVE_H + 65536 * layer : SAVE_FEAT;
                                if (kSplitImageStore) bulk_s2g_stream(save_tile + off + 32768, act_base + 32768, 32768);
                                else bulk_s2g_stream(save_tile + off, act_base, 65536);
                                bulk_commit();
                            }
                        }
                    }
                    if (elected) trace_ev(P.trace, 2 + s, it, ph, 2);  // worker: epilogue done
                    arrive(bar_hi);
                } else {
                    // ddir epilogue: + (bias + per-ray direction bias), ReLU, rgb head (fp32), write preds
                    mbar_wait(bar_h1, accf_par, 6);
                    accf_par ^= 1;
                    if (SAVE && !kDirectSave) {
                        if (elected) bulk_wait_read0();
                        named_bar_sync(1 + s, TILE_M);
                    }
                    const uint64_t ghd = SAVE ? reinterpret_cast<uint64_t>(save_tile + SAVE_HD) - act_base : 0ull;
                    uint64_t r2 = 0ull, g2 = 0ull, b2 = 0ull;
                    uint32_t mask[4];
                    // staged copy lives in shared memory; ragged tiles with many short rays fall back to global
                    const float* db = staged ? (dbs + (int)(ray - ray0) * 128) : (P.dirbias + ray * 128);
                    uint32_t va[32], vb[32];
                    tmem_ld32(t_lane, va);
                    tmem_ld_wait();
                    tmem_ld32(t_lane + 32, vb);
                    ddir_group<SAVE, 0>(va, db, side, r2, g2, b2, mask[0], rs, ghd);
                    tmem_ld_wait();
                    tmem_ld32(t_lane + 64, va);
                    ddir_group<SAVE, 1>(vb, db, side, r2, g2, b2, mask[1], rs, ghd);
                    tmem_ld_wait();
                    tmem_ld32(t_lane + 96, vb);
                    ddir_group<SAVE, 2>(va, db, side, r2, g2, b2, mask[2], rs, ghd);
                    tmem_ld_wait();
                    ddir_group<SAVE, 3>(vb, db, side, r2, g2, b2, mask[3], rs, ghd);
                    float ra, rb, ga, gb, ba, bb;
                    f2_unpack(r2, ra, rb); f2_unpack(g2, ga, gb); f2_unpack(b2, ba, bb);
                    if (elected) trace_ev(P.trace, 2 + s, it, ph, 2);
                    if (valid)
                        P.preds[g_row] = make_float4(ra + rb + side[SIDE_BRGB], ga + gb + side[SIDE_BRGB + 1],
                                                     ba + bb + side[SIDE_BRGB + 2], sig + side[SIDE_BSIG]);
                    tc_fence_before();
                    if (SAVE) {
                        uint4* mp = reinterpret_cast<uint4*>(mask_tile + ((size_t)8 * 128 + row) * 8);
                        mp[0] = make_uint4(mask[0], mask[1], mask[2], mask[3]);
                        mp[1] = make_uint4(0u, 0u, 0u, 0u);
                        if (!kDirectSave) {
                            fence_async();
                            named_bar_sync(1 + s, TILE_M);
                            if (elected) { bulk_s2g_stream(save_tile + SAVE_HD, act_base, 32768); bulk_commit(); }
                        }
                    }
                }
            }
        }
        if (SAVE && !kDirectSave && elected) bulk_wait_all0();
    }

    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync();            // the leader's MMAs read the peer's shared memory: leave together
    if (warp == 9) {
        if (PAIR) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}
}  // namespace

namespace nerf {

int tc_experimental_init();                                              // mlp_tc_experimental.cu
int tc_forward_ts(const tcmlp::FwdParams& P, bool save_acts, cudaStream_t st);

long long* g_trace_buf = nullptr;
int g_pair_mode = 0;   // nerf_debug_pair_mode: 1 = cta_group::2 kernel, 4 = cta_group::2 kernel with TMEM-resident activations

int tc_supported(const nerf_config& c, std::string* why) {
    if (c.num_layers != 8 || c.hidden_dim != 256 || c.skip_layer != 4 || c.l_xyz != 10 || c.l_dir != 4) {
        if (why) *why = "tcgen05 path is specialised to NUM_LAYERS=8, HIDDEN_DIM=256, SKIP_LAYER=4, L_XYZ=10, L_DIR=4";
        return 0;
    }
    return 1;
}

int tc_alloc(nerf_ctx* ctx) {
    for (int net = 0; net < 2; ++net) {
        NERF_CUDA(cudaMalloc(&ctx->w_fwd[net], (size_t)N_CHUNKS * CHUNK_BYTES));
        NERF_CUDA(cudaMalloc(&ctx->side[net], SIDE_FLOATS * sizeof(float)));
    }
    NERF_CUDA(cudaFuncSetAttribute(nerf_mlp_fwd_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    NERF_CUDA(cudaFuncSetAttribute(nerf_mlp_fwd_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    NERF_CUDA(cudaFuncSetAttribute(nerf_mlp_fwd_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    NERF_CUDA(cudaFuncSetAttribute(nerf_mlp_fwd_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    if (int rc = tc_experimental_init()) return rc;
    return NERF_OK;
}

void tc_free(nerf_ctx* ctx) {
    for (int net = 0; net < 2; ++net) {
        cudaFree(ctx->w_fwd[net]);
        cudaFree(ctx->w_bwd[net]);
        cudaFree(ctx->side[net]);
    }
}

static int launch_pack_fwd(nerf_ctx* ctx, int first, int count, bool tick, cudaStream_t st) {
    BlobOffsets off = make_offsets(ctx);
    PackFwdArgs A = {};
    for (int i = 0; i < count; ++i) {
        const int net = first + i;
        A.blob[i] = ctx->params + (int64_t)net * ctx->n_params;
        A.chunks[i] = ctx->w_fwd[net];
        A.side[i] = ctx->side[net];
    }
    A.tick = tick ? ctx->dev_state : nullptr;
    pack_fwd_kernel<<<dim3(num_sms() / 2, count), 256, 0, st>>>(A, off);
    NERF_LAUNCHED();
    for (int i = 0; i < count; ++i) ctx->packed_valid[first + i] = true;
    return NERF_OK;
}

int tc_pack_weights(nerf_ctx* ctx, int net, cudaStream_t st) { return launch_pack_fwd(ctx, net, 1, false, st); }
int tc_pack_all(nerf_ctx* ctx, bool tick_step, cudaStream_t st) { return launch_pack_fwd(ctx, 0, 2, tick_step, st); }

// per-ray ddir biases of the nets in `nets_mask` (bit 0 coarse, bit 1 fine), one launch
int tc_dirbias(nerf_ctx* ctx, const float* d, int64_t B, int nets_mask, cudaStream_t st) {
    DirbiasArgs A = {};
    int count = 0;
    for (int net = 0; net < 2; ++net) {
        if (!(nets_mask & (1 << net))) continue;
        const float* blob = ctx->params + (int64_t)net * ctx->n_params;
        A.wddir[count] = blob + ctx->layers[10].w_off;
        A.bddir[count] = blob + ctx->layers[10].b_off;
        A.out[count] = ctx->fw_dirbias[net];
        ++count;
    }
    if (count == 0 || B == 0) return NERF_OK;
    dirbias_kernel<<<dim3((unsigned)(B < 4 * num_sms() ? B : 4 * num_sms()), count), 128, 0, st>>>(d, B, A);
    NERF_LAUNCHED();
    return NERF_OK;
}

int tc_forward_rays(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t, int64_t B, int N,
                    float* preds, bool save_acts, cudaStream_t st, bool dirbias_ready) {
    if (!ctx->packed_valid[net]) {
        int rc = tc_pack_weights(ctx, net, st);
        if (rc) return rc;
    }
    if (B > ctx->cfg.max_rays) return fail(NERF_ERR_INVALID, "tc_forward_rays: batch exceeds cfg.max_rays");
    if (!dirbias_ready) {
        int rc = tc_dirbias(ctx, d, B, 1 << net, st);
        if (rc) return rc;
    }
    FwdParams P;
    P.o = o; P.d = d; P.t = t; P.N = N;
    P.M = B * (int64_t)N;
    P.n_pairs = ceil_div(P.M, 2 * TILE_M);
    P.w_chunks = ctx->w_fwd[net];
    P.side = ctx->side[net];
    P.dirbias = ctx->fw_dirbias[net];
    P.preds = reinterpret_cast<float4*>(preds);
    P.act_save = save_acts ? reinterpret_cast<uint8_t*>(ctx->act_save[net]) : nullptr;
    P.mask_save = save_acts ? ctx->mask_save[net] : nullptr;
    P.trace = g_trace_buf;
    P.dbg = g_pair_mode >> 3;
    if (save_acts && !ctx->act_save[net]) return fail(NERF_ERR_STATE, "tc_forward_rays: ctx was not created with training=1");
    timing_begin(0, st);
    if (g_pair_mode & 4) {
        // experimental: CTA pairs with the activations resident in tensor memory (mlp_tc_experimental.cu)
        if (int rc = tc_forward_ts(P, save_acts, st)) return rc;
    } else if (g_pair_mode & 1) {
        // CTA pairs: clusters of 2 over 512-row quads
        const int64_t n_quads = (P.n_pairs + 1) / 2;
        const int clusters = (int)(n_quads < num_sms() / 2 ? n_quads : num_sms() / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (save_acts) NERF_CUDA(cudaLaunchKernelEx(&cfg, nerf_mlp_fwd_tc_kernel<true, true>, P));
        else NERF_CUDA(cudaLaunchKernelEx(&cfg, nerf_mlp_fwd_tc_kernel<false, true>, P));
    } else {
        int grid = (int)(P.n_pairs < num_sms() ? P.n_pairs : num_sms());
        if (save_acts) nerf_mlp_fwd_tc_kernel<true, false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(P);
        else nerf_mlp_fwd_tc_kernel<false, false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(P);
    }
    timing_end(0, st);
    NERF_LAUNCHED();
    return NERF_OK;
}

int64_t tc_save_bytes_per_tile() { return SAVE_TILE_BYTES; }
int tc_fwd_chunks() { return N_CHUNKS; }

}  // namespace nerf



// diagnostics: enable (device buffer of 4*3*16*4 int64) / disable the forward-kernel timeline trace
extern "C" int nerf_debug_trace(long long* dev_buf) { nerf::g_trace_buf = dev_buf; return NERF_OK; }


// selects an experimental variant of the fused forward kernel: 0 = default (single CTA), 1 = CTA pair (cta_group::2),
// 4 = CTA pair with the activations resident in tensor memory ("TS" MMAs)
extern "C" int nerf_debug_pair_mode(int on) { nerf::g_pair_mode = on; return NERF_OK; }

