// EXPERIMENTAL forward variants, tcgen05 building-block self-tests and issue-rate probes.  None of this is on the product
// path: the kernels here are reachable only through nerf_debug_pair_mode(4) and the nerf_selftest_* entry points declared in
// include/nerf_b200_debug.h.  They are kept because DESIGN.md quotes their measurements (TMEM-resident activations, CTA
// pairs, MMA issue rates) and because the self-tests pin the descriptor / instruction encodings the product kernels rely on.
#include "common.cuh"
#include "ctx.cuh"
#include "mlp_tc_common.cuh"
#include "mlp_tc_fwd.cuh"
#include "../../include/nerf_b200_debug.h"

using namespace nerf;
using namespace tc5;
using namespace tcmlp;

namespace {

// ------------------------------------------------------------------------------------------------
// "TS" forward kernel: CTA pairs (cta_group::2) with the activations resident in TENSOR MEMORY.
//
// The fused kernels above are bound by shared-memory bandwidth: every MMA re-reads its A tile from shared memory
// and every epilogue writes the next A tile back there.  Here the epilogue writes the next layer's input as packed
// bf16 pairs straight into TMEM (tcgen05.st) and the MMAs take their A operand from TMEM, so shared memory only
// carries the weight stream.  One 128-row tile per CTA, 256 rows per CTA pair (M = 256 MMAs, each CTA staging 64 of
// the 128 weight rows of a chunk):
//   TMEM columns   [0,256)  fp32 accumulator, two 128-column N-halves
//                  [256,384) / [384,512)  activation buffers (128 columns = 256 bf16 K-elements), ping-pong
//   shared memory  positional encoding (the only SS operand: layer 0 and the skip rows of layer 5), a 16-stage ring
//                  of 8 KB weight sub-chunks, the bias/head side table
// Sixteen epilogue warps: warp w owns TMEM lane quarter (w & 3) and 32-column group (w >> 2) of each N-half, so an
// N-half is converted by all 512 threads at once.  Layer l+1's MMAs over K-blocks 0,1 start as soon as N-half 0 of
// layer l has been converted (aready[0]); K-blocks 2,3 follow aready[1]; the conversion of N-half 0 of layer l+1
// overlaps the MMAs of its N-half 1.
// ------------------------------------------------------------------------------------------------
namespace ts {
constexpr int WORKER_WARPS = 16;
constexpr int NI = 1;                                // MMA issuer warps (leader CTA); see the note on NI > 1 at the issuer loop
constexpr int THREADS = (WORKER_WARPS + 1 + NI) * 32;   // workers + producer warp + issuer warps (peer: one forwarder)
constexpr int STAGES_TS = 8;
constexpr int SUB_BYTES = CHUNK_BYTES / 2;           // this CTA's 64 weight rows of a 16 KB chunk
constexpr int SLOT_BYTES = 2 * SUB_BYTES;            // a ring slot = two consecutive K-blocks of one N-half = 8 MMAs
constexpr int SM_ENC = 0;                            // 2 x 16 KB: bf16(enc) K-block, xyz residual K-block
constexpr int SM_RING_TS = 2 * 16384;
constexpr int SM_SIDE_TS = SM_RING_TS + STAGES_TS * SLOT_BYTES;
constexpr int SM_DIRB_TS = SM_SIDE_TS + SIDE_FLOATS * 4;           // 4 rays x 128 floats
constexpr int SM_PART = SM_DIRB_TS + 4 * 512;                      // float4 [128 rows][4 column groups]
constexpr int SM_FULL_TS = SM_PART + 128 * 4 * 16;
constexpr int SM_EMPTY_TS = SM_FULL_TS + 8 * STAGES_TS;
constexpr int SM_PFULL_TS = SM_EMPTY_TS + 8 * STAGES_TS;
constexpr int SM_ACCF_TS = SM_PFULL_TS + 8 * STAGES_TS;            // 2: accumulator N-half complete
constexpr int SM_AREADY_TS = SM_ACCF_TS + 16;                      // 2: activation K-half written (leader's copy counts)
constexpr int SM_TOK_TS = SM_AREADY_TS + 16;                       // NI: issue token of the relay
constexpr int SM_TMEM_TS = SM_TOK_TS + 8 * NI;
constexpr int SMEM_TS = SM_TMEM_TS + 16 + 1024;
static_assert(SM_DIRB_TS % 16 == 0 && SM_FULL_TS % 8 == 0, "alignment");
static_assert(SMEM_TS <= 232448, "exceeds the 227 KB shared memory limit");
constexpr uint32_t T_ACC = 0, T_ABUF = 256;
}  // namespace ts

// bias (+ReLU) + bf16 pack of one 32-column accumulator group; optional sigma head partial sum and ReLU mask
template <bool RELU, bool SIGMA, bool SAVE>
__device__ __forceinline__ void ts_group(const uint32_t (&v)[32], const float* bias, const float* wsig, uint64_t& sig2,
                                         uint32_t& mk, uint32_t (&pk)[16]) {
    const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(bias);
    const ulonglong2* s2 = reinterpret_cast<const ulonglong2*>(wsig);
    uint32_t neg = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const ulonglong2 bb = b2[q];
        float x0, x1, x2, x3;
        f2_unpack(f2_add(f2_pack(v[4 * q], v[4 * q + 1]), bb.x), x0, x1);
        f2_unpack(f2_add(f2_pack(v[4 * q + 2], v[4 * q + 3]), bb.y), x2, x3);
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
}

// 16 channels [16 CQ, 16 CQ + 16) of the positional encoding of one point, packed bf16x2.  sincosf at the first
// octave the quarter needs, exact angle doubling (at most three times) for the following ones.
template <int CQ>
__device__ __forceinline__ void encode_quarter(const float (&p)[3], uint32_t (&E)[8]) {
    constexpr int LO = 16 * CQ;
    constexpr int I0 = (CQ == 0) ? 0 : (CQ == 1) ? 2 : (CQ == 2) ? 4 : 7;
    constexpr int I1 = (CQ == 0) ? 2 : (CQ == 1) ? 4 : (CQ == 2) ? 7 : 9;
    float e[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) e[j] = 0.f;
    if (CQ == 0) { e[0] = p[0]; e[1] = p[1]; e[2] = p[2]; }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float sv, cv;
#pragma unroll
        for (int i = I0; i <= I1; ++i) {
            if (i == I0) sincosf((float)(1 << i) * p[c], &sv, &cv);
            else {
                const float s2 = 2.f * sv * cv;
                const float c2 = fmaf(-2.f * sv, sv, 1.f);
                sv = s2; cv = c2;
            }
            const int is = 3 + 6 * i + c - LO, ic = is + 3;
            if (is >= 0 && is < 16) e[is] = sv;
            if (ic >= 0 && ic < 16) e[ic] = cv;
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) E[q] = pack_bf16x2(e[2 * q], e[2 * q + 1]);
}

// 32 bf16 (64 contiguous bytes after the swizzle) of one row of a saved [128 x 64] image block, straight to global
__device__ __forceinline__ void save_row_half(uint8_t* block, int row, int chunk0, const uint32_t (&pk)[16]) {
    uint8_t* rp = block + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(rp + (((chunk0 + c) ^ (row & 7)) << 4)) =
            make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
}

template <bool SAVE, int CQ>
__device__ __forceinline__ void ts_worker(const FwdParams& P, uint8_t* smem, uint32_t base, uint32_t tmem_base,
                                          const float* side, int q4, int lane, int rank, int unit, int n_units_grid,
                                          int my_units) {
    using namespace ts;
    const int row = 32 * q4 + lane;
    const int wtid = CQ * TILE_M + row;                               // 0..511 among the workers
    const bool elected = (wtid == 0);
    const uint32_t t_lane = tmem_base + (uint32_t(32 * q4) << 16);
    float* dbs = reinterpret_cast<float*>(smem + SM_DIRB_TS);
    float4* part = reinterpret_cast<float4*>(smem + SM_PART) + row * 4;
    const uint32_t accf0 = base + SM_ACCF_TS, accf1 = accf0 + 8;
    const uint32_t ar0 = rank ? map_to_cta(base + SM_AREADY_TS, 0) : base + SM_AREADY_TS;
    const uint32_t ar1 = rank ? map_to_cta(base + SM_AREADY_TS + 8, 0) : base + SM_AREADY_TS + 8;
    // the warp's TMEM / shared-memory writes are complete and fenced; one lane signals the leader's issuer
    auto arrive = [&](uint32_t bar) {
        __syncwarp();
        if (lane == 0) { if (rank) mbar_arrive_cluster(bar); else mbar_arrive(bar); }
    };
    const uint32_t enc_row = base + SM_ENC + (row >> 3) * 1024 + (row & 7) * 128;
    uint32_t accf_par = 0;

    for (int it = 0; it < my_units; ++it) {
        const int64_t wu = unit + (int64_t)it * n_units_grid;
        const int64_t tile = wu * 2 + rank;
        const int64_t g_row = tile * TILE_M + row;
        const bool valid = g_row < P.M;
        const int64_t gr = valid ? g_row : (P.M - 1);
        const int64_t ray = gr / P.N;
        uint8_t* save_tile = SAVE ? P.act_save + tile * SAVE_TILE_BYTES : nullptr;
        uint32_t* mask_tile = SAVE ? P.mask_save + tile * (MASK_TILE_BYTES / 4) : nullptr;

        const int64_t m_first = tile * TILE_M;
        const int64_t ray0 = ((m_first < P.M) ? m_first : (P.M - 1)) / P.N;
        const int64_t m_last = (m_first + TILE_M - 1 < P.M) ? (m_first + TILE_M - 1) : (P.M - 1);
        const int n_rays = (int)(m_last / P.N - ray0) + 1;
        const bool staged = n_rays <= 4;
        if (staged && wtid < n_rays * 32)
            reinterpret_cast<float4*>(dbs)[wtid] = __ldg(reinterpret_cast<const float4*>(P.dirbias + ray0 * 128) + wtid);

        // ---- positional encoding: this thread's 16 channels -> two 16-byte chunks of the SS operand block ----
        {
            const float tv = P.t[gr];
            float p[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) p[c] = __fadd_rn(P.o[ray * 3 + c], __fmul_rn(P.d[ray * 3 + c], tv));
            uint32_t E[8];
            encode_quarter<CQ>(p, E);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(enc_row + (uint32_t((2 * CQ) ^ (row & 7)) << 4)),
                         "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(enc_row + (uint32_t((2 * CQ + 1) ^ (row & 7)) << 4)),
                         "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]) : "memory");
            if (SAVE) {
                uint8_t* rp = save_tile + SAVE_ENC + (row >> 3) * 1024 + (row & 7) * 128;
                *reinterpret_cast<uint4*>(rp + (((2 * CQ) ^ (row & 7)) << 4)) = make_uint4(E[0], E[1], E[2], E[3]);
                *reinterpret_cast<uint4*>(rp + (((2 * CQ + 1) ^ (row & 7)) << 4)) = make_uint4(E[4], E[5], E[6], E[7]);
            }
            if (CQ == 0) {
                float lo[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) lo[c] = p[c] - __bfloat162float(__float2bfloat16_rn(p[c]));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(enc_row + 16384 + (uint32_t(0 ^ (row & 7)) << 4)),
                             "r"(pack_bf16x2(lo[0], lo[1])), "r"(pack_bf16x2(lo[2], 0.f)), "r"(0u), "r"(0u) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(enc_row + 16384 + (uint32_t(1 ^ (row & 7)) << 4)),
                             "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
            }
        }
        tc_fence_before();
        fence_proxy_async_all();
        arrive(ar0);
        arrive(ar1);

        float sig = 0.f;
        int e = 0;                                                    // epilogues done in this tile: writes buffer e & 1
        for (int ph = 0; ph < N_PHASES; ++ph) {
            if (ph == 5) continue;                                    // layer 5 accumulates on through phase 6
            if (ph < 10) {
                const int layer = (ph <= 4) ? ph : (ph == 6 ? 5 : (ph == 7 ? 6 : (ph == 8 ? 7 : 8)));
                const float* bias = side + (layer < 8 ? SIDE_BIAS + layer * H : SIDE_BFEAT) + CQ * 32;
                const float* wsig = side + SIDE_WSIG + CQ * 32;
                const bool relu = layer < 8;
                uint64_t sig2 = 0ull;
                const uint32_t a_dst = t_lane + T_ABUF + (e & 1) * 128 + CQ * 16;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (elected) trace_ev(P.trace, 2, it, ph, h ? 3 : 0);
                    mbar_wait(h ? accf1 : accf0, accf_par, 2 + 4 * h);
                    tc_fence_after();
                    if (elected && h == 0) trace_ev(P.trace, 2, it, ph, 1);
                    uint32_t v[32], pk[16], mk;
                    tmem_ld32(t_lane + T_ACC + h * 128 + CQ * 32, v);
                    tmem_ld_wait();
                    if (ph == 8) ts_group<true, true, SAVE>(v, bias + h * 128, wsig + h * 128, sig2, mk, pk);
                    else if (relu) ts_group<true, false, SAVE>(v, bias + h * 128, wsig, sig2, mk, pk);
                    else ts_group<false, false, SAVE>(v, bias + h * 128, wsig, sig2, mk, pk);
                    tmem_st16(a_dst + h * 64, pk);
                    tmem_st_wait();
                    tc_fence_before();
                    arrive(h ? ar1 : ar0);
                    if (SAVE) {
                        uint8_t* img = save_tile + ((layer < 8) ? SAVE_H + 65536 * layer : SAVE_FEAT);
                        save_row_half(img + (h * 2 + (CQ >> 1)) * 16384, row, (CQ & 1) * 4, pk);
                        if (relu) mask_tile[((size_t)layer * 128 + row) * 8 + h * 4 + CQ] = mk;
                    }
                }
                if (elected) trace_ev(P.trace, 2, it, ph, 2);
                accf_par ^= 1;
                ++e;
                if (ph == 8) {
                    float a, b;
                    f2_unpack(sig2, a, b);
                    sig = a + b;
                }
            } else {
                // ddir epilogue: + (bias + per-ray direction term), ReLU, rgb head partial sums over this thread's 32 columns
                mbar_wait(accf0, accf_par, 2);
                mbar_wait(accf1, accf_par, 6);
                accf_par ^= 1;
                tc_fence_after();
                const float* db = (staged ? (dbs + (int)(ray - ray0) * 128) : (P.dirbias + ray * 128)) + CQ * 32;
                const ulonglong2* d2 = reinterpret_cast<const ulonglong2*>(db);
                const ulonglong2* wr = reinterpret_cast<const ulonglong2*>(side + SIDE_WRGB + CQ * 32);
                const ulonglong2* wg = reinterpret_cast<const ulonglong2*>(side + SIDE_WRGB + 128 + CQ * 32);
                const ulonglong2* wb = reinterpret_cast<const ulonglong2*>(side + SIDE_WRGB + 256 + CQ * 32);
                uint64_t r2 = 0ull, g2 = 0ull, b2 = 0ull;
                uint32_t v[32], pk[16], neg = 0;
                tmem_ld32(t_lane + T_ACC + CQ * 32, v);
                tmem_ld_wait();
                tc_fence_before();
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
                    b2 = f2_fma(x01, c.x, b2); b2 = f2_fma(x23, c.y, b2);
                    if (SAVE) {
                        pk[2 * q] = cvt_bf16x2<false>(x0, x1);
                        pk[2 * q + 1] = cvt_bf16x2<false>(x2, x3);
                    }
                }
                float ra, rb, ga, gb, ba, bb;
                f2_unpack(r2, ra, rb); f2_unpack(g2, ga, gb); f2_unpack(b2, ba, bb);
                part[CQ] = make_float4(ra + rb, ga + gb, ba + bb, sig);
                if (SAVE) {
                    save_row_half(save_tile + SAVE_HD + (CQ >> 1) * 16384, row, (CQ & 1) * 4, pk);
                    uint32_t* mp = mask_tile + ((size_t)8 * 128 + row) * 8;
                    mp[CQ] = signs_to_mask(neg);
                    mp[4 + CQ] = 0u;
                }
                named_bar_sync(1, WORKER_WARPS * 32);
                if (CQ == 0 && valid) {
                    const float4 p0 = part[0], p1 = part[1], p2 = part[2], p3 = part[3];
                    P.preds[g_row] = make_float4((p0.x + p1.x) + (p2.x + p3.x) + side[SIDE_BRGB],
                                                 (p0.y + p1.y) + (p2.y + p3.y) + side[SIDE_BRGB + 1],
                                                 (p0.z + p1.z) + (p2.z + p3.z) + side[SIDE_BRGB + 2],
                                                 (p0.w + p1.w) + (p2.w + p3.w) + side[SIDE_BSIG]);
                }
                if (elected) trace_ev(P.trace, 2, it, ph, 2);
            }
        }
    }
}

template <bool SAVE>
__global__ void __launch_bounds__(ts::THREADS, 1) nerf_mlp_fwd_ts_kernel(const FwdParams P) {
    using namespace ts;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int unit = (int)(blockIdx.x >> 1);
    const int n_units_grid = (int)(gridDim.x >> 1);
    const int64_t n_units = P.n_pairs;                                // 256-row units, one per CTA pair
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES_TS; ++i) {
            mbar_init(base + SM_FULL_TS + 8 * i, 1);
            mbar_init(base + SM_EMPTY_TS + 8 * i, 1);
            mbar_init(base + SM_PFULL_TS + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(base + SM_ACCF_TS + 8 * i, 1);
            mbar_init(base + SM_AREADY_TS + 8 * i, 2 * WORKER_WARPS);   // one arrival per worker warp of both CTAs
        }
        for (int i = 0; i < NI; ++i) mbar_init(base + SM_TOK_TS + 8 * i, 1);
        fence_barrier_init();
        mbar_arrive(base + SM_TOK_TS);                                   // the first issuer holds the token
    }
    float* side = reinterpret_cast<float*>(smem + SM_SIDE_TS);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM_TS);
    if (warp == WORKER_WARPS + 1) tmem_alloc_2cta(base + SM_TMEM_TS, 512);
    for (int i = threadIdx.x; i < SIDE_FLOATS; i += THREADS) side[i] = P.side[i];
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_units = (n_units > unit) ? (int)((n_units - unit + n_units_grid - 1) / n_units_grid) : 0;
    // ring slots per tile: one per (phase, N-half, pair of K-blocks), in weight-stream order
    int slots_per_tile = 0;
    for (int ph = 0; ph < N_PHASES; ++ph) slots_per_tile += c_fwd_prog.chunks[ph] / 2;

    if (warp == WORKER_WARPS) {
        // ---- producer: this CTA's half (64 weight rows) of two consecutive chunks per slot, in stream order ----
        if (lane == 0) {
            int slot = 0;
            uint32_t par = 1;
            const uint8_t* src0 = reinterpret_cast<const uint8_t*>(P.w_chunks) + rank * SUB_BYTES;
            for (int it = 0; it < my_units; ++it)
                for (int c = 0; c < N_CHUNKS; c += 2) {
                    const uint32_t full = base + SM_FULL_TS + 8 * slot, dst = base + SM_RING_TS + slot * SLOT_BYTES;
                    mbar_wait(base + SM_EMPTY_TS + 8 * slot, par, 1);
                    mbar_arrive_expect_tx(full, SLOT_BYTES);
                    bulk_g2s(dst, src0 + (size_t)c * CHUNK_BYTES, SUB_BYTES, full);
                    bulk_g2s(dst + SUB_BYTES, src0 + (size_t)(c + 1) * CHUNK_BYTES, SUB_BYTES, full);
                    if (++slot == STAGES_TS) { slot = 0; par ^= 1; }
                }
        }
    } else if (warp > WORKER_WARPS) {
        const int k = warp - (WORKER_WARPS + 1);                          // issuer index in the relay
        if (lane == 0 && rank != 0) {
            if (k == 0) {
                // ---- peer: forward "my ring slot is full" to the leader ----
                int slot = 0;
                uint32_t par = 0;
                for (int g = 0; g < my_units * slots_per_tile; ++g) {
                    mbar_wait(base + SM_FULL_TS + 8 * slot, par, 7);
                    mbar_arrive_cluster(map_to_cta(base + SM_PFULL_TS + 8 * slot, 0));
                    if (++slot == STAGES_TS) { slot = 0; par ^= 1; }
                }
            }
        } else if (lane == 0) {
            // ---- leader: MMA issuer.  The loop is written as a relay of NI issuer threads (slot seq is issued by thread
            // seq % NI after it has waited for the slot's inputs and for a token), but NI must stay 1: with NI = 4 the
            // kernel ran 1.4x faster and produced WRONG results on hardware -- MMAs issued by different threads into
            // the same accumulator are not ordered by tcgen05.fence + mbarrier hand-off (only completion, i.e.
            // tcgen05.commit + wait, orders them), so a zero-initialising MMA can be overtaken. ----
            const uint32_t desc_hi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));
            const uint32_t lbo_bits = (16u >> 4) << 16;
            const uint32_t enc0 = ((base + SM_ENC) & 0x3FFFF) >> 4;
            const uint32_t ring0 = ((base + SM_RING_TS) & 0x3FFFF) >> 4;
            const uint32_t idesc = make_idesc_bf16(256, 128, 0, 0);
            const uint32_t accf0 = base + SM_ACCF_TS, ar0 = base + SM_AREADY_TS;
            const uint32_t tok_mine = base + SM_TOK_TS + 8 * k, tok_next = base + SM_TOK_TS + 8 * ((k + 1) % NI);
            uint32_t seq = 0, ar_par = 0, tok_par = 0;
            for (int it = 0; it < my_units; ++it) {
                int e = 0;                                            // epilogues so far: TS phases read buffer (e - 1) & 1
                for (int ph = 0; ph < N_PHASES; ++ph) {
                    const int n_ch = c_fwd_prog.chunks[ph], kbs = c_fwd_prog.kb[ph], flags = c_fwd_prog.flags[ph];
                    const int halves = n_ch / kbs;
                    const bool enc = (flags & PH_ENC) != 0, acc_in = (flags & PH_ACC) != 0;
                    const bool consumes = !(enc && acc_in);           // phase 6 only adds the skip rows: nothing new to wait for
                    const bool has_epi = (ph != 5);
                    const uint32_t a_buf = tmem_base + T_ABUF + ((e - 1) & 1) * 128;
                    for (int h = 0; h < halves; ++h) {
                        const uint32_t d_tmem = tmem_base + T_ACC + h * 128;
                        for (int kp = 0; kp < kbs / 2; ++kp, ++seq) {
                            if ((int)(seq % NI) != k) continue;
                            const uint32_t slot = seq % STAGES_TS, ring_par = (seq / STAGES_TS) & 1;
                            if (k == 0) trace_ev(P.trace, 0, it, ph, 0);
                            // the first slot that touches a K-half waits for it; later slots follow it in the relay
                            if (consumes && h == 0) {
                                if (kp == 0) mbar_wait_cluster(ar0, ar_par, 3);
                                if (enc || kp == 1) mbar_wait_cluster(ar0 + 8, ar_par, 5);
                            }
                            mbar_wait(base + SM_FULL_TS + 8 * slot, ring_par, 4);
                            mbar_wait_cluster(base + SM_PFULL_TS + 8 * slot, ring_par, 8);
                            mbar_wait(tok_mine, tok_par, 9);
                            tok_par ^= 1;
                            tc_fence_after();
                            if (k == 0) trace_ev(P.trace, 0, it, ph, 1);
#pragma unroll
                            for (int sub = 0; sub < 2; ++sub) {
                                const int kb = 2 * kp + sub;
                                const uint32_t b_lo = (ring0 + slot * (SLOT_BYTES >> 4) + sub * (SUB_BYTES >> 4)) | lbo_bits;
                                const bool acc0 = (kb > 0) || acc_in;
                                const int n_mma = (enc && kb == 1) ? 1 : 4;
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    if (kk < n_mma) {
                                        const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2 * kk);
                                        const uint32_t accum = (acc0 || kk > 0) ? 1u : 0u;
                                        if (enc) {
                                            const uint32_t a_lo = (enc0 + kb * (16384 >> 4)) | lbo_bits;
                                            const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2 * kk);
                                            mma_bf16_ss_2cta(d_tmem, ad, bd, idesc, accum);
                                        } else {
                                            mma_bf16_ts_2cta(d_tmem, a_buf + kb * 32 + kk * 8, bd, idesc, accum);
                                        }
                                    }
                                }
                            }
                            mma_commit_2cta(base + SM_EMPTY_TS + 8 * slot, 0x3);
                            // the last slot of an N-half: its MMAs depend (accumulator chain) on every earlier one
                            if (has_epi && kp == kbs / 2 - 1) {
                                mma_commit_2cta(accf0 + 8 * h, 0x3);
                                if (halves == 1) mma_commit_2cta(accf0 + 8, 0x3);
                            }
                            tc_fence_before();                        // my MMAs are ordered before the next issuer's
                            mbar_arrive(tok_next);
                            if (k == 0) trace_ev(P.trace, 0, it, ph, 2);
                        }
                    }
                    if (consumes) ar_par ^= 1;
                    if (has_epi) ++e;
                }
            }
        }
    } else {
        const int q4 = warp & 3;
        switch (warp >> 2) {
            case 0: ts_worker<SAVE, 0>(P, smem, base, tmem_base, side, q4, lane, rank, unit, n_units_grid, my_units); break;
            case 1: ts_worker<SAVE, 1>(P, smem, base, tmem_base, side, q4, lane, rank, unit, n_units_grid, my_units); break;
            case 2: ts_worker<SAVE, 2>(P, smem, base, tmem_base, side, q4, lane, rank, unit, n_units_grid, my_units); break;
            default: ts_worker<SAVE, 3>(P, smem, base, tmem_base, side, q4, lane, rank, unit, n_units_grid, my_units); break;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();                       // the leader's MMAs read the peer's shared and tensor memory: leave together
    if (warp == WORKER_WARPS + 1) tmem_dealloc_2cta(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// self-test GEMM (single CTA): validates descriptor / swizzle / TMEM conventions on hardware.
// mode 0: A (128,K) , B (N,K)  row-major fp32 -> C = A * B^T      (K-major operands)
// mode 1: A (K,128) , B (K,N)  row-major fp32 -> C = A^T * B      (MN-major operands)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) selftest_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                               float* __restrict__ C, int N, int K, int mode) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // layout: A image, B image, barrier, tmem slot
    const uint32_t a_bytes = 128 * K * 2, b_bytes = N * K * 2;
    uint8_t* a_img = smem;
    uint8_t* b_img = smem + a_bytes;
    const uint32_t bar = base + a_bytes + b_bytes;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + a_bytes + b_bytes + 16);

    if (mode == 0) {
        // K-major: K-blocks of 64, each [rows x 128 B]
        for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
            int r = i / K, k = i - r * K;
            uint32_t off = (k >> 6) * (128 * 128) + sw128_offset(r, k & 63);
            *reinterpret_cast<__nv_bfloat16*>(a_img + off) = __float2bfloat16(A[i]);
        }
        for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
            int r = i / K, k = i - r * K;
            uint32_t off = (k >> 6) * (N * 128) + sw128_offset(r, k & 63);
            *reinterpret_cast<__nv_bfloat16*>(b_img + off) = __float2bfloat16(B[i]);
        }
    } else {
        // MN-major: blocks of 64 MN-elements, each [K rows x 128 B]
        for (int i = threadIdx.x; i < K * 128; i += blockDim.x) {
            int k = i / 128, m = i - k * 128;
            uint32_t off = (m >> 6) * (K * 128) + sw128_offset(k, m & 63);
            *reinterpret_cast<__nv_bfloat16*>(a_img + off) = __float2bfloat16(A[i]);
        }
        for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
            int k = i / N, n = i - k * N;
            uint32_t off = (n >> 6) * (K * 128) + sw128_offset(k, n & 63);
            *reinterpret_cast<__nv_bfloat16*>(b_img + off) = __float2bfloat16(B[i]);
        }
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(base + a_bytes + b_bytes + 16, 256);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N, mode, mode);
        for (int k16 = 0; k16 < K / 16; ++k16) {
            uint64_t ad, bd;
            if (mode == 0) {
                int kb = k16 >> 2, kk = k16 & 3;
                ad = make_sdesc_sw128(base + kb * (128 * 128) + kk * 32, 16, 1024);
                bd = make_sdesc_sw128(base + a_bytes + kb * (N * 128) + kk * 32, 16, 1024);
            } else {
                ad = make_sdesc_sw128(base + k16 * 2048, K * 128, 1024);
                bd = make_sdesc_sw128(base + a_bytes + k16 * 2048, K * 128, 1024);
            }
            mma_bf16_ss(tmem_base, ad, bd, idesc, k16 > 0 ? 1u : 0u);
        }
        mma_commit(bar);
    }
    mbar_wait(bar, 0, 9);
    tc_fence_after();
    const int row = threadIdx.x;
    for (int cg = 0; cg < N / 32; ++cg) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (uint32_t(32 * warp) << 16) + cg * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) C[(int64_t)row * N + cg * 32 + q] = __uint_as_float(v[q]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
    (void)lane;
}

// self-test of the CTA-pair path: C (256, N) = A (256, K) x B (N, K)^T, K-major operands, N = 128 or 256.
// Each CTA of the pair stages its 128 rows of A and its N/2 rows of B; the leader issues M = 256 MMAs.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
selftest_gemm_2cta_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int N, int K) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    const int NB = N / 2;                                   // B rows held by this CTA
    const uint32_t a_bytes = 128 * K * 2, b_bytes = NB * K * 2;
    uint8_t* a_img = smem;
    uint8_t* b_img = smem + a_bytes;
    const uint32_t bar = base + a_bytes + b_bytes;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + a_bytes + b_bytes + 16);
    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
        int r = i / K, k = i - r * K;
        uint32_t off = (k >> 6) * (128 * 128) + sw128_offset(r, k & 63);
        *reinterpret_cast<__nv_bfloat16*>(a_img + off) = __float2bfloat16(A[(size_t)(rank * 128 + r) * K + k]);
    }
    for (int i = threadIdx.x; i < NB * K; i += blockDim.x) {
        int r = i / K, k = i - r * K;
        uint32_t off = (k >> 6) * (NB * 128) + sw128_offset(r, k & 63);
        *reinterpret_cast<__nv_bfloat16*>(b_img + off) = __float2bfloat16(B[(size_t)(rank * NB + r) * K + k]);
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc_2cta(base + a_bytes + b_bytes + 16, 256);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                          // both CTAs' operands and barriers are ready
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (rank == 0 && threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(256, N, 0, 0);
        for (int k16 = 0; k16 < K / 16; ++k16) {
            int kb = k16 >> 2, kk = k16 & 3;
            uint64_t ad = make_sdesc_sw128(base + kb * (128 * 128) + kk * 32, 16, 1024);
            uint64_t bd = make_sdesc_sw128(base + a_bytes + kb * (NB * 128) + kk * 32, 16, 1024);
            mma_bf16_ss_2cta(tmem_base, ad, bd, idesc, k16 > 0 ? 1u : 0u);
        }
        mma_commit_2cta(bar, 0x3);
    }
    mbar_wait(bar, 0, 9);
    tc_fence_after();
    const int row = rank * 128 + threadIdx.x;
    for (int cg = 0; cg < N / 32; ++cg) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (uint32_t(32 * warp) << 16) + cg * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) C[(int64_t)row * N + cg * 32 + q] = __uint_as_float(v[q]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 0) tmem_dealloc_2cta(tmem_base, 256);
}

// self-test of MMAs whose A operand lives in tensor memory: C (M, N) = A (M, K) x B (N, K)^T with A written to TMEM
// by tcgen05.st as packed bf16 pairs and B K-major in shared memory.  PAIR: M = 256 over a CTA pair (cta_group::2,
// each CTA holds its 128 rows of A in its own TMEM and N/2 rows of B), else M = 128 on one CTA.
template <bool PAIR>
__global__ void __launch_bounds__(128, 1)
selftest_gemm_ts_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int N, int K,
                        int reps, int probe, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    const int NB = PAIR ? N / 2 : N;                        // B rows held by this CTA
    const uint32_t b_bytes = NB * K * 2;
    uint8_t* b_img = smem;
    const uint32_t bar = base + b_bytes;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + b_bytes + 32);
    for (int i = threadIdx.x; i < NB * K; i += blockDim.x) {
        int r = i / K, k = i - r * K;
        uint32_t off = (k >> 6) * (NB * 128) + sw128_offset(r, k & 63);
        *reinterpret_cast<__nv_bfloat16*>(b_img + off) = __float2bfloat16(B[(size_t)(rank * NB + r) * K + k]);
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 8, (1u << 20) - 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        if (PAIR) tmem_alloc_2cta(base + b_bytes + 32, 512); else tmem_alloc(base + b_bytes + 32, 512);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_lane = tmem_base + (uint32_t(32 * warp) << 16);
    constexpr uint32_t A_COL = 256;
    {
        const float* arow = A + (size_t)(rank * 128 + threadIdx.x) * K;
        for (int g = 0; g < K / 32; ++g) {
            uint32_t v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = pack_bf16x2(arow[g * 32 + 2 * q], arow[g * 32 + 2 * q + 1]);
            tmem_st16(t_lane + A_COL + 16 * g, v);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync();
    tc_fence_after();
    if (rank == 0 && threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, N, 0, 0);
        uint64_t bd[16];
#pragma unroll
        for (int k16 = 0; k16 < 16; ++k16) {
            const int kc = (k16 < K / 16) ? k16 : 0, kb = kc >> 2, kk = kc & 3;
            bd[k16] = make_sdesc_sw128(base + kb * (NB * 128) + kk * 32, 16, 1024);
        }
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int k16 = 0; k16 < 16; ++k16) {
                if (k16 < K / 16) {
                    const uint32_t acc = (k16 > 0 || r > 0) ? 1u : 0u;
                    if (PAIR) mma_bf16_ts_2cta(tmem_base, tmem_base + A_COL + k16 * 8, bd[k16], idesc, acc);
                    else mma_bf16_ts(tmem_base, tmem_base + A_COL + k16 * 8, bd[k16], idesc, acc);
                    if (probe >= 2 && (k16 & 3) == 3) {          // probe: cost of a commit per 4 MMAs (barrier never waited on)
                        if (PAIR) mma_commit_2cta(bar + 8, 0x3); else mma_commit(bar + 8);
                    }
                }
            }
        }
        if (PAIR) mma_commit_2cta(bar, 0x3); else mma_commit(bar);
        mbar_wait(bar, 0, 9);
        if (cycles) cycles[0] = clock64() - t0;
    }
    __syncthreads();                                         // nobody spins beside the issuing thread
    mbar_wait(bar, 0, 9);
    tc_fence_after();
    const int row = rank * 128 + threadIdx.x;
    for (int cg = 0; cg < N / 32; ++cg) {
        uint32_t v[32];
        tmem_ld32(t_lane + cg * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) C[(int64_t)row * N + cg * 32 + q] = __uint_as_float(v[q]);
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync();
    if (warp == 0) {
        if (PAIR) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

// MMA issue-rate probe: `reps` back-to-back passes of K=64 (4 MMAs) over fixed smem operands, N columns.
// Reports elapsed SM cycles from first issue to commit completion.  mode 0: K-major, 1: MN-major.
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int reps, int mode, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    const int warp = threadIdx.x >> 5;
    const uint32_t a_bytes = 128 * 64 * 2, b_bytes = N * 64 * 2;
    for (uint32_t i = threadIdx.x; i < (a_bytes + b_bytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    const uint32_t bar = base + a_bytes + b_bytes;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + a_bytes + b_bytes + 16);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(base + a_bytes + b_bytes + 16, 256);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N, mode, mode);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint64_t ad, bd;
                if (mode == 0) {
                    ad = make_sdesc_sw128(base + k * 32, 16, 1024);
                    bd = make_sdesc_sw128(base + a_bytes + k * 32, 16, 1024);
                } else {
                    ad = make_sdesc_sw128(base + k * 2048, 64 * 128, 1024);
                    bd = make_sdesc_sw128(base + a_bytes + k * 2048, 64 * 128, 1024);
                }
                mma_bf16_ss(tmem_base, ad, bd, idesc, 1u);
            }
        }
        mma_commit(bar);
        mbar_wait(bar, 0, 9);
        out[0] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}


// ------------------------------------------------------------------------------------------------
// Collector probe: two 128-row A tiles against ONE B tile (the situation of the forward kernel's two sub-tiles and a
// shared weight chunk).  variant 0: plain SS MMAs; 1: tcgen05.mma.ws with B kept in the collector for the second MMA;
// 2: plain MMAs with A kept in the collector across two N-halves.  C1 = A1 x B^T, C2 = A2 x B^T (128 x N each).
// reps > 1 (with K = 64): rate probe -- descriptors precomputed, nothing but accumulating MMAs in the loop.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) collector_probe_kernel(const float* __restrict__ A1, const float* __restrict__ A2,
                                                                 const float* __restrict__ Bm, float* __restrict__ C1,
                                                                 float* __restrict__ C2, int N, int K, int variant, int reps,
                                                                 long long* __restrict__ cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    const int warp = threadIdx.x >> 5;
    const uint32_t a_bytes = 128 * K * 2, b_bytes = N * K * 2;
    uint8_t* a1_img = smem;
    uint8_t* a2_img = smem + a_bytes;
    uint8_t* b_img = smem + 2 * a_bytes;
    const uint32_t bar = base + 2 * a_bytes + b_bytes;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 2 * a_bytes + b_bytes + 16);
    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
        const int r = i / K, k = i - r * K;
        const uint32_t off = (k >> 6) * (128 * 128) + sw128_offset(r, k & 63);
        *reinterpret_cast<__nv_bfloat16*>(a1_img + off) = __float2bfloat16(A1[i]);
        *reinterpret_cast<__nv_bfloat16*>(a2_img + off) = __float2bfloat16(A2[i]);
    }
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) {          // B as N/128 blocks of [128 n x K], K-blocks of 64 inside
        const int r = i / K, k = i - r * K;
        const uint32_t off = (r >> 7) * (128 * K * 2) + (k >> 6) * (128 * 128) + sw128_offset(r & 127, k & 63);
        *reinterpret_cast<__nv_bfloat16*>(b_img + off) = __float2bfloat16(Bm[i]);
    }
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(base + 2 * a_bytes + b_bytes + 16, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
        long long t0 = clock64();
        if (reps > 1 && K == 64) {
            uint64_t a1d[4], a2d[4], b0d[4], b1d[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                a1d[kk] = make_sdesc_sw128(base + kk * 32, 16, 1024);
                a2d[kk] = make_sdesc_sw128(base + a_bytes + kk * 32, 16, 1024);
                b0d[kk] = make_sdesc_sw128(base + 2 * a_bytes + kk * 32, 16, 1024);
                b1d[kk] = make_sdesc_sw128(base + 2 * a_bytes + (N == 256 ? 128 * K * 2 : 0) + kk * 32, 16, 1024);
            }
            t0 = clock64();
            for (int rep = 0; rep < reps; ++rep) {      // four MMAs per k-step in every variant: (A1, A2) x (B half 0, B half 1)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    if (variant == 1) {
                        mma_bf16_ss_ws(tmem_base, a1d[kk], b0d[kk], idesc, 1u, 0);
                        mma_bf16_ss_ws(tmem_base + 256, a2d[kk], b0d[kk], idesc, 1u, 1);
                        mma_bf16_ss_ws(tmem_base + 128, a1d[kk], b1d[kk], idesc, 1u, 0);
                        mma_bf16_ss_ws(tmem_base + 384, a2d[kk], b1d[kk], idesc, 1u, 1);
                    } else if (variant == 2) {
                        mma_bf16_ss_akeep(tmem_base, a1d[kk], b0d[kk], idesc, 1u, 0);
                        mma_bf16_ss_akeep(tmem_base + 128, a1d[kk], b1d[kk], idesc, 1u, 1);
                        mma_bf16_ss_akeep(tmem_base + 256, a2d[kk], b0d[kk], idesc, 1u, 0);
                        mma_bf16_ss_akeep(tmem_base + 384, a2d[kk], b1d[kk], idesc, 1u, 1);
                    } else if (variant == 3) {      // plain, ONE accumulator
                        mma_bf16_ss(tmem_base, a1d[kk], b0d[kk], idesc, 1u);
                        mma_bf16_ss(tmem_base, a2d[kk], b0d[kk], idesc, 1u);
                        mma_bf16_ss(tmem_base, a1d[kk], b1d[kk], idesc, 1u);
                        mma_bf16_ss(tmem_base, a2d[kk], b1d[kk], idesc, 1u);
                    } else if (variant == 4) {      // plain, two accumulators alternating per MMA
                        mma_bf16_ss(tmem_base, a1d[kk], b0d[kk], idesc, 1u);
                        mma_bf16_ss(tmem_base + 256, a2d[kk], b0d[kk], idesc, 1u);
                        mma_bf16_ss(tmem_base, a1d[kk], b1d[kk], idesc, 1u);
                        mma_bf16_ss(tmem_base + 256, a2d[kk], b1d[kk], idesc, 1u);
                    } else if (variant == 5) {      // plain, accumulator changes once per k-step group (4 MMAs per accumulator)
                        const uint32_t d = tmem_base + (rep & 3) * 128;
                        mma_bf16_ss(d, a1d[kk], b0d[kk], idesc, 1u);
                        mma_bf16_ss(d, a2d[kk], b0d[kk], idesc, 1u);
                        mma_bf16_ss(d, a1d[kk], b1d[kk], idesc, 1u);
                        mma_bf16_ss(d, a2d[kk], b1d[kk], idesc, 1u);
                    } else if (variant == 6) {      // .ws, ONE accumulator pair (keep / reuse alternate, accumulators alternate)
                        mma_bf16_ss_ws(tmem_base, a1d[kk], b0d[kk], idesc, 1u, 0);
                        mma_bf16_ss_ws(tmem_base + 256, a2d[kk], b0d[kk], idesc, 1u, 1);
                        mma_bf16_ss_ws(tmem_base, a1d[kk], b1d[kk], idesc, 1u, 0);
                        mma_bf16_ss_ws(tmem_base + 256, a2d[kk], b1d[kk], idesc, 1u, 1);
                    } else {
                        mma_bf16_ss(tmem_base, a1d[kk], b0d[kk], idesc, 1u);
                        mma_bf16_ss(tmem_base + 256, a2d[kk], b0d[kk], idesc, 1u);
                        mma_bf16_ss(tmem_base + 128, a1d[kk], b1d[kk], idesc, 1u);
                        mma_bf16_ss(tmem_base + 384, a2d[kk], b1d[kk], idesc, 1u);
                    }
                }
            }
        } else {
            for (int nh = 0; nh < N / 128; ++nh) {
                for (int k16 = 0; k16 < K / 16; ++k16) {
                    const int kb = k16 >> 2, kk = k16 & 3;
                    const uint64_t a1 = make_sdesc_sw128(base + kb * (128 * 128) + kk * 32, 16, 1024);
                    const uint64_t a2 = make_sdesc_sw128(base + a_bytes + kb * (128 * 128) + kk * 32, 16, 1024);
                    const uint64_t bd = make_sdesc_sw128(base + 2 * a_bytes + nh * (128 * K * 2) + kb * (128 * 128) + kk * 32, 16, 1024);
                    const uint32_t acc = (k16 > 0) ? 1u : 0u;
                    const uint32_t d1 = tmem_base + nh * 128, d2 = tmem_base + 256 + nh * 128;
                    if (variant == 1) {
                        mma_bf16_ss_ws(d1, a1, bd, idesc, acc, 0);
                        mma_bf16_ss_ws(d2, a2, bd, idesc, acc, 1);
                    } else if (variant == 2) {      // A kept: the same (a, b) product twice would double the result, so keep + plain second tile
                        mma_bf16_ss_akeep(d1, a1, bd, idesc, acc, 0);
                        mma_bf16_ss_akeep(d2, a2, bd, idesc, acc, 0);
                    } else {
                        mma_bf16_ss(d1, a1, bd, idesc, acc);
                        mma_bf16_ss(d2, a2, bd, idesc, acc);
                    }
                }
            }
        }
        mma_commit(bar);
        mbar_wait(bar, 0, 9);
        if (cycles) cycles[0] = clock64() - t0;
    }
    __syncthreads();
    mbar_wait(bar, 0, 9);
    tc_fence_after();
    const int row = threadIdx.x;
    for (int cg = 0; cg < N / 32; ++cg) {
        uint32_t v[32], w[32];
        tmem_ld32(tmem_base + (uint32_t(32 * warp) << 16) + cg * 32, v);
        tmem_ld32(tmem_base + (uint32_t(32 * warp) << 16) + 256 + cg * 32, w);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            C1[(int64_t)row * N + cg * 32 + q] = __uint_as_float(v[q]);
            C2[(int64_t)row * N + cg * 32 + q] = __uint_as_float(w[q]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace

namespace nerf {

int tc_experimental_init() {
    NERF_CUDA(cudaFuncSetAttribute(nerf_mlp_fwd_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts::SMEM_TS));
    NERF_CUDA(cudaFuncSetAttribute(nerf_mlp_fwd_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts::SMEM_TS));
    return NERF_OK;
}

// CTA pairs, activations in tensor memory: one 256-row unit per cluster
int tc_forward_ts(const FwdParams& P, bool save_acts, cudaStream_t st) {
    const int clusters = (int)(P.n_pairs < num_sms() / 2 ? P.n_pairs : num_sms() / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(ts::THREADS);
    cfg.dynamicSmemBytes = ts::SMEM_TS;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (save_acts) NERF_CUDA(cudaLaunchKernelEx(&cfg, nerf_mlp_fwd_ts_kernel<true>, P));
    else NERF_CUDA(cudaLaunchKernelEx(&cfg, nerf_mlp_fwd_ts_kernel<false>, P));
    return NERF_OK;
}

}  // namespace nerf

extern "C" int nerf_selftest_gemm(const float* a, const float* b, float* c, int m, int n, int k, int mode,
                                  void* stream) {
    NERF_CHECK_ARG(a && b && c, "null pointer");
    NERF_CHECK_ARG(m == 128 && (n == 128 || n == 256 || n == 64) && k >= 16 && k <= 256 && (k % 16) == 0,
                   "supported: m=128, n in {64,128,256}, k multiple of 16 up to 256");
    NERF_CHECK_ARG(mode == 1 || (k % 64) == 0, "mode 0 needs k % 64 == 0");
    NERF_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 or 1");
    size_t smem = (size_t)(128 + n) * k * 2 + 64 + 1024;
    NERF_CUDA(cudaFuncSetAttribute(selftest_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selftest_gemm_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, b, c, n, k, mode);
    NERF_LAUNCHED();
    return NERF_OK;
}

extern "C" int nerf_selftest_mma_rate(int n, int reps, int mode, long long* cycles_dev, void* stream) {
    NERF_CHECK_ARG(cycles_dev && (n == 64 || n == 128 || n == 256) && reps >= 1, "bad arguments");
    size_t smem = (size_t)(128 + n) * 64 * 2 + 64 + 1024;
    NERF_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mma_rate_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(n, reps, mode, cycles_dev);
    NERF_LAUNCHED();
    return NERF_OK;
}

extern "C" int nerf_selftest_gemm_2cta(const float* a, const float* b, float* c, int n, int k, void* stream) {
    NERF_CHECK_ARG(a && b && c && (n == 128 || n == 256) && k >= 64 && k <= 256 && (k % 64) == 0, "bad arguments");
    size_t smem = (size_t)(128 + n / 2) * k * 2 + 64 + 1024;
    NERF_CUDA(cudaFuncSetAttribute(selftest_gemm_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selftest_gemm_2cta_kernel<<<2, 128, smem, (cudaStream_t)stream>>>(a, b, c, n, k);
    NERF_LAUNCHED();
    return NERF_OK;
}

// self-test of TMEM-resident A operands ("TS" MMAs): pair = 0 -> C (128, n), pair = 1 -> C (256, n) over a CTA pair
extern "C" int nerf_selftest_gemm_ts(const float* a, const float* b, float* c, int n, int k, int pair, int reps, int probe,
                                     long long* cycles_dev, void* stream) {
    NERF_CHECK_ARG(a && b && c && (n == 128 || n == 256) && k >= 64 && k <= 256 && (k % 64) == 0, "bad arguments");
    NERF_CHECK_ARG(reps >= 1 && (probe == 0 || probe == 2), "reps must be >= 1; probe mode 0 or 2");
    const int nb = pair ? n / 2 : n;
    size_t smem = (size_t)nb * k * 2 + 64 + 1024;
    if (pair) {
        NERF_CUDA(cudaFuncSetAttribute(selftest_gemm_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        NERF_CUDA(cudaLaunchKernelEx(&cfg, selftest_gemm_ts_kernel<true>, a, b, c, n, k, reps, probe, cycles_dev));
    } else {
        NERF_CUDA(cudaFuncSetAttribute(selftest_gemm_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        selftest_gemm_ts_kernel<false><<<1, 128, smem, (cudaStream_t)stream>>>(a, b, c, n, k, reps, probe, cycles_dev);
    }
    NERF_LAUNCHED();
    return NERF_OK;
}


// Probe of the MMA collector variants (see collector_probe_kernel): C1 = A1 x B^T, C2 = A2 x B^T, both (128, n).
extern "C" int nerf_selftest_collector(const float* a1, const float* a2, const float* b, float* c1, float* c2, int n, int k,
                                       int variant, int reps, long long* cycles_dev, void* stream) {
    NERF_CHECK_ARG(a1 && a2 && b && c1 && c2 && (n == 128 || n == 256) && k >= 64 && k <= 256 && (k % 64) == 0, "bad arguments");
    NERF_CHECK_ARG(variant >= 0 && variant <= 6 && reps >= 1 && (variant <= 2 || reps > 1), "variant 0..2 (3..6: rate probes, reps > 1)");
    size_t smem = (size_t)(256 + n) * k * 2 + 64 + 1024;
    NERF_CHECK_ARG(smem <= 227 * 1024, "operands do not fit in shared memory");
    NERF_CUDA(cudaFuncSetAttribute(collector_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    collector_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a1, a2, b, c1, c2, n, k, variant, reps, cycles_dev);
    NERF_LAUNCHED();
    return NERF_OK;
}
