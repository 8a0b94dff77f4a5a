// Shared skeleton of the fused tcgen05 MLP kernels (forward chain and backward dX chain):
// shared-memory map, chunk "programs", the weight producer and the MMA issuer.
//
// A kernel instance processes PAIRS of 128-row sub-tiles (A, B).  Each sub-tile owns a 64 KB
// activation tile in shared memory (the A operand, 4 K-blocks of [128 x 64] bf16, 128B-swizzled)
// and 256 fp32 accumulator columns in TMEM.  Weights arrive as a stream of 16 KB chunks
// ([128 n x 64 k] bf16, K-major, 128B-swizzled) through a STAGES-deep ring; both sub-tiles consume
// every chunk, so L2->SMEM traffic is paid once per 256 rows.  Two MMA issuer threads (one per
// sub-tile) feed the tensor pipe, so one sub-tile's epilogue (CUDA cores) overlaps the other's MMAs.
#pragma once
#include "common.cuh"
#include "tc5.cuh"

namespace tcmlp {

using namespace tc5;

constexpr int H = 256;
constexpr int ENC_X = 63;
constexpr int ENC_D = 27;
constexpr int TILE_M = 128;
constexpr int CHUNK_BYTES = 16384;
constexpr int CHUNK_ELEMS = CHUNK_BYTES / 2;
constexpr int STAGES = 5;
constexpr int MAX_PHASES = 12;

// phase flags
constexpr int PH_ACC = 1;   // first k-block accumulates onto the existing accumulator
constexpr int PH_ENC = 2;   // k-block 1 of this phase holds only 16 meaningful columns (one K=16 MMA)

struct Program {
    int n_phases;
    int n_chunks;
    int chunks[MAX_PHASES];  // chunks per phase
    int kb[MAX_PHASES];      // k-blocks per N-half
    int flags[MAX_PHASES];
};

// fp32 side table (per net) kept in shared memory by both kernels (floats)
constexpr int SIDE_BIAS = 0;        // 8 x 256 trunk biases
constexpr int SIDE_BFEAT = 2048;    // 256
constexpr int SIDE_BDDIR = 2304;    // 128
constexpr int SIDE_WSIG = 2432;     // 256
constexpr int SIDE_WRGB = 2688;     // 3 x 128 ([channel][k])
constexpr int SIDE_BSIG = 3072;     // 1
constexpr int SIDE_BRGB = 3073;     // 3
constexpr int SIDE_FLOATS = 3080;

// shared memory map (bytes, relative to the 1024-aligned base)
constexpr int SM_ACT = 0;                                  // 2 x 65536
constexpr int SM_RING = 131072;                            // STAGES x 16384
constexpr int SM_SIDE = SM_RING + STAGES * CHUNK_BYTES;    // 12320
constexpr int SM_BAR = SM_SIDE + SIDE_FLOATS * 4;
constexpr int SM_FULL = SM_BAR;                            // STAGES mbarriers
constexpr int SM_EMPTY = SM_FULL + 8 * STAGES;             // STAGES
constexpr int SM_ACCF = SM_EMPTY + 8 * STAGES;             // 2 sub-tiles x 2 N-halves
constexpr int SM_ACTR = SM_ACCF + 32;                      // 2 sub-tiles x 2 K-halves
constexpr int SM_PFULL = SM_ACTR + 32;                     // STAGES mbarriers: peer CTA's ring slot is full (pair mode)
constexpr int SM_TMEM = SM_PFULL + 8 * STAGES;             // u32
constexpr int DIRB_ROWS = 5;                               // staged per-ray ddir biases per sub-tile (N >= 32 always fits)
constexpr int SM_TOK = SM_TMEM + 16;                       // 2 mbarriers: issue token of the alternating pair-mode issuers
constexpr int SM_DIRB = (SM_TOK + 16 + 15) & ~15;           // 2 x DIRB_ROWS x 128 floats (float4 aligned)
constexpr int SM_TOTAL = SM_DIRB + 2 * DIRB_ROWS * 512;
constexpr int SMEM_BYTES = SM_TOTAL + 1024;                // + alignment slack
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared memory limit");

constexpr int NUM_THREADS = 352;  // 8 worker warps (4 per sub-tile) + producer warp + 2 MMA issuer warps

// saved activation images per 128-row tile (bytes): written by the forward kernel in training mode,
// read by the weight-gradient kernel.  Every image is the exact shared-memory operand tile.
constexpr int64_t SAVE_ENC = 0;                         // [128 x 64] bf16(enc), col 63 = 0
constexpr int64_t SAVE_H = 16384;                       // + 65536 * i : output of trunk layer i (0..7)
constexpr int64_t SAVE_FEAT = 16384 + 8 * 65536;        // feature (linear) output
constexpr int64_t SAVE_HD = SAVE_FEAT + 65536;          // ddir output, 128 wide (32 KB)
constexpr int64_t SAVE_TILE_BYTES = SAVE_HD + 32768;    // 638976
// ReLU sign masks per tile: [9 layers][128 rows][8 x u32]  (layer 8 = ddir, 4 words used)
constexpr int64_t MASK_TILE_BYTES = 9 * 128 * 32;       // 36864
// gradient (dZ) images per tile, written by the backward chain kernel
constexpr int64_t DZ_Z = 0;                             // + 65536 * l : grad wrt pre-activation of trunk layer l
constexpr int64_t DZ_FEAT = 8 * 65536;
constexpr int64_t DZ_DDIR = 9 * 65536;                  // 128 wide (32 KB)
constexpr int64_t DZ_HEAD = DZ_DDIR + 32768;            // [128 x 64]: cols 0..2 = d rgb_raw, col 3 = d sigma_raw
constexpr int64_t DZ_TILE_BYTES = DZ_HEAD + 16384;      // 638976

// per-net layer offsets inside the flat fp32 blob (floats): d0..d7, sigma, feature, ddir, rgb
struct BlobOffsets {
    int64_t w[12];
    int64_t b[12];
};
template <class Ctx>
inline BlobOffsets make_offsets(const Ctx* ctx) {
    BlobOffsets off;
    for (int i = 0; i < 12; ++i) {
        off.w[i] = ctx->layers[i].w_off;
        off.b[i] = ctx->layers[i].b_off;
    }
    return off;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Per-thread store addresses of one activation-tile row: off[c] is the shared-memory address of 16-byte chunk c
// (8 bf16) of this row inside K-block 0, 128B-swizzled (chunk index XOR (row & 7)); K-block kb adds kb * 16 KB,
// which the compiler folds into the store's immediate offset.  Computed once per thread.
struct RowStore {
    uint32_t off[8];
    __device__ __forceinline__ void init(uint32_t act_base, int row) {
        const uint32_t rb = act_base + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) off[c] = rb + (uint32_t(c ^ (row & 7)) << 4);
    }
    template <int KB>
    __device__ __forceinline__ void store(int c, uint32_t a, uint32_t b, uint32_t cc, uint32_t d) const {
        asm volatile("st.shared.v4.b32 [%0+%5], {%1, %2, %3, %4};" ::"r"(off[c]), "r"(a), "r"(b), "r"(cc), "r"(d),
                     "n"(KB * TILE_M * 128)
                     : "memory");
    }
};

struct Barriers {
    uint32_t full, empty, accf, actr, pfull;
};

// pair = true: the CTA is one half of a cta_group::2 pair; the leader's A-tile barriers also collect the peer's workers
__device__ __forceinline__ void init_barriers(uint32_t base, Barriers& B, bool pair = false, bool shared_chunks = false,
                                              bool one_issuer = false) {
    B.full = base + SM_FULL; B.empty = base + SM_EMPTY; B.accf = base + SM_ACCF; B.actr = base + SM_ACTR;
    B.pfull = base + SM_PFULL;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(B.full + 8 * i, 1);
            mbar_init(B.empty + 8 * i, ((pair && !shared_chunks) || one_issuer) ? 1 : 2);   // commits per slot
            mbar_init(B.pfull + 8 * i, 1);
        }
        mbar_init(base + SM_TOK, 1);
        mbar_init(base + SM_TOK + 8, 1);
        for (int i = 0; i < 4; ++i) {                 // index = sub-tile * 2 + half
            mbar_init(B.accf + 8 * i, 1);
            mbar_init(B.actr + 8 * i, pair ? 2 * TILE_M : TILE_M);
        }
        fence_barrier_init();
    }
}

// ---- weight producer: one thread streams the chunk program `n_tiles` times through the ring ----
__device__ __forceinline__ void producer_loop(uint32_t base, const Barriers& B, const __nv_bfloat16* chunks,
                                              int n_chunks, int total_chunks) {
    int slot = 0, c = 0;
    uint32_t par = 1;  // first lap: slots are free (waiting on parity 1 of a fresh barrier returns at once)
    for (int g = 0; g < total_chunks; ++g) {
        mbar_wait(B.empty + 8 * slot, par, 1);
        mbar_arrive_expect_tx(B.full + 8 * slot, CHUNK_BYTES);
        bulk_g2s_keep(base + SM_RING + slot * CHUNK_BYTES, chunks + (size_t)c * CHUNK_ELEMS, CHUNK_BYTES, B.full + 8 * slot);
        if (++c == n_chunks) c = 0;
        if (++slot == STAGES) { slot = 0; par ^= 1; }
    }
}

// optional timeline trace (block 0 only): trace[((who * 3 + tile) * 16 + phase) * 4 + event] = clock64()
__device__ __forceinline__ void trace_ev(long long* trace, int who, int tile, int ph, int ev) {
    if (trace && blockIdx.x == 0 && tile < 3) trace[((who * 3 + tile) * 16 + ph) * 4 + ev] = clock64();
}

// ---- MMA issuers: one dedicated thread per sub-tile ------------------------------------------------
// Each issuer walks the chunk program for its own sub-tile: wait for the weight chunk, issue the K=16 MMAs,
// release the ring slot with tcgen05.commit (a slot is recycled when BOTH sub-tiles' commits have arrived).
// Hand-offs with the sub-tile's workers are split in halves so that epilogue and MMA overlap:
//   * the A tile is released in two K-halves: actr[lo] (K-blocks 0,1) gates the first chunk of a phase,
//     actr[hi] (K-blocks 2,3) gates k-block kbs/2 of the first N-half;
//   * the accumulator is handed over in two N-halves: accf[h0] after the last chunk of columns 0..127,
//     accf[h1] at the end of the phase (phases with a single N-half commit both at the end).
__device__ __forceinline__ void issuer_loop(uint32_t base, const Barriers& B, uint32_t tmem_base, const Program& prog,
                                            int s, int n_tiles, long long* trace = nullptr) {
    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    // descriptor template: LBO = 16 B (unused for swizzled K-major), SBO = 1024 B, version 1, SWIZZLE_128B
    const uint32_t desc_hi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));
    const uint32_t lbo_bits = (16u >> 4) << 16;
    const uint32_t a_tile = ((base + SM_ACT + s * 65536) & 0x3FFFF) >> 4;
    const uint32_t ring0 = ((base + SM_RING) & 0x3FFFF) >> 4;
    const uint32_t d_base = tmem_base + s * 256;
    const uint32_t bar_lo = B.actr + 16 * s, bar_hi = bar_lo + 8;
    const uint32_t bar_h0 = B.accf + 16 * s, bar_h1 = bar_h0 + 8;
    const int n_phases = prog.n_phases;
    int slot = 0;
    uint32_t ring_par = 0, actr_par = 0;
    for (int tile = 0; tile < n_tiles; ++tile) {
        for (int ph = 0; ph < n_phases; ++ph) {
            const int n_ch = prog.chunks[ph], kbs = prog.kb[ph], flags = prog.flags[ph];
            const int hi_kb = (kbs >= 4) ? (kbs >> 1) : 0;
            trace_ev(trace, s, tile, ph, 0);              // issuer: starts waiting for the A tile
            mbar_wait(bar_lo, actr_par, 3);
            if (hi_kb == 0) mbar_wait(bar_hi, actr_par, 5);
            trace_ev(trace, s, tile, ph, 1);              // issuer: A tile (first half) ready
            int h = 0, kb = 0;
            for (int j = 0; j < n_ch; ++j) {
                if (h == 0 && kb == hi_kb && hi_kb != 0) mbar_wait(bar_hi, actr_par, 5);
                mbar_wait(B.full + 8 * slot, ring_par, 4);
                tc_fence_after();
                const uint32_t a_lo = (a_tile + kb * ((TILE_M * 128) >> 4)) | lbo_bits;
                const uint32_t b_lo = (ring0 + slot * (CHUNK_BYTES >> 4)) | lbo_bits;
                const uint32_t d_tmem = d_base + h * 128;
                const bool acc0 = (kb > 0) || (flags & PH_ACC);
                const int n_mma = ((flags & PH_ENC) && kb == 1) ? 1 : 4;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < n_mma) {
                        const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2 * k);
                        const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2 * k);
                        mma_bf16_ss(d_tmem, ad, bd, idesc, (acc0 || k > 0) ? 1u : 0u);
                    }
                }
                mma_commit(B.empty + 8 * slot);
                if (++slot == STAGES) { slot = 0; ring_par ^= 1; }
                if (++kb == kbs) {
                    kb = 0;
                    if (h == 0 && j + 1 < n_ch) mma_commit(bar_h0);   // columns 0..127 complete, second half follows
                    ++h;
                }
            }
            if (n_ch == kbs) mma_commit(bar_h0);          // single N-half: hand over both barriers at the end
            mma_commit(bar_h1);
            actr_par ^= 1;
            trace_ev(trace, s, tile, ph, 2);              // issuer: all MMAs of the phase issued
        }
    }
}

// ---- weight-stationary issue: ONE issuer thread for both sub-tiles -----------------------------------------------------
// Both sub-tiles multiply their own A tile with the SAME weight chunk.  tcgen05.mma.ws keeps the B operand of an MMA in the
// collector: the pair (sub-tile 0: B_KEEP, sub-tile 1: B_REUSE) reads the 4 KB B slice from shared memory once instead of
// twice -- a quarter of the MMA operand traffic of the kernel, which is bound by exactly that (512 -> 384 KB per phase pair).
// The sub-tiles then advance in lockstep chunk by chunk; the N-half structure still overlaps the epilogue of columns 0..127
// with the MMAs of columns 128..255, and the K-half release overlaps the second epilogue half with the next phase's start.
__device__ __forceinline__ void issuer_loop_ws(uint32_t base, const Barriers& B, uint32_t tmem_base, const Program& prog,
                                               int n_tiles, long long* trace = nullptr, bool plain = false) {
    // executed by ALL lanes of the issuer warp (warp-uniform control flow); MMAs and commits are predicated on lane 0
    const uint32_t elect = (threadIdx.x & 31) == 0 ? 1u : 0u;
    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t desc_hi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));
    const uint32_t lbo_bits = (16u >> 4) << 16;
    const uint32_t a_tile0 = ((base + SM_ACT) & 0x3FFFF) >> 4, a_tile1 = ((base + SM_ACT + 65536) & 0x3FFFF) >> 4;
    const uint32_t ring0 = ((base + SM_RING) & 0x3FFFF) >> 4;
    const int n_phases = prog.n_phases;
    int slot = 0;
    uint32_t ring_par = 0, actr_par = 0;
    for (int tile = 0; tile < n_tiles; ++tile) {
        for (int ph = 0; ph < n_phases; ++ph) {
            const int n_ch = prog.chunks[ph], kbs = prog.kb[ph], flags = prog.flags[ph];
            const int hi_kb = (kbs >= 4) ? (kbs >> 1) : 0;
            if (elect) trace_ev(trace, 0, tile, ph, 0);
            mbar_wait(B.actr, actr_par, 3);                 // K-blocks 0,1 of both A tiles
            mbar_wait(B.actr + 16, actr_par, 3);
            if (hi_kb == 0) { mbar_wait(B.actr + 8, actr_par, 5); mbar_wait(B.actr + 24, actr_par, 5); }
            if (elect) trace_ev(trace, 0, tile, ph, 1);
            int h = 0, kb = 0;
            for (int j = 0; j < n_ch; ++j) {
                if (h == 0 && kb == hi_kb && hi_kb != 0) { mbar_wait(B.actr + 8, actr_par, 5); mbar_wait(B.actr + 24, actr_par, 5); }
                mbar_wait(B.full + 8 * slot, ring_par, 4);
                tc_fence_after();
                const uint32_t a0_lo = (a_tile0 + kb * ((TILE_M * 128) >> 4)) | lbo_bits;
                const uint32_t a1_lo = (a_tile1 + kb * ((TILE_M * 128) >> 4)) | lbo_bits;
                const uint32_t b_lo = (ring0 + slot * (CHUNK_BYTES >> 4)) | lbo_bits;
                const uint32_t d0 = tmem_base + h * 128, d1 = tmem_base + 256 + h * 128;
                const bool acc0 = (kb > 0) || (flags & PH_ACC);
                const int n_mma = ((flags & PH_ENC) && kb == 1) ? 1 : 4;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < n_mma) {
                        const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2 * k);
                        const uint64_t ad0 = ((uint64_t)desc_hi << 32) | (uint64_t)(a0_lo + 2 * k);
                        const uint64_t ad1 = ((uint64_t)desc_hi << 32) | (uint64_t)(a1_lo + 2 * k);
                        const uint32_t acc = (acc0 || k > 0) ? 1u : 0u;
                        if (plain) {     // experiment: the same single-warp issue order with ordinary MMAs
                            mma_bf16_uniform<0>(d0, ad0, bd, idesc, acc, elect);
                            mma_bf16_uniform<0>(d1, ad1, bd, idesc, acc, elect);
                        } else {
                            mma_bf16_uniform<1>(d0, ad0, bd, idesc, acc, elect);
                            mma_bf16_uniform<2>(d1, ad1, bd, idesc, acc, elect);
                        }
                    }
                }
                mma_commit_uniform(B.empty + 8 * slot, elect);
                if (++slot == STAGES) { slot = 0; ring_par ^= 1; }
                if (++kb == kbs) {
                    kb = 0;
                    if (h == 0 && j + 1 < n_ch) { mma_commit_uniform(B.accf, elect); mma_commit_uniform(B.accf + 16, elect); }
                    ++h;
                }
            }
            if (n_ch == kbs) { mma_commit_uniform(B.accf, elect); mma_commit_uniform(B.accf + 16, elect); }
            mma_commit_uniform(B.accf + 8, elect);
            mma_commit_uniform(B.accf + 24, elect);
            actr_par ^= 1;
            if (elect) trace_ev(trace, 0, tile, ph, 2);
        }
    }
}

// Two-warp version of the weight-stationary issue: warp `half` = 0 issues the MMAs of columns 0..127 (N-half 0) of BOTH
// sub-tiles, warp 1 those of columns 128..255, each keeping the shared B slice in its own collector buffer (b0 / b1).  The two
// warps own disjoint accumulator columns and disjoint chunks of the ring (every chunk is consumed by exactly one of them), so a
// single busy warp no longer has to issue all 64 MMAs of a phase (~165 cycles per MMA next to eight epilogue warps).
template <int HALF>
__device__ __forceinline__ void issuer_loop_ws2(uint32_t base, const Barriers& B, uint32_t tmem_base, const Program& prog,
                                                int n_tiles, long long* trace = nullptr) {
    const uint32_t elect = (threadIdx.x & 31) == 0 ? 1u : 0u;
    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t desc_hi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));
    const uint32_t lbo_bits = (16u >> 4) << 16;
    const uint32_t a_tile0 = ((base + SM_ACT) & 0x3FFFF) >> 4, a_tile1 = ((base + SM_ACT + 65536) & 0x3FFFF) >> 4;
    const uint32_t ring0 = ((base + SM_RING) & 0x3FFFF) >> 4;
    // __shfl_sync results are known to be warp-uniform: the descriptor arithmetic below then stays on the uniform datapath
    // instead of going through an ELECT / R2UR.BROADCAST loop per MMA
    const uint32_t d0 = __shfl_sync(0xffffffffu, tmem_base + HALF * 128, 0), d1 = d0 + 256;
    const int n_phases = prog.n_phases;
    uint32_t g = 0;             // ring position of the current phase's first chunk
    uint32_t actr_par = 0;
    for (int tile = 0; tile < n_tiles; ++tile) {
        for (int ph = 0; ph < n_phases; ++ph) {
            const int n_ch = prog.chunks[ph], kbs = prog.kb[ph], flags = prog.flags[ph];
            const bool single = (n_ch == kbs);                // the phase has N-half 0 only (128-wide layer)
            const int hi_kb = (kbs >= 4) ? (kbs >> 1) : 0;
            // every phase of the hand-off barriers has to be observed (also by the warp that has no work in a 128-wide phase):
            // a parity wait on a phase that has not even started yet succeeds at once
            if (elect) trace_ev(trace, HALF, tile, ph, 0);
            mbar_wait(B.actr, actr_par, 3);
            mbar_wait(B.actr + 16, actr_par, 3);
            if (HALF == 1 || hi_kb == 0) { mbar_wait(B.actr + 8, actr_par, 5); mbar_wait(B.actr + 24, actr_par, 5); }
            if (elect) trace_ev(trace, HALF, tile, ph, 1);
            if (HALF == 0 || !single) {
                for (int kb = 0; kb < kbs; ++kb) {
                    if (HALF == 0 && kb == hi_kb && hi_kb != 0) { mbar_wait(B.actr + 8, actr_par, 5); mbar_wait(B.actr + 24, actr_par, 5); }
                    const uint32_t pos = g + HALF * kbs + kb, lap = pos / STAGES, slot = pos - lap * STAGES;
                    mbar_wait(B.full + 8 * slot, lap & 1, 4);
                    tc_fence_after();
                    const uint32_t a0_lo = __shfl_sync(0xffffffffu, (a_tile0 + kb * ((TILE_M * 128) >> 4)) | lbo_bits, 0);
                    const uint32_t a1_lo = a0_lo + (65536 >> 4);
                    const uint32_t b_lo = __shfl_sync(0xffffffffu, (ring0 + slot * (CHUNK_BYTES >> 4)) | lbo_bits, 0);
                    const bool acc0 = (kb > 0) || (flags & PH_ACC);
                    const int n_mma = ((flags & PH_ENC) && kb == 1) ? 1 : 4;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k < n_mma) {
                            const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2 * k);
                            const uint64_t ad0 = ((uint64_t)desc_hi << 32) | (uint64_t)(a0_lo + 2 * k);
                            const uint64_t ad1 = ((uint64_t)desc_hi << 32) | (uint64_t)(a1_lo + 2 * k);
                            const uint32_t acc = (acc0 || k > 0) ? 1u : 0u;
                            mma_bf16_uniform<HALF == 0 ? 1 : 3>(d0, ad0, bd, idesc, acc, elect);
                            mma_bf16_uniform<HALF == 0 ? 2 : 4>(d1, ad1, bd, idesc, acc, elect);
                        }
                    }
                    mma_commit_uniform(B.empty + 8 * slot, elect);
                }
                mma_commit_uniform(B.accf + 8 * HALF, elect);           // this N-half of sub-tile 0 ...
                mma_commit_uniform(B.accf + 16 + 8 * HALF, elect);      // ... and of sub-tile 1
                if (HALF == 0 && single) { mma_commit_uniform(B.accf + 8, elect); mma_commit_uniform(B.accf + 24, elect); }
                if (elect) trace_ev(trace, HALF, tile, ph, 2);
            }
            actr_par ^= 1;
            g += n_ch;
        }
    }
}

// =====================================================================================================
// CTA-pair mode (cta_group::2).  Two CTAs of a cluster (an SM pair) run the same program on 2 x 2 sub-tiles.
// One MMA covers M = 256 rows (128 per CTA) x N = 256 columns; each CTA stages its own A rows and HALF of the
// weight rows (B), so per-SM shared-memory operand traffic is half of the single-CTA kernel's -- the resource
// that bounds it.  Only the leader (cluster rank 0) issues MMAs; hand-offs:
//   * weights: every CTA fills its own ring slot; the peer forwards "slot full" to the leader's pfull barrier;
//     tcgen05.commit multicasts "slot free" to both CTAs;
//   * A tiles: the workers of BOTH CTAs arrive on the leader's actr barriers (count 256);
//   * accumulators: tcgen05.commit multicasts to both CTAs' accf barriers.
// Chunk stream per CTA and phase: k-blocks kb = 0..kbs-1 of N-half `rank` (phases with a single N-half, i.e. the
// 128-wide ddir layer, take rows [64 rank, 64 rank + 64) of each chunk: 8 KB).
// =====================================================================================================
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int code) {
    uint32_t ok = 0, n = 0;
    long long t0 = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if ((++n & 4095u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > TC5_TIMEOUT_CYCLES) {
                printf("tc5: cluster mbarrier wait timeout code=%d block=%d thread=%d parity=%u\n", code, blockIdx.x,
                       threadIdx.x, parity);
                __trap();
            }
        }
    }
}

// Ring order in pair mode ("merged" order): for every phase first the chunks of sub-tile 0, then the same chunks again
// for sub-tile 1.  Each chunk has exactly one consumer, the two sub-tiles alternate on the tensor pipe phase by
// phase, and one sub-tile's whole epilogue runs under the other's MMA phase.
// shared = true: every chunk is loaded ONCE per phase and consumed by sub-tile 0 and then by sub-tile 1 (the slot is released
// by both commits): half the ring-fill traffic; the next phase's chunks stream in behind sub-tile 1's progress.
__device__ __forceinline__ void producer_loop_pair(uint32_t base, const Barriers& B, const __nv_bfloat16* chunks,
                                                   const Program& prog, int rank, int n_tiles, bool shared = false) {
    int slot = 0;
    uint32_t par = 1;
    for (int tile = 0; tile < n_tiles; ++tile) {
        int first = 0;
        for (int ph = 0; ph < prog.n_phases; ++ph) {
            const int n_ch = prog.chunks[ph], kbs = prog.kb[ph];
            const bool single = (n_ch == kbs);                   // one N-half only: split its rows between the CTAs
            const uint32_t bytes = single ? CHUNK_BYTES / 2 : CHUNK_BYTES;
            for (int sub = 0; sub < (shared ? 1 : 2); ++sub) {
                for (int kb = 0; kb < kbs; ++kb) {
                    const int c = first + (single ? kb : rank * kbs + kb);
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(chunks) + (size_t)c * CHUNK_BYTES +
                                         (single ? rank * (CHUNK_BYTES / 2) : 0);
                    mbar_wait(B.empty + 8 * slot, par, 1);
                    mbar_arrive_expect_tx(B.full + 8 * slot, bytes);
                    bulk_g2s_keep(base + SM_RING + slot * CHUNK_BYTES, src, bytes, B.full + 8 * slot);
                    if (++slot == STAGES) { slot = 0; par ^= 1; }
                }
            }
            first += n_ch;
        }
    }
}

// peer CTA: forward "my ring slot is full" to the leader
__device__ __forceinline__ void forwarder_loop_pair(const Barriers& B, int steps_total) {
    int slot = 0;
    uint32_t par = 0;
    for (int g = 0; g < steps_total; ++g) {
        mbar_wait(B.full + 8 * slot, par, 7);
        mbar_arrive_cluster(map_to_cta(B.pfull + 8 * slot, 0));
        if (++slot == STAGES) { slot = 0; par ^= 1; }
    }
}

// leader CTA: one issuer thread per sub-tile pair; sub-tile s owns merged positions [base + s*kbs, base + (s+1)*kbs)
__device__ __forceinline__ void issuer_loop_pair(uint32_t base, const Barriers& B, uint32_t tmem_base, const Program& prog,
                                                 int s, int n_tiles, long long* trace = nullptr, int dbg = 0,
                                                 bool shared = false) {
    const uint32_t desc_hi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));
    const uint32_t lbo_bits = (16u >> 4) << 16;
    const uint32_t a_tile = ((base + SM_ACT + s * 65536) & 0x3FFFF) >> 4;
    const uint32_t ring0 = ((base + SM_RING) & 0x3FFFF) >> 4;
    const uint32_t d_tmem = tmem_base + s * 256;
    const uint32_t bar_lo = B.actr + 16 * s, bar_hi = bar_lo + 8;
    const uint32_t bar_h0 = B.accf + 16 * s, bar_h1 = bar_h0 + 8;
    const uint32_t idesc256 = make_idesc_bf16(256, 256, 0, 0), idesc128 = make_idesc_bf16(256, 128, 0, 0);
    uint32_t g_base = 0;          // merged ring position of the current phase's first chunk
    uint32_t actr_par = 0;
    (void)dbg;
    (void)shared;
    for (int tile = 0; tile < n_tiles; ++tile) {
        for (int ph = 0; ph < prog.n_phases; ++ph) {
            const int n_ch = prog.chunks[ph], kbs = prog.kb[ph], flags = prog.flags[ph];
            const uint32_t idesc = (n_ch == kbs) ? idesc128 : idesc256;
            trace_ev(trace, s, tile, ph, 0);
            mbar_wait_cluster(bar_lo, actr_par, 3);
            mbar_wait_cluster(bar_hi, actr_par, 5);
            trace_ev(trace, s, tile, ph, 1);
            uint32_t w_full = 0, w_pfull = 0;     // diagnostics: cycles spent waiting for weight chunks
            for (int kb = 0; kb < kbs; ++kb) {
                const uint32_t g = g_base + (shared ? 0 : s * kbs) + kb;
                const uint32_t lap = g / STAGES;
                const uint32_t slot = g - lap * STAGES;
                const uint32_t ring_par = lap & 1;
                const long long tw0 = trace ? clock64() : 0;
                mbar_wait(B.full + 8 * slot, ring_par, 4);
                const long long tw1 = trace ? clock64() : 0;
                mbar_wait_cluster(B.pfull + 8 * slot, ring_par, 8);
                if (trace) { w_full += (uint32_t)(tw1 - tw0); w_pfull += (uint32_t)(clock64() - tw1); }
                tc_fence_after();
                const uint32_t a_lo = (a_tile + kb * ((TILE_M * 128) >> 4)) | lbo_bits;
                const uint32_t b_lo = (ring0 + slot * (CHUNK_BYTES >> 4)) | lbo_bits;
                const bool acc0 = (kb > 0) || (flags & PH_ACC);
                const int n_mma = ((flags & PH_ENC) && kb == 1) ? 1 : 4;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < n_mma) {
                        const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2 * k);
                        const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2 * k);
                        mma_bf16_ss_2cta(d_tmem, ad, bd, idesc, (acc0 || k > 0) ? 1u : 0u);
                    }
                }
                mma_commit_2cta(B.empty + 8 * slot, 0x3);      // frees the slot in both CTAs
            }
            mma_commit_2cta(bar_h0, 0x3);                       // all 256 columns complete at once in pair mode
            mma_commit_2cta(bar_h1, 0x3);
            actr_par ^= 1;
            g_base += shared ? kbs : 2 * kbs;
            trace_ev(trace, s, tile, ph, 2);
            if (trace && blockIdx.x == 0 && tile < 3)
                trace[((s * 3 + tile) * 16 + ph) * 4 + 3] = (long long)w_full | ((long long)w_pfull << 32);
        }
    }
}

// Pair mode with SHARED weight chunks: two issuer threads (leader CTA, one per sub-tile) that ALTERNATE on the tensor pipe
// through a token -- sub-tile 0 phase p, sub-tile 1 phase p, sub-tile 0 phase p+1, ... -- so that the MMA windows of the two
// sub-tiles never interleave and each sub-tile's epilogue runs under the other's window.  Every chunk of a phase is loaded
// once per CTA, consumed by both sub-tiles and released by both commits.  A thread waits for everything its window needs
// (the phase's chunk barriers of both CTAs, its A tile) while the OTHER thread issues; holding the token it only issues:
// a cta_group::2 MMA costs ~105 cycles to issue, 16 of them fit under the 2 048 cycles they take to execute.
__device__ __forceinline__ void issuer_loop_pair_shared(uint32_t base, const Barriers& B, uint32_t tmem_base,
                                                        const Program& prog, int s, int n_tiles, long long* trace = nullptr) {
    const uint32_t desc_hi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));
    const uint32_t lbo_bits = (16u >> 4) << 16;
    const uint32_t ring0 = ((base + SM_RING) & 0x3FFFF) >> 4;
    const uint32_t idesc256 = make_idesc_bf16(256, 256, 0, 0), idesc128 = make_idesc_bf16(256, 128, 0, 0);
    const uint32_t a_tile = ((base + SM_ACT + s * 65536) & 0x3FFFF) >> 4;
    const uint32_t d_tmem = tmem_base + s * 256;
    const uint32_t bar_lo = B.actr + 16 * s, bar_hi = bar_lo + 8;
    const uint32_t bar_h0 = B.accf + 16 * s, bar_h1 = bar_h0 + 8;
    const uint32_t tok_mine = base + SM_TOK + 8 * s, tok_other = base + SM_TOK + 8 * (1 - s);
    uint32_t g_base = 0, actr_par = 0, n = 0;
    for (int tile = 0; tile < n_tiles; ++tile) {
        for (int ph = 0; ph < prog.n_phases; ++ph, ++n) {
            const int n_ch = prog.chunks[ph], kbs = prog.kb[ph], flags = prog.flags[ph];
            const uint32_t idesc = (n_ch == kbs) ? idesc128 : idesc256;
            for (int kb = 0; kb < kbs; ++kb) {
                const uint32_t g = g_base + kb, lap = g / STAGES, slot = g - lap * STAGES;
                mbar_wait(B.full + 8 * slot, lap & 1, 4);
                mbar_wait_cluster(B.pfull + 8 * slot, lap & 1, 8);
            }
            trace_ev(trace, s, tile, ph, 0);
            mbar_wait_cluster(bar_lo, actr_par, 3);
            mbar_wait_cluster(bar_hi, actr_par, 5);
            // the token: sub-tile 0 owns it at the start; afterwards each window is handed over by the other thread
            if (s == 1) mbar_wait(tok_mine, n & 1, 10);
            else if (n > 0) mbar_wait(tok_mine, (n - 1) & 1, 10);
            trace_ev(trace, s, tile, ph, 1);
            tc_fence_after();
            for (int kb = 0; kb < kbs; ++kb) {
                const uint32_t g = g_base + kb, lap = g / STAGES, slot = g - lap * STAGES;
                const uint32_t a_lo = (a_tile + kb * ((TILE_M * 128) >> 4)) | lbo_bits;
                const uint32_t b_lo = (ring0 + slot * (CHUNK_BYTES >> 4)) | lbo_bits;
                const bool acc0 = (kb > 0) || (flags & PH_ACC);
                const int n_mma = ((flags & PH_ENC) && kb == 1) ? 1 : 4;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < n_mma) {
                        const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2 * k);
                        const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2 * k);
                        mma_bf16_ss_2cta(d_tmem, ad, bd, idesc, (acc0 || k > 0) ? 1u : 0u);
                    }
                }
                mma_commit_2cta(B.empty + 8 * slot, 0x3);      // the slot is recycled when both sub-tiles' commits arrived
            }
            mma_commit_2cta(bar_h0, 0x3);
            mma_commit_2cta(bar_h1, 0x3);
            mbar_arrive(tok_other);
            trace_ev(trace, s, tile, ph, 2);
            actr_par ^= 1;
            g_base += kbs;
        }
    }
}

}  // namespace tcmlp
