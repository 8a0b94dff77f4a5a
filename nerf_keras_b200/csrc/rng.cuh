// Uniform draws of sample_pdf (data_utils.py:196: `u = tf.random.uniform(shape=[batch, ns_fine])`, always random in the
// reference, also at inference).  Two sources:
//   * an explicit (B, ns_fine) buffer (parity mode: the oracle and the kernels consume the same numbers);
//   * Philox4x32-10 evaluated inside the consuming kernel, keyed by (seed, counter, ray, draw): nothing is written to or
//     read from HBM, the forward resampling kernel and the backward of sample_pdf regenerate identical values, and the
//     counter is read from device memory so that a replayed CUDA graph draws fresh numbers every step.
#pragma once
#include <stdint.h>

namespace nerf {

struct PdfDraws {
    const float* u;                          // explicit draws, or nullptr for the generator
    unsigned long long seed, counter;
    const unsigned long long* counter_dev;   // optional: *counter_dev is added to `counter` (optimiser step)
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}

// u in [0, 1): 24 random mantissa bits, like tf.random.uniform for float32
__device__ __forceinline__ float pdf_draw(const PdfDraws& dr, unsigned long long step, int64_t ray, int nf, int j) {
    if (dr.u) return dr.u[ray * nf + j];
    const unsigned long long idx = (unsigned long long)ray * (unsigned)nf + (unsigned)j;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)step, (uint32_t)(step >> 32)),
                                  make_uint2((uint32_t)dr.seed, (uint32_t)(dr.seed >> 32)));
    return (float)(r.x >> 8) * 5.9604644775390625e-08f;   // 2^-24
}
__device__ __forceinline__ unsigned long long pdf_step(const PdfDraws& dr) {
    return dr.counter + (dr.counter_dev ? *dr.counter_dev : 0ull);
}

}  // namespace nerf
