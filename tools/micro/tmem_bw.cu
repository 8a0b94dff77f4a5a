// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM with 4 / 8 warps.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../nerf_keras_b200/csrc/tc5.cuh"
using namespace tc5;

__global__ void __launch_bounds__(256, 1) tmem_ld_kernel(int reps, long long* out, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = slot;
    const uint32_t t_lane = tb + (uint32_t(32 * (warp & 3)) << 16) + (warp >> 2) * 256;
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int cg = 0; cg < 8; ++cg) {
            uint32_t v[32];
            tmem_ld32(t_lane + cg * 32, v);
            tmem_ld_wait();
            acc += __uint_as_float(v[0]) + __uint_as_float(v[31]);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 12345.f) sink[0] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
    long long* out; float* sink; cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
    for (int threads : {128, 256}) {
        for (int reps : {16, 256}) {
            tmem_ld_kernel<<<1, threads>>>(reps, out, sink);
            long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
            double bytes = (double)threads * 8 * 32 * 4 * reps;
            printf("threads=%d reps=%d: %lld cycles, %.1f B/cycle (%.0f cycles per 128x256 fp32 tile)\n", threads, reps, c, bytes / c,
                   (double)c / reps / (threads / 128));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
