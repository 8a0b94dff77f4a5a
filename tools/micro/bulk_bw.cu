// Micro-benchmark: HBM read bandwidth of cp.async.bulk (UBLKCP) global->shared as a function of copy
// size and number of copies in flight per CTA (1 CTA per SM).  nvcc -arch=sm_100a bulk_bw.cu -o bulk_bw
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../nerf_keras_b200/csrc/tc5.cuh"
using namespace tc5;

// each CTA streams its own contiguous region (region_bytes) or a strided pattern (stride between copies)
// tiled != 0: the weight-gradient kernel's pattern instead -- CTA b reads image (b / 4) of tiles (b % 4), (b % 4) + 4, ..
// of `tiled`-byte tile records (one copy per tile; the records of consecutive copies are 4 x tiled bytes apart)
__global__ void __launch_bounds__(64, 1) bw_kernel(const uint8_t* src, size_t region_bytes, int copy_bytes, int depth,
                                                   size_t stride, int iters, size_t tiled = 0, size_t total = 0) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bar0 = base;                 // depth mbarriers
    const uint32_t buf0 = base + 1024;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) mbar_init(bar0 + 8 * i, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint8_t* p = src + (size_t)blockIdx.x * region_bytes;
        size_t off = 0;
        if (tiled) {
            p = src;
            off = (size_t)(blockIdx.x % 4) * tiled + (size_t)(blockIdx.x / 4) * 65536;
            stride = 4 * tiled;
            region_bytes = total - tiled;
        }
        int slot = 0; uint32_t par = 0;
        // prime
        for (int i = 0; i < depth && i < iters; ++i) {
            mbar_arrive_expect_tx(bar0 + 8 * i, copy_bytes);
            bulk_g2s(buf0 + i * copy_bytes, p + off, copy_bytes, bar0 + 8 * i);
            off += stride; if (off + copy_bytes > region_bytes) off = 0;
        }
        for (int i = 0; i < iters; ++i) {
            mbar_wait(bar0 + 8 * slot, par, 1);
            if (i + depth < iters) {
                mbar_arrive_expect_tx(bar0 + 8 * slot, copy_bytes);
                bulk_g2s(buf0 + slot * copy_bytes, p + off, copy_bytes, bar0 + 8 * slot);
                off += stride; if (off + copy_bytes > region_bytes) off = 0;
            }
            if (++slot == depth) { slot = 0; par ^= 1; }
        }
    }
}

int main(int argc, char** argv) {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    if (argc > 1) sms = atoi(argv[1]);      // CTA count (default: one per SM): per-SM streaming limit at a reduced grid
    const size_t region = 64ull << 20;                 // 64 MiB per CTA -> 9.25 GiB total, far beyond L2
    uint8_t* src; cudaMalloc(&src, region * sms); cudaMemset(src, 1, region * sms);
    cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int sizes[] = {2048, 8192, 16384, 32768, 65536};
    const int inflight_kb[] = {32, 64, 128, 192};
    for (int pattern = 0; pattern < 2; ++pattern)
    for (int s : sizes) for (int kb : inflight_kb) {
        int depth = kb * 1024 / s; if (depth < 1 || depth > 96) continue;
        size_t stride = pattern == 0 ? (size_t)s : (size_t)s * 5 + 81920;   // pattern 1: tile-strided like the image layout
        int iters = (int)((48ull << 20) / s);
        bw_kernel<<<sms, 64, 1024 + 1024 + depth * s, 0>>>(src, region, s, depth, stride, 64);  // warm-up
        cudaEventRecord(a);
        bw_kernel<<<sms, 64, 1024 + 1024 + depth * s, 0>>>(src, region, s, depth, stride, iters);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        cudaError_t e = cudaGetLastError();
        printf("ctas=%d pattern=%s copy=%6d B depth=%3d (%3d KB in flight/SM): %8.1f GB/s %s\n", sms, pattern ? "strided" : "seq", s, depth, kb,
               (double)iters * s * sms / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    // weight-gradient-like pattern: 64 KB images out of 640 KB tile records vs the same copies from image-major storage
    for (size_t rec : {(size_t)655360, (size_t)65536}) for (int depth : {1, 2, 3}) {
        const int s = 65536;
        int iters = (int)((48ull << 20) / s);
        cudaEventRecord(a);
        bw_kernel<<<sms, 64, 1024 + 1024 + depth * s, 0>>>(src, region, s, depth, 0, iters, rec, region * sms);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("ctas=%d pattern=tiles of %zu B copy=%6d B depth=%3d: %8.1f GB/s %s\n", sms, rec, s, depth,
               (double)iters * s * sms / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
