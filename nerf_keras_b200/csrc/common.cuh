// Shared host-side helpers for libnerf_b200.so: error reporting, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/nerf_b200.h"

namespace nerf {

extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define NERF_CHECK_ARG(cond, msg)                                                    \
    do {                                                                             \
        if (!(cond)) return ::nerf::fail(NERF_ERR_INVALID, std::string(__func__) + ": " + (msg)); \
    } while (0)

#define NERF_CUDA(expr)                                                                              \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return ::nerf::fail(NERF_ERR_CUDA, std::string(__func__) + ": " #expr " -> " + cudaGetErrorString(_e)); \
    } while (0)

// every kernel launch goes through this so bench.py can report gpu_launches
#define NERF_LAUNCHED()                        \
    do {                                       \
        ::nerf::g_launches.fetch_add(1);       \
        NERF_CUDA(cudaGetLastError());         \
    } while (0)

// optional per-kernel timing of the fused MLP kernels (bench.py's roofline leg): when enabled every
// fused-MLP launch is bracketed by cudaEvents on its stream; nerf_timing_read() sums them.
void timing_begin(int kind, cudaStream_t st);
void timing_end(int kind, cudaStream_t st);

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

inline int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// grid for a grid-stride elementwise kernel: a multiple of the SM count, capped by the work
inline int stream_grid(int64_t work_items, int threads, int blocks_per_sm = 8) {
    int64_t need = ceil_div(work_items, threads);
    int64_t cap = (int64_t)num_sms() * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

}  // namespace nerf
