// HBM-bound ray / sampling / compositing kernels of the NeRF hot path (sm_100a).
//
// Each kernel restates one function of the reference's data_utils.py (file:line cited at the C-ABI
// entry point in include/nerf_b200.h).  Ray and t-value generation is BIT-EXACT with respect to the
// reference's op order: separately rounded __f*_rn intrinsics, no FMA contraction, true division.
#include "common.cuh"
#include "ctx.cuh"

namespace nerf {
thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};
}  // namespace nerf

using namespace nerf;

// ------------------------------------------------------------------------------------------------
// get_rays  (data_utils.py:23-52)
// Output is produced as flat float4 stores over the (H*W*3) arrays: fully coalesced 16-byte writes,
// 24 algorithmic bytes per ray.  Each thread recomputes the (at most two) pixels its float4 touches.
// ------------------------------------------------------------------------------------------------
struct Pose3x4 {
    float r[3][3];
    float t[3];
};

// direction of pixel (h, w): three separately rounded products, reduced left to right (no FMA)
__device__ __forceinline__ void ray_dir_pixel(int h, int w, float half_w, float half_h, float focal, const Pose3x4& P,
                                              float (&out)[3]) {
    const float tu = __fdiv_rn(__fsub_rn((float)w, half_w), focal);   // (u - W*0.5) / focal
    const float tv = __fdiv_rn(__fsub_rn((float)h, half_h), focal);   // (v - H*0.5) / focal
    const float dc0 = tu, dc1 = -tv, dc2 = -1.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        out[i] = __fadd_rn(__fadd_rn(__fmul_rn(dc0, P.r[i][0]), __fmul_rn(dc1, P.r[i][1])), __fmul_rn(dc2, P.r[i][2]));
}

// One thread computes 4 consecutive pixels (12 floats).  The block's 256 x 12 floats are staged in shared memory
// so that the global stores are fully coalesced float4 writes (thread i writes float4 i, i+256, i+512 of the
// block's contiguous 12 KB span); origins are a repeating 3-float4 pattern and need no staging.
__global__ void __launch_bounds__(256) get_rays_kernel(int width, int64_t n_pix, float half_w, float half_h,
                                                       float focal, Pose3x4 P, float* __restrict__ o,
                                                       float* __restrict__ d) {
    __shared__ float4 stage[256 * 3];
    const int64_t n_grp = n_pix >> 2;
    const float4 opat[3] = {make_float4(P.t[0], P.t[1], P.t[2], P.t[0]), make_float4(P.t[1], P.t[2], P.t[0], P.t[1]),
                            make_float4(P.t[2], P.t[0], P.t[1], P.t[2])};
    float p2[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) p2[i] = __fmul_rn(-1.0f, P.r[i][2]);
    for (int64_t g0 = (int64_t)blockIdx.x * 256; g0 < n_grp; g0 += (int64_t)gridDim.x * 256) {
        const int64_t g = g0 + threadIdx.x;
        if (g < n_grp) {
            const int64_t pix0 = g << 2;
            int h = (int)(pix0 / width);
            int w = (int)(pix0 - (int64_t)h * width);
            float v[12];
            // the row term -tv * R[i][1] and the constant -R[i][2] are shared by the pixels of a row (same roundings
            // as evaluating them per pixel); only tu = (w - W/2) / focal changes along the row
            float p1[3];
            int h_cached = -1;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (h != h_cached) {
                    const float ntv = -__fdiv_rn(__fsub_rn((float)h, half_h), focal);
#pragma unroll
                    for (int i = 0; i < 3; ++i) p1[i] = __fmul_rn(ntv, P.r[i][1]);
                    h_cached = h;
                }
                const float tu = __fdiv_rn(__fsub_rn((float)w, half_w), focal);
#pragma unroll
                for (int i = 0; i < 3; ++i) v[3 * k + i] = __fadd_rn(__fadd_rn(__fmul_rn(tu, P.r[i][0]), p1[i]), p2[i]);
                if (++w == width) { w = 0; ++h; }
            }
            stage[threadIdx.x * 3] = make_float4(v[0], v[1], v[2], v[3]);
            stage[threadIdx.x * 3 + 1] = make_float4(v[4], v[5], v[6], v[7]);
            stage[threadIdx.x * 3 + 2] = make_float4(v[8], v[9], v[10], v[11]);
        }
        __syncthreads();
        const int64_t n_valid = ((n_grp - g0) < 256 ? (n_grp - g0) : 256) * 3;   // float4s produced by this block
        float4* dp = reinterpret_cast<float4*>(d) + g0 * 3;
        float4* op = reinterpret_cast<float4*>(o) + g0 * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int j = threadIdx.x + 256 * k;
            if (j < n_valid) {
                dp[j] = stage[j];
                op[j] = opat[j % 3];
            }
        }
        __syncthreads();
    }
    // tail (n_pix % 4 pixels)
    if (blockIdx.x == 0 && threadIdx.x < (n_pix & 3)) {
        const int64_t pix = (n_grp << 2) + threadIdx.x;
        const int h = (int)(pix / width), w = (int)(pix - (int64_t)h * width);
        float dd[3];
        ray_dir_pixel(h, w, half_w, half_h, focal, P, dd);
#pragma unroll
        for (int i = 0; i < 3; ++i) { d[pix * 3 + i] = dd[i]; o[pix * 3 + i] = P.t[i]; }
    }
}

extern "C" int nerf_get_rays(int height, int width, float focal, const float* pose, float* o, float* d, void* stream) {
    NERF_CHECK_ARG(height > 0 && width > 0, "height/width must be positive");
    NERF_CHECK_ARG(pose && o && d, "null pointer");
    Pose3x4 P;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) P.r[i][j] = pose[i * 4 + j];
        P.t[i] = pose[i * 4 + 3];
    }
    int64_t n_pix = (int64_t)height * width;
    // `width * 0.5` is a Python float (double) that TF converts to float32 when it meets the f32 tensor
    float half_w = (float)((double)width * 0.5), half_h = (float)((double)height * 0.5);
    int threads = 256;
    int grid = stream_grid(ceil_div(n_pix, 4), threads);
    get_rays_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(width, n_pix, half_w, half_h, focal, P, o, d);
    NERF_LAUNCHED();
    return NERF_OK;
}

// ------------------------------------------------------------------------------------------------
// ndc_rays (extension) -- same op order as oracle.ndc_rays
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ndc_rays_kernel(float sx, float sy, float near_p, const float* __restrict__ oi,
                                                       const float* __restrict__ di, float* __restrict__ oo,
                                                       float* __restrict__ dout, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float ox = oi[3 * i], oy = oi[3 * i + 1], oz = oi[3 * i + 2];
        float dx = di[3 * i], dy = di[3 * i + 1], dz = di[3 * i + 2];
        float t = __fdiv_rn(-__fadd_rn(near_p, oz), dz);
        ox = __fadd_rn(ox, __fmul_rn(t, dx));
        oy = __fadd_rn(oy, __fmul_rn(t, dy));
        oz = __fadd_rn(oz, __fmul_rn(t, dz));
        float two_n = __fmul_rn(2.0f, near_p);
        float oxz = __fdiv_rn(ox, oz), oyz = __fdiv_rn(oy, oz);
        oo[3 * i] = __fmul_rn(sx, oxz);
        oo[3 * i + 1] = __fmul_rn(sy, oyz);
        oo[3 * i + 2] = __fadd_rn(1.0f, __fdiv_rn(two_n, oz));
        dout[3 * i] = __fmul_rn(sx, __fsub_rn(__fdiv_rn(dx, dz), oxz));
        dout[3 * i + 1] = __fmul_rn(sy, __fsub_rn(__fdiv_rn(dy, dz), oyz));
        dout[3 * i + 2] = __fdiv_rn(-two_n, oz);
    }
}

extern "C" int nerf_ndc_rays(int height, int width, float focal, float near_plane, const float* o_in,
                             const float* d_in, float* o_out, float* d_out, int64_t n, void* stream) {
    NERF_CHECK_ARG(o_in && d_in && o_out && d_out && n >= 0, "bad arguments");
    if (n == 0) return NERF_OK;
    float sx = -(focal / (float)((double)width * 0.5));
    float sy = -(focal / (float)((double)height * 0.5));
    ndc_rays_kernel<<<stream_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(sx, sy, near_plane, o_in, d_in, o_out,
                                                                           d_out, n);
    NERF_LAUNCHED();
    return NERF_OK;
}

// ------------------------------------------------------------------------------------------------
// generate_t_vals  (data_utils.py:119-138): TF linspace (exact ends, start + delta*i inside, separate
// mul/add) + shared jitter (u * (far-near)) / N, broadcast to (B,N).  4*N algorithmic bytes per ray.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tval_at(int n, int N, float near_f, float far_f, float delta) {
    if (n == 0) return near_f;
    if (n == N - 1) return far_f;
    return __fadd_rn(near_f, __fmul_rn(delta, (float)n));
}

__global__ void __launch_bounds__(256) t_vals_kernel(float near_f, float far_f, float range_f, int N, int64_t B,
                                                     const float* __restrict__ u, int u_per_ray,
                                                     float* __restrict__ t) {
    const float delta = (N > 1) ? __fdiv_rn(__fsub_rn(far_f, near_f), (float)(N - 1)) : 0.0f;
    const float fN = (float)N;
    const int64_t n_el = B * N;
    if ((N & 3) == 0 && !u_per_ray && (blockDim.x % (N >> 2)) == 0) {
        // every ray gets the same row (the reference's single shared jitter vector, Q1): a thread always writes the
        // same column group, so its float4 is computed once and the kernel is a pure streaming store
        const int gpr = N >> 2;                          // float4 groups per row
        const int cg = threadIdx.x % gpr;
        float r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float x = tval_at(cg * 4 + k, N, near_f, far_f, delta);
            if (u) x = __fadd_rn(x, __fdiv_rn(__fmul_rn(u[cg * 4 + k], range_f), fN));
            r[k] = x;
        }
        const float4 val = make_float4(r[0], r[1], r[2], r[3]);
        const int64_t n_vec = n_el >> 2;
        for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n_vec; v += (int64_t)gridDim.x * blockDim.x)
            reinterpret_cast<float4*>(t)[v] = val;       // (grid stride is a multiple of gpr: column group is invariant)
    } else if ((N & 3) == 0) {
        const int64_t n_vec = n_el >> 2;
        for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n_vec;
             v += (int64_t)gridDim.x * blockDim.x) {
            int64_t e0 = v << 2;
            int n0 = (int)(e0 % N);
            float r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float x = tval_at(n0 + k, N, near_f, far_f, delta);
                if (u) {
                    float uu = u_per_ray ? u[e0 + k] : u[n0 + k];
                    x = __fadd_rn(x, __fdiv_rn(__fmul_rn(uu, range_f), fN));
                }
                r[k] = x;
            }
            reinterpret_cast<float4*>(t)[v] = make_float4(r[0], r[1], r[2], r[3]);
        }
    } else {
        for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_el;
             e += (int64_t)gridDim.x * blockDim.x) {
            int n = (int)(e % N);
            float x = tval_at(n, N, near_f, far_f, delta);
            if (u) {
                float uu = u_per_ray ? u[e] : u[n];
                x = __fadd_rn(x, __fdiv_rn(__fmul_rn(uu, range_f), fN));
            }
            t[e] = x;
        }
    }
}

extern "C" int nerf_generate_t_vals(double near_plane, double far_plane, int64_t batch, int num_samples,
                                    const float* u, int u_per_ray, float* t, void* stream) {
    NERF_CHECK_ARG(batch >= 0 && num_samples >= 1 && t, "bad arguments");
    if (batch == 0) return NERF_OK;
    // near/far are Python floats (doubles) in the reference: linspace casts each to f32, while
    // `(far - near)` is evaluated in double before it meets the f32 noise tensor (data_utils.py:131-133)
    float range_f = (float)(far_plane - near_plane);
    int64_t work = (num_samples & 3) == 0 ? batch * num_samples / 4 : batch * num_samples;
    t_vals_kernel<<<stream_grid(work, 256), 256, 0, (cudaStream_t)stream>>>((float)near_plane, (float)far_plane,
                                                                            range_f, num_samples, batch, u,
                                                                            u_per_ray, t);
    NERF_LAUNCHED();
    return NERF_OK;
}

// ------------------------------------------------------------------------------------------------
// sample_rays  (data_utils.py:55-73)   rays = o + (d * t), dirs = broadcast(d)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_rays_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                                          const float* __restrict__ t, int64_t B, int N,
                                                          float* __restrict__ rays, float* __restrict__ dirs) {
    const int64_t n_el = B * N * 3;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_el; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t s = e / 3;
        int c = (int)(e - s * 3);
        int64_t b = s / N;
        float dv = d[b * 3 + c];
        rays[e] = __fadd_rn(o[b * 3 + c], __fmul_rn(dv, t[s]));
        dirs[e] = dv;
    }
}

extern "C" int nerf_sample_rays(const float* o, const float* d, const float* t, int64_t batch, int num_samples,
                                float* rays, float* dirs, void* stream) {
    NERF_CHECK_ARG(o && d && t && rays && dirs && batch >= 0 && num_samples >= 1, "bad arguments");
    if (batch == 0) return NERF_OK;
    sample_rays_kernel<<<stream_grid(batch * num_samples * 3, 256), 256, 0, (cudaStream_t)stream>>>(
        o, d, t, batch, num_samples, rays, dirs);
    NERF_LAUNCHED();
    return NERF_OK;
}

// ------------------------------------------------------------------------------------------------
// encode_position  (data_utils.py:7-21)  accurate fp32 sinf/cosf (arguments reach ~2^9 * |x|)
// one output element per thread-iteration: stores are fully coalesced over the (n, 3+6L) array
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) encode_position_kernel(const float* __restrict__ x, int64_t n, int L,
                                                              float* __restrict__ out) {
    const int C = 3 + 6 * L;
    const int64_t n_el = n * C;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_el; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t p = e / C;
        int c = (int)(e - p * C);
        float v;
        if (c < 3) {
            v = x[p * 3 + c];
        } else {
            int q = c - 3;
            int i = q / 6;
            int r = q - i * 6;
            int comp = r % 3;
            float arg = __fmul_rn(exp2f((float)i), x[p * 3 + comp]);
            v = (r >= 3) ? cosf(arg) : sinf(arg);
        }
        out[e] = v;
    }
}

extern "C" int nerf_encode_position(const float* x, int64_t n, int L, float* out, void* stream) {
    NERF_CHECK_ARG(x && out && n >= 0 && L >= 0 && L <= 24, "bad arguments");
    if (n == 0) return NERF_OK;
    encode_position_kernel<<<stream_grid(n * (3 + 6 * L), 256), 256, 0, (cudaStream_t)stream>>>(x, n, L, out);
    NERF_LAUNCHED();
    return NERF_OK;
}

// ------------------------------------------------------------------------------------------------
// volume_render  (data_utils.py:75-98)   one warp per ray; lanes interleave over samples so the
// float4 pred loads are contiguous; exclusive cumprod = per-32-chunk warp scan with a running carry.
// Algorithmic bytes: 24 per sample (16 preds + 4 t + 4 weights) + 20 per ray.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_incl_scan_mul(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v *= n;
    }
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
// forward compositing only needs sigmoid to ~1e-6 absolute (rgb tolerance 1e-5): ex2.approx based exp + fast divide
__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// NCH = number of 32-sample chunks per ray, held entirely in registers; the NEXT ray's chunks are loaded before the
// current ray is processed, so each warp keeps two rays (2 * N * 20 bytes) in flight -- compositing is latency-
// bound on HBM otherwise (one 640-byte chunk in flight per warp caps the kernel near 3.5 TB/s).
template <int NCH>
__global__ void __launch_bounds__(256) volume_render_kernel(const float4* __restrict__ preds,
                                                            const float* __restrict__ t, int64_t B, int N,
                                                            float* __restrict__ rgb, float* __restrict__ depth,
                                                            float* __restrict__ weights, float* __restrict__ acc) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float4 cp[NCH], np_[NCH];
    float ct[NCH], nt[NCH];
    auto load = [&](int64_t r, float4 (&p)[NCH], float (&tt)[NCH]) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int n = c * 32 + lane;
            const bool ok = (r < B) && (n < N);
            p[c] = ok ? __ldg(preds + r * N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
            tt[c] = ok ? __ldg(t + r * N + n) : 0.f;
        }
    };
    load(ray, cp, ct);
    for (; ray < B; ray += warps_total) {
        load(ray + warps_total, np_, nt);
        float carry = 1.0f;  // exclusive transmittance entering this chunk
        float sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int n = c * 32 + lane;
            const bool ok = n < N;
            const float4 pr = cp[c];
            const float tn = ct[c];
            // t[n+1]: neighbour lane, or lane 0 of the next chunk for lane 31
            float tn1 = __shfl_down_sync(0xffffffffu, tn, 1);
            const float t_first_next = __shfl_sync(0xffffffffu, ct[(c + 1 < NCH) ? c + 1 : c], 0);
            if (lane == 31) tn1 = t_first_next;
            const float delta = (n == N - 1) ? 1e10f : (tn1 - tn);
            const float sigma = fmaxf(pr.w, 0.0f);
            const float alpha = 1.0f - expf(-sigma * delta);
            const float x = ok ? ((1.0f - alpha) + 1e-10f) : 1.0f;
            const float incl = warp_incl_scan_mul(x, lane);
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0f;
            const float T = carry * excl;
            const float w = ok ? alpha * T : 0.f;
            carry *= __shfl_sync(0xffffffffu, incl, 31);
            if (ok) {
                if (weights) weights[ray * N + n] = w;
                sr = fmaf(w, sigmoidf_fast(pr.x), sr);
                sg = fmaf(w, sigmoidf_fast(pr.y), sg);
                sb = fmaf(w, sigmoidf_fast(pr.z), sb);
                sd = fmaf(w, tn, sd);
                sa += w;
            }
        }
        sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sd = warp_sum(sd); sa = warp_sum(sa);
        if (lane == 0) {
            if (rgb) { rgb[ray * 3] = sr; rgb[ray * 3 + 1] = sg; rgb[ray * 3 + 2] = sb; }
            if (depth) depth[ray] = sd;
            if (acc) acc[ray] = sa;
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) { cp[c] = np_[c]; ct[c] = nt[c]; }
    }
}

extern "C" int nerf_volume_render(const float* preds, const float* t, int64_t batch, int num_samples, float* rgb,
                                  float* depth, float* weights, float* acc, void* stream) {
    NERF_CHECK_ARG(preds && t && batch >= 0 && num_samples >= 1, "bad arguments");
    NERF_CHECK_ARG(num_samples <= 512, "num_samples > 512 is not supported by the compositing kernel");
    if (batch == 0) return NERF_OK;
    const int threads = 256;
    const int grid = stream_grid(batch * 32, threads);
    const float4* p4 = reinterpret_cast<const float4*>(preds);
    cudaStream_t st = (cudaStream_t)stream;
    const int nch = (num_samples + 31) / 32;
#define NERF_VR(K) volume_render_kernel<K><<<grid, threads, 0, st>>>(p4, t, batch, num_samples, rgb, depth, weights, acc)
    if (nch <= 1) NERF_VR(1);
    else if (nch == 2) NERF_VR(2);
    else if (nch == 3) NERF_VR(3);
    else if (nch == 4) NERF_VR(4);
    else if (nch <= 6) NERF_VR(6);
    else if (nch <= 8) NERF_VR(8);
    else NERF_VR(16);
#undef NERF_VR
    NERF_LAUNCHED();
    return NERF_OK;
}

// ------------------------------------------------------------------------------------------------
// volume_render backward.  Inputs: preds, t, d_rgb (B,3) = dL/dC, optional
// d_w_extra (B,N) added to g_n = dL/dw_n (the Q5 term for the coarse net).  Output d_preds (B,N,4)
// and optionally d_delta (B,N) = dL/d(delta_n) (needed only for the Q5 path of the fine net).
// Division-free reverse affine scan: R_n = g_{n+1} a_{n+1} + x_{n+1} R_{n+1}; dL/dx_n = T_n R_n.
// ------------------------------------------------------------------------------------------------
// Same register-resident / next-ray-prefetch structure as the forward kernel; the transmittances of the forward
// sweep stay in registers for the reverse sweep.
template <int NCH>
__global__ void __launch_bounds__(256, NCH <= 2 ? 3 : (NCH <= 6 ? 2 : 1)) volume_render_bwd_kernel(const float4* __restrict__ preds,
                                                                const float* __restrict__ t,
                                                                const float* __restrict__ d_rgb,
                                                                const float* __restrict__ d_w_extra, int64_t B,
                                                                int N, float4* __restrict__ d_preds,
                                                                float* __restrict__ d_delta,
                                                                float* __restrict__ g_brgb, float* __restrict__ g_bsig) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float4 cp[NCH], np_[NCH];
    float ct[NCH], nt[NCH];
    float4 hb = make_float4(0.f, 0.f, 0.f, 0.f);     // bias gradients of the rgb / sigma heads: sums of d_preds
    auto load = [&](int64_t r, float4 (&p)[NCH], float (&tt)[NCH]) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int n = c * 32 + lane;
            const bool ok = (r < B) && (n < N);
            p[c] = ok ? __ldg(preds + r * N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
            tt[c] = ok ? __ldg(t + r * N + n) : 0.f;
        }
    };
    load(ray, cp, ct);
    for (; ray < B; ray += warps_total) {
        load(ray + warps_total, np_, nt);
        const float dr = __ldg(d_rgb + ray * 3), dg = __ldg(d_rgb + ray * 3 + 1), db = __ldg(d_rgb + ray * 3 + 2);
        // exp(-sigma delta) of the forward sweep is kept for the reverse sweep when the ray fits in few chunks; long rays
        // recompute it (six more live registers per lane cost the 192-sample kernel more occupancy than the exp saves)
        constexpr bool KEEP_E = NCH <= 3;
        float Tc[NCH], dl[NCH], ec[KEEP_E ? NCH : 1];
        // forward sweep: transmittance
        float carry = 1.0f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int n = c * 32 + lane;
            const bool ok = n < N;
            const float tn = ct[c];
            float tn1 = __shfl_down_sync(0xffffffffu, tn, 1);
            const float t_first_next = __shfl_sync(0xffffffffu, ct[(c + 1 < NCH) ? c + 1 : c], 0);
            if (lane == 31) tn1 = t_first_next;
            const float delta = (n == N - 1) ? 1e10f : (tn1 - tn);
            dl[c] = delta;
            const float sigma = fmaxf(cp[c].w, 0.0f);
            const float e_fw = expf(-sigma * delta);
            if (KEEP_E) ec[c] = e_fw;
            const float alpha = 1.0f - e_fw;
            const float x = ok ? ((1.0f - alpha) + 1e-10f) : 1.0f;
            const float incl = warp_incl_scan_mul(x, lane);
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0f;
            Tc[c] = carry * excl;
            carry *= __shfl_sync(0xffffffffu, incl, 31);
        }
        // reverse sweep: R_n = b_n + a_n R_{n+1}, with a_n = x_{n+1}, b_n = g_{n+1} alpha_{n+1}
        float Rcarry = 0.0f;
#pragma unroll
        for (int c = NCH - 1; c >= 0; --c) {
            const int n = c * 32 + lane;
            const bool ok = n < N;
            const float4 pr = cp[c];
            const float delta = dl[c];
            const float sigma = fmaxf(pr.w, 0.0f);
            const float e = KEEP_E ? ec[KEEP_E ? c : 0] : expf(-sigma * delta);
            const float alpha = 1.0f - e;
            const float x = ok ? ((1.0f - alpha) + 1e-10f) : 1.0f;
            // colours to ~1e-6 absolute (ex2.approx based), as in the forward kernel: d_preds carries them times w <= 1
            const float cr = sigmoidf_fast(pr.x), cg = sigmoidf_fast(pr.y), cb = sigmoidf_fast(pr.z);
            float g = dr * cr + dg * cg + db * cb;
            if (d_w_extra && ok) g += d_w_extra[ray * N + n];
            if (!ok) g = 0.f;
            const float galpha = ok ? g * alpha : 0.f;
            // element n contributes the affine map f_n(R) = galpha_n + x_n * R  (maps R_n -> R_{n-1});
            // suffix composition within the chunk (lanes to the right applied first)
            float a = x, b = galpha;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float a2 = __shfl_down_sync(0xffffffffu, a, o);
                const float b2 = __shfl_down_sync(0xffffffffu, b, o);
                if (lane + o < 32) {  // f(R) = b + a * (b2 + a2 R)
                    b = b + a * b2;
                    a = a * a2;
                }
            }
            const float a_ex = __shfl_down_sync(0xffffffffu, a, 1);
            const float b_ex = __shfl_down_sync(0xffffffffu, b, 1);
            const float Rn = (lane == 31) ? Rcarry : (b_ex + a_ex * Rcarry);
            const float a0 = __shfl_sync(0xffffffffu, a, 0), b0 = __shfl_sync(0xffffffffu, b, 0);
            const float T = ok ? Tc[c] : 0.f;
            const float w = alpha * T;
            const float dLdx = T * Rn;
            const float dLdalpha = g * T - dLdx;
            const float dsig = dLdalpha * delta * e;
            const float ds_raw = (pr.w > 0.0f) ? dsig : 0.0f;
            if (ok) {
                const float4 dpv =
                    make_float4(w * dr * cr * (1.0f - cr), w * dg * cg * (1.0f - cg), w * db * cb * (1.0f - cb), ds_raw);
                d_preds[ray * N + n] = dpv;
                hb.x += dpv.x; hb.y += dpv.y; hb.z += dpv.z; hb.w += dpv.w;
                if (d_delta) d_delta[ray * N + n] = (n == N - 1) ? 0.0f : dLdalpha * sigma * e;
            }
            Rcarry = b0 + a0 * Rcarry;
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) { cp[c] = np_[c]; ct[c] = nt[c]; }
    }
    if (g_brgb) {
        // one atomic per block and component (block reduce through shared memory)
        __shared__ float4 s_hb[8];
        hb.x = warp_sum(hb.x); hb.y = warp_sum(hb.y); hb.z = warp_sum(hb.z); hb.w = warp_sum(hb.w);
        if (lane == 0) s_hb[threadIdx.x >> 5] = hb;
        __syncthreads();
        if (threadIdx.x < 4) {
            float a = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += reinterpret_cast<const float*>(&s_hb[w])[threadIdx.x];
            atomicAdd(threadIdx.x < 3 ? g_brgb + threadIdx.x : g_bsig, a);
        }
    }
}

// g_brgb (3 floats) / g_bsig (1 float), optional: the bias gradients of the rgb and sigma heads are ACCUMULATED there
extern "C" int nerf_volume_render_bwd_heads(const float* preds, const float* t, const float* d_rgb, const float* d_w_extra,
                                            int64_t batch, int num_samples, float* d_preds, float* d_delta, float* g_brgb,
                                            float* g_bsig, void* stream);
extern "C" int nerf_volume_render_bwd(const float* preds, const float* t, const float* d_rgb, const float* d_w_extra,
                                      int64_t batch, int num_samples, float* d_preds, float* d_delta, void* stream) {
    return nerf_volume_render_bwd_heads(preds, t, d_rgb, d_w_extra, batch, num_samples, d_preds, d_delta, nullptr, nullptr,
                                        stream);
}
extern "C" int nerf_volume_render_bwd_heads(const float* preds, const float* t, const float* d_rgb, const float* d_w_extra,
                                            int64_t batch, int num_samples, float* d_preds, float* d_delta, float* g_brgb,
                                            float* g_bsig, void* stream) {
    NERF_CHECK_ARG(preds && t && d_rgb && d_preds && batch >= 0 && num_samples >= 1, "bad arguments");
    NERF_CHECK_ARG(num_samples <= 512, "num_samples > 512 is not supported by the compositing kernel");
    if (batch == 0) return NERF_OK;
    const int threads = 256;
    const int grid = stream_grid(batch * 32, threads);
    const float4* p4 = reinterpret_cast<const float4*>(preds);
    float4* dp4 = reinterpret_cast<float4*>(d_preds);
    cudaStream_t st = (cudaStream_t)stream;
    const int nch = (num_samples + 31) / 32;
#define NERF_VRB(K) \
    volume_render_bwd_kernel<K><<<grid, threads, 0, st>>>(p4, t, d_rgb, d_w_extra, batch, num_samples, dp4, d_delta, g_brgb, g_bsig)
    if (nch <= 1) NERF_VRB(1);
    else if (nch == 2) NERF_VRB(2);
    else if (nch == 3) NERF_VRB(3);
    else if (nch == 4) NERF_VRB(4);
    else if (nch <= 6) NERF_VRB(6);
    else if (nch <= 8) NERF_VRB(8);
    else NERF_VRB(16);
#undef NERF_VRB
    NERF_LAUNCHED();
    return NERF_OK;
}

// ------------------------------------------------------------------------------------------------
// sample_pdf (data_utils.py:172-223) and the fused resample+merge (models.py:165-167).
// One warp per ray.  cdf (nc+1) and t_mid (nc-1) live in shared memory; each lane inverts the CDF for
// its share of the u draws with a binary search (searchsorted side="right").
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_incl_scan_add(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// builds cdf[0..nc] in smem from weights; all lanes participate
__device__ __forceinline__ void build_cdf(const float* __restrict__ w, int nc, float* cdf, int lane) {
    float s = 0.f;
    for (int n = lane; n < nc; n += 32) s += (w[n] + 1e-5f);
    s = warp_sum(s);
    float carry = 0.f;
    if (lane == 0) cdf[0] = 0.f;
    for (int base = 0; base < nc; base += 32) {
        int n = base + lane;
        float pdf = (n < nc) ? __fdiv_rn(w[n] + 1e-5f, s) : 0.f;
        float incl = warp_incl_scan_add(pdf, lane);
        if (n < nc) cdf[n + 1] = carry + incl;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
}

__device__ __forceinline__ float invert_cdf(const float* cdf, const float* tmid, int nc, float u) {
    // searchsorted(cdf, u, side="right"): first index i in [0, nc+1] with cdf[i] > u
    int lo = 0, hi = nc + 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cdf[mid] > u) hi = mid; else lo = mid + 1;
    }
    int idx = lo;
    int below = max(0, idx - 1);
    int above = min(nc, idx);
    float cb = cdf[below], ca = cdf[above];
    float tb = tmid[min(nc - 2, below)], ta = tmid[min(nc - 2, above)];
    float den = ca - cb;
    if (den < 1e-5f) den = 1.0f;
    float f = __fdiv_rn(u - cb, den);
    return __fadd_rn(tb, __fmul_rn(f, ta - tb));
}

__global__ void __launch_bounds__(128) sample_pdf_kernel(const float* __restrict__ t_mid,
                                                         const float* __restrict__ weights,
                                                         const float* __restrict__ u, int64_t B, int nc, int nf,
                                                         float* __restrict__ samples) {
    extern __shared__ float smem_sp[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* cdf = smem_sp + (size_t)wib * (2 * nc + 2);
    float* tm = cdf + nc + 1;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < B; ray += warps_total) {
        build_cdf(weights + ray * nc, nc, cdf, lane);
        for (int n = lane; n < nc - 1; n += 32) tm[n] = t_mid[ray * (nc - 1) + n];
        __syncwarp();
        for (int j = lane; j < nf; j += 32) samples[ray * nf + j] = invert_cdf(cdf, tm, nc, u[ray * nf + j]);
        __syncwarp();
    }
}

extern "C" int nerf_sample_pdf(const float* t_mid, const float* weights, const float* u, int64_t batch, int nc,
                               int ns_fine, float* samples, void* stream) {
    NERF_CHECK_ARG(t_mid && weights && u && samples && batch >= 0 && nc >= 2 && ns_fine >= 1, "bad arguments");
    if (batch == 0) return NERF_OK;
    int threads = 128;
    size_t smem = (size_t)(threads / 32) * (2 * nc + 2) * sizeof(float);
    NERF_CHECK_ARG(smem <= 48 * 1024, "nc too large");
    sample_pdf_kernel<<<stream_grid(batch * 32, threads), threads, smem, (cudaStream_t)stream>>>(
        t_mid, weights, u, batch, nc, ns_fine, samples);
    NERF_LAUNCHED();
    return NERF_OK;
}

// fused: t_mid = 0.5*(t[1:]+t[:-1]); t_fine = sample_pdf(...); t_all = sort(concat([t, t_fine]))
// bitonic sort of P = next_pow2(nc+nf) (value, source index) pairs per warp in shared memory.
__global__ void __launch_bounds__(128) resample_merge_kernel(const float* __restrict__ t,
                                                             const float* __restrict__ weights,
                                                             const PdfDraws dr, int64_t B, int nc, int nf,
                                                             int P, float* __restrict__ t_all,
                                                             int32_t* __restrict__ src_idx) {
    const unsigned long long dstep = pdf_step(dr);
    extern __shared__ float smem_rm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int per_warp = (2 * nc + 2) + 2 * P;
    float* cdf = smem_rm + (size_t)wib * per_warp;
    float* tm = cdf + nc + 1;
    float* key = tm + nc + 1;
    int* val = reinterpret_cast<int*>(key + P);
    const int na = nc + nf;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < B; ray += warps_total) {
        const float* tr = t + ray * nc;
        build_cdf(weights + ray * nc, nc, cdf, lane);
        for (int n = lane; n < nc; n += 32) {
            float tn = tr[n];
            key[n] = tn;
            val[n] = n;
            if (n < nc - 1) tm[n] = __fmul_rn(0.5f, __fadd_rn(tr[n + 1], tn));
        }
        __syncwarp();
        for (int j = lane; j < nf; j += 32) {
            key[nc + j] = invert_cdf(cdf, tm, nc, pdf_draw(dr, dstep, ray, nf, j));
            val[nc + j] = nc + j;
        }
        for (int j = na + lane; j < P; j += 32) {
            key[j] = __int_as_float(0x7f800000);  // +inf padding
            val[j] = -1;
        }
        __syncwarp();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = lane; i < P; i += 32) {
                    int ixj = i ^ j;
                    if (ixj > i) {
                        bool up = ((i & k) == 0);
                        float a = key[i], b = key[ixj];
                        if ((a > b) == up) {
                            key[i] = b; key[ixj] = a;
                            int va = val[i]; val[i] = val[ixj]; val[ixj] = va;
                        }
                    }
                }
                __syncwarp();
            }
        }
        for (int j = lane; j < na; j += 32) {
            t_all[ray * na + j] = key[j];
            if (src_idx) src_idx[ray * na + j] = val[j];
        }
        __syncwarp();
    }
}

// Fast path of the fused resample + merge: the coarse samples of a ray are already sorted (generate_t_vals is
// monotone), so only the nf fine samples need sorting.  They are sorted IN REGISTERS (FI keys + source indices per lane,
// bitonic network: partner distances < 32 through warp shuffles, >= 32 between a lane's own registers) and the two
// sorted sequences are merged by rank: a fine sample lands at (its sorted rank) + #(coarse <= it), a coarse sample at
// (its index) + #(fine < it) -- ~1.1 K warp instructions per ray instead of ~10 K for the 256-key shared-memory
// bitonic sort below.  Pairs move together, so equal keys cannot break the permutation.  A ray whose coarse samples
// are NOT sorted takes the rank-counting path (all keys, ties broken by index).
template <int FI>
__global__ void __launch_bounds__(128) resample_merge_sorted_kernel(const float* __restrict__ t,
                                                                    const float* __restrict__ weights,
                                                                    const PdfDraws dr, int64_t B, int nc, int nf,
                                                                    float* __restrict__ t_all,
                                                                    int32_t* __restrict__ src_idx) {
    constexpr int P = 32 * FI;
    const unsigned long long dstep = pdf_step(dr);
    extern __shared__ float smem_rm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int per_warp = (2 * nc + 2) + nc + P;
    float* cdf = smem_rm + (size_t)wib * per_warp;
    float* tm = cdf + nc + 1;
    float* ck = tm + nc + 1;          // coarse keys
    float* fk = ck + nc;              // fine keys: raw, then sorted
    const int na = nc + nf;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < B; ray += warps_total) {
        const float* tr = t + ray * nc;
        float* out_t = t_all + ray * na;
        int32_t* out_i = src_idx ? src_idx + ray * na : nullptr;
        build_cdf(weights + ray * nc, nc, cdf, lane);
        bool ok = true;
        for (int n = lane; n < nc; n += 32) {
            const float tn = tr[n];
            ck[n] = tn;
            if (n < nc - 1) {
                const float tn1 = tr[n + 1];
                tm[n] = __fmul_rn(0.5f, __fadd_rn(tn1, tn));
                ok = ok && (tn <= tn1);
            }
        }
        const bool sorted = __all_sync(0xffffffffu, ok);
        __syncwarp();
        float key[FI];
        int idx[FI];
#pragma unroll
        for (int r = 0; r < FI; ++r) {
            const int j = r * 32 + lane;
            key[r] = (j < nf) ? invert_cdf(cdf, tm, nc, pdf_draw(dr, dstep, ray, nf, j)) : __int_as_float(0x7f800000);
            idx[r] = (j < nf) ? nc + j : -1;
        }
        if (!sorted) {
            // rank counting over all keys, ties broken by source index
#pragma unroll
            for (int r = 0; r < FI; ++r) fk[r * 32 + lane] = key[r];
            __syncwarp();
            for (int e = lane; e < na; e += 32) {
                const float x = (e < nc) ? ck[e] : fk[e - nc];
                int pos = 0;
                for (int k = 0; k < na; ++k) {
                    const float y = (k < nc) ? ck[k] : fk[k - nc];
                    pos += (y < x || (y == x && k < e)) ? 1 : 0;
                }
                out_t[pos] = x;
                if (out_i) out_i[pos] = e;
            }
            __syncwarp();
            continue;
        }
        // ---- bitonic sort of P (key, idx) pairs, element e = r * 32 + lane ----
#pragma unroll
        for (int k = 2; k <= P; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                if (j >= 32) {
#pragma unroll
                    for (int r = 0; r < FI; ++r) {
                        const int rp = r ^ (j >> 5);
                        if (rp > r) {
                            const bool up = (((r * 32) & k) == 0);          // bits >= 5 of e come from r
                            const bool sw = up ? (key[r] > key[rp]) : (key[r] <= key[rp]);
                            if (sw) {
                                const float tk = key[r]; key[r] = key[rp]; key[rp] = tk;
                                const int ti = idx[r]; idx[r] = idx[rp]; idx[rp] = ti;
                            }
                        }
                    }
                } else {
                    const bool lower = (lane & j) == 0;
#pragma unroll
                    for (int r = 0; r < FI; ++r) {
                        const float ok_ = __shfl_xor_sync(0xffffffffu, key[r], j);
                        const int oi = __shfl_xor_sync(0xffffffffu, idx[r], j);
                        const bool up = ((((r * 32) + lane) & k) == 0);
                        const bool le = lower ? (key[r] <= ok_) : (ok_ <= key[r]);   // (lower element <= upper element)
                        const bool take = up ? !le : le;                               // ascending: swap when lower > upper
                        if (take) { key[r] = ok_; idx[r] = oi; }
                    }
                }
            }
        }
        // sorted fine keys to shared memory for the coarse samples' rank search
#pragma unroll
        for (int r = 0; r < FI; ++r) fk[r * 32 + lane] = key[r];
        __syncwarp();
#pragma unroll
        for (int r = 0; r < FI; ++r) {
            const int e = r * 32 + lane;
            if (e < nf) {
                int lo = 0, hi = nc;                                    // #(coarse <= key): upper bound
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (ck[mid] <= key[r]) lo = mid + 1; else hi = mid;
                }
                out_t[e + lo] = key[r];
                if (out_i) out_i[e + lo] = idx[r];
            }
        }
        for (int n = lane; n < nc; n += 32) {
            const float c = ck[n];
            int lo = 0, hi = nf;                                        // #(fine < c): lower bound
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (fk[mid] < c) lo = mid + 1; else hi = mid;
            }
            out_t[n + lo] = c;
            if (out_i) out_i[n + lo] = n;
        }
        __syncwarp();
    }
}

namespace nerf {
int resample_merge(const float* t, const float* weights, PdfDraws dr, int64_t batch, int nc, int nf, float* t_all,
                   int32_t* src_idx, cudaStream_t st) {
    if (batch == 0) return NERF_OK;
    int threads = 128;
    if (nf <= 256 && nc <= 1024) {
        // fine samples sorted in registers, merged with the (sorted) coarse samples by rank
        const int fi = nf <= 32 ? 1 : nf <= 64 ? 2 : nf <= 128 ? 4 : 8;
        size_t smem_fast = (size_t)(threads / 32) * ((2 * nc + 2) + nc + 32 * fi) * sizeof(float);
        const int grid = stream_grid(batch * 32, threads);
        if (smem_fast <= 48 * 1024) {
            if (fi == 1) resample_merge_sorted_kernel<1><<<grid, threads, smem_fast, st>>>(t, weights, dr, batch, nc, nf, t_all, src_idx);
            else if (fi == 2) resample_merge_sorted_kernel<2><<<grid, threads, smem_fast, st>>>(t, weights, dr, batch, nc, nf, t_all, src_idx);
            else if (fi == 4) resample_merge_sorted_kernel<4><<<grid, threads, smem_fast, st>>>(t, weights, dr, batch, nc, nf, t_all, src_idx);
            else resample_merge_sorted_kernel<8><<<grid, threads, smem_fast, st>>>(t, weights, dr, batch, nc, nf, t_all, src_idx);
            NERF_LAUNCHED();
            return NERF_OK;
        }
    }
    int P = 1;
    while (P < nc + nf) P <<= 1;
    size_t smem = (size_t)(threads / 32) * ((2 * nc + 2) + 2 * P) * sizeof(float);
    NERF_CHECK_ARG(smem <= 96 * 1024, "nc+nf too large");
    if (smem > 48 * 1024)
        NERF_CUDA(cudaFuncSetAttribute(resample_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    resample_merge_kernel<<<stream_grid(batch * 32, threads), threads, smem, st>>>(t, weights, dr, batch, nc, nf, P, t_all,
                                                                                  src_idx);
    NERF_LAUNCHED();
    return NERF_OK;
}
}  // namespace nerf

// test hook: the uniforms the in-kernel generator hands to sample_pdf for (seed, counter)
__global__ void __launch_bounds__(256) pdf_draws_kernel(PdfDraws dr, int64_t B, int nf, float* __restrict__ out) {
    const unsigned long long step = pdf_step(dr);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B * nf; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = pdf_draw(dr, step, i / nf, nf, (int)(i % nf));
}
extern "C" int nerf_debug_pdf_draws(uint64_t seed, uint64_t counter, int64_t batch, int nf, float* out, void* stream) {
    NERF_CHECK_ARG(out && batch >= 1 && nf >= 1, "bad arguments");
    pdf_draws_kernel<<<stream_grid(batch * nf, 256), 256, 0, (cudaStream_t)stream>>>(PdfDraws{nullptr, seed, counter, nullptr},
                                                                                   batch, nf, out);
    NERF_LAUNCHED();
    return NERF_OK;
}

extern "C" int nerf_resample_merge(const float* t, const float* weights, const float* u, int64_t batch, int nc, int nf,
                                   float* t_all, int32_t* src_idx, void* stream) {
    NERF_CHECK_ARG(t && weights && u && t_all && batch >= 0 && nc >= 2 && nf >= 1, "bad arguments");
    return nerf::resample_merge(t, weights, PdfDraws{u, 0ull, 0ull, nullptr}, batch, nc, nf, t_all, src_idx,
                                (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// metrics (models.py:98-102,110): mse_c, mse_f over B*3 elements, psnr = -10 log10(mse_f).
// Deterministic two-stage reduce is unnecessary at B*3 <= a few 100k: one 1024-thread block.
// Also emits d_rgb_c / d_rgb_f = 2 (rgb - img) / (3B) when requested (train step).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) metrics_kernel(const float* __restrict__ img, const float* __restrict__ rc,
                                                       const float* __restrict__ rf, int64_t n_el,
                                                       float* __restrict__ metrics, float* __restrict__ d_rc,
                                                       float* __restrict__ d_rf, float* __restrict__ sums) {
    __shared__ float sc[32], sf[32];
    float ac = 0.f, af = 0.f;
    const float scale = 2.0f / (float)n_el;
    for (int64_t i = threadIdx.x; i < n_el; i += blockDim.x) {
        float im = img[i];
        float a = rc[i] - im, b = rf[i] - im;
        ac += a * a;
        af += b * b;
        if (d_rc) d_rc[i] = a * scale;
        if (d_rf) d_rf[i] = b * scale;
    }
    ac = warp_sum(ac); af = warp_sum(af);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sc[w] = ac; sf[w] = af; }
    __syncthreads();
    if (w == 0) {
        ac = sc[lane]; af = sf[lane];
        ac = warp_sum(ac); af = warp_sum(af);
        if (lane == 0) {
            float mc = ac / (float)n_el, mf = af / (float)n_el;
            const float ps = -10.0f * log10f(mf);
            metrics[0] = mc;
            metrics[1] = mf;
            metrics[2] = ps;
            if (sums) { sums[0] += mc; sums[1] += mf; sums[2] += ps; sums[3] += 1.0f; }   // keras.metrics.Mean (models.py:113-115)
        }
    }
}

// rgb_f == rgb_c is the single-net shape (NS_FINE = 0): loss and psnr then describe the only net there is
extern "C" int nerf_metrics_grad_sums(const float* images, const float* rgb_c, const float* rgb_f, int64_t batch,
                                      float* metrics_dev, float* d_rgb_c, float* d_rgb_f, float* sums, void* stream) {
    NERF_CHECK_ARG(images && rgb_c && rgb_f && metrics_dev && batch >= 1, "bad arguments");
    metrics_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(images, rgb_c, rgb_f, batch * 3, metrics_dev, d_rgb_c,
                                                         d_rgb_f, sums);
    NERF_LAUNCHED();
    return NERF_OK;
}
extern "C" int nerf_metrics_grad(const float* images, const float* rgb_c, const float* rgb_f, int64_t batch,
                                 float* metrics_dev, float* d_rgb_c, float* d_rgb_f, void* stream) {
    return nerf_metrics_grad_sums(images, rgb_c, rgb_f, batch, metrics_dev, d_rgb_c, d_rgb_f, nullptr, stream);
}

extern "C" int nerf_metrics(const float* images, const float* rgb_c, const float* rgb_f, int64_t batch,
                            float* metrics_dev, void* stream) {
    return nerf_metrics_grad(images, rgb_c, rgb_f, batch, metrics_dev, nullptr, nullptr, stream);
}

// ------------------------------------------------------------------------------------------------
// Keras Adam (train_lego.py:149-151): m += (g-m)(1-b1); v += (g^2-v)(1-b2); p -= a_t m/(sqrt(v)+eps)
// one flat multi-tensor launch: 16 B read + 12 B written per parameter.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                   float alpha_t, float one_minus_b1, float one_minus_b2, float eps,
                                                   float grad_scale, const nerf_dev_state* __restrict__ ds) {
    if (ds) {
        // step count and learning rate live in device memory (CUDA-graph replays): alpha_t of update t = step + 1
        __shared__ float s_alpha;
        if (threadIdx.x == 0) {
            const double tt = (double)(ds->step + 1ull);
            s_alpha = (float)((double)ds->lr * sqrt(1.0 - pow(0.999, tt)) / (1.0 - pow(0.9, tt)));
        }
        __syncthreads();
        alpha_t = s_alpha;
    }
    const int64_t n_vec = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float gk = ga[k] * grad_scale;
            ma[k] = ma[k] + (gk - ma[k]) * one_minus_b1;
            va[k] = va[k] + (gk * gk - va[k]) * one_minus_b2;
            pa[k] = pa[k] - alpha_t * ma[k] / (sqrtf(va[k]) + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        int64_t i = (n_vec << 2) + threadIdx.x;
        float gk = g[i] * grad_scale;
        float mk = m[i] + (gk - m[i]) * one_minus_b1;
        float vk = v[i] + (gk * gk - v[i]) * one_minus_b2;
        m[i] = mk; v[i] = vk;
        p[i] = p[i] - alpha_t * mk / (sqrtf(vk) + eps);
    }
}

extern "C" int nerf_adam_flat(float* params, const float* grads, float* m, float* v, int64_t n, int64_t step,
                              float lr, float grad_scale, void* stream) {
    NERF_CHECK_ARG(params && grads && m && v && n >= 1 && step >= 1, "bad arguments");
    const double b1 = 0.9, b2 = 0.999;
    double alpha = (double)lr * sqrt(1.0 - pow(b2, (double)step)) / (1.0 - pow(b1, (double)step));
    adam_kernel<<<stream_grid(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(
        params, grads, m, v, n, (float)alpha, (float)(1.0 - b1), (float)(1.0 - b2), 1e-7f, grad_scale, nullptr);
    NERF_LAUNCHED();
    return NERF_OK;
}

namespace nerf {
// the same update with t and the learning rate read from the context's device state (graph-replayable)
int adam_from_dev_state(float* params, const float* grads, float* m, float* v, int64_t n, const nerf_dev_state* ds,
                        float grad_scale, cudaStream_t st) {
    adam_kernel<<<stream_grid(n / 4 + 1, 256), 256, 0, st>>>(params, grads, m, v, n, 0.f, (float)(1.0 - 0.9),
                                                             (float)(1.0 - 0.999), 1e-7f, grad_scale, ds);
    NERF_LAUNCHED();
    return NERF_OK;
}
bool stream_is_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
    return cs != cudaStreamCaptureStatusNone;
}
}  // namespace nerf

// ------------------------------------------------------------------------------------------------
// kernel timing hooks
// ------------------------------------------------------------------------------------------------
#include <mutex>
#include <vector>
namespace nerf {
static bool g_timing_on = false;
static std::mutex g_timing_mu;
struct TimedSpan { int kind; cudaEvent_t a, b; };
static std::vector<TimedSpan> g_spans;
static std::vector<cudaEvent_t> g_open[8];
void timing_begin(int kind, cudaStream_t st) {
    if (!g_timing_on || stream_is_capturing(st)) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_open[kind & 7].push_back(e);
}
void timing_end(int kind, cudaStream_t st) {
    if (!g_timing_on || stream_is_capturing(st)) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    std::lock_guard<std::mutex> lk(g_timing_mu);
    auto& open = g_open[kind & 7];
    if (open.empty()) { cudaEventDestroy(e); return; }
    g_spans.push_back({kind, open.back(), e});
    open.pop_back();
}
}  // namespace nerf

extern "C" int nerf_timing_enable(int on) {
    std::lock_guard<std::mutex> lk(nerf::g_timing_mu);
    nerf::g_timing_on = on != 0;
    return NERF_OK;
}
// sums (and clears) the recorded spans of `kind`: total milliseconds and number of launches.
// Synchronises on the recorded events.
extern "C" int nerf_timing_read(int kind, double* total_ms, int64_t* launches) {
    NERF_CHECK_ARG(total_ms && launches, "null pointer");
    std::lock_guard<std::mutex> lk(nerf::g_timing_mu);
    double tot = 0;
    int64_t n = 0;
    std::vector<nerf::TimedSpan> keep;
    for (auto& s : nerf::g_spans) {
        if (s.kind != kind) { keep.push_back(s); continue; }
        cudaEventSynchronize(s.b);
        float ms = 0;
        cudaEventElapsedTime(&ms, s.a, s.b);
        tot += ms; ++n;
        cudaEventDestroy(s.a); cudaEventDestroy(s.b);
    }
    nerf::g_spans.swap(keep);
    *total_ms = tot; *launches = n;
    return NERF_OK;
}

extern "C" const char* nerf_last_error(void) { return nerf::g_last_error.c_str(); }
extern "C" int nerf_version(void) { return 100; }
extern "C" int64_t nerf_launch_count(void) { return nerf::g_launches.load(); }
