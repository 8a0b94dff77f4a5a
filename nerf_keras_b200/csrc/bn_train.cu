// BATCH_NORM=true training (models.py:30-33, 49-52 with training=True): BatchNormalization after every trunk Dense and
// after the direction Dense uses the statistics of the CURRENT batch over all (ray, sample) rows, which couples every
// sample of a batch between consecutive layers.  That does not fit the fused per-tile tcgen05 kernels, so this path is
// layer by layer on fp32 activations: tcgen05 GEMMs on split bf16 operands (gemm_tc.cu: fp32-grade results, no library
// GEMM) between hand-written statistics / normalise / backward kernels, with the sampling, compositing, resampling and
// metric kernels of the main path around it.
// It serves the reference's BN configs, which are small-batch (256-512 rays, 48-192 samples per ray); it is not the
// benchmark path.  Inference with BN uses the fused kernels (the moving statistics fold into W, b on the host).
//
//   forward per BN layer   z = x W (+ b, which batch-norm cancels);  mu, var = batch moments (biased);
//                          zhat = (z - mu) rsqrt(var + eps);  a = relu(gamma zhat + beta)
//   moving statistics      mean <- 0.99 mean + 0.01 (mu + b),  var <- 0.99 var + 0.01 var_batch        (Keras defaults)
//   backward               dy = da * 1[a > 0];  dgamma = sum dy zhat;  dbeta = sum dy;
//                          dz = gamma rstd (dy - dbeta / M - zhat dgamma / M);  dW = x^T dz;  dx = dz W^T;  db = 0
// Gradient semantics: stop-gradient on the fine sample positions (the original NeRF's); the reference's un-stopped term
//  is not carried through this path.
#include <vector>

#include "common.cuh"

namespace nerf {
// gemm_tc.cu: tcgen05 MMAs on split bf16 operands, fp32 accumulation in tensor memory; skinny head shapes on CUDA cores
int tc_gemm_f32(cudaStream_t st, bool ta, bool tb, int64_t M, int N, int K, const float* A, int64_t lda, const float* B,
                int64_t ldb, float beta, float* C, int64_t ldc, bool precise);
}  // namespace nerf
extern "C" int nerf_metrics_grad(const float*, const float*, const float*, int64_t, float*, float*, float*, void*);
extern "C" int nerf_volume_render_bwd(const float* preds, const float* t, const float* d_rgb, const float* d_w_extra, int64_t batch,
                                      int num_samples, float* d_preds, float* d_delta, void* stream);

namespace {
using namespace nerf;

constexpr float BN_EPS = 1e-3f, BN_MOMENTUM = 0.99f;

struct Lay { int fi, fo; int64_t w, b; };

struct Arch {
    int L, H, skip, ex, ed, nbn;
    std::vector<Lay> lay;      // d0..d(L-1), sigma, feature, ddir, rgb
    int64_t n_params;
};

Arch make_arch(const nerf_config& c) {
    Arch a;
    a.L = c.num_layers; a.H = c.hidden_dim; a.skip = c.skip_layer;
    a.ex = 3 + 6 * c.l_xyz; a.ed = 3 + 6 * c.l_dir;
    a.nbn = a.L * a.H + a.H / 2;
    int64_t off = 0;
    auto add = [&](int fi, int fo) { a.lay.push_back(Lay{fi, fo, off, off + (int64_t)fi * fo}); off += (int64_t)fi * fo + fo; };
    int fan_in = a.ex;
    for (int i = 0; i < a.L; ++i) {
        add(fan_in, a.H);
        fan_in = a.H;
        if (i % a.skip == 0 && i > 0) fan_in = a.H + a.ex;
    }
    add(fan_in, 1); add(fan_in, a.H); add(a.H + a.ed, a.H / 2); add(a.H / 2, 3);
    a.n_params = off;
    return a;
}

// row-major C (M x N) = op(A) op(B) + beta C on the tensor cores (gemm_tc.cu); no library GEMM is involved
// forward products (precise = true) use the three-way operand split: their results decide ReLU signs
int gemm(cudaStream_t st, bool ta, bool tb, int64_t M, int N, int K, const float* A, int lda, const float* B, int ldb,
         float beta, float* C, int ldc, bool precise = false) {
    return nerf::tc_gemm_f32(st, ta, tb, M, N, K, A, lda, B, ldb, beta, C, ldc, precise);
}
#define GEMM(...) do { int _rc = gemm(__VA_ARGS__); if (_rc) return _rc; } while (0)

// ---- column reductions over M rows (double accumulation), one block per 32 columns ----------------------------
// mode 0: s0 = sum x            s1 = sum x^2
// mode 1: s0 = sum dy           s1 = sum dy * zhat     with dy = x * (a > 0)
__global__ void __launch_bounds__(256) col_reduce_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ a,
                                                         const float* __restrict__ zhat, int64_t M, int C, int mode,
                                                         double* __restrict__ s0, double* __restrict__ s1) {
    __shared__ double sh0[8][32], sh1[8][32];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
    const int64_t rows_per = (M + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = blockIdx.y * rows_per, r1 = (r0 + rows_per < M) ? r0 + rows_per : M;
    double a0 = 0.0, a1 = 0.0;
    if (col < C) {
        // four rows per iteration, all loads issued before the first use (one dependent HBM round trip per row made this
        // kernel 7x slower than its bytes)
        for (int64_t r = r0 + ry; r < r1; r += 32) {
            float v[4], av[4], zv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t rr = r + 8 * q;
                const bool ok = rr < r1;
                v[q] = ok ? x[rr * ldx + col] : 0.f;
                av[q] = (ok && mode != 0) ? a[rr * C + col] : 1.f;
                zv[q] = (ok && mode != 0) ? zhat[rr * C + col] : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (mode == 0) { a0 += v[q]; a1 += (double)v[q] * v[q]; }
                else {
                    const float dy = (av[q] > 0.f) ? v[q] : 0.f;
                    a0 += dy; a1 += (double)dy * zv[q];
                }
            }
        }
    }
    sh0[ry][threadIdx.x & 31] = a0; sh1[ry][threadIdx.x & 31] = a1;
    __syncthreads();
    if (ry == 0 && col < C) {
        for (int k = 1; k < 8; ++k) { a0 += sh0[k][threadIdx.x]; a1 += sh1[k][threadIdx.x]; }
        atomicAdd(s0 + col, a0);
        if (s1) atomicAdd(s1 + col, a1);
    }
}

// batch moments from the column sums; moving statistics (the true pre-activation is z + b)
__global__ void bn_finalize_kernel(const double* __restrict__ s0, const double* __restrict__ s1, int64_t M, int C,
                                   const float* __restrict__ bias, float* __restrict__ mu, float* __restrict__ rstd,
                                   float* __restrict__ mov_mean, float* __restrict__ mov_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = s0[c] / (double)M;
    double v = s1[c] / (double)M - m * m;
    if (v < 0.0) v = 0.0;
    mu[c] = (float)m;
    rstd[c] = rsqrtf((float)v + BN_EPS);
    if (mov_mean) {
        mov_mean[c] = BN_MOMENTUM * mov_mean[c] + (1.0f - BN_MOMENTUM) * ((float)m + bias[c]);
        mov_var[c] = BN_MOMENTUM * mov_var[c] + (1.0f - BN_MOMENTUM) * (float)v;
    }
}

// z -> zhat (in place), a = relu(gamma zhat + beta)
__global__ void __launch_bounds__(256) bn_apply_kernel(float* __restrict__ z, float* __restrict__ a, int64_t n, int C,
                                                       const float* __restrict__ mu, const float* __restrict__ rstd,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const float zh = (z[i] - mu[c]) * rstd[c];
        z[i] = zh;
        a[i] = fmaxf(fmaf(gamma[c], zh, beta[c]), 0.f);
    }
}

// dz = gamma rstd (dy - dbeta / M - zhat dgamma / M), dy = da * (a > 0); also emits dgamma / dbeta as floats
__global__ void __launch_bounds__(256) bn_bwd_kernel(const float* __restrict__ da, const float* __restrict__ a,
                                                     const float* __restrict__ zhat, int64_t n, int C, int64_t M,
                                                     const double* __restrict__ s_dy, const double* __restrict__ s_dyz,
                                                     const float* __restrict__ gamma, const float* __restrict__ rstd,
                                                     float* __restrict__ dz) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    if ((C & 3) == 0 && stride % C == 0 && (n & 3) == 0) {
        // four consecutive columns per thread, the SAME four in every iteration (the grid stride is a multiple of C): the
        // per-column constants (two double divisions each) leave the loop
        const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
        const int c0 = (int)(i0 % C);
        float gr[4], mdy[4], mdz[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            gr[q] = gamma[c0 + q] * rstd[c0 + q];
            mdy[q] = (float)(s_dy[c0 + q] / (double)M);
            mdz[q] = (float)(s_dyz[c0 + q] / (double)M);
        }
        for (int64_t i = i0; i < n; i += stride) {
            const float4 d4 = *reinterpret_cast<const float4*>(da + i), a4 = *reinterpret_cast<const float4*>(a + i),
                         z4 = *reinterpret_cast<const float4*>(zhat + i);
            float4 o;
            o.x = gr[0] * (((a4.x > 0.f) ? d4.x : 0.f) - mdy[0] - z4.x * mdz[0]);
            o.y = gr[1] * (((a4.y > 0.f) ? d4.y : 0.f) - mdy[1] - z4.y * mdz[1]);
            o.z = gr[2] * (((a4.z > 0.f) ? d4.z : 0.f) - mdy[2] - z4.z * mdz[2]);
            o.w = gr[3] * (((a4.w > 0.f) ? d4.w : 0.f) - mdy[3] - z4.w * mdz[3]);
            *reinterpret_cast<float4*>(dz + i) = o;
        }
        return;
    }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const float dy = (a[i] > 0.f) ? da[i] : 0.f;
        const float mdy = (float)(s_dy[c] / (double)M), mdz = (float)(s_dyz[c] / (double)M);
        dz[i] = gamma[c] * rstd[c] * (dy - mdy - zhat[i] * mdz);
    }
}

__global__ void store_sums_kernel(const double* __restrict__ s, int C, float* __restrict__ dst) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) dst[c] += (float)s[c];
}

// preds (M,4) += [b_rgb, b_sigma]
__global__ void head_bias_add_kernel(float4* __restrict__ preds, int64_t M, const float* __restrict__ brgb, const float* __restrict__ bsig) {
    const float r = brgb[0], g = brgb[1], b = brgb[2], s = bsig[0];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = preds[i];
        preds[i] = make_float4(p.x + r, p.y + g, p.z + b, p.w + s);
    }
}

__global__ void bias_rows_kernel(float* __restrict__ x, int64_t n, int C, const float* __restrict__ bias) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] += bias[i % C];
}

struct Bump {
    uint8_t* p; size_t left;
    float* f(int64_t n) { return reinterpret_cast<float*>(take((size_t)n * 4)); }
    double* d(int64_t n) { return reinterpret_cast<double*>(take((size_t)n * 8)); }
    uint8_t* take(size_t bytes) {
        bytes = (bytes + 255) & ~size_t(255);
        if (bytes > left) return nullptr;
        uint8_t* r = p; p += bytes; left -= bytes; return r;
    }
};

struct NetBuf {                 // activations of one net's forward, kept for its backward
    int64_t M;
    float *pts, *dirs, *enc, *dirc;
    std::vector<float*> a, zh;  // per trunk layer: post-ReLU output, normalised pre-activation
    float *feat, *zh_d, *a_d, *preds;
    std::vector<float*> mu, rstd;   // per BN layer (L trunk + ddir)
};

size_t net_floats(const Arch& A, int64_t M) {
    return (size_t)M * (3 + 3 + A.ex + A.ed + 2 * (size_t)A.L * A.H + A.H + 2 * (A.H / 2) + 4) + 2 * (size_t)A.nbn + 64 * 40;
}

int alloc_net(const Arch& A, int64_t M, Bump& ws, NetBuf& nb) {
    nb.M = M;
    nb.pts = ws.f(M * 3); nb.dirs = ws.f(M * 3); nb.enc = ws.f(M * A.ex); nb.dirc = ws.f(M * A.ed);
    for (int l = 0; l < A.L; ++l) { nb.a.push_back(ws.f(M * A.H)); nb.zh.push_back(ws.f(M * A.H)); }
    nb.feat = ws.f(M * A.H); nb.zh_d = ws.f(M * (A.H / 2)); nb.a_d = ws.f(M * (A.H / 2)); nb.preds = ws.f(M * 4);
    for (int l = 0; l <= A.L; ++l) { nb.mu.push_back(ws.f(A.H)); nb.rstd.push_back(ws.f(A.H)); }
    if (!nb.preds || !nb.rstd.back()) return fail(NERF_ERR_INVALID, "nerf_bn_forward_backward: workspace too small");
    return NERF_OK;
}

inline int egrid(int64_t n) { int64_t g = (n + 255) / 256; int64_t cap = 8 * (int64_t)num_sms(); return (int)(g < cap ? (g < 1 ? 1 : g) : cap); }

int batch_norm_fwd(cudaStream_t st, float* z, float* a, int64_t M, int C, const float* bias, const float* gamma, const float* beta,
                   float* mov_mean, float* mov_var, float* mu, float* rstd, double* sums) {
    NERF_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)C * 8, st));
    dim3 grid((C + 31) / 32, (unsigned)((M + 511) / 512 < 2048 ? (M + 511) / 512 : 2048));     // 512 rows per block
    col_reduce_kernel<<<grid, 256, 0, st>>>(z, C, nullptr, nullptr, M, C, 0, sums, sums + C);
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, sums + C, M, C, bias, mu, rstd, mov_mean, mov_var);
    bn_apply_kernel<<<egrid(M * C), 256, 0, st>>>(z, a, M * C, C, mu, rstd, gamma, beta);
    NERF_LAUNCHED();
    return NERF_OK;
}

// da (M,C) -> dz (M,C) through relu + batch norm; accumulates dgamma / dbeta
int batch_norm_bwd(cudaStream_t st, const float* da, const float* a, const float* zhat, int64_t M, int C, const float* gamma,
                   const float* rstd, float* dz, float* dgamma, float* dbeta, double* sums) {
    NERF_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)C * 8, st));
    dim3 grid((C + 31) / 32, (unsigned)((M + 511) / 512 < 2048 ? (M + 511) / 512 : 2048));     // 512 rows per block
    col_reduce_kernel<<<grid, 256, 0, st>>>(da, C, a, zhat, M, C, 1, sums, sums + C);
    bn_bwd_kernel<<<egrid(M * C), 256, 0, st>>>(da, a, zhat, M * C, C, M, sums, sums + C, gamma, rstd, dz);
    store_sums_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, C, dbeta);
    store_sums_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums + C, C, dgamma);
    NERF_LAUNCHED();
    return NERF_OK;
}

int col_sums(cudaStream_t st, const float* x, int ldx, int64_t M, int C, float* dst, double* sums) {
    NERF_CUDA(cudaMemsetAsync(sums, 0, (size_t)C * 8, st));
    dim3 grid((C + 31) / 32, (unsigned)((M + 511) / 512 < 2048 ? (M + 511) / 512 : 2048));     // 512 rows per block
    col_reduce_kernel<<<grid, 256, 0, st>>>(x, ldx, nullptr, nullptr, M, C, 0, sums, nullptr);
    store_sums_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, C, dst);
    NERF_LAUNCHED();
    return NERF_OK;
}

// one net, forward with batch statistics.  bn = [gamma | beta | mean | var], each nbn floats; mean / var updated in place.
int net_forward(cudaStream_t st, const nerf_config& cfg, const Arch& A, const float* P, float* bn, const float* o, const float* d,
                const float* t, int64_t B, int N, NetBuf& nb, double* sums) {
    const int64_t M = nb.M;
    const int H = A.H, L = A.L;
    int rc;
    if ((rc = nerf_sample_rays(o, d, t, B, N, nb.pts, nb.dirs, st))) return rc;
    if ((rc = nerf_encode_position(nb.pts, M, cfg.l_xyz, nb.enc, st))) return rc;
    if ((rc = nerf_encode_position(nb.dirs, M, cfg.l_dir, nb.dirc, st))) return rc;
    float *gamma = bn, *beta = bn + A.nbn, *mmean = bn + 2 * A.nbn, *mvar = bn + 3 * A.nbn;
    const float* x = nb.enc;
    int xk = A.ex;
    for (int l = 0; l < L; ++l) {
        const Lay& ly = A.lay[l];
        float* z = nb.zh[l];
        GEMM(st, false, false, M, H, xk, x, xk, P + ly.w, H, 0.f, z, H, true);
        if (ly.fi > xk)   // skip connection: [h, enc] @ W = h @ W[:H] + enc @ W[H:]
            GEMM(st, false, false, M, H, A.ex, nb.enc, A.ex, P + ly.w + (int64_t)xk * H, H, 1.f, z, H, true);
        if ((rc = batch_norm_fwd(st, z, nb.a[l], M, H, P + ly.b, gamma + l * H, beta + l * H, mmean + l * H, mvar + l * H,
                                 nb.mu[l], nb.rstd[l], sums)))
            return rc;
        x = nb.a[l]; xk = H;
    }
    const Lay &ls = A.lay[L], &lf = A.lay[L + 1], &ld = A.lay[L + 2], &lr = A.lay[L + 3];
    // with the skip after the LAST trunk layer excluded by construction (i < L), the head input is h_L only when
    // (L-1) % skip != 0; the reference's 8 / 4 and every shipped config satisfy that
    if (ls.fi != H) return fail(NERF_ERR_INVALID, "BN training: a skip connection into the heads is not supported");
    GEMM(st, false, false, M, 1, H, x, H, P + ls.w, 1, 0.f, nb.preds + 3, 4);                 // sigma -> preds[:, 3]
    GEMM(st, false, false, M, H, H, x, H, P + lf.w, H, 0.f, nb.feat, H, true);
    bias_rows_kernel<<<egrid(M * H), 256, 0, st>>>(nb.feat, M * H, H, P + lf.b);
    GEMM(st, false, false, M, H / 2, H, nb.feat, H, P + ld.w, H / 2, 0.f, nb.zh_d, H / 2, true);
    GEMM(st, false, false, M, H / 2, A.ed, nb.dirc, A.ed, P + ld.w + (int64_t)H * (H / 2), H / 2, 1.f, nb.zh_d, H / 2, true);
    if ((rc = batch_norm_fwd(st, nb.zh_d, nb.a_d, M, H / 2, P + ld.b, gamma + L * H, beta + L * H, mmean + L * H, mvar + L * H,
                             nb.mu[L], nb.rstd[L], sums)))
        return rc;
    GEMM(st, false, false, M, 3, H / 2, nb.a_d, H / 2, P + lr.w, 3, 0.f, nb.preds, 4);        // rgb -> preds[:, 0:3]
    head_bias_add_kernel<<<egrid(M), 256, 0, st>>>(reinterpret_cast<float4*>(nb.preds), M, P + lr.b, P + ls.b);
    NERF_LAUNCHED();
    return NERF_OK;
}

// one net, backward from d_preds (M,4).  G: this net's gradient blob (zeroed by the caller, accumulated into),
// bng = [dgamma | dbeta].  dA / dZ / dF: (M, H) temporaries.
int net_backward(cudaStream_t st, const Arch& A, const float* P, const float* bn, const float* d_preds, const NetBuf& nb, float* G,
                 float* bng, float* dA, float* dZ, float* dF, double* sums) {
    const int64_t M = nb.M;
    const int H = A.H, L = A.L, Hh = A.H / 2;
    const int Mi = (int)M;
    const float* gamma = bn;
    float *dgamma = bng, *dbeta = bng + A.nbn;
    const Lay &ls = A.lay[L], &lf = A.lay[L + 1], &ld = A.lay[L + 2], &lr = A.lay[L + 3];
    const float* a_last = nb.a[L - 1];
    const float* d_sig = d_preds + 3;                                                         // (M,1) view, ld 4
    int rc;
    // rgb head:  db += sum d_rgb;  dW += a_d^T d_rgb;  da_d = d_rgb W^T
    if ((rc = col_sums(st, d_preds, 4, M, 3, G + lr.b, sums))) return rc;
    GEMM(st, true, false, Hh, 3, Mi, nb.a_d, Hh, d_preds, 4, 1.f, G + lr.w, 3);
    GEMM(st, false, true, M, Hh, 3, d_preds, 4, P + lr.w, 3, 0.f, dA, Hh);
    // direction layer: relu + batch norm backward, dW (feature rows, direction rows), d feature
    if ((rc = batch_norm_bwd(st, dA, nb.a_d, nb.zh_d, M, Hh, gamma + L * H, nb.rstd[L], dZ, dgamma + L * H, dbeta + L * H, sums)))
        return rc;
    GEMM(st, true, false, H, Hh, Mi, nb.feat, H, dZ, Hh, 1.f, G + ld.w, Hh);
    GEMM(st, true, false, A.ed, Hh, Mi, nb.dirc, A.ed, dZ, Hh, 1.f, G + ld.w + (int64_t)H * Hh, Hh);
    GEMM(st, false, true, M, H, Hh, dZ, Hh, P + ld.w, Hh, 0.f, dF, H);
    // feature layer (linear) and sigma head share the trunk output
    if ((rc = col_sums(st, dF, H, M, H, G + lf.b, sums))) return rc;
    GEMM(st, true, false, H, H, Mi, a_last, H, dF, H, 1.f, G + lf.w, H);
    if ((rc = col_sums(st, d_sig, 4, M, 1, G + ls.b, sums))) return rc;
    GEMM(st, true, false, H, 1, Mi, a_last, H, d_sig, 4, 1.f, G + ls.w, 1);
    GEMM(st, false, true, M, H, H, dF, H, P + lf.w, H, 0.f, dA, H);
    GEMM(st, false, true, M, H, 1, d_sig, 4, P + ls.w, 1, 1.f, dA, H);                        // += d_sigma (x) W_sigma
    // trunk, last layer first
    for (int l = L - 1; l >= 0; --l) {
        const Lay& ly = A.lay[l];
        if ((rc = batch_norm_bwd(st, dA, nb.a[l], nb.zh[l], M, H, gamma + l * H, nb.rstd[l], dZ, dgamma + l * H, dbeta + l * H, sums)))
            return rc;
        const float* x = (l == 0) ? nb.enc : nb.a[l - 1];
        const int xk = (l == 0) ? A.ex : H;
        GEMM(st, true, false, xk, H, Mi, x, xk, dZ, H, 1.f, G + ly.w, H);
        if (ly.fi > xk) GEMM(st, true, false, A.ex, H, Mi, nb.enc, A.ex, dZ, H, 1.f, G + ly.w + (int64_t)xk * H, H);
        if (l > 0) GEMM(st, false, true, M, H, H, dZ, H, P + ly.w, H, 0.f, dA, H);
    }
    return NERF_OK;
}

}  // namespace

extern "C" int64_t nerf_bn_param_count(const nerf_config* cfg) {
    if (!cfg) return -1;
    return (int64_t)cfg->num_layers * cfg->hidden_dim + cfg->hidden_dim / 2;
}

extern "C" int64_t nerf_bn_workspace_bytes(const nerf_config* cfg, int64_t batch) {
    if (!cfg || batch < 1) return -1;
    const Arch A = make_arch(*cfg);
    const int64_t Mc = batch * cfg->ns_coarse, Mf = batch * (int64_t)(cfg->ns_coarse + cfg->ns_fine);
    size_t fl = net_floats(A, Mc) + net_floats(A, Mf) + 3 * (size_t)Mf * A.H + 4 * (size_t)(Mc + Mf) + 2 * (size_t)Mf + 32 * (size_t)batch;
    return (int64_t)(fl * 4 + 4 * (size_t)A.H * 8 + (1 << 20));
}

// One training forward + backward with BATCH_NORM=true.
//   params   [coarse | fine] Dense kernels and biases, nerf_param_count() floats each (un-folded)
//   bn       [coarse | fine] x [gamma | beta | moving mean | moving var], nerf_bn_param_count() floats each; the moving
//            statistics are updated in place
//   grads    [coarse | fine] gradient of mean(loss_coarse + loss_fine) wrt params (overwritten)
//   bn_grads [coarse | fine] x [dgamma | dbeta] (overwritten)
//   metrics  [mse_coarse, mse_fine, psnr_fine]
extern "C" int nerf_bn_forward_backward(const nerf_config* cfg, const float* params, float* bn, const float* images, const float* o,
                                        const float* d, const float* t, const float* u_pdf, int64_t batch, float* grads,
                                        float* bn_grads, float* metrics, void* workspace, int64_t workspace_bytes, void* stream) {
    NERF_CHECK_ARG(cfg && params && bn && images && o && d && t && u_pdf && grads && bn_grads && metrics && workspace && batch >= 1,
                   "bad arguments");
    NERF_CHECK_ARG(workspace_bytes >= nerf_bn_workspace_bytes(cfg, batch), "workspace smaller than nerf_bn_workspace_bytes()");
    cudaStream_t st = (cudaStream_t)stream;
    const Arch A = make_arch(*cfg);
    const int Nc = cfg->ns_coarse, Nf = cfg->ns_fine, Na = Nc + Nf;
    const int64_t B = batch, Mc = B * Nc, Mf = B * (int64_t)Na;
    Bump ws{reinterpret_cast<uint8_t*>(workspace), (size_t)workspace_bytes};
    NetBuf nc_, nf_;
    int rc;
    if ((rc = alloc_net(A, Mc, ws, nc_))) return rc;
    if ((rc = alloc_net(A, Mf, ws, nf_))) return rc;
    float *dA = ws.f(Mf * A.H), *dZ = ws.f(Mf * A.H), *dF = ws.f(Mf * A.H);
    float *dpc = ws.f(Mc * 4), *dpf = ws.f(Mf * 4);
    float *w_c = ws.f(Mc), *w_f = ws.f(Mf), *t_all = ws.f(Mf);
    float *rgb_c = ws.f(B * 3), *rgb_f = ws.f(B * 3), *dep_c = ws.f(B), *dep_f = ws.f(B), *drgb_c = ws.f(B * 3), *drgb_f = ws.f(B * 3);
    double* sums = ws.d(4 * A.H);
    if (!sums) return fail(NERF_ERR_INVALID, "nerf_bn_forward_backward: workspace too small");
    const float *Pc = params, *Pf = params + A.n_params;
    float *bn_c = bn, *bn_f = bn + 4 * A.nbn;

    if ((rc = net_forward(st, *cfg, A, Pc, bn_c, o, d, t, B, Nc, nc_, sums))) return rc;
    if ((rc = nerf_volume_render(nc_.preds, t, B, Nc, rgb_c, dep_c, w_c, nullptr, st))) return rc;
    if ((rc = nerf_resample_merge(t, w_c, u_pdf, B, Nc, Nf, t_all, nullptr, st))) return rc;
    if ((rc = net_forward(st, *cfg, A, Pf, bn_f, o, d, t_all, B, Na, nf_, sums))) return rc;
    if ((rc = nerf_volume_render(nf_.preds, t_all, B, Na, rgb_f, dep_f, w_f, nullptr, st))) return rc;
    if ((rc = nerf_metrics_grad(images, rgb_c, rgb_f, B, metrics, drgb_c, drgb_f, st))) return rc;

    NERF_CUDA(cudaMemsetAsync(grads, 0, 2 * (size_t)A.n_params * 4, st));
    NERF_CUDA(cudaMemsetAsync(bn_grads, 0, 4 * (size_t)A.nbn * 4, st));
    if ((rc = nerf_volume_render_bwd(nf_.preds, t_all, drgb_f, nullptr, B, Na, dpf, nullptr, st))) return rc;
    if ((rc = net_backward(st, A, Pf, bn_f, dpf, nf_, grads + A.n_params, bn_grads + 2 * A.nbn, dA, dZ, dF, sums))) return rc;
    if ((rc = nerf_volume_render_bwd(nc_.preds, t, drgb_c, nullptr, B, Nc, dpc, nullptr, st))) return rc;
    if ((rc = net_backward(st, A, Pc, bn_c, dpc, nc_, grads, bn_grads, dA, dZ, dF, sums))) return rc;
    return NERF_OK;
}
