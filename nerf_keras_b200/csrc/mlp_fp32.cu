// fp32 CUDA-core MLP path (NERF_PRECISION_FP32): layer-by-layer Dense kernels that restate
// models.py:24-62 exactly (x @ W + b, ReLU, [h, enc] skip concat, sigma/feature/ddir/rgb heads).
// This is the tight-parity / debug path (and the on-GPU cross-check for the tcgen05 kernel); the
// product path is mlp_tc.cu.  Activations are materialised per chunk of samples in a ctx workspace.
#include "common.cuh"
#include "ctx.cuh"

using namespace nerf;

// Y[m, n] = act(sum_k X[m, k] W[k, n] + b[n]);  X row stride ldx, Y row stride ldy, W (K,N) row-major.
// 64x64 output tile per 256-thread CTA, 4x4 register micro-tile, BK = 16.
template <bool RELU>
__global__ void __launch_bounds__(256) linear_f32_kernel(const float* __restrict__ X, int ldx,
                                                         const float* __restrict__ W, const float* __restrict__ bias,
                                                         float* __restrict__ Y, int ldy, int64_t M, int K, int N) {
    __shared__ float Xs[16][64 + 4];
    __shared__ float Ws[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * 64;
    const int n0 = blockIdx.y * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        // X tile: 64 rows x 16 k  (thread loads 4 elements)
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            int r = i >> 4, kk = i & 15;
            int64_t m = m0 + r;
            int k = k0 + kk;
            Xs[kk][r] = (m < M && k < K) ? X[m * ldx + k] : 0.f;
        }
        for (int i = threadIdx.x; i < 16 * 64; i += 256) {
            int kk = i >> 6, c = i & 63;
            int k = k0 + kk, n = n0 + c;
            Ws[kk][c] = (k < K && n < N) ? W[(int64_t)k * N + n] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = Xs[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Ws[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + bias[n];
            if (RELU) v = fmaxf(v, 0.f);
            Y[m * ldy + n] = v;
        }
    }
}

static int linear_f32(const float* X, int ldx, const float* W, const float* b, float* Y, int ldy, int64_t M, int K,
                      int N, bool relu, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(M, 64), (unsigned)ceil_div(N, 64));
    if (relu)
        linear_f32_kernel<true><<<grid, 256, 0, st>>>(X, ldx, W, b, Y, ldy, M, K, N);
    else
        linear_f32_kernel<false><<<grid, 256, 0, st>>>(X, ldx, W, b, Y, ldy, M, K, N);
    NERF_LAUNCHED();
    return NERF_OK;
}

// copy src (n, c) [row stride lds] into dst columns [col0, col0+c) of rows with stride ldd
__global__ void __launch_bounds__(256) copy_cols_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst,
                                                        int ldd, int col0, int64_t n, int c) {
    const int64_t n_el = n * c;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_el; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e / c;
        int k = (int)(e - r * c);
        dst[r * ldd + col0 + k] = src[r * lds + k];
    }
}

// sample_rays + encode_position x2 for samples [s0, s0+n) of a (B,N) ray batch (models.py:152-154)
__global__ void __launch_bounds__(256) encode_rays_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                                          const float* __restrict__ t, int N, int64_t s0, int64_t n,
                                                          int Lx, int Ld, float* __restrict__ enc_x,
                                                          float* __restrict__ enc_d) {
    const int Cx = 3 + 6 * Lx, Cd = 3 + 6 * Ld, C = Cx + Cd;
    const int64_t n_el = n * C;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_el; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t s = e / C;
        int c = (int)(e - s * C);
        int64_t g = s0 + s;
        int64_t ray = g / N;
        bool is_dir = c >= Cx;
        int cc = is_dir ? c - Cx : c;
        int comp, i = -1, r = 0;
        if (cc < 3) comp = cc;
        else { int q = cc - 3; i = q / 6; r = q - i * 6; comp = r % 3; }
        float dv = d[ray * 3 + comp];
        float base = is_dir ? dv : __fadd_rn(o[ray * 3 + comp], __fmul_rn(dv, t[g]));
        float v = base;
        if (i >= 0) {
            float arg = __fmul_rn(exp2f((float)i), base);
            v = (r >= 3) ? cosf(arg) : sinf(arg);
        }
        if (is_dir) enc_d[s * Cd + cc] = v; else enc_x[s * Cx + cc] = v;
    }
}

namespace nerf {

// runs the fp32 MLP over n samples given encoded inputs (row-major, dense); chunked by the workspace size
int mlp_fp32_forward_encoded(nerf_ctx* ctx, int net, const float* enc_x, const float* enc_d, int64_t n, float* preds,
                             cudaStream_t st) {
    const nerf_config& c = ctx->cfg;
    const int H = c.hidden_dim, Cx = 3 + 6 * c.l_xyz, Cd = 3 + 6 * c.l_dir, L = c.num_layers;
    int rc = ensure_fp32_workspace(ctx);
    if (rc) return rc;
    const int64_t chunk = ctx->fp32_chunk;
    const float* Wb = ctx->params + (int64_t)net * ctx->n_params;
    float* bufA = ctx->ws_a;           // [chunk, H + Cx]
    float* bufB = ctx->ws_b;           // [chunk, H + Cx]
    float* bufC = ctx->ws_c;           // [chunk, H + Cd]
    const int ldab = H + Cx, ldc = H + Cd;
    for (int64_t s0 = 0; s0 < n; s0 += chunk) {
        int64_t m = (n - s0 < chunk) ? n - s0 : chunk;
        const float* x = enc_x + s0 * Cx;
        int ldx = Cx, K = Cx;
        float* cur = bufA;
        float* nxt = bufB;
        for (int i = 0; i < L; ++i) {
            const LayerInfo& li = ctx->layers[i];
            int r2 = linear_f32(x, ldx, Wb + li.w_off, Wb + li.b_off, cur, ldab, m, K, H, true, st);
            if (r2) return r2;
            K = H;
            if (i % c.skip_layer == 0 && i > 0) {  // x = concat([x, ray_input])  (models.py:38-39)
                copy_cols_kernel<<<stream_grid(m * Cx, 256), 256, 0, st>>>(enc_x + s0 * Cx, Cx, cur, ldab, H, m, Cx);
                NERF_LAUNCHED();
                K = H + Cx;
            }
            x = cur; ldx = ldab;
            float* tmp = cur; cur = nxt; nxt = tmp;
        }
        const LayerInfo& ls = ctx->layers[L], &lf = ctx->layers[L + 1], &ld = ctx->layers[L + 2], &lr = ctx->layers[L + 3];
        float* pr = preds + s0 * 4;
        int r2 = linear_f32(x, ldx, Wb + ls.w_off, Wb + ls.b_off, pr + 3, 4, m, K, 1, false, st);          // sigma
        if (r2) return r2;
        r2 = linear_f32(x, ldx, Wb + lf.w_off, Wb + lf.b_off, bufC, ldc, m, K, H, false, st);              // feature
        if (r2) return r2;
        copy_cols_kernel<<<stream_grid(m * Cd, 256), 256, 0, st>>>(enc_d + s0 * Cd, Cd, bufC, ldc, H, m, Cd);
        NERF_LAUNCHED();
        r2 = linear_f32(bufC, ldc, Wb + ld.w_off, Wb + ld.b_off, cur, ldab, m, H + Cd, H / 2, true, st);   // ddir
        if (r2) return r2;
        r2 = linear_f32(cur, ldab, Wb + lr.w_off, Wb + lr.b_off, pr, 4, m, H / 2, 3, false, st);           // rgb
        if (r2) return r2;
    }
    return NERF_OK;
}

int mlp_fp32_forward_rays(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t, int64_t B, int N,
                          float* preds, cudaStream_t st) {
    const nerf_config& c = ctx->cfg;
    const int Cx = 3 + 6 * c.l_xyz, Cd = 3 + 6 * c.l_dir;
    int rc = ensure_fp32_workspace(ctx);
    if (rc) return rc;
    const int64_t chunk = ctx->fp32_chunk;
    const int64_t n = B * N;
    for (int64_t s0 = 0; s0 < n; s0 += chunk) {
        int64_t m = (n - s0 < chunk) ? n - s0 : chunk;
        encode_rays_kernel<<<stream_grid(m * (Cx + Cd), 256), 256, 0, st>>>(o, d, t, N, s0, m, c.l_xyz, c.l_dir,
                                                                           ctx->ws_encx, ctx->ws_encd);
        NERF_LAUNCHED();
        rc = mlp_fp32_forward_encoded(ctx, net, ctx->ws_encx, ctx->ws_encd, m, preds + s0 * 4, st);
        if (rc) return rc;
    }
    return NERF_OK;
}

}  // namespace nerf
