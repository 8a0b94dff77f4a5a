// Context management and the reference-level entry points (models.py) of libnerf_b200.so.
#include "ctx.cuh"

using namespace nerf;

extern "C" int nerf_volume_render_bwd_heads(const float*, const float*, const float*, const float*, int64_t, int, float*,
                                            float*, float*, float*, void*);
extern "C" int nerf_metrics_grad_sums(const float*, const float*, const float*, int64_t, float*, float*, float*, float*,
                                      void*);
extern "C" int nerf_adam_flat(float*, const float*, float*, float*, int64_t, int64_t, float, float, void*);

namespace nerf {
int64_t tc_save_bytes_per_tile();
int tc_train_alloc(nerf_ctx* ctx);
int tc_backward(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t, int64_t B, int N,
                const float* d_preds, cudaStream_t st, int flags);
int adam_from_dev_state(float* params, const float* grads, float* m, float* v, int64_t n, const nerf_dev_state* ds,
                        float grad_scale, cudaStream_t st);
}  // namespace nerf

static int build_layers(const nerf_config& c, std::vector<LayerInfo>& layers, int64_t& n_params) {
    const int ex = 3 + 6 * c.l_xyz, ed = 3 + 6 * c.l_dir, Hd = c.hidden_dim;
    layers.clear();
    int64_t off = 0;
    auto add = [&](int fi, int fo) {
        LayerInfo li;
        li.fan_in = fi; li.fan_out = fo;
        li.w_off = off; off += (int64_t)fi * fo;
        li.b_off = off; off += fo;
        layers.push_back(li);
    };
    int fan_in = ex;
    for (int i = 0; i < c.num_layers; ++i) {
        add(fan_in, Hd);
        fan_in = Hd;
        if (i % c.skip_layer == 0 && i > 0) fan_in = Hd + ex;  // models.py:38-39
    }
    add(fan_in, 1);          // sigma   models.py:42
    add(fan_in, Hd);         // feature models.py:45
    add(Hd + ed, Hd / 2);    // ddir    models.py:48-54
    add(Hd / 2, 3);          // rgb     models.py:57
    n_params = off;
    return NERF_OK;
}

static int validate_cfg(const nerf_config* c) {
    NERF_CHECK_ARG(c != nullptr, "null config");
    NERF_CHECK_ARG(c->num_layers >= 1 && c->num_layers <= 16, "NUM_LAYERS out of range");
    NERF_CHECK_ARG(c->hidden_dim >= 2 && c->hidden_dim <= 1024 && c->hidden_dim % 2 == 0, "HIDDEN_DIM out of range");
    NERF_CHECK_ARG(c->skip_layer >= 1, "SKIP_LAYER must be >= 1");
    NERF_CHECK_ARG(c->l_xyz >= 0 && c->l_xyz <= 16 && c->l_dir >= 0 && c->l_dir <= 16, "L_XYZ/L_DIR out of range");
    NERF_CHECK_ARG(c->ns_coarse >= 2 && c->ns_fine >= 0, "NS_COARSE must be >= 2 and NS_FINE >= 0");
    NERF_CHECK_ARG(c->max_rays >= 1, "max_rays must be >= 1");
    NERF_CHECK_ARG(c->batch_norm == 0, "BATCH_NORM=true is not supported by the B200 path (see DESIGN.md)");
    // the skip concat after the LAST trunk layer would change the head fan-in; the reference configs never do this
    NERF_CHECK_ARG(!((c->num_layers - 1) % c->skip_layer == 0 && c->num_layers - 1 > 0),
                   "skip connection on the last trunk layer is not supported");
    return NERF_OK;
}

extern "C" int64_t nerf_param_count(const nerf_config* cfg) {
    if (validate_cfg(cfg)) return -1;
    std::vector<LayerInfo> layers;
    int64_t n = 0;
    build_layers(*cfg, layers, n);
    return n;
}

extern "C" int nerf_create(const nerf_config* cfg, nerf_ctx** out) {
    NERF_CHECK_ARG(out != nullptr, "null out pointer");
    int rc = validate_cfg(cfg);
    if (rc) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(NERF_ERR_CUDA, "nerf_create: no CUDA device available (this library has no CPU fallback)");
    nerf_ctx* ctx = new nerf_ctx();
    ctx->cfg = *cfg;
    cudaGetDevice(&ctx->device);
    build_layers(*cfg, ctx->layers, ctx->n_params);
    const int64_t np2 = 2 * ctx->n_params;
    const int64_t R = cfg->max_rays, Nc = cfg->ns_coarse, Na = cfg->ns_coarse + cfg->ns_fine;
#define ALLOC(ptr, bytes)                                                          \
    do {                                                                           \
        cudaError_t _e = cudaMalloc((void**)&(ptr), (size_t)(bytes));              \
        if (_e != cudaSuccess) {                                                   \
            nerf_destroy(ctx);                                                     \
            return fail(NERF_ERR_CUDA, std::string("nerf_create: cudaMalloc failed: ") + cudaGetErrorString(_e)); \
        }                                                                          \
    } while (0)
    ALLOC(ctx->params, np2 * 4);
    cudaMemset(ctx->params, 0, np2 * 4);
    ALLOC(ctx->fw_pred_c, R * Nc * 16);
    ALLOC(ctx->fw_pred_f, R * Na * 16);
    ALLOC(ctx->fw_w_c, R * Nc * 4);
    ALLOC(ctx->fw_w_f, R * Na * 4);
    ALLOC(ctx->fw_t_all, R * Na * 4);
    ALLOC(ctx->fw_src_idx, R * Na * 4);
    ALLOC(ctx->fw_rgb_c, R * 12);
    ALLOC(ctx->fw_rgb_f, R * 12);
    ALLOC(ctx->fw_dirbias[0], R * (cfg->hidden_dim / 2) * 4);
    ALLOC(ctx->fw_dirbias[1], R * (cfg->hidden_dim / 2) * 4);
    ALLOC(ctx->dev_state, sizeof(nerf_dev_state));
    ALLOC(ctx->metric_sums, 4 * sizeof(float));
    cudaMemset(ctx->metric_sums, 0, 4 * sizeof(float));
    {
        nerf_dev_state h = {};
        h.lr = cfg->learning_rate;
        cudaMemcpy(ctx->dev_state, &h, sizeof(h), cudaMemcpyHostToDevice);
    }
    if (cfg->training) {
        ALLOC(ctx->grads, np2 * 4);
        ALLOC(ctx->adam_m, np2 * 4);
        ALLOC(ctx->adam_v, np2 * 4);
        cudaMemset(ctx->grads, 0, np2 * 4);
        cudaMemset(ctx->adam_m, 0, np2 * 4);
        cudaMemset(ctx->adam_v, 0, np2 * 4);
        // padded to the 256-row tile pairs the kernels read; the pad stays zero (no gradient from padding rows)
        ALLOC(ctx->tr_dpred_c, (R * Nc + 512) * 16);
        ALLOC(ctx->tr_dpred_f, (R * Na + 512) * 16);
        cudaMemset(ctx->tr_dpred_c, 0, (R * Nc + 512) * 16);
        cudaMemset(ctx->tr_dpred_f, 0, (R * Na + 512) * 16);
        ALLOC(ctx->tr_drgb_c, R * 12);
        ALLOC(ctx->tr_drgb_f, R * 12);
        ALLOC(ctx->tr_ddelta_f, R * Na * 4);
        ALLOC(ctx->tr_dtp_f, R * Na * 4);
        ALLOC(ctx->tr_dw_extra, R * Nc * 4);
    }
#undef ALLOC
    if (tc_supported(*cfg, nullptr)) {
        rc = tc_alloc(ctx);
        if (rc) { nerf_destroy(ctx); return rc; }
        if (cfg->training) {
            rc = tc_train_alloc(ctx);
            if (rc) { nerf_destroy(ctx); return rc; }
        }
    }
    *out = ctx;
    return NERF_OK;
}

extern "C" int nerf_destroy(nerf_ctx* ctx) {
    if (!ctx) return NERF_OK;
    cudaFree(ctx->params); cudaFree(ctx->grads); cudaFree(ctx->adam_m); cudaFree(ctx->adam_v);
    cudaFree(ctx->ws_a); cudaFree(ctx->ws_b); cudaFree(ctx->ws_c); cudaFree(ctx->ws_encx); cudaFree(ctx->ws_encd);
    cudaFree(ctx->fw_pred_c); cudaFree(ctx->fw_pred_f); cudaFree(ctx->fw_w_c); cudaFree(ctx->fw_w_f);
    cudaFree(ctx->fw_t_all); cudaFree(ctx->fw_src_idx); cudaFree(ctx->fw_rgb_c); cudaFree(ctx->fw_rgb_f);
    cudaFree(ctx->fw_dirbias[0]); cudaFree(ctx->fw_dirbias[1]); cudaFree(ctx->dev_state); cudaFree(ctx->metric_sums);
    cudaFree(ctx->tr_dpred_c); cudaFree(ctx->tr_dpred_f); cudaFree(ctx->tr_drgb_c); cudaFree(ctx->tr_drgb_f);
    cudaFree(ctx->tr_ddirsum[0]); cudaFree(ctx->tr_ddirsum[1]); cudaFree(ctx->tr_ddelta_f); cudaFree(ctx->tr_dtp_f); cudaFree(ctx->tr_dw_extra);
    cudaFree(ctx->w_ig); cudaFree(ctx->far_t); cudaFree(ctx->far_pred);
    for (int n = 0; n < 2; ++n) {
        cudaFree(ctx->act_save[n]); cudaFree(ctx->dz_save[n]); cudaFree(ctx->mask_save[n]); cudaFree(ctx->chain_progress[n]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    tc_free(ctx);
    delete ctx;
    return NERF_OK;
}

extern "C" int nerf_set_weights(nerf_ctx* ctx, int net, const float* blob, int64_t n, void* stream) {
    NERF_CHECK_ARG(ctx && blob, "null pointer");
    NERF_CHECK_ARG(net == 0 || net == 1, "net must be NERF_NET_COARSE or NERF_NET_FINE");
    NERF_CHECK_ARG(n == ctx->n_params, "blob size does not match nerf_param_count()");
    NERF_CUDA(cudaMemcpyAsync(ctx->params + (int64_t)net * ctx->n_params, blob, n * 4, cudaMemcpyDefault,
                              (cudaStream_t)stream));
    ctx->weights_set[net] = true;
    ctx->packed_valid[net] = false;
    ctx->bwd_packed_valid = false;
    return NERF_OK;
}

extern "C" int nerf_get_weights(nerf_ctx* ctx, int net, float* blob, int64_t n, void* stream) {
    NERF_CHECK_ARG(ctx && blob, "null pointer");
    NERF_CHECK_ARG(net == 0 || net == 1, "net must be NERF_NET_COARSE or NERF_NET_FINE");
    NERF_CHECK_ARG(n == ctx->n_params, "blob size does not match nerf_param_count()");
    NERF_CUDA(cudaMemcpyAsync(blob, ctx->params + (int64_t)net * ctx->n_params, n * 4, cudaMemcpyDefault,
                              (cudaStream_t)stream));
    return NERF_OK;
}

extern "C" int nerf_grad_buffer(nerf_ctx* ctx, float** grads, int64_t* n) {
    NERF_CHECK_ARG(ctx && grads && n, "null pointer");
    if (!ctx->grads) return fail(NERF_ERR_STATE, "nerf_grad_buffer: ctx was not created with training=1");
    *grads = ctx->grads;
    *n = 2 * ctx->n_params;
    return NERF_OK;
}

namespace nerf {
int ensure_fp32_workspace(nerf_ctx* ctx) {
    if (ctx->ws_a) return NERF_OK;
    const nerf_config& c = ctx->cfg;
    const int64_t chunk = 65536;
    const int Cx = 3 + 6 * c.l_xyz, Cd = 3 + 6 * c.l_dir, Hd = c.hidden_dim;
    NERF_CUDA(cudaMalloc(&ctx->ws_a, chunk * (Hd + Cx) * 4));
    NERF_CUDA(cudaMalloc(&ctx->ws_b, chunk * (Hd + Cx) * 4));
    NERF_CUDA(cudaMalloc(&ctx->ws_c, chunk * (Hd + Cd) * 4));
    NERF_CUDA(cudaMalloc(&ctx->ws_encx, chunk * Cx * 4));
    NERF_CUDA(cudaMalloc(&ctx->ws_encd, chunk * Cd * 4));
    ctx->fp32_chunk = chunk;
    return NERF_OK;
}
}  // namespace nerf

static int check_ready(nerf_ctx* ctx, int net) {
    if (!ctx->weights_set[net]) return fail(NERF_ERR_STATE, "weights of this net were never set (nerf_set_weights)");
    return NERF_OK;
}

// ---- exact far sigma (nerf_set_exact_far_sigma): last sample of every ray through the fp32 MLP ------------------------
__global__ void __launch_bounds__(256) far_gather_kernel(const float* __restrict__ t, int64_t B, int N, float* __restrict__ t_last) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < B; r += (int64_t)gridDim.x * blockDim.x)
        t_last[r] = t[r * N + (N - 1)];
}
__global__ void __launch_bounds__(256) far_patch_kernel(const float4* __restrict__ far_pred, int64_t B, int N,
                                                        float4* __restrict__ preds) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < B; r += (int64_t)gridDim.x * blockDim.x)
        preds[r * N + (N - 1)].w = far_pred[r].w;
}
static int patch_far_sigma(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t, int64_t B, int N,
                           float* preds, cudaStream_t st) {
    if (!ctx->far_t) {
        NERF_CUDA(cudaMalloc(&ctx->far_t, (size_t)ctx->cfg.max_rays * 4));
        NERF_CUDA(cudaMalloc(&ctx->far_pred, (size_t)ctx->cfg.max_rays * 16));
    }
    far_gather_kernel<<<stream_grid(B, 256), 256, 0, st>>>(t, B, N, ctx->far_t);
    NERF_LAUNCHED();
    int rc = mlp_fp32_forward_rays(ctx, net, o, d, ctx->far_t, B, 1, ctx->far_pred, st);
    if (rc) return rc;
    far_patch_kernel<<<stream_grid(B, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(ctx->far_pred), B, N,
                                                          reinterpret_cast<float4*>(preds));
    NERF_LAUNCHED();
    return NERF_OK;
}

extern "C" int nerf_set_exact_far_sigma(nerf_ctx* ctx, int on) {
    NERF_CHECK_ARG(ctx != nullptr, "null ctx");
    ctx->exact_far_sigma = on != 0;
    return NERF_OK;
}

extern "C" int nerf_set_backward_overlap(nerf_ctx* ctx, int wgrad_sms) {
    NERF_CHECK_ARG(ctx != nullptr, "null ctx");
    NERF_CHECK_ARG(wgrad_sms == 0 || (wgrad_sms >= 13 && wgrad_sms <= num_sms() / 2), "wgrad_sms: 0 or 13 .. SMs / 2");
    if (wgrad_sms && !ctx->side_stream) return fail(NERF_ERR_STATE, "nerf_set_backward_overlap: ctx was not created with training=1");
    ctx->wgrad_ctas = wgrad_sms;
    return NERF_OK;
}

extern "C" int nerf_mlp_forward_encoded(nerf_ctx* ctx, int net, const float* rays_enc, const float* dirs_enc,
                                        int64_t n, float* preds, void* stream) {
    NERF_CHECK_ARG(ctx && rays_enc && dirs_enc && preds && n >= 0, "bad arguments");
    NERF_CHECK_ARG(net == 0 || net == 1, "net must be 0 or 1");
    int rc = check_ready(ctx, net);
    if (rc) return rc;
    if (n == 0) return NERF_OK;
    return mlp_fp32_forward_encoded(ctx, net, rays_enc, dirs_enc, n, preds, (cudaStream_t)stream);
}

extern "C" int nerf_mlp_forward_rays(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t,
                                     int64_t batch, int num_samples, int precision, float* preds, void* stream) {
    NERF_CHECK_ARG(ctx && o && d && t && preds && batch >= 0 && num_samples >= 1, "bad arguments");
    NERF_CHECK_ARG(net == 0 || net == 1, "net must be 0 or 1");
    int rc = check_ready(ctx, net);
    if (rc) return rc;
    if (batch == 0) return NERF_OK;
    if (precision == NERF_PRECISION_FP32)
        return mlp_fp32_forward_rays(ctx, net, o, d, t, batch, num_samples, preds, (cudaStream_t)stream);
    NERF_CHECK_ARG(precision == NERF_PRECISION_BF16_TC, "unknown precision");
    std::string why;
    if (!tc_supported(ctx->cfg, &why)) return fail(NERF_ERR_INVALID, why);
    rc = tc_forward_rays(ctx, net, o, d, t, batch, num_samples, preds, false, (cudaStream_t)stream);
    if (rc == NERF_OK && ctx->exact_far_sigma)
        rc = patch_far_sigma(ctx, net, o, d, t, batch, num_samples, preds, (cudaStream_t)stream);
    return rc;
}

// NeRFTrainer.forward_pass (models.py:151-176).  NS_FINE = 0 (single-net benchmark shape) stops after the coarse net.
static int forward_pass_impl(nerf_ctx* ctx, const float* o, const float* d, const float* t, const PdfDraws& draws,
                             int64_t B, int precision, bool save, const nerf_forward_out* out, cudaStream_t st) {
    const nerf_config& c = ctx->cfg;
    const int Nc = c.ns_coarse, Nf = c.ns_fine, Na = Nc + Nf;
    nerf_forward_out z = {};
    if (out) z = *out;
    float* pred_c = z.pred_c ? z.pred_c : ctx->fw_pred_c;
    float* pred_f = z.pred_f ? z.pred_f : ctx->fw_pred_f;
    float* w_c = z.w_c ? z.w_c : ctx->fw_w_c;
    float* w_f = z.w_f ? z.w_f : ctx->fw_w_f;
    float* t_all = z.t_all ? z.t_all : ctx->fw_t_all;
    float* rgb_c = z.rgb_c ? z.rgb_c : ctx->fw_rgb_c;
    float* rgb_f = z.rgb_f ? z.rgb_f : ctx->fw_rgb_f;
    int rc;
    const bool tc = precision != NERF_PRECISION_FP32;
    if (tc && (rc = tc_dirbias(ctx, d, B, Nf > 0 ? 3 : 1, st))) return rc;     // ddir biases of both nets, one launch
    auto mlp = [&](int net, const float* tt, int N, float* preds) -> int {
        if (!tc) return mlp_fp32_forward_rays(ctx, net, o, d, tt, B, N, preds, st);
        int r = tc_forward_rays(ctx, net, o, d, tt, B, N, preds, save, st, true);
        if (r == NERF_OK && ctx->exact_far_sigma && !save) r = patch_far_sigma(ctx, net, o, d, tt, B, N, preds, st);
        return r;
    };
    if ((rc = mlp(NERF_NET_COARSE, t, Nc, pred_c))) return rc;                                       // :152-157
    // the coarse weights are only needed by the resampling: skip the store when nobody asked for them
    if ((rc = nerf_volume_render(pred_c, t, B, Nc, rgb_c, z.depth_c, w_c, z.acc_c, st))) return rc;  // :164
    if (Nf == 0) return NERF_OK;
    if ((rc = resample_merge(t, w_c, draws, B, Nc, Nf, t_all, ctx->fw_src_idx, st))) return rc;      // :165-167
    if ((rc = mlp(NERF_NET_FINE, t_all, Na, pred_f))) return rc;                                     // :169-173
    if ((rc = nerf_volume_render(pred_f, t_all, B, Na, rgb_f, z.depth_f, w_f, z.acc_f, st))) return rc;  // :175
    return NERF_OK;
}

extern "C" int nerf_forward_pass(nerf_ctx* ctx, const float* o, const float* d, const float* t, const float* u_pdf,
                                 int64_t batch, int precision, const nerf_forward_out* out, void* stream) {
    NERF_CHECK_ARG(ctx && o && d && t && batch >= 0, "bad arguments");
    NERF_CHECK_ARG(batch <= ctx->cfg.max_rays, "batch exceeds cfg.max_rays");
    NERF_CHECK_ARG(precision == NERF_PRECISION_FP32 || precision == NERF_PRECISION_BF16_TC, "unknown precision");
    int rc = check_ready(ctx, 0);
    if (rc) return rc;
    if (ctx->cfg.ns_fine > 0 && (rc = check_ready(ctx, 1))) return rc;
    if (precision == NERF_PRECISION_BF16_TC) {
        std::string why;
        if (!tc_supported(ctx->cfg, &why)) return fail(NERF_ERR_INVALID, why);
    }
    if (batch == 0) return NERF_OK;
    // in-kernel draws of an inference pass: keyed by a per-call counter in its own half of the counter space
    PdfDraws dr = {u_pdf, ctx->seed, u_pdf ? 0ull : ((1ull << 62) + ctx->render_draws++), nullptr};
    return forward_pass_impl(ctx, o, d, t, dr, batch, precision, false, out, (cudaStream_t)stream);
}

// NeRFTrainer.train_step up to (not including) apply_gradients  (models.py:88-106)
extern "C" int nerf_train_phases(nerf_ctx* ctx, const float* images, const float* o, const float* d, const float* t,
                                 const float* u_pdf, int64_t batch, float* metrics_dev, int phases, void* stream) {
    NERF_CHECK_ARG(ctx && images && o && d && t && metrics_dev && batch >= 1, "bad arguments");
    NERF_CHECK_ARG(batch <= ctx->cfg.max_rays, "batch exceeds cfg.max_rays");
    NERF_CHECK_ARG((phases & 3) != 0 && (phases & ~7) == 0, "phases: bit 0 forward + fine backward, bit 1 coarse backward, bit 2 accumulate");
    if (!ctx->cfg.training || !ctx->grads) return fail(NERF_ERR_STATE, "ctx was not created with training=1");
    std::string why;
    if (!tc_supported(ctx->cfg, &why)) return fail(NERF_ERR_INVALID, why);
    const bool single = ctx->cfg.ns_fine == 0;
    int rc = check_ready(ctx, 0);
    if (rc) return rc;
    if (!single && (rc = check_ready(ctx, 1))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int Nc = ctx->cfg.ns_coarse, Na = Nc + ctx->cfg.ns_fine;
    const int64_t np = ctx->n_params;
    const bool q5 = !ctx->cfg.stop_grad_samples && !single;   // reference semantics: gradient flows through the fine sample positions
    // uniform draws of sample_pdf: explicit, or Philox keyed by the optimiser step held in device memory
    const PdfDraws dr = {u_pdf, ctx->seed, 0ull, u_pdf ? nullptr : &ctx->dev_state->step};
    const LayerInfo& Lrgb = ctx->layers[ctx->cfg.num_layers + 3];
    const LayerInfo& Lsig = ctx->layers[ctx->cfg.num_layers];
    if (phases & NERF_PHASE_FORWARD_FINE) {
        if ((rc = forward_pass_impl(ctx, o, d, t, dr, batch, NERF_PRECISION_BF16_TC, true, nullptr, st))) return rc;
        if ((rc = nerf_metrics_grad_sums(images, ctx->fw_rgb_c, single ? ctx->fw_rgb_c : ctx->fw_rgb_f, batch, metrics_dev,
                                         ctx->tr_drgb_c, single ? nullptr : ctx->tr_drgb_f, ctx->metric_sums, st)))
            return rc;
        if (!(phases & 4)) NERF_CUDA(cudaMemsetAsync(ctx->grads, 0, 2 * np * 4, st));
        if (!single) {
            // fine net first: its input gradient feeds the coarse net through sort + sample_pdf (models.py:165-175)
            float* gf = ctx->grads + np;
            if ((rc = nerf_volume_render_bwd_heads(ctx->fw_pred_f, ctx->fw_t_all, ctx->tr_drgb_f, nullptr, batch, Na,
                                                   ctx->tr_dpred_f, q5 ? ctx->tr_ddelta_f : nullptr, gf + Lrgb.b_off,
                                                   gf + Lsig.b_off, st)))
                return rc;
            // q5: the fine net's input gradient (dtp) comes out of its weight-gradient kernel as a by-product
            if ((rc = tc_backward(ctx, NERF_NET_FINE, o, d, ctx->fw_t_all, batch, Na, ctx->tr_dpred_f, st, q5 ? 7 : 3))) return rc;
        }
    }
    if (phases & NERF_PHASE_COARSE) {
        const float* d_w_extra = nullptr;
        if (q5) {
            if ((rc = sample_pdf_backward(t, ctx->fw_w_c, dr, ctx->fw_src_idx, ctx->tr_dtp_f, ctx->tr_ddelta_f, batch, Nc,
                                          ctx->cfg.ns_fine, ctx->tr_dw_extra, st)))
                return rc;
            d_w_extra = ctx->tr_dw_extra;
        }
        float* gc = ctx->grads;
        if ((rc = nerf_volume_render_bwd_heads(ctx->fw_pred_c, t, ctx->tr_drgb_c, d_w_extra, batch, Nc, ctx->tr_dpred_c,
                                               nullptr, gc + Lrgb.b_off, gc + Lsig.b_off, st)))
            return rc;
        if ((rc = tc_backward(ctx, NERF_NET_COARSE, o, d, t, batch, Nc, ctx->tr_dpred_c, st, 3))) return rc;
    }
    return NERF_OK;
}

extern "C" int nerf_train_forward_backward(nerf_ctx* ctx, const float* images, const float* o, const float* d,
                                           const float* t, const float* u_pdf, int64_t batch, float* metrics_dev,
                                           void* stream) {
    return nerf_train_phases(ctx, images, o, d, t, u_pdf, batch, metrics_dev, 3, stream);
}

// keras Adam.apply_gradients on [coarse | fine]; t and the learning rate come from device memory, the step counter is
// bumped by the re-pack launch that follows (forward images of both nets), then the transposed images are rebuilt:
// every operand image is current when the next step starts and no later call has to pack anything.
extern "C" int nerf_adam_step(nerf_ctx* ctx, float grad_scale, void* stream) {
    NERF_CHECK_ARG(ctx != nullptr, "null ctx");
    if (!ctx->grads) return fail(NERF_ERR_STATE, "nerf_adam_step: ctx was not created with training=1");
    cudaStream_t st = (cudaStream_t)stream;
    ctx->adam_step += 1;
    int rc = adam_from_dev_state(ctx->params, ctx->grads, ctx->adam_m, ctx->adam_v, 2 * ctx->n_params, ctx->dev_state,
                                 grad_scale, st);
    if (rc) return rc;
    if (tc_supported(ctx->cfg, nullptr)) {
        if ((rc = tc_pack_all(ctx, true, st))) return rc;
        if ((rc = tc_pack_backward(ctx, st))) return rc;
    } else {
        return fail(NERF_ERR_STATE, "nerf_adam_step: training runs on the tcgen05 path only (8x256, skip 4, L 10/4)");
    }
    return NERF_OK;
}

extern "C" int nerf_set_seed(nerf_ctx* ctx, uint64_t seed) {
    NERF_CHECK_ARG(ctx != nullptr, "null ctx");
    ctx->seed = seed;
    ctx->render_draws = 0;
    NERF_CUDA(cudaMemcpy(&ctx->dev_state->seed, &seed, sizeof(seed), cudaMemcpyHostToDevice));
    return NERF_OK;
}

extern "C" int nerf_set_learning_rate(nerf_ctx* ctx, float lr, void* stream) {
    NERF_CHECK_ARG(ctx != nullptr && lr >= 0.f, "bad arguments");
    ctx->cfg.learning_rate = lr;
    // the value is read from pageable host memory: synchronous with respect to the host, ordered on `stream`
    NERF_CUDA(cudaMemcpyAsync(&ctx->dev_state->lr, &ctx->cfg.learning_rate, sizeof(float), cudaMemcpyHostToDevice,
                              (cudaStream_t)stream));
    NERF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return NERF_OK;
}

extern "C" int nerf_get_optimizer_state(nerf_ctx* ctx, float* m, float* v, int64_t* step, void* stream) {
    NERF_CHECK_ARG(ctx && m && v && step, "null pointer");
    if (!ctx->grads) return fail(NERF_ERR_STATE, "ctx was not created with training=1");
    cudaStream_t st = (cudaStream_t)stream;
    NERF_CUDA(cudaMemcpyAsync(m, ctx->adam_m, 2 * ctx->n_params * 4, cudaMemcpyDefault, st));
    NERF_CUDA(cudaMemcpyAsync(v, ctx->adam_v, 2 * ctx->n_params * 4, cudaMemcpyDefault, st));
    // the authoritative count is the device one (graph replays do not pass through nerf_adam_step on the host)
    unsigned long long s = 0;
    NERF_CUDA(cudaMemcpyAsync(&s, &ctx->dev_state->step, sizeof(s), cudaMemcpyDeviceToHost, st));
    NERF_CUDA(cudaStreamSynchronize(st));
    ctx->adam_step = (int64_t)s;
    *step = ctx->adam_step;
    return NERF_OK;
}

extern "C" int nerf_set_optimizer_state(nerf_ctx* ctx, const float* m, const float* v, int64_t step, void* stream) {
    NERF_CHECK_ARG(ctx && m && v && step >= 0, "bad arguments");
    if (!ctx->grads) return fail(NERF_ERR_STATE, "ctx was not created with training=1");
    cudaStream_t st = (cudaStream_t)stream;
    NERF_CUDA(cudaMemcpyAsync(ctx->adam_m, m, 2 * ctx->n_params * 4, cudaMemcpyDefault, st));
    NERF_CUDA(cudaMemcpyAsync(ctx->adam_v, v, 2 * ctx->n_params * 4, cudaMemcpyDefault, st));
    ctx->adam_step = step;
    const unsigned long long s = (unsigned long long)step;
    NERF_CUDA(cudaMemcpyAsync(&ctx->dev_state->step, &s, sizeof(s), cudaMemcpyHostToDevice, st));
    NERF_CUDA(cudaStreamSynchronize(st));
    return NERF_OK;
}

extern "C" int nerf_metric_sums(nerf_ctx* ctx, float** sums_dev) {
    NERF_CHECK_ARG(ctx && sums_dev, "null pointer");
    *sums_dev = ctx->metric_sums;
    return NERF_OK;
}

// NeRFTrainer.test_step (models.py:122-145): metrics of one batch, added to the running sums of the context
extern "C" int nerf_metrics_accumulate(nerf_ctx* ctx, const float* images, const float* rgb_c, const float* rgb_f,
                                       int64_t batch, float* metrics_dev, void* stream) {
    NERF_CHECK_ARG(ctx != nullptr, "null ctx");
    return nerf_metrics_grad_sums(images, rgb_c, rgb_f, batch, metrics_dev, nullptr, nullptr, ctx->metric_sums, stream);
}

// Diagnostics: gradients of sum(preds * d_preds) wrt one net's weights for given rays / t-values
// (forward with saved activations + the full tcgen05 backward).  Result lands in the ctx gradient buffer.
extern "C" int nerf_debug_mlp_grads(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t,
                                    int64_t batch, int num_samples, const float* d_preds, float* preds, void* stream) {
    NERF_CHECK_ARG(ctx && o && d && t && d_preds && preds && batch >= 1, "bad arguments");
    NERF_CHECK_ARG(net == 0 || net == 1, "net must be 0 or 1");
    if (!ctx->cfg.training || !ctx->grads) return fail(NERF_ERR_STATE, "ctx was not created with training=1");
    const int n_max = net == 0 ? ctx->cfg.ns_coarse : ctx->cfg.ns_coarse + ctx->cfg.ns_fine;
    NERF_CHECK_ARG(batch * (int64_t)num_samples <= (int64_t)ctx->cfg.max_rays * n_max, "batch too large for the workspace");
    NERF_CHECK_ARG(batch <= ctx->cfg.max_rays, "batch exceeds cfg.max_rays");
    int rc = check_ready(ctx, net);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = tc_forward_rays(ctx, net, o, d, t, batch, num_samples, preds, true, st))) return rc;
    NERF_CUDA(cudaMemsetAsync(ctx->grads, 0, 2 * ctx->n_params * 4, st));
    float* dp = net == 0 ? ctx->tr_dpred_c : ctx->tr_dpred_f;   // tile-padded staging buffer
    NERF_CUDA(cudaMemcpyAsync(dp, d_preds, (size_t)batch * num_samples * 16, cudaMemcpyDeviceToDevice, st));
    // the fine net also leaves its input gradient in the workspace (nerf_debug_fused_input_grad reads it)
    return tc_backward(ctx, net, o, d, t, batch, num_samples, dp, st, net == 1 ? 4 : 0);
}

// Diagnostics: after nerf_debug_mlp_grads(NERF_NET_FINE, ...) -- the same quantity as nerf_debug_input_grad, as produced
// by the weight-gradient kernel's fused path (what the training step uses); device-to-device copy of batch * num_samples floats.
extern "C" int nerf_debug_fused_input_grad(nerf_ctx* ctx, int64_t batch, int num_samples, float* dtp, void* stream) {
    NERF_CHECK_ARG(ctx && dtp && batch >= 1 && num_samples >= 1, "bad arguments");
    if (!ctx->tr_dtp_f) return fail(NERF_ERR_STATE, "ctx was not created with training=1");
    NERF_CHECK_ARG(batch * (int64_t)num_samples <= (int64_t)ctx->cfg.max_rays * (ctx->cfg.ns_coarse + ctx->cfg.ns_fine), "too large");
    NERF_CUDA(cudaMemcpyAsync(dtp, ctx->tr_dtp_f, (size_t)batch * num_samples * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return NERF_OK;
}

// Diagnostics: after nerf_debug_mlp_grads(net, ...) -- dtp[m] = < d_ray, d(sum(preds * d_preds)) / d pts[m] >.
extern "C" int nerf_debug_input_grad(nerf_ctx* ctx, int net, const float* o, const float* d, const float* t,
                                     int64_t batch, int num_samples, float* dtp, void* stream) {
    NERF_CHECK_ARG(ctx && o && d && t && dtp && batch >= 1, "bad arguments");
    if (!ctx->cfg.training || !ctx->grads) return fail(NERF_ERR_STATE, "ctx was not created with training=1");
    return tc_input_grad(ctx, net, o, d, t, batch, num_samples, dtp, (cudaStream_t)stream);
}
