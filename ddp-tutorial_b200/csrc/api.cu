// extern "C" surface of libb2ddp.so -- see include/b2ddp.h for the contract and the
// reference line each entry replaces.  Nothing throws across this boundary.
#include <stdarg.h>
#include <stdlib.h>
#include <new>
#include <algorithm>

#include "model.cuh"

namespace b2 {
thread_local char g_err[1024] = {0};
int64_t g_launches = 0;
bool g_pdl = [] { const char *e = getenv("B2_PDL"); return !(e && e[0] == '0'); }();
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace b2

using namespace b2;

struct b2_model { Model *m; };

#define GUARD_BEGIN try {
#define GUARD_END                                                       \
    } catch (const std::bad_alloc &) {                                  \
        set_error("out of host memory"); return B2_ERR_STATE;           \
    } catch (...) {                                                     \
        set_error("unexpected C++ exception"); return B2_ERR_STATE;     \
    }

extern "C" {

int b2_version(void) { return 100; }
const char *b2_last_error(void) { return g_err; }
int64_t b2_launch_count(void) { return g_launches; }

/* ---- in-graph kernel trace (common.cuh: trace_mark) ---- */
static unsigned long long *g_trace_buf = nullptr;
static unsigned int *g_trace_cnt = nullptr;
static unsigned int g_trace_cap = 0;
static int trace_install_all(unsigned long long *buf, unsigned int *cnt, unsigned int cap) {
    B2_TRY(trace_install_comm(buf, cnt, cap));
    B2_TRY(trace_install_simt(buf, cnt, cap));
    B2_TRY(trace_install_umma(buf, cnt, cap));
    B2_TRY(trace_install_elem(buf, cnt, cap));
    return 0;
}
int b2_trace_start(int capacity) {
    B2_CHECK(capacity > 0 && capacity <= (1 << 20), B2_ERR_ARG, "trace_start: capacity out of range");
    B2_CUDA(cudaDeviceSynchronize());
    if (g_trace_buf) { trace_install_all(nullptr, g_trace_cnt, 0); cudaFree(g_trace_buf); cudaFree(g_trace_cnt); g_trace_buf = nullptr; }
    B2_CUDA(cudaMalloc(&g_trace_buf, (size_t)capacity * 16));
    B2_CUDA(cudaMalloc(&g_trace_cnt, sizeof(unsigned int)));
    B2_CUDA(cudaMemset(g_trace_cnt, 0, sizeof(unsigned int)));
    g_trace_cap = (unsigned int)capacity;
    int r = trace_install_all(g_trace_buf, g_trace_cnt, g_trace_cap);
    if (r) set_error("trace_start: cudaMemcpyToSymbol failed");
    return r;
}
int b2_trace_stop(unsigned long long *records, int max_records, int *n_out) {
    B2_CHECK(g_trace_buf && records && n_out, B2_ERR_STATE, "trace_stop: no trace running / NULL argument");
    B2_CUDA(cudaDeviceSynchronize());
    unsigned int n = 0;
    B2_CUDA(cudaMemcpy(&n, g_trace_cnt, sizeof(n), cudaMemcpyDeviceToHost));
    if (n > g_trace_cap) n = g_trace_cap;
    if ((int)n > max_records) n = (unsigned int)max_records;
    B2_CUDA(cudaMemcpy(records, g_trace_buf, (size_t)n * 16, cudaMemcpyDeviceToHost));
    *n_out = (int)n;
    trace_install_all(nullptr, g_trace_cnt, 0);
    cudaFree(g_trace_buf); cudaFree(g_trace_cnt);
    g_trace_buf = nullptr; g_trace_cnt = nullptr; g_trace_cap = 0;
    return 0;
}

int b2_model_create(int kind, int batch, int num_classes, int precision, int device, b2_model **out) {
    GUARD_BEGIN
    B2_CHECK(out, B2_ERR_ARG, "b2_model_create: out is NULL");
    Model *m = nullptr;
    B2_TRY(Model::create(kind, batch, num_classes, precision, device, &m));
    *out = new b2_model{m};
    return 0;
    GUARD_END
}
int b2_model_destroy(b2_model *h) {
    if (!h) return 0;
    delete h->m;
    delete h;
    return 0;
}
int64_t b2_model_param_count(const b2_model *h) { return h ? h->m->n_params : 0; }
int b2_model_num_tensors(const b2_model *h) { return h ? (int)h->m->params.size() : 0; }
int b2_model_tensor_info(const b2_model *h, int i, const char **name, int64_t *offset, int *rank, int64_t shape[4],
                         int64_t stride[4]) {
    B2_CHECK(h && i >= 0 && i < (int)h->m->params.size(), B2_ERR_ARG, "tensor_info: index %d out of range", i);
    const ParamInfo &p = h->m->params[i];
    if (name) *name = p.name.c_str();
    if (offset) *offset = p.offset;
    if (rank) *rank = p.rank;
    for (int k = 0; k < 4; ++k) { if (shape) shape[k] = p.shape[k]; if (stride) stride[k] = p.stride[k]; }
    return 0;
}
int64_t b2_model_buffer_count(const b2_model *h) { return h ? h->m->n_buffers : 0; }
int b2_model_num_buffers(const b2_model *h) { return h ? (int)h->m->buffers.size() : 0; }
int b2_model_buffer_info(const b2_model *h, int i, const char **name, int64_t *offset, int64_t *numel) {
    B2_CHECK(h && i >= 0 && i < (int)h->m->buffers.size(), B2_ERR_ARG, "buffer_info: index %d out of range", i);
    const BufferInfo &b = h->m->buffers[i];
    if (name) *name = b.name.c_str();
    if (offset) *offset = b.offset;
    if (numel) *numel = b.numel;
    return 0;
}
int64_t b2_model_steps_tracked(const b2_model *h) { return h ? h->m->steps_tracked : 0; }
int b2_model_bind(b2_model *h, float *params, float *grads, float *buffers) {
    B2_CHECK(h && params && grads, B2_ERR_ARG, "b2_model_bind: NULL argument");
    B2_CHECK(((uintptr_t)params % 16) == 0 && ((uintptr_t)grads % 16) == 0, B2_ERR_ARG, "b2_model_bind: vectors must be 16-byte aligned");
    h->m->p_params = params; h->m->p_grads = grads; h->m->p_buffers = buffers;
    h->m->weights_dirty = true;
    h->m->plans_ready = false;
    h->m->drop_graphs();
    return 0;
}
int b2_model_bind_bf16(b2_model *h, void *bf16_params) {
    B2_CHECK(h, B2_ERR_ARG, "NULL model");
    B2_CHECK(h->m->precision == B2_PREC_BF16, B2_ERR_STATE, "bind_bf16: model is not in bf16 mode");
    B2_CHECK(((uintptr_t)bf16_params % 128) == 0, B2_ERR_ARG, "bind_bf16: mirror must be 128-byte aligned");
    h->m->p_params_bf16 = bf16_params ? (bf16 *)bf16_params : h->m->own_bf16;
    h->m->weights_dirty = true; h->m->plans_ready = false; h->m->drop_graphs();
    return 0;
}
int b2_model_mark_weights_clean(b2_model *h) {
    B2_CHECK(h, B2_ERR_ARG, "NULL model");
    h->m->weights_dirty = false;
    return 0;
}
int b2_model_sync_weights(b2_model *h, void *stream) {
    GUARD_BEGIN
    B2_CHECK(h, B2_ERR_ARG, "b2_model_sync_weights: NULL model");
    return h->m->sync_weights((cudaStream_t)stream);
    GUARD_END
}
int b2_forward(b2_model *h, const float *x, float *logits, void *stream) {
    GUARD_BEGIN
    B2_CHECK(h && x, B2_ERR_ARG, "b2_forward: NULL argument");
    return h->m->forward(x, logits, (cudaStream_t)stream);
    GUARD_END
}
int b2_backward(b2_model *h, const float *dlogits, void *stream) {
    GUARD_BEGIN
    B2_CHECK(h && dlogits, B2_ERR_ARG, "b2_backward: NULL argument");
    return h->m->backward(dlogits, (cudaStream_t)stream);
    GUARD_END
}
int b2_model_mark_weights_dirty(b2_model *h) {
    B2_CHECK(h, B2_ERR_ARG, "NULL model");
    h->m->weights_dirty = true;
    return 0;
}
int b2_model_set_graphs(b2_model *h, int enable) {
    B2_CHECK(h, B2_ERR_ARG, "NULL model");
    h->m->use_graphs = enable != 0;
    if (!enable) h->m->drop_graphs();
    return 0;
}
int b2_model_set_ddp(b2_model *h, int enable, float *momentum_buf, float lr, float momentum, int algo, int64_t bucket_elems) {
    GUARD_BEGIN
    B2_CHECK(h, B2_ERR_ARG, "NULL model");
    return h->m->set_ddp(enable != 0, momentum_buf, lr, momentum, algo, bucket_elems, false);
    GUARD_END
}
int b2_model_set_fused_sgd(b2_model *h, int enable, float *momentum_buf, float lr, float momentum, int64_t bucket_elems) {
    GUARD_BEGIN
    B2_CHECK(h, B2_ERR_ARG, "NULL model");
    return h->m->set_ddp(enable != 0, momentum_buf, lr, momentum, 0, bucket_elems, true);
    GUARD_END
}
int b2_model_ddp_first_step(b2_model *h, int first) {
    B2_CHECK(h, B2_ERR_ARG, "NULL model");
    if (h->m->ddp_first != (first ? 1 : 0)) { h->m->ddp_first = first ? 1 : 0; h->m->drop_graphs(); h->m->eager_runs_bwd = 0; }
    return 0;
}
int b2_model_num_buckets(const b2_model *h) { return h ? (int)h->m->buckets.size() : 0; }
int b2_model_bucket_plan(b2_model *h, int64_t bucket_elems, int max_buckets, int64_t *offsets, int64_t *counts, int *after_op) {
    GUARD_BEGIN
    B2_CHECK(h && bucket_elems > 0, B2_ERR_ARG, "bucket_plan: bad arguments");
    std::vector<Model::Bucket> saved = h->m->buckets;
    h->m->compute_buckets(bucket_elems);
    const int n = (int)h->m->buckets.size();
    for (int i = 0; i < n && i < max_buckets; ++i) {
        if (offsets) offsets[i] = h->m->buckets[i].off;
        if (counts) counts[i] = h->m->buckets[i].n;
        if (after_op) after_op[i] = h->m->buckets[i].after_op;
    }
    h->m->buckets = saved;
    return n;
    GUARD_END
}
int64_t b2_model_workspace_bytes(const b2_model *h) { return h ? (int64_t)h->m->slab_bytes : 0; }
int b2_model_op_info(const b2_model *h, int op, int *type, int geom[11], int flags[4]) {
    B2_CHECK(h && op >= 0 && op < (int)h->m->ops.size(), B2_ERR_ARG, "op_info: op %d out of range", op);
    const Op &o = h->m->ops[op];
    if (type) *type = o.type;
    if (geom) { const ConvGeom &g = o.g; int v[11] = {g.N, g.Hi, g.Wi, g.Ci, g.Ho, g.Wo, g.Co, g.kh, g.kw, g.stride, g.pad}; for (int i = 0; i < 11; ++i) geom[i] = v[i]; }
    if (flags) { flags[0] = o.use_umma; flags[1] = o.stem_umma; flags[2] = o.bn_fused; flags[3] = o.is_fc; }
    return 0;
}
int b2_model_profile(b2_model *h, int enable) {
    B2_CHECK(h, B2_ERR_ARG, "NULL model");
    h->m->profiling = enable != 0;
    return 0;
}
int b2_model_profile_collect(b2_model *h, double *ms, double *flops, double *bytes, int *launches) {
    GUARD_BEGIN
    B2_CHECK(h && ms && flops && bytes && launches, B2_ERR_ARG, "profile_collect: NULL argument");
    return h->m->prof_collect(ms, flops, bytes, launches);
    GUARD_END
}
int b2_model_time_conv_kernels(b2_model *h, int iters, double *ms3, double *flops3, int *launches3, void *stream) {
    GUARD_BEGIN
    B2_CHECK(h && ms3 && flops3 && launches3 && iters > 0, B2_ERR_ARG, "time_conv_kernels: bad arguments");
    return h->m->time_conv_kernels(iters, ms3, flops3, launches3, (cudaStream_t)stream);
    GUARD_END
}
int b2_model_check(b2_model *h) {
    GUARD_BEGIN
    B2_CHECK(h, B2_ERR_ARG, "NULL model");
    int e = 0;
    B2_CUDA(cudaMemcpy(&e, h->m->fb_err, sizeof(int), cudaMemcpyDeviceToHost));
    B2_CHECK(e == 0, B2_ERR_TIMEOUT, "model: a fused conv+BatchNorm grid barrier timed out (grid not co-resident?)");
    return 0;
    GUARD_END
}
int b2_model_num_ops(const b2_model *h) { return h ? (int)h->m->ops.size() : 0; }
int b2_model_read_activation(b2_model *h, int op, int want_grad, float *out, int64_t capacity, int64_t *numel,
                             int shape[4], void *stream) {
    GUARD_BEGIN
    B2_CHECK(h && op >= 0 && op < (int)h->m->ops.size(), B2_ERR_ARG, "read_activation: op %d out of range", op);
    const Op &o = h->m->ops[op];
    const Tensor &t = h->m->tensors[o.out];
    if (numel) *numel = t.numel;
    if (shape) { shape[0] = t.N; shape[1] = t.C; shape[2] = t.H; shape[3] = t.W; }
    if (!out) return 0;
    B2_CHECK(capacity >= t.numel, B2_ERR_ARG, "read_activation: capacity %lld < %lld", (long long)capacity, (long long)t.numel);
    const void *src = want_grad ? t.grad : t.val;
    if (o.type == OP_MAXPOOL && o.out_nchw) {   // already NCHW-flat
        B2_CHECK(t.dt == DT_F32, B2_ERR_UNSUPPORTED, "read_activation: flat bf16 tensor");
        B2_CUDA(cudaMemcpyAsync(out, src, t.numel * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return 0;
    }
    return op_nhwc_to_nchw_f32(src, t.dt, out, t.N, t.C, t.H, t.W, (cudaStream_t)stream);
    GUARD_END
}

int b2_ce_loss_grad(const float *logits, const int64_t *labels, int batch, int classes, float *loss_out, float *dlogits,
                    void *stream) {
    GUARD_BEGIN
    B2_CHECK(logits && labels && loss_out, B2_ERR_ARG, "b2_ce_loss_grad: NULL argument");
    return op_ce_loss_grad(logits, labels, batch, classes, loss_out, dlogits, nullptr, (cudaStream_t)stream);
    GUARD_END
}
int b2_ce_loss_grad_running(const float *logits, const int64_t *labels, int batch, int classes, float *loss_out, float *dlogits,
                            float *running2, void *stream) {
    GUARD_BEGIN
    B2_CHECK(logits && labels && loss_out, B2_ERR_ARG, "b2_ce_loss_grad_running: NULL argument");
    return op_ce_loss_grad(logits, labels, batch, classes, loss_out, dlogits, running2, (cudaStream_t)stream);
    GUARD_END
}

int b2_sgd(float *p, const float *g, float *m, int64_t n, float lr, float momentum, float grad_scale, int first_step,
           void *p_bf16, void *stream) {
    GUARD_BEGIN
    B2_CHECK(n >= 0, B2_ERR_ARG, "b2_sgd: negative length");
    if (n == 0) return 0;
    B2_CHECK(p && g && m, B2_ERR_ARG, "b2_sgd: NULL argument");
    return op_sgd(p, g, m, n, lr, momentum, grad_scale, first_step, (bf16 *)p_bf16, (cudaStream_t)stream);
    GUARD_END
}

// ---- input pipeline / eval -------------------------------------------------------
int b2_preprocess_u8(const void *x_u8_nhwc, float *y_nchw, int N, int H, int W, int C, const float *mean3, const float *std3,
                     void *stream) {
    GUARD_BEGIN
    B2_CHECK(x_u8_nhwc && y_nchw && mean3 && std3, B2_ERR_ARG, "b2_preprocess_u8: NULL argument");
    return op_u8_to_normalized((const uint8_t *)x_u8_nhwc, y_nchw, N, H, W, C, mean3, std3, (cudaStream_t)stream);
    GUARD_END
}
int b2_argmax_count(const float *logits, const int64_t *labels, int batch, int classes, long long *counts, void *stream) {
    GUARD_BEGIN
    B2_CHECK(logits && labels && counts, B2_ERR_ARG, "b2_argmax_count: NULL argument");
    return op_argmax_count(logits, labels, batch, classes, counts, (cudaStream_t)stream);
    GUARD_END
}

// ---- kernel-level test entry: one conv pass through the tcgen05 path or the SIMT path -------
int b2_conv_test(int N, int Hi, int Wi, int Ci, int Co, int k, int stride, int pad, int mode, int use_umma, int acc,
                 const void *x_bf16, const float *w_master, const void *dy_bf16, void *out, void *stream) {
    GUARD_BEGIN
    cudaStream_t s = (cudaStream_t)stream;
    ConvGeom g{};
    g.N = N; g.Hi = Hi; g.Wi = Wi; g.Ci = Ci; g.Co = Co; g.kh = g.kw = k; g.stride = stride; g.pad = pad;
    g.Ho = (Hi + 2 * pad - k) / stride + 1; g.Wo = (Wi + 2 * pad - k) / stride + 1;
    const int taps = k * k;
    bf16 *w = nullptr, *wt = nullptr; float *ws = nullptr;
    size_t wsf = std::max(umma_wgrad_workspace(g), conv_simt_wgrad_workspace(g)) + 64;
    B2_CUDA(cudaMalloc(&w, (size_t)Co * taps * Ci * 2));
    B2_CUDA(cudaMalloc(&ws, wsf * 4));
    int rc = op_f32_to_bf16(w_master, w, (int64_t)Co * taps * Ci, s);
    if (rc == 0) {
        if (use_umma) {
            if (!umma_conv_eligible(g)) { set_error("conv_test: shape not eligible for the tcgen05 path"); rc = B2_ERR_UNSUPPORTED; }
            else if (mode == 0) { UmmaConvPlan pl; rc = umma_plan_fprop(g, (const bf16 *)x_bf16, w, (bf16 *)out, &pl); if (!rc) rc = umma_conv_launch(pl, s); }
            else if (mode == 1) { UmmaConvPlan pl; rc = umma_plan_dgrad(g, (const bf16 *)dy_bf16, w, (bf16 *)out, acc != 0, &pl); if (!rc) rc = umma_conv_launch(pl, s); }
            else {
                UmmaWgradPlan pl; rc = umma_plan_wgrad(g, (const bf16 *)x_bf16, (const bf16 *)dy_bf16, (float *)out, ws, wsf, acc != 0 /* deterministic */, &pl);
                if (!rc && (pl.p.atomic || pl.need_zero)) { cudaError_t ze = cudaMemsetAsync(out, 0, (size_t)Co * taps * Ci * 4, s); if (ze != cudaSuccess) rc = B2_ERR_CUDA; }
                if (!rc) rc = umma_wgrad_launch(pl, s);
            }
        } else {
            if (mode == 0) rc = conv_simt_fprop(g, x_bf16, DT_BF16, w, DT_BF16, g.K(), nullptr, false, out, DT_BF16, s);
            else if (mode == 1) rc = conv_simt_dgrad(g, dy_bf16, DT_BF16, w, DT_BF16, g.K(), out, acc != 0, s);
            else rc = conv_simt_wgrad(g, x_bf16, dy_bf16, DT_BF16, (float *)out, ws, wsf, s);
        }
    }
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(w); cudaFree(wt); cudaFree(ws);
    if (rc == 0 && e != cudaSuccess) { set_error("conv_test: %s", cudaGetErrorString(e)); rc = B2_ERR_CUDA; }
    return rc;
    GUARD_END
}


// ---- kernel-level test entry: fprop with the fused BatchNorm(+residual, ReLU) epilogue -------------
// (the launch bench.py times for every conv+BN pair).  x: bf16 NHWC, w: fp32 O,H,W,I, res: bf16 NHWC of the
// output shape or NULL, gamma/beta: fp32 [Co].  Outputs: z = conv(x) (bf16), y = relu?(bn(z) + res) (bf16),
// mean / invstd (fp32 [Co]) of the training-mode batch statistics.
int b2_conv_bn_test(int N, int Hi, int Wi, int Ci, int Co, int k, int stride, int pad, int relu, const void *x_bf16,
                    const float *w_master, const void *res_bf16, const float *gamma, const float *beta, void *z_out, void *y_out,
                    float *mean_out, float *invstd_out, int *info3, long long *timeline13, void *stream) {
    GUARD_BEGIN
    cudaStream_t s = (cudaStream_t)stream;
    ConvGeom g{};
    g.N = N; g.Hi = Hi; g.Wi = Wi; g.Ci = Ci; g.Co = Co; g.kh = g.kw = k; g.stride = stride; g.pad = pad;
    g.Ho = (Hi + 2 * pad - k) / stride + 1; g.Wo = (Wi + 2 * pad - k) / stride + 1;
    B2_CHECK(umma_conv_eligible(g), B2_ERR_UNSUPPORTED, "conv_bn_test: shape not eligible for the tcgen05 path");
    bf16 *w = nullptr; float *scratch = nullptr;
    const size_t wn = (size_t)Co * k * k * Ci, sums = 2 * (size_t)kBnSumCopies * Co;
    B2_CUDA(cudaMalloc(&w, wn * 2));
    B2_CUDA(cudaMalloc(&scratch, (sums + 8 + 32) * 4));
    B2_CUDA(cudaMemsetAsync(scratch, 0, (sums + 8 + 32) * 4, s));
    int rc = op_f32_to_bf16(w_master, w, (int64_t)wn, s);
    UmmaConvPlan pl;
    if (!rc) rc = umma_plan_fprop(g, (const bf16 *)x_bf16, w, (bf16 *)z_out, &pl);
    if (!rc) rc = umma_conv_check_coresident(pl);
    if (!rc) {
        ConvParams &q = pl.p;
        q.stats = scratch;
        q.fuse_bn = 1; q.fb_y = (bf16 *)y_out; q.fb_res = (const bf16 *)res_bf16; q.fb_gamma = gamma; q.fb_beta = beta;
        q.fb_save_mean = mean_out; q.fb_save_invstd = invstd_out; q.fb_running_mean = nullptr; q.fb_running_var = nullptr;
        q.fb_sync = (unsigned int *)(scratch + sums); q.fb_err = (int *)(scratch + sums + 4);
        q.fb_count = (float)g.M(); q.fb_inv_count = 1.0 / (double)g.M(); q.fb_eps = 1e-5f; q.fb_momentum = 0.1f; q.fb_relu = relu ? 1 : 0;
        if (timeline13) q.timeline = (long long *)(scratch + sums + 8);
        if (info3) { info3[0] = pl.bn_tile; info3[1] = pl.grid; info3[2] = pl.p.total_tiles; }
        rc = umma_conv_launch(pl, s);
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (timeline13 && e == cudaSuccess) e = cudaMemcpy(timeline13, scratch + sums + 8, 13 * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(w); cudaFree(scratch);
    if (rc == 0 && e != cudaSuccess) { set_error("conv_bn_test: %s", cudaGetErrorString(e)); rc = B2_ERR_CUDA; }
    return rc;
    GUARD_END
}

// ---- CUDA graphs -------------------------------------------------------------------
int b2_graph_begin(void *stream) {
    B2_CUDA(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
    return 0;
}
int b2_graph_end(void *stream, void **graph_exec) {
    B2_CHECK(graph_exec, B2_ERR_ARG, "b2_graph_end: NULL out");
    cudaGraph_t g = nullptr;
    B2_CUDA(cudaStreamEndCapture((cudaStream_t)stream, &g));
    cudaGraphExec_t e = nullptr;
    cudaError_t r = cudaGraphInstantiate(&e, g, 0);
    cudaGraphDestroy(g);
    if (r != cudaSuccess) { set_error("cudaGraphInstantiate: %s", cudaGetErrorString(r)); return B2_ERR_CUDA; }
    *graph_exec = (void *)e;
    return 0;
}
int b2_graph_launch(void *graph_exec, void *stream) {
    B2_CHECK(graph_exec, B2_ERR_ARG, "b2_graph_launch: NULL graph");
    B2_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
    return 0;
}
int b2_graph_destroy(void *graph_exec) {
    if (graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)graph_exec);
    return 0;
}

}  // extern "C"
