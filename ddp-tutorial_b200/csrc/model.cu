// Static-graph executor for the three topologies on the reference's hot path:
//   Net          /root/reference/cifar_example_ddp.py:23-40
//   ResNet-18/50 torchvision/models/resnet.py:166-285 (substituted at cifar_example.py:60)
// forward = net(inputs) (cifar_example_ddp.py:104), backward = loss.backward() (:106).
// There is no autograd: the backward schedule is fixed when the model is built, every
// gradient lands in the caller's flat fp32 grad vector (the DDP bucket, reducer.hpp:347).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <string>
#include <algorithm>

#include "model.cuh"

namespace b2 {

// =================================================================================
// graph construction
// =================================================================================
int Model::add_tensor(int N, int H, int W, int C, int dt) {
    Tensor t{};
    t.N = N; t.H = H; t.W = W; t.C = C; t.dt = dt;
    tensors.push_back(t);
    return (int)tensors.size() - 1;
}

int64_t Model::add_param(const std::string &name, int rank, const int64_t *shape, const int64_t *stride, int64_t numel) {
    ParamInfo p{};
    p.name = name; p.rank = rank; p.numel = numel;
    for (int i = 0; i < 4; ++i) { p.shape[i] = i < rank ? shape[i] : 1; p.stride[i] = i < rank ? stride[i] : 0; }
    p.offset = n_params;
    n_params += (numel + 63) / 64 * 64;  // every tensor starts 128-byte aligned in the bf16 mirror (TMA base)
    params.push_back(p);
    return p.offset;
}
int64_t Model::add_buffer(const std::string &name, int64_t numel) {
    BufferInfo b{};
    b.name = name; b.numel = numel; b.offset = n_buffers;
    n_buffers += (numel + 3) / 4 * 4;
    buffers.push_back(b);
    return b.offset;
}

// conv weight stored O,H,W,I; torch's logical O,I,H,W shape is a strided view of it
int64_t Model::add_conv_weight(const std::string &name, int Co, int Ci, int kh, int kw) {
    int64_t shape[4] = {Co, Ci, kh, kw};
    int64_t stride[4] = {(int64_t)kh * kw * Ci, 1, (int64_t)kw * Ci, Ci};
    return add_param(name, 4, shape, stride, (int64_t)Co * Ci * kh * kw);
}
int64_t Model::add_vec(const std::string &name, int n) {
    int64_t shape[1] = {n}, stride[1] = {1};
    return add_param(name, 1, shape, stride, n);
}
int64_t Model::add_mat(const std::string &name, int rows, int cols) {
    int64_t shape[2] = {rows, cols}, stride[2] = {cols, 1};
    return add_param(name, 2, shape, stride, (int64_t)rows * cols);
}

int Model::op_conv(int in, int64_t w_off, int64_t b_off, int Co, int k, int stride, int pad, bool relu, int out_dt) {
    const Tensor &ti = tensors[in];
    Op o{};
    o.type = OP_CONV; o.in = in; o.res = -1;
    o.g.N = ti.N; o.g.Hi = ti.H; o.g.Wi = ti.W; o.g.Ci = ti.C;
    o.g.kh = o.g.kw = k; o.g.stride = stride; o.g.pad = pad;
    o.g.Ho = (ti.H + 2 * pad - k) / stride + 1; o.g.Wo = (ti.W + 2 * pad - k) / stride + 1; o.g.Co = Co;
    o.w_off = w_off; o.b_off = b_off; o.relu = relu;
    o.out = add_tensor(ti.N, o.g.Ho, o.g.Wo, Co, out_dt < 0 ? ti.dt : out_dt);
    ops.push_back(o);
    return o.out;
}
int Model::op_bn(int in, int64_t g_off, int64_t b_off, int64_t rm_off, int64_t rv_off, bool relu, int res) {
    const Tensor &ti = tensors[in];
    Op o{};
    o.type = OP_BN; o.in = in; o.res = res; o.relu = relu;
    o.w_off = g_off; o.b_off = b_off; o.rm_off = rm_off; o.rv_off = rv_off;
    o.out = add_tensor(ti.N, ti.H, ti.W, ti.C, ti.dt);
    ops.push_back(o);
    return o.out;
}
int Model::op_maxpool(int in, int k, int stride, int pad, bool out_nchw_flat) {
    const Tensor &ti = tensors[in];
    Op o{};
    o.type = OP_MAXPOOL; o.in = in; o.res = -1;
    o.g.N = ti.N; o.g.Hi = ti.H; o.g.Wi = ti.W; o.g.Ci = ti.C; o.g.Co = ti.C;
    o.g.kh = o.g.kw = k; o.g.stride = stride; o.g.pad = pad;
    o.g.Ho = (ti.H + 2 * pad - k) / stride + 1; o.g.Wo = (ti.W + 2 * pad - k) / stride + 1;
    o.out_nchw = out_nchw_flat;
    if (out_nchw_flat) o.out = add_tensor(ti.N, 1, 1, ti.C * o.g.Ho * o.g.Wo, ti.dt);
    else o.out = add_tensor(ti.N, o.g.Ho, o.g.Wo, ti.C, ti.dt);
    ops.push_back(o);
    return o.out;
}
int Model::op_avgpool(int in) {
    const Tensor &ti = tensors[in];
    Op o{};
    o.type = OP_AVGPOOL; o.in = in; o.res = -1;
    o.out = add_tensor(ti.N, 1, 1, ti.C, DT_F32);
    ops.push_back(o);
    return o.out;
}

// ---- Net (cifar_example_ddp.py:23-40): always fp32 -- its GEMMs (K=75/150, N=6/16) are far
// below tensor-core tile sizes
void Model::build_net() {
    int64_t c1w = add_conv_weight("conv1.weight", 6, 3, 5, 5), c1b = add_vec("conv1.bias", 6);
    int64_t c2w = add_conv_weight("conv2.weight", 16, 6, 5, 5), c2b = add_vec("conv2.bias", 16);
    int64_t f1w = add_mat("fc1.weight", 120, 400), f1b = add_vec("fc1.bias", 120);
    int64_t f2w = add_mat("fc2.weight", 84, 120), f2b = add_vec("fc2.bias", 84);
    int64_t f3w = add_mat("fc3.weight", num_classes, 84), f3b = add_vec("fc3.bias", num_classes);
    int x = add_tensor(batch, 32, 32, 3, DT_F32);
    input_tensor = x;
    Op in{}; in.type = OP_INPUT; in.in = -1; in.out = x; in.res = -1; ops.push_back(in);
    int t = op_conv(x, c1w, c1b, 6, 5, 1, 0, true, -1);
    t = op_maxpool(t, 2, 2, 0, false);
    t = op_conv(t, c2w, c2b, 16, 5, 1, 0, true, -1);
    t = op_maxpool(t, 2, 2, 0, true);                  // torch.flatten(x, 1) of NCHW: (c,h,w) order
    t = op_conv(t, f1w, f1b, 120, 1, 1, 0, true, -1);  // Linear == 1x1 conv on [B,1,1,K]
    t = op_conv(t, f2w, f2b, 84, 1, 1, 0, true, -1);
    t = op_conv(t, f3w, f3b, num_classes, 1, 1, 0, false, DT_F32);
    logits_tensor = t;
}

struct BnIds { int64_t g, b, rm, rv; };
static BnIds add_bn_params(Model &m, const std::string &prefix, int C) {
    BnIds r;
    r.g = m.add_vec(prefix + ".weight", C);
    r.b = m.add_vec(prefix + ".bias", C);
    r.rm = m.add_buffer(prefix + ".running_mean", C);
    r.rv = m.add_buffer(prefix + ".running_var", C);
    return r;
}

// ---- torchvision ResNet (resnet.py:166-285); BasicBlock :59-105, Bottleneck :108-163
void Model::build_resnet(bool bottleneck, const int layers[4]) {
    const int adt = precision == B2_PREC_BF16 ? DT_BF16 : DT_F32;
    int64_t c1w = add_conv_weight("conv1.weight", 64, 3, 7, 7);
    BnIds b1 = add_bn_params(*this, "bn1", 64);
    int x = add_tensor(batch, 32, 32, 3, adt);
    input_tensor = x;
    Op in{}; in.type = OP_INPUT; in.in = -1; in.out = x; in.res = -1; ops.push_back(in);
    int t = op_conv(x, c1w, -1, 64, 7, 2, 3, false, -1);
    t = op_bn(t, b1.g, b1.b, b1.rm, b1.rv, true, -1);
    t = op_maxpool(t, 3, 2, 1, false);
    int inplanes = 64;
    const int expansion = bottleneck ? 4 : 1;
    for (int L = 0; L < 4; ++L) {
        int planes = 64 << L;
        for (int bi = 0; bi < layers[L]; ++bi) {
            int stride = (bi == 0 && L > 0) ? 2 : 1;
            std::string pre = "layer" + std::to_string(L + 1) + "." + std::to_string(bi);
            bool down = (stride != 1) || (inplanes != planes * expansion);
            // parameters in torch registration order: conv1,bn1,conv2,bn2,(conv3,bn3),downsample
            int64_t w1, w2, w3 = -1, wd = -1;
            BnIds n1, n2, n3{}, nd{};
            if (!bottleneck) {
                w1 = add_conv_weight(pre + ".conv1.weight", planes, inplanes, 3, 3); n1 = add_bn_params(*this, pre + ".bn1", planes);
                w2 = add_conv_weight(pre + ".conv2.weight", planes, planes, 3, 3);   n2 = add_bn_params(*this, pre + ".bn2", planes);
            } else {
                w1 = add_conv_weight(pre + ".conv1.weight", planes, inplanes, 1, 1); n1 = add_bn_params(*this, pre + ".bn1", planes);
                w2 = add_conv_weight(pre + ".conv2.weight", planes, planes, 3, 3);   n2 = add_bn_params(*this, pre + ".bn2", planes);
                w3 = add_conv_weight(pre + ".conv3.weight", planes * 4, planes, 1, 1); n3 = add_bn_params(*this, pre + ".bn3", planes * 4);
            }
            if (down) {
                wd = add_conv_weight(pre + ".downsample.0.weight", planes * expansion, inplanes, 1, 1);
                nd = add_bn_params(*this, pre + ".downsample.1", planes * expansion);
            }
            // the block's first conv comes first (in backward it is then the LAST writer of the block
            // input's gradient); the downsample branch comes before the block's last conv, whose
            // kernel also applies the final BatchNorm + residual + ReLU
            int idn = t, o;
            auto downsample = [&]() {
                if (!down) return;
                int d = op_conv(t, wd, -1, planes * expansion, 1, stride, 0, false, -1);
                idn = op_bn(d, nd.g, nd.b, nd.rm, nd.rv, false, -1);
            };
            if (!bottleneck) {
                o = op_conv(t, w1, -1, planes, 3, stride, 1, false, -1);
                o = op_bn(o, n1.g, n1.b, n1.rm, n1.rv, true, -1);
                downsample();          // before conv2: conv2's kernel applies bn2 and needs the residual
                o = op_conv(o, w2, -1, planes, 3, 1, 1, false, -1);
                o = op_bn(o, n2.g, n2.b, n2.rm, n2.rv, true, idn);
            } else {
                o = op_conv(t, w1, -1, planes, 1, 1, 0, false, -1);
                o = op_bn(o, n1.g, n1.b, n1.rm, n1.rv, true, -1);
                o = op_conv(o, w2, -1, planes, 3, stride, 1, false, -1);
                o = op_bn(o, n2.g, n2.b, n2.rm, n2.rv, true, -1);
                downsample();
                o = op_conv(o, w3, -1, planes * 4, 1, 1, 0, false, -1);
                o = op_bn(o, n3.g, n3.b, n3.rm, n3.rv, true, idn);
            }
            t = o;
            inplanes = planes * expansion;
        }
    }
    int64_t fw = add_mat("fc.weight", num_classes, inplanes), fb = add_vec("fc.bias", num_classes);
    t = op_avgpool(t);
    t = op_conv(t, fw, fb, num_classes, 1, 1, 0, false, DT_F32);
    logits_tensor = t;
}

// =================================================================================
// finalisation: backward schedule flags, memory
// =================================================================================
int Model::finalize() {
    // which gradients are first-written vs accumulated, in reverse op order
    std::vector<int> written(tensors.size(), 0);
    for (int i = (int)ops.size() - 1; i >= 0; --i) {
        Op &o = ops[i];
        if (o.type == OP_INPUT) continue;
        o.need_dgrad = ops_produce(o.in);
        if (o.need_dgrad) { o.acc_in = written[o.in] != 0; written[o.in] = 1; }
        if (o.res >= 0) { o.acc_res = written[o.res] != 0; written[o.res] = 1; }
    }
    // ---- device memory: one slab for activations + grads + per-op scratch
    size_t total = 0;
    auto reserve = [&](size_t bytes) { size_t off = total; total += (bytes + 255) / 256 * 256; return off; };
    std::vector<size_t> voff(tensors.size()), goff(tensors.size());
    for (size_t i = 0; i < tensors.size(); ++i) {
        Tensor &t = tensors[i];
        t.numel = (int64_t)t.N * t.H * t.W * t.C;
        voff[i] = reserve(t.numel * dt_size(t.dt));
        goff[i] = reserve(t.numel * dt_size(t.dt));
    }
    struct Scratch { size_t partial, mean, invstd, coef, idx, col, wpad, dwpad; };
    std::vector<Scratch> so(ops.size());
    size_t wg_floats = 0;
    for (size_t i = 0; i < ops.size(); ++i) {
        Op &o = ops[i];
        if (o.type == OP_BN) {
            const Tensor &t = tensors[o.in];
            int64_t M = (int64_t)t.N * t.H * t.W;
            o.bn.row_blocks = bn_row_blocks(M, t.C);
            so[i].partial = reserve((size_t)o.bn.row_blocks * 2 * t.C * 4);
            so[i].mean = reserve(t.C * 4); so[i].invstd = reserve(t.C * 4); so[i].coef = reserve(3 * t.C * 4);
        } else if (o.type == OP_MAXPOOL) {
            so[i].idx = reserve((size_t)o.g.N * o.g.Ho * o.g.Wo * o.g.Ci);
        } else if (o.type == OP_CONV) {
            wg_floats = std::max(wg_floats, conv_simt_wgrad_workspace(o.g));
            o.wdt = DT_F32; o.wpitch = o.g.K();
            o.is_fc = tensors[o.in].dt == DT_F32 && tensors[o.out].dt == DT_F32 && o.g.kh == 1 && o.g.Hi == 1 && o.g.Wi == 1 &&
                      o.g.stride == 1 && o.g.pad == 0 && o.g.Co <= 16 && kind != B2_MODEL_NET;
            if (tensors[o.in].dt == DT_BF16 && o.in == input_tensor && conv_impl == 1 && !o.relu && o.b_off < 0 && o.g.Co % 64 == 0 &&
                o.g.kh == 7 && o.g.Ci == 3) {
                o.kpad = (o.g.Ci * o.g.kh * 8 + 63) / 64 * 64;      // GEMM-K order (c, kh, kw padded to 8)
                ConvGeom v{};
                v.N = o.g.N; v.Hi = v.Ho = o.g.Ho; v.Wi = v.Wo = o.g.Wo; v.Ci = o.kpad; v.Co = o.g.Co;
                v.kh = v.kw = 1; v.stride = 1; v.pad = 0;
                if (umma_conv_eligible(v)) {
                    o.stem_umma = true; o.vg = v; o.wdt = DT_BF16;
                    so[i].col = reserve((size_t)v.M() * o.kpad * 2);
                    so[i].wpad = reserve((size_t)v.Co * o.kpad * 2);
                    so[i].dwpad = reserve((size_t)v.Co * o.kpad * 4);
                    wg_floats = std::max(wg_floats, umma_wgrad_workspace(v));
                }
            }
            if (!o.stem_umma && tensors[o.in].dt == DT_BF16) {
                o.wdt = DT_BF16;
                o.use_umma = conv_impl == 1 && tensors[o.out].dt == DT_BF16 && !o.relu && o.b_off < 0 && umma_conv_eligible(o.g);
                if (o.use_umma) wg_floats = std::max(wg_floats, umma_wgrad_workspace(o.g));
            }
        }
    }
    // classifier head: avgpool immediately followed by its only consumer, the fc layer, and no ReLU in between
    for (size_t i = 0; i + 1 < ops.size(); ++i)
        if (ops[i].type == OP_AVGPOOL && ops[i + 1].type == OP_CONV && ops[i + 1].is_fc && ops[i + 1].in == ops[i].out && !ops[i + 1].relu) {
            ops[i].head_fused = true; ops[i + 1].head_fused = true;
        }
    // fused BatchNorm: producer conv runs on the tcgen05 path (its epilogue emits the statistics)
    size_t sums_floats = 0;
    std::vector<size_t> sum_off(ops.size(), 0);
    for (size_t i = 0; i < ops.size(); ++i) {
        Op &o = ops[i];
        if (o.type != OP_BN) continue;
        for (const Op &c : ops)
            if (c.type == OP_CONV && c.out == o.in && (c.use_umma || c.stem_umma)) o.bn_fused = true;
        if (o.bn_fused) { sum_off[i] = sums_floats; sums_floats += 4 * (size_t)kBnSumCopies * tensors[o.in].C + 4; }
    }
    size_t sums_off = reserve(sums_floats * 4 + 256);
    size_t err_off = reserve(256);

    size_t wg_off = reserve(wg_floats * 4 + 256);
    size_t bf_off = precision == B2_PREC_BF16 ? reserve((size_t)n_params * 2 + 256) : 0;
    wgrad_floats = wg_floats;
    slab_bytes = total;
    if (device < 0) return 0;        // metadata-only instance (layout / schedule queries on a host without a GPU)
    B2_CUDA(cudaMalloc(&slab, total));
    B2_CUDA(cudaMemset(slab, 0, total));
    slab_bytes = total;
    char *base = (char *)slab;
    for (size_t i = 0; i < tensors.size(); ++i) { tensors[i].val = base + voff[i]; tensors[i].grad = base + goff[i]; }
    for (size_t i = 0; i < ops.size(); ++i) {
        Op &o = ops[i];
        if (o.type == OP_BN) {
            o.bn.partial = (float *)(base + so[i].partial); o.bn.mean = (float *)(base + so[i].mean);
            o.bn.invstd = (float *)(base + so[i].invstd); o.bn.coef = (float *)(base + so[i].coef);
        } else if (o.type == OP_MAXPOOL) {
            o.idx = (uint8_t *)(base + so[i].idx);
        } else if (o.type == OP_CONV && o.stem_umma) {
            o.col = (bf16 *)(base + so[i].col); o.wpad = (bf16 *)(base + so[i].wpad); o.dwpad = (float *)(base + so[i].dwpad);
        }
    }
    wgrad_ws = (float *)(base + wg_off);
    bn_sums = (float *)(base + sums_off); bn_sums_bytes = sums_floats * 4;
    fb_err = (int *)(base + err_off);

    for (size_t i = 0; i < ops.size(); ++i)
        if (ops[i].type == OP_BN && ops[i].bn_fused) {
            ops[i].fsum = bn_sums + sum_off[i]; ops[i].bsum = ops[i].fsum + 2 * (size_t)kBnSumCopies * tensors[ops[i].in].C;
            ops[i].fb_counter = (unsigned int *)(ops[i].bsum + 2 * (size_t)kBnSumCopies * tensors[ops[i].in].C);
        }
    if (precision == B2_PREC_BF16) { own_bf16 = (bf16 *)(base + bf_off); p_params_bf16 = own_bf16; }
    return 0;
}

bool Model::ops_produce(int tensor) const {
    // does any differentiable op produce this tensor (i.e. is a gradient w.r.t. it needed)?
    for (const Op &o : ops)
        if (o.out == tensor) return o.type != OP_INPUT;
    return false;
}

// params each op produces gradients for, as (offset, padded length) in the flat vector
static void op_param_ranges(const Model &m, const Op &o, std::vector<std::pair<int64_t, int64_t>> &out) {
    auto find = [&](int64_t off) {
        for (const ParamInfo &p : m.params) if (p.offset == off) { out.push_back({p.offset, (p.numel + 63) / 64 * 64}); return; }
    };
    if (o.type == OP_CONV) { find(o.w_off); if (o.b_off >= 0) find(o.b_off); }
    else if (o.type == OP_BN) { find(o.w_off); find(o.b_off); }
}

void Model::compute_buckets(int64_t bucket_elems) {
    buckets.clear();
    std::vector<char> done(params.size(), 0);
    int64_t suffix = n_params;           // everything at offsets >= suffix has its gradient
    int64_t emitted = n_params;
    // The stem's backward (max-pool backward, its BatchNorm backward over the largest activation of the net, the 7x7
    // wgrad) is a ~90 us serial tail with no convolution running beside it.  Whatever the cap says, a bucket closes
    // right before that tail, so that only the stem's own few parameters wait for the end of backward (a one-shot,
    // latency-bound exchange) while the bucket before it is exchanged UNDER the tail (measured with tools/step_trace.py:
    // the remainder bucket of a 16 MB cap, 8.3 MB, was 37 us of exposed exchange at the end of every step).
    int tail_first = -1;                 // first op after the stem's BatchNorm
    for (size_t i = 0; i < ops.size(); ++i) if (ops[i].type == OP_BN) { tail_first = (int)i + 1; break; }
    for (int i = (int)ops.size() - 1; i >= 0; --i) {
        std::vector<std::pair<int64_t, int64_t>> r;
        op_param_ranges(*this, ops[i], r);
        for (auto &pr : r)
            for (size_t k = 0; k < params.size(); ++k) if (params[k].offset == pr.first) done[k] = 1;
        // largest complete suffix (params are stored in ascending offset order)
        suffix = n_params;
        for (int k = (int)params.size() - 1; k >= 0 && done[k]; --k) suffix = params[k].offset;
        if (emitted - suffix >= bucket_elems || ((i == 0 || i == tail_first) && emitted > suffix)) {
            buckets.push_back(Bucket{suffix, emitted - suffix, i});
            emitted = suffix;
        }
    }
    if (emitted > 0) buckets.push_back(Bucket{0, emitted, 0});
}

int Model::set_ddp(bool enable, float *mom, float lr, float momentum, int algo, int64_t bucket_elems, bool local) {
    drop_graphs();
    eager_runs_bwd = 0;
    ddp_on = false;
    if (!enable) return 0;
    ddp_local = local;
    B2_CHECK(local || comm_ready(), B2_ERR_STATE, "set_ddp: communicator not initialised / no buffers registered");
    B2_CHECK(mom, B2_ERR_ARG, "set_ddp: NULL momentum buffer");
    ddp_mom = mom; ddp_lr = lr; ddp_momentum = momentum; ddp_algo = algo;
    if (!ddp_stream) {
        B2_CUDA(cudaStreamCreateWithFlags(&ddp_stream, cudaStreamNonBlocking));
        B2_CUDA(cudaEventCreateWithFlags(&ddp_fork, cudaEventDisableTiming));
        B2_CUDA(cudaEventCreateWithFlags(&ddp_join, cudaEventDisableTiming));
    }
    compute_buckets(bucket_elems > 0 ? bucket_elems : (int64_t)2 << 20);
    ddp_on = true;
    return 0;
}

Model::~Model() {
    drop_graphs();
    if (ddp_stream) { cudaStreamDestroy(ddp_stream); cudaEventDestroy(ddp_fork); cudaEventDestroy(ddp_join); }
    if (wg_stream) { cudaStreamDestroy(wg_stream); cudaEventDestroy(wg_fork); cudaEventDestroy(wg_join); }
    if (cap_stream) cudaStreamDestroy(cap_stream);
    if (slab) cudaFree(slab);
}

// =================================================================================
// profiling (events around each op; only when enabled -- never in the timed bench region)
// =================================================================================
int Model::prof_begin(int cat, double flops, double bytes, cudaStream_t s) {
    if (!profiling) return -1;
    ProfRec r{};
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    r.cat = cat; r.flops = flops; r.bytes = bytes;
    cudaEventRecord(r.a, s);
    prof.push_back(r);
    return (int)prof.size() - 1;
}
void Model::prof_end(int h, cudaStream_t s) {
    if (h >= 0) cudaEventRecord(prof[h].b, s);
}
int Model::prof_collect(double ms[PC_COUNT], double flops[PC_COUNT], double bytes[PC_COUNT], int launches[PC_COUNT]) {
    for (int i = 0; i < PC_COUNT; ++i) { ms[i] = 0; flops[i] = 0; bytes[i] = 0; launches[i] = 0; }
    B2_CUDA(cudaDeviceSynchronize());
    for (ProfRec &r : prof) {
        float t = 0.f;
        B2_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.cat] += t; flops[r.cat] += r.flops; bytes[r.cat] += r.bytes; launches[r.cat] += 1;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    prof.clear();
    return 0;
}

int Model::time_conv_kernels(int iters, double ms[3], double flops[3], int launches[3], cudaStream_t s) {
    B2_CHECK(plans_ready, B2_ERR_STATE, "time_conv_kernels: run a forward first");
    for (int i = 0; i < 3; ++i) { ms[i] = 0; flops[i] = 0; launches[i] = 0; }
    cudaEvent_t a, b;
    B2_CUDA(cudaEventCreate(&a)); B2_CUDA(cudaEventCreate(&b));
    // every launch starts from zeroed BatchNorm sums (as in a real step: the fused conv+BN kernels then pay
    // their grid barrier and normalise real statistics) and from zeroed gradients (atomic split-K weight
    // gradients); the memsets are part of the timed region.  Running statistics are saved and restored.
    float *saved_buffers = nullptr;
    if (p_buffers && n_buffers) {
        B2_CUDA(cudaMalloc(&saved_buffers, (size_t)n_buffers * 4));
        B2_CUDA(cudaMemcpyAsync(saved_buffers, p_buffers, (size_t)n_buffers * 4, cudaMemcpyDeviceToDevice, s));
    }
    auto timed = [&](int cls, double fl, auto &&launch) -> int {
        for (int w = 0; w < 2; ++w) { if (bn_sums_bytes) B2_CUDA(cudaMemsetAsync(bn_sums, 0, bn_sums_bytes, s)); B2_TRY(launch()); }
        B2_CUDA(cudaEventRecord(a, s));
        for (int it = 0; it < iters; ++it) { if (bn_sums_bytes) B2_CUDA(cudaMemsetAsync(bn_sums, 0, bn_sums_bytes, s)); B2_TRY(launch()); }
        B2_CUDA(cudaEventRecord(b, s));
        B2_CUDA(cudaEventSynchronize(b));
        float t = 0.f;
        B2_CUDA(cudaEventElapsedTime(&t, a, b));
        ms[cls] += t / iters; flops[cls] += fl; launches[cls] += 1;
        return 0;
    };
    for (Op &o : ops) {
        if (o.type != OP_CONV || !(o.use_umma || o.stem_umma)) continue;
        if (o.stem_umma && o.stem_gather) { B2_CHECK(cur_x, B2_ERR_STATE, "time_conv_kernels: no input image yet"); o.fplan.p.stem_x = cur_x; o.wplan.p.stem_x = cur_x; }
        B2_TRY(timed(0, o.fplan.flops, [&] { return umma_conv_launch(o.fplan, s); }));
        if (o.use_umma && o.need_dgrad) B2_TRY(timed(1, o.dplan.flops, [&] { return umma_conv_launch(o.dplan, s); }));
        B2_TRY(timed(2, o.wplan.flops, [&] { return umma_wgrad_launch(o.wplan, s); }));
    }
    if (saved_buffers) {
        B2_CUDA(cudaMemcpyAsync(p_buffers, saved_buffers, (size_t)n_buffers * 4, cudaMemcpyDeviceToDevice, s));
        B2_CUDA(cudaStreamSynchronize(s));
        cudaFree(saved_buffers);
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    return 0;
}

// =================================================================================
// execution
// =================================================================================
int Model::sync_weights(cudaStream_t s) {
    B2_CHECK(p_params, B2_ERR_STATE, "model: parameters not bound");
    if (p_params_bf16) B2_TRY(op_f32_to_bf16(p_params, p_params_bf16, n_params, s));
    weights_dirty = false;
    return 0;
}

int Model::build_plans() {
    zero_grads_each_step = false;
    for (Op &o : ops) {
        if (o.type == OP_CONV && o.stem_umma) {
            Tensor &to = tensors[o.out];
            B2_TRY(umma_plan_fprop(o.vg, o.col, o.wpad, (bf16 *)to.val, &o.fplan));
            // gather the im2col rows in shared memory straight from the fp32 image (no 50 MB im2col round trip)
            static const bool gather = [] { const char *e = getenv("B2_STEM_GATHER"); return !(e && e[0] == '0'); }();
            o.stem_gather = gather && umma_stem_supported(o.g, o.vg, o.kpad) && o.fplan.bn_tile == 64;
            o.fplan.stem = o.stem_gather;
            B2_TRY(umma_plan_wgrad(o.vg, o.col, (const bf16 *)to.grad, o.dwpad, wgrad_ws, wgrad_floats, deterministic, &o.wplan));
            if (o.wplan.p.atomic) zero_grads_each_step = true;
            o.wplan.stem = o.stem_gather;
            continue;
        }
        if (o.type != OP_CONV || !o.use_umma) continue;
        Tensor &ti = tensors[o.in], &to = tensors[o.out];
        const bf16 *wb = p_params_bf16 + o.w_off;
        B2_TRY(umma_plan_fprop(o.g, (const bf16 *)ti.val, wb, (bf16 *)to.val, &o.fplan));
        if (o.need_dgrad) {
            // a 1x1 stride-2 dgrad only touches even pixels: zero-fill first when it is the first writer
            bool sparse = o.g.stride == 2 && o.g.kh == 1;
            B2_TRY(umma_plan_dgrad(o.g, (const bf16 *)to.grad, wb, (bf16 *)ti.grad, o.acc_in || sparse, &o.dplan));
        }
        B2_TRY(umma_plan_wgrad(o.g, (const bf16 *)ti.val, (const bf16 *)to.grad, p_grads + o.w_off, wgrad_ws, wgrad_floats, deterministic, &o.wplan));
        if (o.wplan.p.atomic) zero_grads_each_step = true;
        B2_TRY(umma_wgrad_zero_dead(o.wplan, 0));      // once: dead-tap gradients are exactly zero and never rewritten
    }
    B2_CUDA(cudaStreamSynchronize(0));
    for (Op &b : ops) {
        if (b.type != OP_BN || !b.bn_fused) continue;
        b.fwd_in_conv = false;
        static const bool allow_fb = [] { const char *e = getenv("B2_FUSE_BN_FWD"); return !(e && e[0] == '0'); }();
        for (Op &c : ops)
            if (c.type == OP_CONV && c.out == b.in) {
                ConvParams &q = c.fplan.p;
                q.stats = b.fsum;
                // conv + BatchNorm(+residual, ReLU) in one launch: the kernel is persistent (grid <= #SMs), so its
                // grid barrier only needs one CTA per SM to be resident -- verified with the occupancy API
                if (allow_fb && umma_conv_check_coresident(c.fplan) == 0) {
                    const Tensor &tx = tensors[b.in];
                    q.fuse_bn = 1;
                    q.fb_y = (bf16 *)tensors[b.out].val;
                    q.fb_res = b.res >= 0 ? (const bf16 *)tensors[b.res].val : nullptr;
                    q.fb_gamma = p_params + b.w_off; q.fb_beta = p_params + b.b_off;
                    q.fb_save_mean = b.bn.mean; q.fb_save_invstd = b.bn.invstd;
                    q.fb_running_mean = p_buffers ? p_buffers + b.rm_off : nullptr;
                    q.fb_running_var = p_buffers ? p_buffers + b.rv_off : nullptr;
                    q.fb_sync = b.fb_counter; q.fb_err = fb_err;
                    q.fb_count = (float)((int64_t)tx.N * tx.H * tx.W); q.fb_inv_count = 1.0 / (double)((int64_t)tx.N * tx.H * tx.W); q.fb_eps = 1e-5f; q.fb_momentum = 0.1f;
                    q.fb_relu = b.relu ? 1 : 0;
                    b.fwd_in_conv = true;
                }
            }
        // last writer of d(b.out) in backward order = its first consumer in forward order
        b.bwd_sums_fused = false;
        for (Op &c : ops) {
            bool consumes = (c.in == b.out) || (c.res == b.out);
            if (!consumes) continue;
            bool covers_all = c.g.stride == 1 || c.g.kh >= c.g.stride;      // a 1x1 stride-2 dgrad only touches even pixels
            // the dgrad that finalises d(b.out) reads x / y row-coalesced in its staged epilogue and adds the
            // BatchNorm-backward sums there: no separate reduce pass over (x, y, dy)  (B2_FUSE_BN_BWD=0 disables)
            static const bool enable = [] { const char *e = getenv("B2_FUSE_BN_BWD"); return !(e && e[0] == '0'); }();
            if (enable && c.type == OP_CONV && c.use_umma && c.need_dgrad && c.in == b.out && covers_all) {
                ConvParams &q = c.dplan.p;
                q.bn_bsum = b.bsum; q.bn_x = (const bf16 *)tensors[b.in].val; q.bn_y = (const bf16 *)tensors[b.out].val;
                q.bn_mean = b.bn.mean; q.bn_invstd = b.bn.invstd; q.bn_relu = b.relu ? 1 : 0;
                b.bwd_sums_fused = true;
            }
            break;      // only the first consumer matters
        }
    }
    plans_ready = true;
    return 0;
}

void Model::drop_graphs() {
    for (auto &g : fwd_graphs) cudaGraphExecDestroy(g.exec);
    for (auto &g : bwd_graphs) cudaGraphExecDestroy(g.exec);
    fwd_graphs.clear(); bwd_graphs.clear();
}

int Model::run_graphed(std::vector<GraphEntry> &cache, const void *a, const void *b, bool is_fwd, cudaStream_t s) {
    for (GraphEntry &g : cache)
        if (g.a == a && g.b == b) {
            B2_CUDA(cudaGraphLaunch(g.exec, s));
            g_launches += g.launches;
            return 0;
        }
    if (!cap_stream) {
        // the forward / backward main path is captured on a HIGH-priority stream: its kernel nodes then win the SMs
        // that free up over the concurrently running weight-gradient / optimizer branches (default priority)
        int lo = 0, hi = 0;
        B2_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        B2_CUDA(cudaStreamCreateWithPriority(&cap_stream, cudaStreamNonBlocking, hi));
    }
    if (cache.size() >= 8) { cudaGraphExecDestroy(cache.front().exec); cache.erase(cache.begin()); }
    const int64_t l0 = g_launches;
    B2_CUDA(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
    int rc = is_fwd ? forward_impl((const float *)a, (float *)const_cast<void *>(b), cap_stream)
                    : backward_impl((const float *)a, cap_stream);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(cap_stream, &graph);
    const int64_t captured = g_launches - l0;
    g_launches = l0;
    if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) { set_error("graph capture failed: %s", cudaGetErrorString(e)); return B2_ERR_CUDA; }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { set_error("cudaGraphInstantiate: %s", cudaGetErrorString(e)); return B2_ERR_CUDA; }
    cache.push_back(GraphEntry{a, b, exec, captured});
    B2_CUDA(cudaGraphLaunch(exec, s));
    g_launches += captured;
    return 0;
}

int Model::forward(const float *x_nchw, float *logits, cudaStream_t s) {
    B2_CHECK(device >= 0 && slab, B2_ERR_STATE, "model: metadata-only instance (created with device < 0) cannot execute");
    B2_CHECK(p_params && p_grads, B2_ERR_STATE, "model: parameters/gradients not bound (b2_model_bind)");
    if (!plans_ready) B2_TRY(build_plans());
    // the first two calls run eagerly (lazy module loading, function attributes), then replay
    if (weights_dirty) B2_TRY(sync_weights(s));        // never inside a graph: depends on host state
    if (use_graphs && !profiling && eager_runs_fwd >= 2) {
        B2_TRY(run_graphed(fwd_graphs, x_nchw, logits, true, s));
        ++steps_tracked;
        return 0;
    }
    ++eager_runs_fwd;
    B2_TRY(forward_impl(x_nchw, logits, s));
    ++steps_tracked;
    return 0;
}

int Model::forward_impl(const float *x_nchw, float *logits, cudaStream_t s) {
    cur_x = x_nchw;
    if (bn_sums_bytes) B2_CUDA(cudaMemsetAsync(bn_sums, 0, bn_sums_bytes, s));
    // split-K weight gradients accumulate with red.global.add: the flat gradient vector must be zero when
    // backward starts.  The previous step's optimizer has consumed it (stream order), so zero it now on the
    // weight-gradient stream, under the forward pass (joined before forward returns: graph-capturable).
    const bool zero_side = zero_grads_each_step && wgrad_async && !profiling;
    if (zero_grads_each_step) {
        if (zero_side) {
            B2_TRY(ensure_side_streams());
            B2_CUDA(cudaEventRecord(wg_fork, s));
            B2_CUDA(cudaStreamWaitEvent(wg_stream, wg_fork, 0));
            B2_CUDA(cudaMemsetAsync(p_grads, 0, (size_t)n_params * 4, wg_stream));
            for (Op &o : ops)        // the stem accumulates into its own padded dW buffer
                if (o.type == OP_CONV && o.stem_umma && o.wplan.p.atomic)
                    B2_CUDA(cudaMemsetAsync(o.dwpad, 0, (size_t)o.vg.Co * o.kpad * 4, wg_stream));
            B2_CUDA(cudaEventRecord(wg_join, wg_stream));
        } else {
            B2_CUDA(cudaMemsetAsync(p_grads, 0, (size_t)n_params * 4, s));
            for (Op &o : ops)
                if (o.type == OP_CONV && o.stem_umma && o.wplan.p.atomic)
                    B2_CUDA(cudaMemsetAsync(o.dwpad, 0, (size_t)o.vg.Co * o.kpad * 4, s));
        }
    }

    for (Op &o : ops) {
        Tensor &to = tensors[o.out];
        switch (o.type) {
        case OP_INPUT: {
            bool needed = false;
            for (const Op &c : ops) if ((c.in == o.out || c.res == o.out) && !(c.type == OP_CONV && c.stem_umma)) needed = true;
            if (!needed) break;
            int ph = prof_begin(PC_OTHER, 0, to.numel * (4.0 + dt_size(to.dt)), s);
            B2_TRY(op_nchw_to_nhwc(x_nchw, to.val, to.dt, to.N, to.C, to.H, to.W, s));
            prof_end(ph, s);
            break;
        }
        case OP_CONV: {
            Tensor &ti = tensors[o.in];
            if (o.stem_umma) {
                int ph = prof_begin(PC_CONV_FPROP, 2.0 * o.g.M() * o.g.Co * o.g.K(), (double)o.vg.M() * o.kpad * 4 + to.numel * 2.0, s);
                if (o.stem_gather) o.fplan.p.stem_x = cur_x;
                else B2_TRY(op_stem_im2col(cur_x, o.col, o.g.N, o.g.Ci, o.g.Hi, o.g.Wi, o.g.kh, o.g.stride, o.g.pad, o.g.Ho, o.g.Wo, o.kpad, s));
                B2_TRY(op_stem_pack_weights(p_params + o.w_off, o.wpad, o.g.Co, o.g.Ci, o.g.kh, o.kpad, s));
                B2_TRY(umma_conv_launch(o.fplan, s));
                prof_end(ph, s);
                break;
            }
            if (o.is_fc) {
                int ph = prof_begin(PC_OTHER, 0, 0, s);
                if (o.head_fused) {          // avgpool + fc in one kernel (the avgpool op before this one was skipped)
                    const Op &ap = ops[&o - &ops[0] - 1];
                    const Tensor &tp = tensors[ap.in];
                    B2_TRY(op_head_forward(tp.val, tp.dt, p_params + o.w_off, o.b_off >= 0 ? p_params + o.b_off : nullptr, (float *)ti.val,
                                           (float *)to.val, tp.N, tp.H * tp.W, tp.C, o.g.Co, s));
                } else
                B2_TRY(op_fc_forward((const float *)ti.val, p_params + o.w_off, o.b_off >= 0 ? p_params + o.b_off : nullptr,
                                     (float *)to.val, o.g.N, o.g.Ci, o.g.Co, o.relu, s));
                prof_end(ph, s);
                if (o.out == logits_tensor && logits)
                    B2_CUDA(cudaMemcpyAsync(logits, to.val, to.numel * 4, cudaMemcpyDeviceToDevice, s));
                break;
            }
            const void *w = o.wdt == DT_BF16 ? (const void *)(p_params_bf16 + o.w_off) : (const void *)(p_params + o.w_off);
            const float *bias = o.b_off >= 0 ? p_params + o.b_off : nullptr;
            int ph = prof_begin(PC_CONV_FPROP, o.use_umma ? o.fplan.flops : 2.0 * o.g.M() * o.g.Co * o.g.K(),
                                ti.numel * (double)dt_size(ti.dt) + to.numel * (double)dt_size(to.dt) + (double)o.g.Co * o.g.K() * dt_size(o.wdt), s);
            if (o.use_umma) B2_TRY(umma_conv_launch(o.fplan, s));
            else B2_TRY(conv_simt_fprop(o.g, ti.val, ti.dt, w, o.wdt, o.wpitch, bias, o.relu, to.val, to.dt, s));
            prof_end(ph, s);
            if (o.out == logits_tensor && logits)
                B2_CUDA(cudaMemcpyAsync(logits, to.val, to.numel * 4, cudaMemcpyDeviceToDevice, s));
            break;
        }
        case OP_BN: {
            Tensor &ti = tensors[o.in];
            int64_t M = (int64_t)ti.N * ti.H * ti.W;
            if (o.bn_fused && o.fwd_in_conv) break;      // done inside the producing conv kernel
            int ph = prof_begin(PC_BN_FWD, 0, ti.numel * (double)dt_size(ti.dt) * (o.res >= 0 ? 4.0 : 3.0), s);
            if (o.bn_fused)
                B2_TRY(op_bn_forward_fused(ti.val, to.val, o.res >= 0 ? tensors[o.res].val : nullptr, ti.dt, M, ti.C, o.fsum,
                                           p_params + o.w_off, p_params + o.b_off, p_buffers ? p_buffers + o.rm_off : nullptr,
                                           p_buffers ? p_buffers + o.rv_off : nullptr, 1e-5f, 0.1f, o.relu, o.bn, s));
            else
            B2_TRY(op_bn_forward(ti.val, to.val, o.res >= 0 ? tensors[o.res].val : nullptr, ti.dt, M, ti.C,
                                 p_params + o.w_off, p_params + o.b_off, p_buffers ? p_buffers + o.rm_off : nullptr,
                                 p_buffers ? p_buffers + o.rv_off : nullptr, 1e-5f, 0.1f, o.relu, o.bn, s));
            prof_end(ph, s);
            break;
        }
        case OP_MAXPOOL: {
            Tensor &ti = tensors[o.in];
            int ph = prof_begin(PC_POOL, 0, (ti.numel + to.numel) * (double)dt_size(ti.dt) + to.numel, s);
            if (ti.dt == DT_BF16 && !o.out_nchw && o.g.Ci % 8 == 0)
                B2_TRY(op_maxpool_forward_bf16x8((const bf16 *)ti.val, (bf16 *)to.val, o.idx, o.g.N, o.g.Hi, o.g.Wi, o.g.Ci, o.g.kh,
                                                 o.g.stride, o.g.pad, o.g.Ho, o.g.Wo, s));
            else
                B2_TRY(op_maxpool_forward(ti.val, to.val, o.idx, ti.dt, o.g.N, o.g.Hi, o.g.Wi, o.g.Ci, o.g.kh, o.g.stride,
                                          o.g.pad, o.g.Ho, o.g.Wo, o.out_nchw, s));
            prof_end(ph, s);
            break;
        }
        case OP_AVGPOOL: {
            if (o.head_fused) break;         // done by the fc op's fused kernel
            Tensor &ti = tensors[o.in];
            B2_TRY(op_avgpool_forward(ti.val, ti.dt, (float *)to.val, ti.N, ti.H * ti.W, ti.C, s));
            break;
        }
        }
    }
    if (zero_side) B2_CUDA(cudaStreamWaitEvent(s, wg_join, 0));
    return 0;
}

int Model::backward(const float *dlogits, cudaStream_t s) {
    B2_CHECK(device >= 0 && slab, B2_ERR_STATE, "model: metadata-only instance (created with device < 0) cannot execute");
    B2_CHECK(p_params && p_grads, B2_ERR_STATE, "model: parameters/gradients not bound (b2_model_bind)");
    // the stem's weight gradient gathers from the input image of the matching forward: part of the graph key
    if (use_graphs && !profiling && eager_runs_bwd >= 2) return run_graphed(bwd_graphs, dlogits, cur_x, false, s);
    ++eager_runs_bwd;
    return backward_impl(dlogits, s);
}

int Model::ensure_side_streams() {
    if (!wg_stream) {
        B2_CUDA(cudaStreamCreateWithFlags(&wg_stream, cudaStreamNonBlocking));
        B2_CUDA(cudaEventCreateWithFlags(&wg_fork, cudaEventDisableTiming));
        B2_CUDA(cudaEventCreateWithFlags(&wg_join, cudaEventDisableTiming));
    }
    return 0;
}

int Model::backward_impl(const float *dlogits, cudaStream_t s) {
    Tensor &tl = tensors[logits_tensor];
    B2_CUDA(cudaMemcpyAsync(tl.grad, dlogits, tl.numel * 4, cudaMemcpyDeviceToDevice, s));
    size_t next_bucket = 0;
    bool forked = false, wg_used = false;
    const bool async_wg = wgrad_async && !profiling;
    if (async_wg) B2_TRY(ensure_side_streams());
    // stream for this op's weight-gradient kernels: forks from `s` at the call (dY is final there)
    auto wgrad_stream = [&]() -> cudaStream_t {
        if (!async_wg) return s;
        cudaEventRecord(wg_fork, s);
        cudaStreamWaitEvent(wg_stream, wg_fork, 0);
        wg_used = true;
        return wg_stream;
    };
    auto flush_buckets = [&](int i) -> int {
        // every bucket whose last gradient was produced by ops[i]: exchange + apply on the side stream
        while (ddp_on && next_bucket < buckets.size() && buckets[next_bucket].after_op >= i) {
            const Bucket &bk = buckets[next_bucket++];
            B2_CUDA(cudaEventRecord(ddp_fork, s));
            B2_CUDA(cudaStreamWaitEvent(ddp_stream, ddp_fork, 0));
            if (wg_used) {                                  // the bucket also needs the async weight gradients
                B2_CUDA(cudaEventRecord(wg_join, wg_stream));
                B2_CUDA(cudaStreamWaitEvent(ddp_stream, wg_join, 0));
            }
            if (ddp_local)
                B2_TRY(op_sgd(p_params + bk.off, p_grads + bk.off, ddp_mom + bk.off, bk.n, ddp_lr, ddp_momentum, 1.f, ddp_first,
                              p_params_bf16 ? p_params_bf16 + bk.off : nullptr, ddp_stream));
            else
                B2_TRY(comm_allreduce_sgd_range(ddp_mom, bk.off, bk.n, ddp_lr, ddp_momentum, ddp_first, ddp_algo,
                                                p_params_bf16, ddp_stream));
            forked = true;
        }
        return 0;
    };
    for (int i = (int)ops.size() - 1; i >= 0; --i) {
        Op &o = ops[i];
        if (o.type == OP_INPUT) { B2_TRY(flush_buckets(i)); continue; }
        Tensor &to = tensors[o.out];
        Tensor &ti = tensors[o.in];
        switch (o.type) {
        case OP_CONV: {
            if (o.stem_umma) {
                int ph = prof_begin(PC_CONV_WGRAD, 2.0 * o.g.M() * o.g.Co * o.g.K(), (double)o.vg.M() * o.kpad * 2 + to.numel * 2.0, s);
                cudaStream_t ws = wgrad_stream();
                if (o.stem_gather) o.wplan.p.stem_x = cur_x;
                B2_TRY(umma_wgrad_launch(o.wplan, ws));
                B2_TRY(op_stem_unpack_wgrad(o.dwpad, p_grads + o.w_off, o.g.Co, o.g.Ci, o.g.kh, o.kpad, ws));
                prof_end(ph, s);
                break;
            }
            if (o.relu) B2_TRY(op_relu_mask(to.val, to.grad, to.dt, to.numel, s));
            if (o.is_fc && o.head_fused) {
                // weight gradient off the critical path; dgrad + average-pool backward in one kernel
                int ph = prof_begin(PC_OTHER, 0, 0, s);
                B2_TRY(op_fc_wgrad((const float *)ti.val, (const float *)to.grad, p_grads + o.w_off, o.b_off >= 0 ? p_grads + o.b_off : nullptr,
                                   o.g.N, o.g.Ci, o.g.Co, wgrad_stream()));
                const Op &ap = ops[i - 1];
                Tensor &tp = tensors[ap.in];
                B2_TRY(op_head_dgrad(p_params + o.w_off, (const float *)to.grad, (float *)ti.grad, tp.grad, tp.dt, tp.N, tp.H * tp.W, tp.C,
                                     o.g.Co, ap.acc_in, s));
                prof_end(ph, s);
                break;
            }
            if (o.is_fc) {
                int ph = prof_begin(PC_OTHER, 0, 0, s);
                B2_TRY(op_fc_backward((const float *)ti.val, p_params + o.w_off, (const float *)to.grad, p_grads + o.w_off,
                                      o.b_off >= 0 ? p_grads + o.b_off : nullptr, o.need_dgrad ? (float *)ti.grad : nullptr,
                                      o.g.N, o.g.Ci, o.g.Co, o.acc_in, s));
                prof_end(ph, s);
                break;
            }
            if (o.b_off >= 0) B2_TRY(op_bias_grad(to.grad, to.dt, o.g.M(), o.g.Co, p_grads + o.b_off, s));
            B2_CHECK(to.dt == ti.dt, B2_ERR_STATE, "conv backward: mixed activation types");
            const double cflops = 2.0 * o.g.M() * o.g.Co * o.g.K();
            const double cbytes = ti.numel * (double)dt_size(ti.dt) + to.numel * (double)dt_size(to.dt);
            int ph = prof_begin(PC_CONV_WGRAD, o.use_umma ? o.wplan.flops : cflops, cbytes + (double)o.g.Co * o.g.K() * 4, s);
            {
                cudaStream_t ws = wgrad_stream();
                if (o.use_umma) B2_TRY(umma_wgrad_launch(o.wplan, ws));
                else B2_TRY(conv_simt_wgrad(o.g, ti.val, to.grad, ti.dt, p_grads + o.w_off, wgrad_ws, wgrad_floats, ws));
            }
            prof_end(ph, s);
            if (o.need_dgrad) {
                const void *w = o.wdt == DT_BF16 ? (const void *)(p_params_bf16 + o.w_off) : (const void *)(p_params + o.w_off);
                ph = prof_begin(PC_CONV_DGRAD, o.use_umma ? o.dplan.flops : cflops, cbytes + (double)o.g.Co * o.g.K() * dt_size(o.wdt), s);
                if (o.use_umma) {
                    if (o.g.stride == 2 && o.g.kh == 1 && !o.acc_in)
                        B2_CUDA(cudaMemsetAsync(ti.grad, 0, ti.numel * dt_size(ti.dt), s));
                    B2_TRY(umma_conv_launch(o.dplan, s));
                } else {
                    B2_TRY(conv_simt_dgrad(o.g, to.grad, to.dt, w, o.wdt, o.wpitch, ti.grad, o.acc_in, s));
                }
                prof_end(ph, s);
            }
            break;
        }
        case OP_BN: {
            int64_t M = (int64_t)ti.N * ti.H * ti.W;
            int ph = prof_begin(PC_BN_BWD, 0, ti.numel * (double)dt_size(ti.dt) * (o.res >= 0 ? 7.0 : 6.0), s);
            if (o.bn_fused)
                B2_TRY(op_bn_backward_fused(ti.val, to.val, to.grad, ti.grad, o.res >= 0 ? tensors[o.res].grad : nullptr, ti.dt, M,
                                            ti.C, o.bsum, p_params + o.w_off, p_grads + o.w_off, p_grads + o.b_off, o.relu,
                                            o.acc_in, o.acc_res, o.bwd_sums_fused, o.bn, s));
            else
            B2_TRY(op_bn_backward(ti.val, to.val, to.grad, ti.grad, o.res >= 0 ? tensors[o.res].grad : nullptr, ti.dt, M,
                                  ti.C, p_params + o.w_off, p_grads + o.w_off, p_grads + o.b_off, o.relu, o.acc_in,
                                  o.acc_res, o.bn, s));
            prof_end(ph, s);
            break;
        }
        case OP_MAXPOOL:
            if (o.need_dgrad) {
                if (ti.dt == DT_BF16 && !o.out_nchw && o.g.Ci % 8 == 0)
                    B2_TRY(op_maxpool_backward_bf16x8((const bf16 *)to.grad, o.idx, (bf16 *)ti.grad, o.g.N, o.g.Hi, o.g.Wi, o.g.Ci,
                                                      o.g.kh, o.g.stride, o.g.pad, o.g.Ho, o.g.Wo, o.acc_in, s));
                else
                    B2_TRY(op_maxpool_backward(to.grad, o.idx, ti.grad, ti.dt, o.g.N, o.g.Hi, o.g.Wi, o.g.Ci, o.g.kh,
                                               o.g.stride, o.g.pad, o.g.Ho, o.g.Wo, o.out_nchw, o.acc_in, s));
            }
            break;
        case OP_AVGPOOL:
            if (o.head_fused) break;         // done by the fc op's fused dgrad kernel
            B2_TRY(op_avgpool_backward((const float *)to.grad, ti.grad, ti.dt, ti.N, ti.H * ti.W, ti.C, o.acc_in, s));
            break;
        }
        B2_TRY(flush_buckets(i));
    }
    if (wg_used) {                                  // join the weight-gradient stream
        B2_CUDA(cudaEventRecord(wg_join, wg_stream));
        B2_CUDA(cudaStreamWaitEvent(s, wg_join, 0));
    }
    if (forked) {                                   // join: the caller's stream continues after the exchange
        B2_CUDA(cudaEventRecord(ddp_join, ddp_stream));
        B2_CUDA(cudaStreamWaitEvent(s, ddp_join, 0));
    }
    return 0;
}


int Model::create(int kind, int batch, int num_classes, int precision, int device, Model **out) {
    B2_CHECK(batch > 0 && num_classes > 0, B2_ERR_ARG, "model: batch and num_classes must be positive");
    B2_CHECK(precision == B2_PREC_FP32 || precision == B2_PREC_BF16, B2_ERR_ARG, "model: unknown precision %d", precision);
    if (device >= 0) B2_CUDA(cudaSetDevice(device));
    Model *m = new Model();
    m->kind = kind; m->batch = batch; m->num_classes = num_classes; m->precision = precision; m->device = device;
    if (kind == B2_MODEL_NET) { m->precision = B2_PREC_FP32; m->build_net(); }
    else if (kind == B2_MODEL_RESNET18) { const int l[4] = {2, 2, 2, 2}; m->build_resnet(false, l); }
    else if (kind == B2_MODEL_RESNET50) { const int l[4] = {3, 4, 6, 3}; m->build_resnet(true, l); }
    else { delete m; set_error("model: unknown kind %d", kind); return B2_ERR_ARG; }
    int r = m->finalize();
    if (r != 0) { delete m; return r; }
    *out = m;
    return 0;
}

}  // namespace b2
