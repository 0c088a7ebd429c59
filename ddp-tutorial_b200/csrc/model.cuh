// Static training graph: tensors, ops, parameter table.
#pragma once
#include <string>
#include <vector>
#include "ops.cuh"
#include "conv_umma.cuh"

namespace b2 {

enum OpType { OP_INPUT = 0, OP_CONV = 1, OP_BN = 2, OP_MAXPOOL = 3, OP_AVGPOOL = 4 };

struct Tensor {
    int N, H, W, C;      // NHWC (or [N,1,1,features])
    int dt;              // DT_F32 / DT_BF16
    int64_t numel;
    void *val, *grad;
};

struct Op {
    int type;
    int in, out, res;            // tensor ids (res: residual added before the ReLU of a BN op)
    ConvGeom g;                  // conv / pool geometry
    int64_t w_off, b_off;        // conv: weight/bias; bn: gamma/beta (offsets into the flat param vector)
    int64_t rm_off, rv_off;      // bn: running mean / var (offsets into the flat buffer vector)
    bool relu, out_nchw;
    bool need_dgrad, acc_in, acc_res;
    BnWork bn;
    bool bn_fused;               // statistics come from the producing conv's epilogue
    float *fsum, *bsum;          // [copies][2][C] each, inside the per-step zeroed arena
    bool fwd_in_conv;            // forward done inside the producing conv kernel (grid-barrier fusion)
    unsigned int *fb_counter;    // [0] arrivals, [1] generation of the fused conv+BN grid barrier
    bool bwd_sums_fused;         // the dgrad that finalises this BN's output gradient also accumulates bsum
    uint8_t *idx;                // maxpool argmax
    int wdt, wpitch;             // conv: weight operand type / row pitch as the kernels see it
    bool use_umma;               // tcgen05 path (bf16, channel counts multiples of 64)
    bool is_fc;                  // small fp32 classifier head -> dedicated kernels
    bool head_fused;             // avgpool + fc pair executed as one forward kernel / one dgrad kernel (set on both ops)
    // stem (Ci = 3): bf16 im2col rows + 1x1 tcgen05 GEMM over them
    bool stem_umma;
    bool stem_gather;            // ... with the im2col rows gathered in shared memory (no HBM im2col buffer)
    ConvGeom vg;                 // the virtual 1x1 conv over the im2col tensor [N,Ho,Wo,kpad]
    int kpad;
    bf16 *col, *wpad;            // [M][kpad], [Co][kpad]
    float *dwpad;                // [Co][kpad]
    UmmaConvPlan fplan, dplan;
    UmmaWgradPlan wplan;
};

// per-launch-group profiling (bench.py's roofline leg): CUDA events around every op
enum ProfCat { PC_CONV_FPROP = 0, PC_CONV_DGRAD, PC_CONV_WGRAD, PC_BN_FWD, PC_BN_BWD, PC_POOL, PC_OTHER, PC_COUNT };
struct ProfRec { cudaEvent_t a, b; int cat; double flops, bytes; };

struct ParamInfo { std::string name; int64_t offset, numel; int rank; int64_t shape[4], stride[4]; };
struct BufferInfo { std::string name; int64_t offset, numel; };

struct Model {
    int kind = 0, batch = 0, num_classes = 0, precision = 0, device = 0;
    std::vector<Tensor> tensors;
    std::vector<Op> ops;
    std::vector<ParamInfo> params;
    std::vector<BufferInfo> buffers;
    int64_t n_params = 0, n_buffers = 0;
    int input_tensor = -1, logits_tensor = -1;
    const float *cur_x = nullptr;    // the caller's NCHW input of the forward being enqueued
    float *p_params = nullptr, *p_grads = nullptr, *p_buffers = nullptr;
    bf16 *p_params_bf16 = nullptr;   // bf16 mirror of the whole flat parameter vector (same offsets)
    bf16 *own_bf16 = nullptr;        // model-owned default for the mirror
    void *slab = nullptr; size_t slab_bytes = 0;
    float *wgrad_ws = nullptr; size_t wgrad_floats = 0;
    int *fb_err = nullptr;           // set by a fused conv+BN kernel whose grid barrier timed out
    float *bn_sums = nullptr; size_t bn_sums_bytes = 0;    // all fused-BN sum buffers, zeroed once per step
    bool weights_dirty = true;
    bool plans_ready = false;
    int conv_impl = 1;           // 0 = SIMT everywhere, 1 = tcgen05 where eligible
    // split-K weight gradients add into the flat gradient vector with red.global.add (no partials, no reduce
    // kernel); the vector is then zeroed once per step, on the weight-gradient stream while forward runs.
    // deterministic = true restores the fixed-order partials + reduce form (bit-reproducible run to run).
    bool deterministic = false;
    bool zero_grads_each_step = false;
    int ensure_side_streams();
    int build_plans();
    // CUDA-graph replay of forward / backward (launch-latency removal): one graph per distinct
    // (input, output) pointer pair, captured on a private stream, replayed on the caller's
    struct GraphEntry { const void *a; const void *b; cudaGraphExec_t exec; int64_t launches; };
    bool use_graphs = true;
    int eager_runs_fwd = 0, eager_runs_bwd = 0;
    cudaStream_t cap_stream = nullptr;
    std::vector<GraphEntry> fwd_graphs, bwd_graphs;
    int forward_impl(const float *x_nchw, float *logits, cudaStream_t s);
    int backward_impl(const float *dlogits, cudaStream_t s);
    int run_graphed(std::vector<GraphEntry> &cache, const void *a, const void *b, bool is_fwd, cudaStream_t s);
    void drop_graphs();
    // DDP inside backward: gradient buckets (contiguous suffixes of the flat vector, in the
    // order backward completes them) exchanged + applied on a side stream while backward runs
    struct Bucket { int64_t off, n; int after_op; };   // launch after ops[after_op] has run backward
    std::vector<Bucket> buckets;
    bool ddp_on = false;
    bool ddp_local = false;          // world size 1: the bucketed update is the plain SGD kernel (no exchange)
    float *ddp_mom = nullptr; float ddp_lr = 0.f, ddp_momentum = 0.f; int ddp_algo = 0, ddp_first = 1;
    // weight gradients are off the backward critical path (only the optimizer consumes them):
    // they run on their own stream, forked after each layer's dY is final, joined at the end
    cudaStream_t wg_stream = nullptr;
    cudaEvent_t wg_fork = nullptr, wg_join = nullptr;
    bool wgrad_async = true;
    cudaStream_t ddp_stream = nullptr;
    cudaEvent_t ddp_fork = nullptr, ddp_join = nullptr;
    int set_ddp(bool enable, float *mom, float lr, float momentum, int algo, int64_t bucket_elems, bool local);
    void compute_buckets(int64_t bucket_elems);
    int64_t steps_tracked = 0;
    bool profiling = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    int prof_begin(int cat, double flops, double bytes, cudaStream_t s);
    void prof_end(int h, cudaStream_t s);
    int prof_collect(double ms[PC_COUNT], double flops[PC_COUNT], double bytes[PC_COUNT], int launches[PC_COUNT]);
    // per-class (fprop, dgrad, wgrad) average kernel duration of the tcgen05 conv launches of one
    // step: every plan is launched `iters` times back to back between two events
    int time_conv_kernels(int iters, double ms[3], double flops[3], int launches[3], cudaStream_t s);

    static int create(int kind, int batch, int num_classes, int precision, int device, Model **out);
    ~Model();
    int forward(const float *x_nchw, float *logits, cudaStream_t s);
    int backward(const float *dlogits, cudaStream_t s);
    int sync_weights(cudaStream_t s);

    // construction helpers
    int add_tensor(int N, int H, int W, int C, int dt);
    int64_t add_param(const std::string &name, int rank, const int64_t *shape, const int64_t *stride, int64_t numel);
    int64_t add_buffer(const std::string &name, int64_t numel);
    int64_t add_conv_weight(const std::string &name, int Co, int Ci, int kh, int kw);
    int64_t add_vec(const std::string &name, int n);
    int64_t add_mat(const std::string &name, int rows, int cols);
    int op_conv(int in, int64_t w_off, int64_t b_off, int Co, int k, int stride, int pad, bool relu, int out_dt);
    int op_bn(int in, int64_t g_off, int64_t b_off, int64_t rm_off, int64_t rv_off, bool relu, int res);
    int op_maxpool(int in, int k, int stride, int pad, bool out_nchw_flat);
    int op_avgpool(int in);
    void build_net();
    void build_resnet(bool bottleneck, const int layers[4]);
    int finalize();
    bool ops_produce(int tensor) const;
};

}  // namespace b2
