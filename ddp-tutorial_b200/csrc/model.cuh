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
    uint8_t *idx;                // maxpool argmax
    int wdt, wpitch;             // conv: weight operand type / row pitch as the kernels see it
    bf16 *wcopy;                 // bf16 mode: compute copy of the weights [Co][wpitch]
    bool use_umma;               // tcgen05 path (bf16, channel counts multiples of 64)
    bf16 *wt;                    // tcgen05 dgrad operand: [Ci][taps*Co]
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
    float *p_params = nullptr, *p_grads = nullptr, *p_buffers = nullptr;
    void *slab = nullptr; size_t slab_bytes = 0;
    float *wgrad_ws = nullptr; size_t wgrad_floats = 0;
    bool weights_dirty = true;
    bool plans_ready = false;
    int conv_impl = 1;           // 0 = SIMT everywhere, 1 = tcgen05 where eligible
    int build_plans();
    int64_t steps_tracked = 0;
    bool profiling = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    int prof_begin(int cat, double flops, double bytes, cudaStream_t s);
    void prof_end(int h, cudaStream_t s);
    int prof_collect(double ms[PC_COUNT], double flops[PC_COUNT], double bytes[PC_COUNT], int launches[PC_COUNT]);

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
