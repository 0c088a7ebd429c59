// Launcher prototypes for the non-GEMM kernels and the SIMT (CUDA-core) conv path.
#pragma once
#include "common.cuh"

namespace b2 {

// dt: activation storage type, 0 = fp32, 1 = bf16
enum { DT_F32 = 0, DT_BF16 = 1 };
static inline size_t dt_size(int dt) { return dt == DT_F32 ? 4 : 2; }

// ---- layout ------------------------------------------------------------------
int op_nchw_to_nhwc(const float *x, void *y, int dt, int N, int C, int H, int W, cudaStream_t s);
int op_nhwc_to_nchw_f32(const void *x, int dt, float *y, int N, int C, int H, int W, cudaStream_t s);
int op_f32_to_bf16(const float *x, bf16 *y, int64_t n, cudaStream_t s);

// ---- batch norm (training mode, batch statistics) ---------------------------------
struct BnWork {           // per-BN-layer scratch, all fp32 device pointers
    float *partial;       // [bn_rows_blocks][2][C]
    float *mean, *invstd; // [C]
    float *coef;          // [3][C]  (a, c1, c2) for backward
    int row_blocks;
};
int bn_row_blocks(int64_t M, int C);
int op_bn_forward(const void *x, void *y, const void *res, int dt, int64_t M, int C,
                  const float *gamma, const float *beta, float *running_mean, float *running_var,
                  float eps, float momentum, bool relu, const BnWork &w, cudaStream_t s);
// dy: grad wrt y (post relu/residual).  dx written (acc_dx=false) or accumulated.
// dres (may be null): gradient of the residual input, written or accumulated.
int op_bn_backward(const void *x, const void *y, const void *dy, void *dx, void *dres, int dt,
                   int64_t M, int C, const float *gamma, float *dgamma, float *dbeta, bool relu,
                   bool acc_dx, bool acc_dres, const BnWork &w, cudaStream_t s);

// Fused variants (bf16 tcgen05 path): the producing conv's epilogue already accumulated the
// per-channel sum / sum of squares into fsum[2][C] (atomics, zeroed once per step), so forward is
// ONE kernel (finalize per block in shared memory + apply); backward is a reduce (atomics into
// bsum[2][C]) plus an apply that derives its coefficients per block.
int op_bn_forward_fused(const void *x, void *y, const void *res, int dt, int64_t M, int C, const float *fsum,
                        const float *gamma, const float *beta, float *running_mean, float *running_var, float eps,
                        float momentum, bool relu, const BnWork &w, cudaStream_t s);
// sums_ready: bsum was already accumulated by the dgrad that produced dy (skip the reduce pass)
int op_bn_backward_fused(const void *x, const void *y, const void *dy, void *dx, void *dres, int dt, int64_t M, int C,
                         float *bsum, const float *gamma, float *dgamma, float *dbeta, bool relu, bool acc_dx,
                         bool acc_dres, bool sums_ready, const BnWork &w, cudaStream_t s);

// ---- pooling -------------------------------------------------------------------
int op_maxpool_forward(const void *x, void *y, uint8_t *idx, int dt, int N, int Hi, int Wi, int C,
                       int k, int stride, int pad, int Ho, int Wo, bool out_nchw, cudaStream_t s);
int op_maxpool_backward(const void *dy, const uint8_t *idx, void *dx, int dt, int N, int Hi, int Wi,
                        int C, int k, int stride, int pad, int Ho, int Wo, bool out_nchw, bool acc,
                        cudaStream_t s);
// x/dx are NHWC of type xdt; y/dy are fp32 [N][C] (the classifier head runs in fp32)
int op_avgpool_forward(const void *x, int xdt, float *y, int N, int HW, int C, cudaStream_t s);
int op_avgpool_backward(const float *dy, void *dx, int xdt, int N, int HW, int C, bool acc, cudaStream_t s);

// 8-channel vectorised bf16 max pooling (C % 8 == 0, NHWC in and out)
int op_maxpool_forward_bf16x8(const bf16 *x, bf16 *y, uint8_t *idx, int N, int Hi, int Wi, int C, int k, int stride,
                              int pad, int Ho, int Wo, cudaStream_t s);
int op_maxpool_backward_bf16x8(const bf16 *dy, const uint8_t *idx, bf16 *dx, int N, int Hi, int Wi, int C, int k,
                               int stride, int pad, int Ho, int Wo, bool acc, cudaStream_t s);

// ---- stem: fp32 NCHW image -> bf16 im2col rows [N*Ho*Wo][kpad] for a k x k / stride / pad conv,
// column order (kh, kw, c) = the O,H,W,I weight order; columns >= k*k*C are zero
int op_stem_im2col(const float *x_nchw, bf16 *col, int N, int C, int H, int W, int k, int stride, int pad, int Ho,
                   int Wo, int kpad, cudaStream_t s);
// The stem's GEMM-K order is (c, kh, kw padded to 8): one 16-byte im2col chunk = 8 consecutive input
// pixels of one (channel, filter row).  pack: fp32 O,H,W,I master -> bf16 [Co][kpad] in that order;
// unpack: fp32 dW [Co][kpad] in that order -> O,H,W,I gradient rows.
int op_stem_pack_weights(const float *master, bf16 *wpad, int Co, int C, int k, int kpad, cudaStream_t s);
int op_stem_unpack_wgrad(const float *dwpad, float *grad, int Co, int C, int k, int kpad, cudaStream_t s);

// ---- small fully-connected head in fp32 (classifier: N <= 32 outputs) ------------------------
int op_fc_forward(const float *x, const float *w, const float *bias, float *y, int B, int K, int N, bool relu, cudaStream_t s);
// dw [N][K], db [N] (may be null), dx [B][K] (may be null; acc adds)
int op_fc_backward(const float *x, const float *w, const float *dy, float *dw, float *db, float *dx, int B, int K, int N,
                   bool acc_dx, cudaStream_t s);

// fused head: feat = mean over HW of x (written, fp32 [B][C]); y = feat W^T + bias  /  dfeat = dy W (written), dx (+)= dfeat / HW
int op_head_forward(const void *x, int xdt, const float *w, const float *bias, float *feat, float *y, int B, int HW, int C, int N,
                    cudaStream_t s);
int op_head_dgrad(const float *w, const float *dy, float *dfeat, void *dx, int xdt, int B, int HW, int C, int N, bool acc, cudaStream_t s);
int op_fc_wgrad(const float *x, const float *dy, float *dw, float *db, int B, int K, int N, cudaStream_t s);

// ---- pointwise ------------------------------------------------------------------
// dy = y > 0 ? dy : 0  (in place)
int op_relu_mask(const void *y, void *dy, int dt, int64_t n, cudaStream_t s);
// db[c] = sum_m dy[m,c]
int op_bias_grad(const void *dy, int dt, int64_t M, int C, float *db, cudaStream_t s);

// ---- SIMT implicit-GEMM convolution (fp32 FMA, any shape) ---------------------------
// weights are O,H,W,I (K = kh*kw*Ci contiguous per output channel) with row pitch wpitch
// elements; wdt is the weight storage type.
int conv_simt_fprop(const ConvGeom &g, const void *x, int dt, const void *w, int wdt, int wpitch,
                    const float *bias, bool relu, void *y, int ydt, cudaStream_t s);
int conv_simt_dgrad(const ConvGeom &g, const void *dy, int dt, const void *w, int wdt, int wpitch,
                    void *dx, bool acc, cudaStream_t s);
// dw fp32 [Co][K] (pitch K); deterministic split over the pixel dimension through `partial`
// (capacity in floats); returns the number of splits it used
size_t conv_simt_wgrad_workspace(const ConvGeom &g);
int conv_simt_wgrad(const ConvGeom &g, const void *x, const void *dy, int dt, float *dw,
                    float *partial, size_t partial_floats, cudaStream_t s);

// ---- loss / optimizer ------------------------------------------------------------------
int op_ce_loss_grad(const float *logits, const int64_t *labels, int B, int C, float *loss,
                    float *dlogits, float *running, cudaStream_t s);
// pb (optional): bf16 mirror of p, written in the same pass
int op_sgd(float *p, const float *g, float *m, int64_t n, float lr, float momentum, float gscale,
           int first, bf16 *pb, cudaStream_t s);

// ---- input pipeline / eval -------------------------------------------
// uint8 HWC images [N][H][W][C] (CIFAR's on-disk layout) -> fp32 NCHW, ToTensor + Normalize(mean, std)
// per channel: y = (x/255 - mean[c]) / std[c]   (cifar_example_ddp.py:61-63)
int op_u8_to_normalized(const uint8_t *x, float *y, int N, int H, int W, int C, const float *mean3, const float *std3, cudaStream_t s);
// counts[0] += #(argmax(logits[b]) == labels[b]), counts[1] += B   (first maximum wins, like torch.max)
int op_argmax_count(const float *logits, const int64_t *labels, int B, int C, long long *counts, cudaStream_t s);

// ---- communicator hooks used by the model's bucketed backward (comm.cu) ------------------
bool comm_ready();
int comm_allreduce_sgd_range(float *m, int64_t off, int64_t n, float lr, float momentum, int first, int algo, bf16 *mirror,
                             cudaStream_t s);

}  // namespace b2
