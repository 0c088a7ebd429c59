// HBM-bound kernels of the training step: layout conversion, BatchNorm (train mode),
// pooling, ReLU masks, bias gradients, cross-entropy + softmax-grad, SGD-momentum.
//
// Reference semantics (what torch dispatches for the cited reference lines):
//   BatchNorm2d train fwd/bwd : torchvision resnet.py:82,198 under cifar_example.py:77-79
//   max_pool2d                : cifar_example_ddp.py:27,34-35
//   cross_entropy (mean)      : cifar_example_ddp.py:85,105  (functional.py:3421-3504)
//   SGD momentum              : cifar_example_ddp.py:86,107  (sgd.py:_single_tensor_sgd)
#include <stdlib.h>
#define B2_TRACE_TU_ID 4
#include "ops.cuh"

namespace b2 {

#define DISPATCH_DT(dt, T, ...)                    \
    if ((dt) == DT_F32) { typedef float T; __VA_ARGS__; } \
    else { typedef bf16 T; __VA_ARGS__; }

static inline int grid_for(int64_t n, int block, int max_blocks = kNumSMs * 16) {
    int64_t b = ceil_div(n, block);
    if (b > max_blocks) b = max_blocks;
    if (b < 1) b = 1;
    return (int)b;
}

// ---------------------------------------------------------------------------------
// layout
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void k_nchw_to_nhwc(const float *__restrict__ x, T *__restrict__ y, int N, int C, int H, int W) {
    pdl_sync(); B2_TRACE_BEGIN();
    int64_t total = (int64_t)N * C * H * W;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = i % C; int64_t r = i / C;
        int w = r % W; r /= W;
        int h = r % H; int n = r / H;
        y[i] = from_f<T>(x[(((int64_t)n * C + c) * H + h) * W + w]);
    }
}
int op_nchw_to_nhwc(const float *x, void *y, int dt, int N, int C, int H, int W, cudaStream_t s) {
    int64_t total = (int64_t)N * C * H * W;
    DISPATCH_DT(dt, T, B2_LAUNCH((k_nchw_to_nhwc<T>), grid_for(total, 256), 256, 0, s, x, (T *)y, N, C, H, W));
    return 0;
}
template <typename T>
__global__ void k_nhwc_to_nchw(const T *__restrict__ x, float *__restrict__ y, int N, int C, int H, int W) {
    pdl_sync(); B2_TRACE_BEGIN();
    int64_t total = (int64_t)N * C * H * W;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int w = i % W; int64_t r = i / W;
        int h = r % H; r /= H;
        int c = r % C; int n = r / C;
        y[i] = to_f<T>(x[(((int64_t)n * H + h) * W + w) * C + c]);
    }
}
int op_nhwc_to_nchw_f32(const void *x, int dt, float *y, int N, int C, int H, int W, cudaStream_t s) {
    int64_t total = (int64_t)N * C * H * W;
    DISPATCH_DT(dt, T, B2_LAUNCH((k_nhwc_to_nchw<T>), grid_for(total, 256), 256, 0, s, (const T *)x, y, N, C, H, W));
    return 0;
}
__global__ void k_f32_to_bf16(const float *__restrict__ x, bf16 *__restrict__ y, int64_t n) {
    pdl_sync(); B2_TRACE_BEGIN();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = __float2bfloat16_rn(x[i]);
}
int op_f32_to_bf16(const float *x, bf16 *y, int64_t n, cudaStream_t s) {
    B2_LAUNCH(k_f32_to_bf16, grid_for(n, 256), 256, 0, s, x, y, n);
    return 0;
}
// ---------------------------------------------------------------------------------
// BatchNorm.  x is [M][C] (NHWC flattened); per-channel statistics over M rows.
// Block = 32 channels x 8 row lanes; grid = (C/32, row_blocks).
// ---------------------------------------------------------------------------------
int bn_row_blocks(int64_t M, int C) {
    int gx = (C + 31) / 32;
    int64_t want = (kNumSMs * 4 + gx - 1) / gx;
    int64_t maxb = ceil_div(M, 8 * 4);   // at least 4 rows per thread
    if (want > maxb) want = maxb;
    if (want < 1) want = 1;
    return (int)want;
}

template <typename T, bool BWD>
__global__ void __launch_bounds__(256)
k_bn_reduce(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ dy,
            const float *__restrict__ mean, const float *__restrict__ invstd, bool relu,
            int64_t M, int C, float *__restrict__ partial) {
    pdl_sync(); B2_TRACE_BEGIN();
    // FWD: partial = {sum x, sum x^2};  BWD: partial = {sum g, sum g*xhat}
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const int64_t rows_per_block = (M + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = blockIdx.y * rows_per_block;
    int64_t r1 = r0 + rows_per_block; if (r1 > M) r1 = M;
    float s0 = 0.f, s1 = 0.f;
    if (c < C) {
        float mu = 0.f, is = 0.f;
        if (BWD) { mu = mean[c]; is = invstd[c]; }
        for (int64_t r = r0 + ty; r < r1; r += 8) {
            float v = to_f<T>(x[r * C + c]);
            if (!BWD) { s0 += v; s1 += v * v; }
            else {
                float g = to_f<T>(dy[r * C + c]);
                if (relu && !(to_f<T>(y[r * C + c]) > 0.f)) g = 0.f;
                s0 += g; s1 += g * ((v - mu) * is);
            }
        }
    }
    __shared__ float sh0[8][33], sh1[8][33];
    sh0[ty][tx] = s0; sh1[ty][tx] = s1;
    __syncthreads();
    if (ty == 0 && c < C) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { a += sh0[j][tx]; b += sh1[j][tx]; }
        partial[((int64_t)blockIdx.y * 2 + 0) * C + c] = a;
        partial[((int64_t)blockIdx.y * 2 + 1) * C + c] = b;
    }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per channel: lanes stride over the row-block partials, fp64 combine
__global__ void __launch_bounds__(128)
k_bn_finalize_fwd(const float *__restrict__ partial, int nblk, int64_t M, int C, float eps,
                  float momentum, float *__restrict__ mean, float *__restrict__ invstd,
                  float *running_mean, float *running_var) {
    pdl_sync(); B2_TRACE_BEGIN();
    const int c = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    double s0 = 0.0, s1 = 0.0;
    for (int b = lane; b < nblk; b += 32) { s0 += partial[((int64_t)b * 2) * C + c]; s1 += partial[((int64_t)b * 2 + 1) * C + c]; }
    s0 = warp_sum_d(s0); s1 = warp_sum_d(s1);
    if (lane != 0) return;
    double mu = s0 / (double)M;
    double var = s1 / (double)M - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)mu;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
        double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
k_bn_apply(const T *__restrict__ x, T *__restrict__ y, const T *__restrict__ res,
           const float *__restrict__ mean, const float *__restrict__ invstd,
           const float *__restrict__ gamma, const float *__restrict__ beta, bool relu, int64_t total, int C) {
    pdl_sync(); B2_TRACE_BEGIN();
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * V; i < total; i += (int64_t)gridDim.x * blockDim.x * V) {
        int c0 = i % C;
        __align__(16) T xv[V], rv[V], ov[V];
        *(uint4 *)xv = *(const uint4 *)(x + i);
        if (res) *(uint4 *)rv = *(const uint4 *)(res + i);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            int c = c0 + j;
            float v = ((to_f<T>(xv[j]) - mean[c]) * invstd[c]) * gamma[c] + beta[c];
            if (res) v += to_f<T>(rv[j]);
            if (relu) v = v > 0.f ? v : 0.f;
            ov[j] = from_f<T>(v);
        }
        *(uint4 *)(y + i) = *(uint4 *)ov;
    }
}

int op_bn_forward(const void *x, void *y, const void *res, int dt, int64_t M, int C, const float *gamma,
                  const float *beta, float *running_mean, float *running_var, float eps, float momentum,
                  bool relu, const BnWork &w, cudaStream_t s) {
    B2_CHECK(C % 8 == 0, B2_ERR_ARG, "bn: C=%d must be a multiple of 8", C);
    dim3 grid((C + 31) / 32, w.row_blocks);
    DISPATCH_DT(dt, T, B2_LAUNCH((k_bn_reduce<T, false>), grid, 256, 0, s, (const T *)x, (const T *)nullptr,
                                 (const T *)nullptr, (const float *)nullptr, (const float *)nullptr, false, M, C, w.partial));
    B2_LAUNCH(k_bn_finalize_fwd, (C + 3) / 4, 128, 0, s, w.partial, w.row_blocks, M, C, eps, momentum, w.mean,
              w.invstd, running_mean, running_var);
    int64_t total = M * C;
    if (dt == DT_F32) {
        B2_LAUNCH((k_bn_apply<float, 4>), grid_for(total / 4, 256), 256, 0, s, (const float *)x, (float *)y,
                  (const float *)res, w.mean, w.invstd, gamma, beta, relu, total, C);
    } else {
        B2_LAUNCH((k_bn_apply<bf16, 8>), grid_for(total / 8, 256), 256, 0, s, (const bf16 *)x, (bf16 *)y,
                  (const bf16 *)res, w.mean, w.invstd, gamma, beta, relu, total, C);
    }
    return 0;
}

__global__ void __launch_bounds__(128)
k_bn_finalize_bwd(const float *__restrict__ partial, int nblk, int64_t M, int C,
                  const float *__restrict__ gamma, const float *__restrict__ invstd,
                  float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ coef) {
    pdl_sync(); B2_TRACE_BEGIN();
    const int c = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    double s0 = 0.0, s1 = 0.0;
    for (int b = lane; b < nblk; b += 32) { s0 += partial[((int64_t)b * 2) * C + c]; s1 += partial[((int64_t)b * 2 + 1) * C + c]; }
    s0 = warp_sum_d(s0); s1 = warp_sum_d(s1);
    if (lane != 0) return;
    dbeta[c] = (float)s0;
    dgamma[c] = (float)s1;
    coef[c] = gamma[c] * invstd[c];
    coef[C + c] = (float)(s0 / (double)M);
    coef[2 * C + c] = (float)(s1 / (double)M);
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
k_bn_bwd_apply(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ dy, T *__restrict__ dx,
               T *__restrict__ dres, const float *__restrict__ mean, const float *__restrict__ invstd,
               const float *__restrict__ coef, bool relu, bool acc_dx, bool acc_dres, int64_t total, int C) {
    pdl_sync(); B2_TRACE_BEGIN();
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * V; i < total; i += (int64_t)gridDim.x * blockDim.x * V) {
        int c0 = i % C;
        __align__(16) T xv[V], yv[V], gv[V], ov[V], rv[V], pv[V];
        *(uint4 *)xv = *(const uint4 *)(x + i);
        *(uint4 *)gv = *(const uint4 *)(dy + i);
        if (relu) *(uint4 *)yv = *(const uint4 *)(y + i);
        if (acc_dx) *(uint4 *)pv = *(const uint4 *)(dx + i);
        if (dres && acc_dres) *(uint4 *)rv = *(const uint4 *)(dres + i);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            int c = c0 + j;
            float g = to_f<T>(gv[j]);
            if (relu && !(to_f<T>(yv[j]) > 0.f)) g = 0.f;
            float xh = (to_f<T>(xv[j]) - mean[c]) * invstd[c];
            float d = coef[c] * (g - coef[C + c] - xh * coef[2 * C + c]);
            if (acc_dx) d += to_f<T>(pv[j]);
            ov[j] = from_f<T>(d);
            if (dres) {
                float r = g;
                if (acc_dres) r += to_f<T>(rv[j]);
                rv[j] = from_f<T>(r);
            }
        }
        *(uint4 *)(dx + i) = *(uint4 *)ov;
        if (dres) *(uint4 *)(dres + i) = *(uint4 *)rv;
    }
}

int op_bn_backward(const void *x, const void *y, const void *dy, void *dx, void *dres, int dt, int64_t M, int C,
                   const float *gamma, float *dgamma, float *dbeta, bool relu, bool acc_dx, bool acc_dres,
                   const BnWork &w, cudaStream_t s) {
    dim3 grid((C + 31) / 32, w.row_blocks);
    DISPATCH_DT(dt, T, B2_LAUNCH((k_bn_reduce<T, true>), grid, 256, 0, s, (const T *)x, (const T *)y, (const T *)dy,
                                 w.mean, w.invstd, relu, M, C, w.partial));
    B2_LAUNCH(k_bn_finalize_bwd, (C + 3) / 4, 128, 0, s, w.partial, w.row_blocks, M, C, gamma, w.invstd, dgamma,
              dbeta, w.coef);
    int64_t total = M * C;
    if (dt == DT_F32) {
        B2_LAUNCH((k_bn_bwd_apply<float, 4>), grid_for(total / 4, 256), 256, 0, s, (const float *)x, (const float *)y,
                  (const float *)dy, (float *)dx, (float *)dres, w.mean, w.invstd, w.coef, relu, acc_dx, acc_dres, total, C);
    } else {
        B2_LAUNCH((k_bn_bwd_apply<bf16, 8>), grid_for(total / 8, 256), 256, 0, s, (const bf16 *)x, (const bf16 *)y,
                  (const bf16 *)dy, (bf16 *)dx, (bf16 *)dres, w.mean, w.invstd, w.coef, relu, acc_dx, acc_dres, total, C);
    }
    return 0;
}

// ---------------------------------------------------------------------------------
// pooling
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void k_maxpool_fwd(const T *__restrict__ x, T *__restrict__ y, uint8_t *__restrict__ idx, int N, int Hi,
                              int Wi, int C, int k, int stride, int pad, int Ho, int Wo, bool out_nchw) {
    pdl_sync(); B2_TRACE_BEGIN();
    int64_t total = (int64_t)N * Ho * Wo * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = i % C; int64_t r = i / C;
        int ow = r % Wo; r /= Wo;
        int oh = r % Ho; int n = r / Ho;
        float best = -INFINITY; int bi = 255;
        for (int a = 0; a < k; ++a) {
            int hi = oh * stride - pad + a;
            if (hi < 0 || hi >= Hi) continue;
            for (int b = 0; b < k; ++b) {
                int wi = ow * stride - pad + b;
                if (wi < 0 || wi >= Wi) continue;
                float v = to_f<T>(x[(((int64_t)n * Hi + hi) * Wi + wi) * C + c]);
                if (v > best || bi == 255) { best = v; bi = a * k + b; }   // first maximum wins (torch)
            }
        }
        int64_t o = out_nchw ? (((int64_t)n * C + c) * Ho + oh) * Wo + ow : i;
        y[o] = from_f<T>(best);
        idx[i] = (uint8_t)bi;
    }
}
int op_maxpool_forward(const void *x, void *y, uint8_t *idx, int dt, int N, int Hi, int Wi, int C, int k, int stride,
                       int pad, int Ho, int Wo, bool out_nchw, cudaStream_t s) {
    int64_t total = (int64_t)N * Ho * Wo * C;
    DISPATCH_DT(dt, T, B2_LAUNCH((k_maxpool_fwd<T>), grid_for(total, 256), 256, 0, s, (const T *)x, (T *)y, idx, N, Hi,
                                 Wi, C, k, stride, pad, Ho, Wo, out_nchw));
    return 0;
}
template <typename T>
__global__ void k_maxpool_bwd(const T *__restrict__ dy, const uint8_t *__restrict__ idx, T *__restrict__ dx, int N,
                              int Hi, int Wi, int C, int k, int stride, int pad, int Ho, int Wo, bool out_nchw, bool acc) {
    pdl_sync(); B2_TRACE_BEGIN();
    int64_t total = (int64_t)N * Hi * Wi * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = i % C; int64_t r = i / C;
        int wi = r % Wi; r /= Wi;
        int hi = r % Hi; int n = r / Hi;
        float g = 0.f;
        // windows (oh,ow) with oh*stride - pad <= hi <= oh*stride - pad + k - 1
        int oh_lo = (hi + pad - k + 1 + stride - 1); oh_lo = oh_lo < 0 ? 0 : oh_lo / stride;
        int oh_hi = (hi + pad) / stride; if (oh_hi > Ho - 1) oh_hi = Ho - 1;
        int ow_lo = (wi + pad - k + 1 + stride - 1); ow_lo = ow_lo < 0 ? 0 : ow_lo / stride;
        int ow_hi = (wi + pad) / stride; if (ow_hi > Wo - 1) ow_hi = Wo - 1;
        for (int oh = oh_lo; oh <= oh_hi; ++oh)
            for (int ow = ow_lo; ow <= ow_hi; ++ow) {
                int a = hi - (oh * stride - pad), b = wi - (ow * stride - pad);
                int64_t oi = (((int64_t)n * Ho + oh) * Wo + ow) * C + c;
                if (idx[oi] == (uint8_t)(a * k + b)) {
                    int64_t o = out_nchw ? (((int64_t)n * C + c) * Ho + oh) * Wo + ow : oi;
                    g += to_f<T>(dy[o]);
                }
            }
        if (acc) g += to_f<T>(dx[i]);
        dx[i] = from_f<T>(g);
    }
}
int op_maxpool_backward(const void *dy, const uint8_t *idx, void *dx, int dt, int N, int Hi, int Wi, int C, int k,
                        int stride, int pad, int Ho, int Wo, bool out_nchw, bool acc, cudaStream_t s) {
    int64_t total = (int64_t)N * Hi * Wi * C;
    DISPATCH_DT(dt, T, B2_LAUNCH((k_maxpool_bwd<T>), grid_for(total, 256), 256, 0, s, (const T *)dy, idx, (T *)dx, N, Hi,
                                 Wi, C, k, stride, pad, Ho, Wo, out_nchw, acc));
    return 0;
}

template <typename T>
__global__ void k_avgpool_fwd(const T *__restrict__ x, float *__restrict__ y, int N, int HW, int C) {
    pdl_sync(); B2_TRACE_BEGIN();
    int64_t total = (int64_t)N * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = i % C; int64_t n = i / C;
        float s = 0.f;
        for (int p = 0; p < HW; ++p) s += to_f<T>(x[(n * HW + p) * C + c]);
        y[i] = s / (float)HW;
    }
}
int op_avgpool_forward(const void *x, int xdt, float *y, int N, int HW, int C, cudaStream_t s) {
    DISPATCH_DT(xdt, T, B2_LAUNCH((k_avgpool_fwd<T>), grid_for((int64_t)N * C, 256), 256, 0, s, (const T *)x, y, N, HW, C));
    return 0;
}
template <typename T>
__global__ void k_avgpool_bwd(const float *__restrict__ dy, T *__restrict__ dx, int N, int HW, int C, bool acc) {
    pdl_sync(); B2_TRACE_BEGIN();
    int64_t total = (int64_t)N * HW * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = i % C; int64_t n = i / ((int64_t)HW * C);
        float g = dy[n * C + c] / (float)HW;
        if (acc) g += to_f<T>(dx[i]);
        dx[i] = from_f<T>(g);
    }
}
int op_avgpool_backward(const float *dy, void *dx, int xdt, int N, int HW, int C, bool acc, cudaStream_t s) {
    DISPATCH_DT(xdt, T, B2_LAUNCH((k_avgpool_bwd<T>), grid_for((int64_t)N * HW * C, 256), 256, 0, s, dy, (T *)dx, N, HW, C, acc));
    return 0;
}

// ---------------------------------------------------------------------------------
// pointwise
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void k_relu_mask(const T *__restrict__ y, T *__restrict__ dy, int64_t n) {
    pdl_sync(); B2_TRACE_BEGIN();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (!(to_f<T>(y[i]) > 0.f)) dy[i] = from_f<T>(0.f);
}
int op_relu_mask(const void *y, void *dy, int dt, int64_t n, cudaStream_t s) {
    DISPATCH_DT(dt, T, B2_LAUNCH((k_relu_mask<T>), grid_for(n, 256), 256, 0, s, (const T *)y, (T *)dy, n));
    return 0;
}

template <typename T>
__global__ void __launch_bounds__(256) k_bias_grad(const T *__restrict__ dy, int64_t M, int C, float *__restrict__ db) {
    pdl_sync(); B2_TRACE_BEGIN();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float s = 0.f;
    if (c < C)
        for (int64_t r = ty; r < M; r += 8) s += to_f<T>(dy[r * C + c]);
    __shared__ float sh[8][33];
    sh[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < C) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) a += sh[j][tx];
        db[c] = a;
    }
}
int op_bias_grad(const void *dy, int dt, int64_t M, int C, float *db, cudaStream_t s) {
    DISPATCH_DT(dt, T, B2_LAUNCH((k_bias_grad<T>), (C + 31) / 32, 256, 0, s, (const T *)dy, M, C, db));
    return 0;
}

// ---------------------------------------------------------------------------------
// cross entropy (mean reduction) + softmax gradient, one kernel, one block
// ---------------------------------------------------------------------------------
// Rows are staged through shared memory so that the global reads are coalesced (a row is only
// C = 10 floats); one thread then owns one row.  `running` (optional, float[2]) accumulates
// (sum of losses, count) on the device -- the reference's running_loss (cifar_example_ddp.py:110)
// without a host read every step.  A label outside [0, C) traps (torch raises a device assert).
constexpr int kCeRows = 512;
__global__ void __launch_bounds__(kCeRows)
k_ce_loss_grad(const float *__restrict__ logits, const int64_t *__restrict__ labels, int B, int C, int tile_rows,
               float *__restrict__ loss, float *__restrict__ dlogits, float *__restrict__ running) {
    pdl_sync(); B2_TRACE_BEGIN();
    extern __shared__ float tile[];                 // [tile_rows][C], tile_rows <= kCeRows
    __shared__ float sh[kCeRows / 32];
    float local = 0.f;
    const float invB = 1.f / (float)B;
    for (int r0 = 0; r0 < B; r0 += tile_rows) {
        const int rows = min(tile_rows, B - r0);
        const float *src = logits + (int64_t)r0 * C;
        for (int i = threadIdx.x; i < rows * C; i += kCeRows) tile[i] = src[i];
        __syncthreads();
        if ((int)threadIdx.x < rows) {
            float *row = tile + threadIdx.x * C;
            float mx = -INFINITY;
            for (int j = 0; j < C; ++j) mx = fmaxf(mx, row[j]);
            float se = 0.f;
            for (int j = 0; j < C; ++j) se += expf(row[j] - mx);
            const float lse = logf(se);
            const long long lab = labels[r0 + threadIdx.x];
            if (lab < 0 || lab >= C) {
                printf("b2 cross_entropy: label %lld of row %d is outside [0, %d)\n", lab, r0 + (int)threadIdx.x, C);
                __trap();
            }
            local += -(row[lab] - mx - lse);
            if (dlogits)
                for (int j = 0; j < C; ++j) row[j] = (expf(row[j] - mx - lse) - (j == (int)lab ? 1.f : 0.f)) * invB;
        }
        __syncthreads();
        if (dlogits) {
            float *dst = dlogits + (int64_t)r0 * C;
            for (int i = threadIdx.x; i < rows * C; i += kCeRows) dst[i] = tile[i];
        }
        __syncthreads();
    }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < kCeRows / 32; ++w) t += sh[w];        // fixed order: deterministic
        t *= invB;
        *loss = t;
        if (running) { running[0] += t; running[1] += 1.f; }
    }
}
int op_ce_loss_grad(const float *logits, const int64_t *labels, int B, int C, float *loss, float *dlogits, float *running,
                    cudaStream_t s) {
    B2_CHECK(B > 0 && C > 0, B2_ERR_ARG, "ce: empty batch (B=%d, C=%d)", B, C);
    B2_CHECK(C <= 10240, B2_ERR_UNSUPPORTED, "ce: at most 10240 classes (row tile in shared memory), got %d", C);
    int tile_rows = 10240 / C; if (tile_rows > kCeRows) tile_rows = kCeRows;        // <= 40 KB of shared memory
    B2_LAUNCH(k_ce_loss_grad, 1, kCeRows, (size_t)tile_rows * C * sizeof(float), s, logits, labels, B, C, tile_rows, loss, dlogits, running);
    return 0;
}

// ---------------------------------------------------------------------------------
// SGD with momentum: 20 B/param of HBM traffic (16 on the first step)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_sgd(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, int64_t n, float lr, float mom,
      float gscale, int first, bf16 *__restrict__ pb) {
    pdl_sync(); B2_TRACE_BEGIN();
    int64_t n4 = n >> 2;
    float4 *p4 = (float4 *)p; const float4 *g4 = (const float4 *)g; float4 *m4 = (float4 *)m;
    // torch semantics (sgd.py:_single_tensor_sgd): buf.mul_(momentum).add_(grad) is two
    // roundings; param.add_(buf, alpha=-lr) is one fused multiply-add.
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 gv = g4[i], pv = p4[i], mv;
        if (gscale != 1.f) { gv.x = __fmul_rn(gv.x, gscale); gv.y = __fmul_rn(gv.y, gscale); gv.z = __fmul_rn(gv.z, gscale); gv.w = __fmul_rn(gv.w, gscale); }
        if (first) mv = gv;
        else {
            mv = m4[i];
            mv.x = __fadd_rn(__fmul_rn(mv.x, mom), gv.x); mv.y = __fadd_rn(__fmul_rn(mv.y, mom), gv.y);
            mv.z = __fadd_rn(__fmul_rn(mv.z, mom), gv.z); mv.w = __fadd_rn(__fmul_rn(mv.w, mom), gv.w);
        }
        pv.x = fmaf(-lr, mv.x, pv.x); pv.y = fmaf(-lr, mv.y, pv.y); pv.z = fmaf(-lr, mv.z, pv.z); pv.w = fmaf(-lr, mv.w, pv.w);
        m4[i] = mv; p4[i] = pv;
        if (pb) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
            uint2 pk; pk.x = *(uint32_t *)&lo; pk.y = *(uint32_t *)&hi;
            ((uint2 *)pb)[i] = pk;
        }
    }
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gv = gscale != 1.f ? __fmul_rn(g[i], gscale) : g[i];
        float mv = first ? gv : __fadd_rn(__fmul_rn(m[i], mom), gv);
        m[i] = mv; float pn = fmaf(-lr, mv, p[i]); p[i] = pn;
        if (pb) pb[i] = __float2bfloat16_rn(pn);
    }
    B2_TRACE_END();
}
int op_sgd(float *p, const float *g, float *m, int64_t n, float lr, float momentum, float gscale, int first, bf16 *pb,
           cudaStream_t s) {
    if (n <= 0) return 0;
    B2_CHECK(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0), B2_ERR_ARG,
             "sgd: p/g/m must be 16-byte aligned");
    B2_CHECK(((uintptr_t)pb % 8) == 0, B2_ERR_ARG, "sgd: bf16 mirror must be 8-byte aligned");
    B2_LAUNCH(k_sgd, grid_for(ceil_div(n, 4), 256, kNumSMs * 8), 256, 0, s, p, g, m, n, lr, momentum, gscale, first, pb);
    return 0;
}

}  // namespace b2

// =====================================================================================
// round-1 additions: vectorised bf16 max pooling, stem im2col, fp32 classifier head
// =====================================================================================
namespace b2 {

static inline int grid_for2(int64_t n, int block, int max_blocks = kNumSMs * 16) {
    int64_t b = ceil_div(n, block);
    if (b > max_blocks) b = max_blocks;
    if (b < 1) b = 1;
    return (int)b;
}

template <int KK, int SS, int PP>        // compile-time window / stride / pad (0 = use the runtime arguments)
__global__ void __launch_bounds__(256)
k_maxpool_fwd_bf16x8(const bf16 *__restrict__ x, bf16 *__restrict__ y, uint8_t *__restrict__ idx, int N, int Hi, int Wi,
                     int C, int k_, int stride_, int pad_, int Ho, int Wo) {
    pdl_sync(); B2_TRACE_BEGIN();
    const int k = KK ? KK : k_, stride = KK ? SS : stride_, pad = KK ? PP : pad_;
    const int C8 = C >> 3;
    const int64_t total = (int64_t)N * Ho * Wo * C8;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c8 = i % C8; int64_t r = i / C8;
        int ow = r % Wo; r /= Wo;
        int oh = r % Ho; int n = r / Ho;
        float best[8]; int bi[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; bi[e] = 255; }
        if (KK) {
            // compile-time window: all KK*KK (predicated) loads first, then the compare chain in the same order
            constexpr int KW = KK ? KK : 1;
            uint4 raw[KW * KW]; bool ok[KW * KW];
#pragma unroll
            for (int a = 0; a < KW; ++a)
#pragma unroll
                for (int b = 0; b < KW; ++b) {
                    const int hi = oh * SS - PP + a, wi = ow * SS - PP + b;
                    ok[a * KW + b] = hi >= 0 && hi < Hi && wi >= 0 && wi < Wi;
                    raw[a * KW + b] = make_uint4(0u, 0u, 0u, 0u);
                    if (ok[a * KW + b]) raw[a * KW + b] = *(const uint4 *)(x + (((int64_t)n * Hi + hi) * Wi + wi) * C + c8 * 8);
                }
#pragma unroll
            for (int j = 0; j < KW * KW; ++j) {
                if (!ok[j]) continue;
                const bf16 *v = (const bf16 *)&raw[j];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float f = __bfloat162float(v[e]);
                    if (f > best[e] || bi[e] == 255) { best[e] = f; bi[e] = j; }
                }
            }
        } else
        for (int a = 0; a < k; ++a) {
            int hi = oh * stride - pad + a;
            if (hi < 0 || hi >= Hi) continue;
            for (int b = 0; b < k; ++b) {
                int wi = ow * stride - pad + b;
                if (wi < 0 || wi >= Wi) continue;
                uint4 raw = *(const uint4 *)(x + (((int64_t)n * Hi + hi) * Wi + wi) * C + c8 * 8);
                const bf16 *v = (const bf16 *)&raw;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float f = __bfloat162float(v[e]);
                    if (f > best[e] || bi[e] == 255) { best[e] = f; bi[e] = a * k + b; }
                }
            }
        }
        __align__(16) bf16 o[8]; __align__(8) uint8_t ix[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { o[e] = __float2bfloat16_rn(best[e]); ix[e] = (uint8_t)bi[e]; }
        *(uint4 *)(y + i * 8) = *(uint4 *)o;
        *(uint2 *)(idx + i * 8) = *(uint2 *)ix;
    }
}
int op_maxpool_forward_bf16x8(const bf16 *x, bf16 *y, uint8_t *idx, int N, int Hi, int Wi, int C, int k, int stride,
                              int pad, int Ho, int Wo, cudaStream_t s) {
    B2_CHECK(C % 8 == 0, B2_ERR_ARG, "maxpool x8: C must be a multiple of 8");
    int64_t total = (int64_t)N * Ho * Wo * (C / 8);
    if (k == 3 && stride == 2 && pad == 1) { B2_LAUNCH((k_maxpool_fwd_bf16x8<3, 2, 1>), grid_for2(total, 256), 256, 0, s, x, y, idx, N, Hi, Wi, C, k, stride, pad, Ho, Wo); }
    else { B2_LAUNCH((k_maxpool_fwd_bf16x8<0, 0, 0>), grid_for2(total, 256), 256, 0, s, x, y, idx, N, Hi, Wi, C, k, stride, pad, Ho, Wo); }
    return 0;
}

template <int KK, int SS, int PP>
__global__ void __launch_bounds__(256)
k_maxpool_bwd_bf16x8(const bf16 *__restrict__ dy, const uint8_t *__restrict__ idx, bf16 *__restrict__ dx, int N, int Hi,
                     int Wi, int C, int k_, int stride_, int pad_, int Ho, int Wo, bool acc) {
    pdl_sync(); B2_TRACE_BEGIN();
    const int k = KK ? KK : k_, stride = KK ? SS : stride_, pad = KK ? PP : pad_;
    const int C8 = C >> 3;
    const int64_t total = (int64_t)N * Hi * Wi * C8;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c8 = i % C8; int64_t r = i / C8;
        int wi = r % Wi; r /= Wi;
        int hi = r % Hi; int n = r / Hi;
        float g[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] = 0.f;
        if (KK == 3 && SS == 2 && PP == 1) {
            // 3x3 / stride 2 / pad 1: pixel (hi, wi) lies in window row hi>>1 (at row (hi&1)+1 of it) and, for odd hi,
            // also in window row (hi>>1)+1 (at its row 0); same for columns.  All (up to) four candidate windows are
            // loaded at once instead of one dependent round trip per window.
            const int ohA = hi >> 1, owA = wi >> 1;
            const bool rowA = ohA < Ho, colA = owA < Wo;
            const bool rowB = (hi & 1) && (ohA + 1 < Ho), colB = (wi & 1) && (owA + 1 < Wo);
            const int rA = (hi & 1) + 1, cA = (wi & 1) + 1;
            const bool ok[4] = {rowA && colA, rowA && colB, rowB && colA, rowB && colB};
            const int code[4] = {rA * 3 + cA, rA * 3, cA, 0};
            uint2 iraw[4]; uint4 draw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                iraw[j] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);          // 255 never equals a window position (0..8)
                draw[j] = make_uint4(0u, 0u, 0u, 0u);
                if (ok[j]) {
                    const int64_t o = ((((int64_t)n * Ho + (ohA + (j >> 1))) * Wo + (owA + (j & 1))) * C8 + c8) * 8;
                    iraw[j] = *(const uint2 *)(idx + o);
                    draw[j] = *(const uint4 *)(dy + o);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint8_t *ix = (const uint8_t *)&iraw[j]; const bf16 *dv = (const bf16 *)&draw[j];
#pragma unroll
                for (int e = 0; e < 8; ++e) if (ix[e] == code[j]) g[e] += __bfloat162float(dv[e]);
            }
        } else {
        int oh_lo = hi + pad - k + stride; oh_lo = oh_lo < 0 ? 0 : oh_lo / stride;
        int oh_hi = (hi + pad) / stride; if (oh_hi > Ho - 1) oh_hi = Ho - 1;
        int ow_lo = wi + pad - k + stride; ow_lo = ow_lo < 0 ? 0 : ow_lo / stride;
        int ow_hi = (wi + pad) / stride; if (ow_hi > Wo - 1) ow_hi = Wo - 1;
        for (int oh = oh_lo; oh <= oh_hi; ++oh)
            for (int ow = ow_lo; ow <= ow_hi; ++ow) {
                const int mine = (hi - (oh * stride - pad)) * k + (wi - (ow * stride - pad));
                const int64_t o = ((((int64_t)n * Ho + oh) * Wo + ow) * C8 + c8) * 8;
                uint2 iraw = *(const uint2 *)(idx + o);
                uint4 draw = *(const uint4 *)(dy + o);
                const uint8_t *ix = (const uint8_t *)&iraw; const bf16 *dv = (const bf16 *)&draw;
#pragma unroll
                for (int e = 0; e < 8; ++e) if (ix[e] == mine) g[e] += __bfloat162float(dv[e]);
            }
        }
        __align__(16) bf16 o8[8];
        if (acc) {
            uint4 old = *(const uint4 *)(dx + i * 8);
            const bf16 *ov = (const bf16 *)&old;
#pragma unroll
            for (int e = 0; e < 8; ++e) g[e] += __bfloat162float(ov[e]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) o8[e] = __float2bfloat16_rn(g[e]);
        *(uint4 *)(dx + i * 8) = *(uint4 *)o8;
    }
}
int op_maxpool_backward_bf16x8(const bf16 *dy, const uint8_t *idx, bf16 *dx, int N, int Hi, int Wi, int C, int k,
                               int stride, int pad, int Ho, int Wo, bool acc, cudaStream_t s) {
    B2_CHECK(C % 8 == 0, B2_ERR_ARG, "maxpool x8: C must be a multiple of 8");
    int64_t total = (int64_t)N * Hi * Wi * (C / 8);
    if (k == 3 && stride == 2 && pad == 1) { B2_LAUNCH((k_maxpool_bwd_bf16x8<3, 2, 1>), grid_for2(total, 256), 256, 0, s, dy, idx, dx, N, Hi, Wi, C, k, stride, pad, Ho, Wo, acc); }
    else { B2_LAUNCH((k_maxpool_bwd_bf16x8<0, 0, 0>), grid_for2(total, 256), 256, 0, s, dy, idx, dx, N, Hi, Wi, C, k, stride, pad, Ho, Wo, acc); }
    return 0;
}

// ---- stem im2col.  One block per image: the whole fp32 image (CH x H x W, 12 KB for CIFAR) is
// staged in shared memory with coalesced loads, then every thread emits 16-byte chunks of the bf16
// im2col rows.  GEMM-K order (c, kh, kw8): chunk j < CH*KS is (c = j / KS, kh = j % KS) and holds the
// 8 consecutive input pixels wi = wo*stride - pad + 0..7 of row hi (the 8th has a zero weight); chunks
// beyond CH*KS are zero.  No per-element index arithmetic.
template <int KS, int CH>
__global__ void __launch_bounds__(256)
k_stem_im2col(const float *__restrict__ x, bf16 *__restrict__ col, int N, int H, int W, int stride, int pad,
              int Ho, int Wo, int kpad) {
    pdl_sync(); B2_TRACE_BEGIN();
    static_assert(KS <= 8, "one filter row per 8-wide chunk");
    extern __shared__ float simg[];                 // [CH][H][W + 1]: odd row pitch, lanes on different rows hit different banks
    const int n = blockIdx.x;
    const int WP = W + 1;
    const float *src = x + (int64_t)n * CH * H * W;
    for (int i = threadIdx.x; i < CH * H * W; i += blockDim.x) { const int r = i / W; simg[r * WP + (i - r * W)] = src[i]; }
    __syncthreads();
    const int chunks = kpad >> 3;
    bf16 *obase = col + (int64_t)n * Ho * Wo * kpad;
    const int items = Ho * Wo * chunks;
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int j = i % chunks, pix = i / chunks;
        const int ho = pix / Wo, wo = pix - ho * Wo;
        __align__(16) bf16 o[8];
        const int c = j / KS, kh = j - c * KS;
        const int hi = ho * stride - pad + kh, wi0 = wo * stride - pad;
        const bool live = j < CH * KS && hi >= 0 && hi < H;
        const float *row = simg + (c * H + (live ? hi : 0)) * WP;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int wi = wi0 + e;
            o[e] = __float2bfloat16_rn((live && wi >= 0 && wi < W) ? row[wi] : 0.f);
        }
        *(uint4 *)(obase + (int64_t)pix * kpad + j * 8) = *(uint4 *)o;
    }
}
int op_stem_im2col(const float *x_nchw, bf16 *col, int N, int C, int H, int W, int k, int stride, int pad, int Ho, int Wo,
                   int kpad, cudaStream_t s) {
    B2_CHECK(kpad % 8 == 0 && kpad >= C * k * 8, B2_ERR_ARG, "stem im2col: bad kpad %d", kpad);
    B2_CHECK(k == 7 && C == 3 && (C * H * W) % 4 == 0, B2_ERR_UNSUPPORTED, "stem im2col: only the 7x7 / 3-channel ResNet stem is instantiated");
    size_t smem = (size_t)C * H * (W + 1) * sizeof(float);
    B2_LAUNCH((k_stem_im2col<7, 3>), N, 256, smem, s, x_nchw, col, N, H, W, stride, pad, Ho, Wo, kpad);
    return 0;
}

__global__ void k_stem_pack_weights(const float *__restrict__ m, bf16 *__restrict__ w, int Co, int C, int k, int kpad) {
    pdl_sync(); B2_TRACE_BEGIN();
    const int total = Co * kpad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i / kpad, kk = i - co * kpad;
        const int kw = kk & 7, j = kk >> 3;
        const int c = j / k, kh = j - c * k;
        float v = 0.f;
        if (j < C * k && kw < k) v = m[((int64_t)co * k * k + kh * k + kw) * C + c];
        w[i] = __float2bfloat16_rn(v);
    }
}
int op_stem_pack_weights(const float *master, bf16 *wpad, int Co, int C, int k, int kpad, cudaStream_t s) {
    B2_LAUNCH(k_stem_pack_weights, (Co * kpad + 255) / 256, 256, 0, s, master, wpad, Co, C, k, kpad);
    return 0;
}
__global__ void k_stem_unpack_wgrad(const float *__restrict__ dwpad, float *__restrict__ g, int Co, int C, int k, int kpad) {
    pdl_sync(); B2_TRACE_BEGIN();
    const int K = k * k * C, total = Co * K;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i / K, r = i - co * K;
        const int c = r % C, tap = r / C;
        const int kh = tap / k, kw = tap - kh * k;
        g[i] = dwpad[(int64_t)co * kpad + (c * k + kh) * 8 + kw];
    }
}
int op_stem_unpack_wgrad(const float *dwpad, float *grad, int Co, int C, int k, int kpad, cudaStream_t s) {
    B2_LAUNCH(k_stem_unpack_wgrad, (Co * k * k * C + 255) / 256, 256, 0, s, dwpad, grad, Co, C, k, kpad);
    return 0;
}

// ---- classifier head, fp32.  forward: one warp per sample; N <= 32 outputs
// one warp per sample; the weight matrix is staged in shared memory in K-chunks of 512 (every
// block reading the same 20 KB straight from L2 serialises on the hot lines: 32 us -> ~4 us)
template <int NMAX>
__global__ void __launch_bounds__(256)
k_fc_fwd(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias, float *__restrict__ y,
         int B, int K, int N, bool relu) {
    pdl_sync(); B2_TRACE_BEGIN();
    constexpr int KC = 512;
    __shared__ float sw[NMAX * KC];
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const bool active = warp < B;
    const float *xr = x + (int64_t)(active ? warp : 0) * K;
    float acc[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
    for (int k0 = 0; k0 < K; k0 += KC) {
        const int kc = K - k0 < KC ? K - k0 : KC;
        __syncthreads();
        for (int i = threadIdx.x; i < N * kc; i += blockDim.x) { int n = i / kc, kk = i - n * kc; sw[n * KC + kk] = w[(int64_t)n * K + k0 + kk]; }
        __syncthreads();
        if (active)
            for (int kk = lane; kk < kc; kk += 32) {
                const float xv = xr[k0 + kk];
#pragma unroll
                for (int n = 0; n < NMAX; ++n)
                    if (n < N) acc[n] = fmaf(xv, sw[n * KC + kk], acc[n]);
            }
    }
    if (!active) return;
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
        if (n < N) {
            float s = warp_sum(acc[n]);
            if (lane == 0) {
                if (bias) s += bias[n];
                if (relu) s = s > 0.f ? s : 0.f;
                y[(int64_t)warp * N + n] = s;
            }
        }
    }
}
int op_fc_forward(const float *x, const float *w, const float *bias, float *y, int B, int K, int N, bool relu, cudaStream_t s) {
    B2_CHECK(N <= 16, B2_ERR_UNSUPPORTED, "fc head: at most 16 outputs");
    B2_LAUNCH((k_fc_fwd<16>), (int)ceil_div((int64_t)B * 32, 256), 256, 0, s, x, w, bias, y, B, K, N, relu);
    return 0;
}
// dW[n][k] = sum_b dy[b][n] x[b][k]: block = 32 k-columns x 8 batch lanes, fixed-order smem combine
__global__ void __launch_bounds__(256)
k_fc_wgrad(const float *__restrict__ x, const float *__restrict__ dy, float *__restrict__ dw, float *__restrict__ db,
           int B, int K, int N) {
    pdl_sync(); B2_TRACE_BEGIN();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int kk = blockIdx.x * 32 + tx, n = blockIdx.y;
    float s = 0.f, sb = 0.f;
    if (kk < K) {
        int b = ty;
        for (; b + 24 < B; b += 32) {
            float d0 = dy[(int64_t)b * N + n], d1 = dy[(int64_t)(b + 8) * N + n], d2 = dy[(int64_t)(b + 16) * N + n], d3 = dy[(int64_t)(b + 24) * N + n];
            float x0 = x[(int64_t)b * K + kk], x1 = x[(int64_t)(b + 8) * K + kk], x2 = x[(int64_t)(b + 16) * K + kk], x3 = x[(int64_t)(b + 24) * K + kk];
            s = fmaf(d0, x0, s); s = fmaf(d1, x1, s); s = fmaf(d2, x2, s); s = fmaf(d3, x3, s);
            sb += d0 + d1 + d2 + d3;
        }
        for (; b < B; b += 8) { float d = dy[(int64_t)b * N + n]; s = fmaf(d, x[(int64_t)b * K + kk], s); sb += d; }
    }
    __shared__ float sh[8][33], shb[8];
    sh[ty][tx] = s;
    if (tx == 0) shb[ty] = sb;
    __syncthreads();
    if (ty == 0 && kk < K) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) a += sh[j][tx];
        dw[(int64_t)n * K + kk] = a;
        if (db && kk == 0) { float t = 0.f; for (int j = 0; j < 8; ++j) t += shb[j]; db[n] = t; }
    }
}
__global__ void __launch_bounds__(256)
k_fc_dgrad(const float *__restrict__ w, const float *__restrict__ dy, float *__restrict__ dx, int B, int K, int N, bool acc) {
    pdl_sync(); B2_TRACE_BEGIN();
    const int64_t total = (int64_t)B * K;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int kk = i % K; int64_t b = i / K;
        float s = 0.f;
        for (int n = 0; n < N; ++n) s = fmaf(dy[b * N + n], w[(int64_t)n * K + kk], s);
        if (acc) s += dx[i];
        dx[i] = s;
    }
}
// ---- classifier head, fused: adaptive average pool (1x1) + Linear forward; Linear dgrad + average-pool backward ----
// (torchvision resnet.py:278-280: avgpool -> flatten -> fc).  One warp per image; same summation order as the
// separate k_avgpool_fwd + k_fc_fwd pair, so the fp32 parity mode is unchanged bit for bit.
template <typename T, int NMAX>
__global__ void __launch_bounds__(256)
k_head_fwd(const T *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias, float *__restrict__ feat,
           float *__restrict__ y, int B, int HW, int C, int N) {
    pdl_sync(); B2_TRACE_BEGIN();
    // The weight rows are C*4 bytes apart (2 KB for ResNet-18): read straight from global, the N rows of one channel
    // block collide in the L1 sets and every access goes to L2 (measured 28-47 us for this 5 MFLOP layer).  So the
    // block first copies W [N][C] into shared memory with coalesced 16-byte loads; then one warp per image, lanes over
    // channels (conflict-free shared reads), N accumulators per lane, butterfly at the end.
    extern __shared__ float sw[];                       // [N][C]
    {
        const int n4 = N * C / 4;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) ((float4 *)sw)[i] = __ldg((const float4 *)w + i);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    float acc[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
    if (HW == 1) {
        // 1x1 feature map (32x32 inputs): 16 independent activation loads in flight per lane, then the multiply-adds
        const T *xr = x + (int64_t)b * C;
        for (int c0 = lane; c0 < C; c0 += 512) {
            float xv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { const int c = c0 + 32 * j; xv[j] = c < C ? to_f<T>(xr[c]) : 0.f; }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int c = c0 + 32 * j;
                if (c < C) {
                    feat[(int64_t)b * C + c] = xv[j];
#pragma unroll
                    for (int n = 0; n < NMAX; ++n)
                        if (n < N) acc[n] = fmaf(xv[j], sw[n * C + c], acc[n]);
                }
            }
        }
    } else {
        for (int c = lane; c < C; c += 32) {
            float s = 0.f;
            for (int p = 0; p < HW; ++p) s += to_f<T>(x[((int64_t)b * HW + p) * C + c]);
            s = s / (float)HW;
            feat[(int64_t)b * C + c] = s;
#pragma unroll
            for (int n = 0; n < NMAX; ++n)
                if (n < N) acc[n] = fmaf(s, sw[n * C + c], acc[n]);
        }
    }
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
        if (n < N) {
            const float s = warp_sum(acc[n]);
            if (lane == 0) y[(int64_t)b * N + n] = bias ? s + bias[n] : s;
        }
    }
}
int op_head_forward(const void *x, int xdt, const float *w, const float *bias, float *feat, float *y, int B, int HW, int C, int N,
                    cudaStream_t s) {
    B2_CHECK(N <= 16 && (N * C) % 4 == 0 && (size_t)N * C * 4 <= 160 * 1024, B2_ERR_UNSUPPORTED, "fc head: %d x %d weights do not fit the shared-memory copy", N, C);
    static const bool attr_set = [] {
        cudaFuncSetAttribute(k_head_fwd<float, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(k_head_fwd<bf16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        return true;
    }();
    (void)attr_set;
    DISPATCH_DT(xdt, T, B2_LAUNCH((k_head_fwd<T, 16>), (int)ceil_div(B, 8), 256, (size_t)N * C * sizeof(float), s, (const T *)x, w, bias, feat, y, B, HW, C, N));
    return 0;
}
template <typename T>
__global__ void __launch_bounds__(256)
k_head_dgrad(const float *__restrict__ w, const float *__restrict__ dy, float *__restrict__ dfeat, T *__restrict__ dx, int B, int HW,
             int C, int N, bool acc) {
    pdl_sync(); B2_TRACE_BEGIN();
    const int64_t total = (int64_t)B * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = i % C; const int64_t b = i / C;
        float s = 0.f;
        for (int n = 0; n < N; ++n) s = fmaf(dy[b * N + n], __ldg(w + (int64_t)n * C + c), s);
        dfeat[i] = s;
        const float g = s / (float)HW;
        for (int p = 0; p < HW; ++p) {
            const int64_t j = (b * HW + p) * C + c;
            dx[j] = from_f<T>(acc ? g + to_f<T>(dx[j]) : g);
        }
    }
}
int op_head_dgrad(const float *w, const float *dy, float *dfeat, void *dx, int xdt, int B, int HW, int C, int N, bool acc, cudaStream_t s) {
    DISPATCH_DT(xdt, T, B2_LAUNCH((k_head_dgrad<T>), grid_for((int64_t)B * C, 256), 256, 0, s, w, dy, dfeat, (T *)dx, B, HW, C, N, acc));
    return 0;
}
int op_fc_wgrad(const float *x, const float *dy, float *dw, float *db, int B, int K, int N, cudaStream_t s) {
    B2_LAUNCH(k_fc_wgrad, dim3((unsigned)ceil_div(K, 32), (unsigned)N), 256, 0, s, x, dy, dw, db, B, K, N);
    return 0;
}
int op_fc_backward(const float *x, const float *w, const float *dy, float *dw, float *db, float *dx, int B, int K, int N,
                   bool acc_dx, cudaStream_t s) {
    dim3 grid((unsigned)ceil_div(K, 32), (unsigned)N);
    B2_LAUNCH(k_fc_wgrad, grid, 256, 0, s, x, dy, dw, db, B, K, N);
    if (dx) B2_LAUNCH(k_fc_dgrad, grid_for2((int64_t)B * K, 256), 256, 0, s, w, dy, dx, B, K, N, acc_dx);
    return 0;
}

}  // namespace b2

// =====================================================================================
// fused BatchNorm (statistics produced by the conv epilogue)
// =====================================================================================
namespace b2 {

template <typename T, int V>
__global__ void __launch_bounds__(256)
k_bn_apply_fused(const T *__restrict__ x, T *__restrict__ y, const T *__restrict__ res, const float *__restrict__ fsum,
                 int64_t M, const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ save_mean,
                 float *__restrict__ save_invstd, float *running_mean, float *running_var, float eps, float momentum,
                 bool relu, int64_t total, int C) {
    pdl_sync(); B2_TRACE_BEGIN();
    extern __shared__ float sm[];             // scale[C], shift[C]
    float *scale = sm, *shift = sm + C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t0 = 0.f, t1 = 0.f;
#pragma unroll
        for (int k = 0; k < kBnSumCopies; ++k) { t0 += fsum[(2 * k) * C + c]; t1 += fsum[(2 * k + 1) * C + c]; }
        double mu = (double)t0 / (double)M;
        double var = (double)t1 / (double)M - mu * mu;
        if (var < 0.0) var = 0.0;
        float is = (float)(1.0 / sqrt(var + (double)eps));
        float g = gamma[c];
        scale[c] = is * g;
        shift[c] = beta[c] - (float)mu * is * g;
        if (blockIdx.x == 0) {
            save_mean[c] = (float)mu; save_invstd[c] = is;
            if (running_mean) {
                double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
            }
        }
    }
    __syncthreads();
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * V; i < total; i += (int64_t)gridDim.x * blockDim.x * V) {
        int c0 = i % C;
        __align__(16) T xv[V], rv[V], ov[V];
        *(uint4 *)xv = *(const uint4 *)(x + i);
        if (res) *(uint4 *)rv = *(const uint4 *)(res + i);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float v = fmaf(to_f<T>(xv[j]), scale[c0 + j], shift[c0 + j]);
            if (res) v += to_f<T>(rv[j]);
            if (relu) v = v > 0.f ? v : 0.f;
            ov[j] = from_f<T>(v);
        }
        *(uint4 *)(y + i) = *(uint4 *)ov;
    }
}

int op_bn_forward_fused(const void *x, void *y, const void *res, int dt, int64_t M, int C, const float *fsum,
                        const float *gamma, const float *beta, float *running_mean, float *running_var, float eps,
                        float momentum, bool relu, const BnWork &w, cudaStream_t s) {
    B2_CHECK(dt == DT_BF16 && C % 8 == 0, B2_ERR_ARG, "bn fused: bf16 activations with C %% 8 == 0 only");
    int64_t total = M * C;
    B2_CHECK(total < ((int64_t)1 << 32), B2_ERR_UNSUPPORTED, "bn fused: tensor of %lld elements (32-bit indexing)", (long long)total);
    int64_t blocks = ceil_div(total / 8, 256); if (blocks > kNumSMs * 2) blocks = kNumSMs * 2;
    B2_LAUNCH((k_bn_apply_fused<bf16, 8>), (int)blocks, 256, 2 * C * sizeof(float), s, (const bf16 *)x, (bf16 *)y, (const bf16 *)res,
              fsum, M, gamma, beta, w.mean, w.invstd, running_mean, running_var, eps, momentum, relu, total, C);
    return 0;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_bn_bwd_reduce_atomic(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ dy,
                       const float *__restrict__ mean, const float *__restrict__ invstd, bool relu, int64_t M, int C,
                       int opb /* channel octets per block: min(C/8, 32) */, float *__restrict__ bsum) {
    pdl_sync(); B2_TRACE_BEGIN();
    // each thread owns 8 consecutive channels (16-byte loads); 256/opb row lanes per block
    const int tx = threadIdx.x % opb, ty = threadIdx.x / opb, lanes = 256 / opb;
    const int c0 = (blockIdx.x * opb + tx) * 8;
    const int64_t rows_per_block = (M + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = blockIdx.y * rows_per_block;
    int64_t r1 = r0 + rows_per_block; if (r1 > M) r1 = M;
    float s0[8], s1[8], mu[8], is[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s0[e] = 0.f; s1[e] = 0.f; mu[e] = mean[c0 + e]; is[e] = invstd[c0 + e]; }
    constexpr int RU = 2;                      // rows in flight per thread (the loop is latency-bound otherwise)
    for (int64_t rb = r0 + ty; rb < r1; rb += lanes * RU) {
        __align__(16) T xv[RU][8], yv[RU][8], gv[RU][8];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int64_t r = rb + (int64_t)u * lanes;
            if (r < r1) {
                *(uint4 *)xv[u] = *(const uint4 *)(x + r * C + c0);
                *(uint4 *)gv[u] = *(const uint4 *)(dy + r * C + c0);
                if (relu) *(uint4 *)yv[u] = *(const uint4 *)(y + r * C + c0);
            }
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            if (rb + (int64_t)u * lanes >= r1) break;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float g = to_f<T>(gv[u][e]);
                if (relu && !(to_f<T>(yv[u][e]) > 0.f)) g = 0.f;
                s0[e] += g; s1[e] += g * ((to_f<T>(xv[u][e]) - mu[e]) * is[e]);
            }
        }
    }
    __shared__ float sh[256 * 17];
    float *mine = sh + threadIdx.x * 17;       // thread (ty, tx) sits at index ty*opb + tx == threadIdx.x
#pragma unroll
    for (int e = 0; e < 8; ++e) { mine[e] = s0[e]; mine[8 + e] = s1[e]; }
    __syncthreads();
    for (int stride = lanes >> 1; stride > 0; stride >>= 1) {      // tree over the row lanes, all threads busy
        if (ty < stride) {
            const float *other = sh + (threadIdx.x + stride * opb) * 17;
#pragma unroll
            for (int e = 0; e < 16; ++e) mine[e] += other[e];
        }
        __syncthreads();
    }
    if (ty == 0) {
        float *dst = bsum + (size_t)(blockIdx.y % kBnSumCopies) * 2 * C;     // spread same-address atomics over copies
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            atomicAdd(dst + c0 + e, mine[e]);
            atomicAdd(dst + C + c0 + e, mine[8 + e]);
        }
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
k_bn_bwd_apply_fused(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ dy, T *__restrict__ dx,
                     T *__restrict__ dres, const float *__restrict__ mean, const float *__restrict__ invstd,
                     const float *__restrict__ bsum, const float *__restrict__ gamma, float *__restrict__ dgamma,
                     float *__restrict__ dbeta, int64_t M, bool relu, bool acc_dx, bool acc_dres, int64_t total, int C) {
    pdl_sync(); B2_TRACE_BEGIN();
    extern __shared__ float sm[];     // a[C], b[C], k[C]: dx = a*g + b*x + k   (per channel)
    float *ca = sm, *cb = sm + C, *ck = sm + 2 * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int k = 0; k < kBnSumCopies; ++k) { sg += bsum[(2 * k) * C + c]; sgx += bsum[(2 * k + 1) * C + c]; }
        float is = invstd[c], mu = mean[c], a = gamma[c] * is;
        float c1 = sg / (float)M, c2 = sgx / (float)M;
        // dx = a*(g - c1 - xhat*c2), xhat = (x - mu)*is
        ca[c] = a; cb[c] = -a * c2 * is; ck[c] = -a * c1 + a * c2 * is * mu;
        if (blockIdx.x == 0) { dbeta[c] = sg; dgamma[c] = sgx; }
    }
    __syncthreads();
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * V; i < total; i += (int64_t)gridDim.x * blockDim.x * V) {
        const int c0 = (int)((unsigned)i % (unsigned)C);      // total < 2^32 (host check): 32-bit division
        __align__(16) T xv[V], yv[V], gv[V], ov[V], rv[V], pv[V];
        *(uint4 *)xv = *(const uint4 *)(x + i);
        *(uint4 *)gv = *(const uint4 *)(dy + i);
        if (relu) *(uint4 *)yv = *(const uint4 *)(y + i);
        if (acc_dx) *(uint4 *)pv = *(const uint4 *)(dx + i);
        if (dres && acc_dres) *(uint4 *)rv = *(const uint4 *)(dres + i);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            int c = c0 + j;
            float g = to_f<T>(gv[j]);
            if (relu && !(to_f<T>(yv[j]) > 0.f)) g = 0.f;
            float d = fmaf(ca[c], g, fmaf(cb[c], to_f<T>(xv[j]), ck[c]));
            if (acc_dx) d += to_f<T>(pv[j]);
            ov[j] = from_f<T>(d);
            if (dres) {
                float r = g;
                if (acc_dres) r += to_f<T>(rv[j]);
                rv[j] = from_f<T>(r);
            }
        }
        *(uint4 *)(dx + i) = *(uint4 *)ov;
        if (dres) *(uint4 *)(dres + i) = *(uint4 *)rv;
    }
}

// Small feature maps (M rows <= 4096): ONE kernel, one block per 8-channel octet: reduce the
// sums over all rows, derive the coefficients, apply (second read comes from L2).  No atomics.
template <typename T>
__global__ void __launch_bounds__(256)
k_bn_bwd_small(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ dy, T *__restrict__ dx,
               T *__restrict__ dres, const float *__restrict__ mean, const float *__restrict__ invstd,
               const float *__restrict__ gamma, float *__restrict__ dgamma, float *__restrict__ dbeta, int64_t M, int C,
               bool relu, bool acc_dx, bool acc_dres) {
    pdl_sync(); B2_TRACE_BEGIN();
    const int c0 = blockIdx.x * 8;
    float s0[8], s1[8], mu[8], is[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s0[e] = 0.f; s1[e] = 0.f; mu[e] = mean[c0 + e]; is[e] = invstd[c0 + e]; }
    for (int64_t r = threadIdx.x; r < M; r += blockDim.x) {
        __align__(16) T xv[8], yv[8], gv[8];
        *(uint4 *)xv = *(const uint4 *)(x + r * C + c0);
        *(uint4 *)gv = *(const uint4 *)(dy + r * C + c0);
        if (relu) *(uint4 *)yv = *(const uint4 *)(y + r * C + c0);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float g = to_f<T>(gv[e]);
            if (relu && !(to_f<T>(yv[e]) > 0.f)) g = 0.f;
            s0[e] += g; s1[e] += g * ((to_f<T>(xv[e]) - mu[e]) * is[e]);
        }
    }
    __shared__ float red[8][16];
    __shared__ float tot[16];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s0[e] = warp_sum(s0[e]); s1[e] = warp_sum(s1[e]); }
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { red[wid][e] = s0[e]; red[wid][8 + e] = s1[e]; }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        float t = 0.f;
        for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
        tot[threadIdx.x] = t;
        if (threadIdx.x < 8) dbeta[c0 + threadIdx.x] = t; else dgamma[c0 + threadIdx.x - 8] = t;
    }
    __syncthreads();
    float ca[8], cb[8], ck[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        float a = gamma[c0 + e] * is[e], c1 = tot[e] / (float)M, c2 = tot[8 + e] / (float)M;
        ca[e] = a; cb[e] = -a * c2 * is[e]; ck[e] = -a * c1 + a * c2 * is[e] * mu[e];
    }
    for (int64_t r = threadIdx.x; r < M; r += blockDim.x) {
        __align__(16) T xv[8], yv[8], gv[8], ov[8], rv[8], pv[8];
        *(uint4 *)xv = *(const uint4 *)(x + r * C + c0);
        *(uint4 *)gv = *(const uint4 *)(dy + r * C + c0);
        if (relu) *(uint4 *)yv = *(const uint4 *)(y + r * C + c0);
        if (acc_dx) *(uint4 *)pv = *(const uint4 *)(dx + r * C + c0);
        if (dres && acc_dres) *(uint4 *)rv = *(const uint4 *)(dres + r * C + c0);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float g = to_f<T>(gv[e]);
            if (relu && !(to_f<T>(yv[e]) > 0.f)) g = 0.f;
            float d = fmaf(ca[e], g, fmaf(cb[e], to_f<T>(xv[e]), ck[e]));
            if (acc_dx) d += to_f<T>(pv[e]);
            ov[e] = from_f<T>(d);
            if (dres) { float rr = g; if (acc_dres) rr += to_f<T>(rv[e]); rv[e] = from_f<T>(rr); }
        }
        *(uint4 *)(dx + r * C + c0) = *(uint4 *)ov;
        if (dres) *(uint4 *)(dres + r * C + c0) = *(uint4 *)rv;
    }
}

int op_bn_backward_fused(const void *x, const void *y, const void *dy, void *dx, void *dres, int dt, int64_t M, int C,
                         float *bsum, const float *gamma, float *dgamma, float *dbeta, bool relu, bool acc_dx,
                         bool acc_dres, bool sums_ready, const BnWork &w, cudaStream_t s) {
    B2_CHECK(dt == DT_BF16 && C % 8 == 0, B2_ERR_ARG, "bn fused: bf16 activations with C %% 8 == 0 only");
    if (!sums_ready && M <= 1024 && C >= 512) {      // measured (r01): wins only for the 1x1-spatial stage
        B2_LAUNCH((k_bn_bwd_small<bf16>), C / 8, 256, 0, s, (const bf16 *)x, (const bf16 *)y, (const bf16 *)dy, (bf16 *)dx,
                  (bf16 *)dres, w.mean, w.invstd, gamma, dgamma, dbeta, M, C, relu, acc_dx, acc_dres);
        return 0;
    }
    if (!sums_ready) {
        // (only BatchNorms whose output gradient is not finalised by a tcgen05 dgrad: the stem's, the last block's)
        int opb = C / 8 < 32 ? C / 8 : 32;
        B2_CHECK((C / 8) % opb == 0 && 256 % opb == 0, B2_ERR_ARG, "bn fused: unsupported channel count %d", C);
        int gx = (C / 8) / opb;
        int lanes = 256 / opb;
        int64_t want = (kNumSMs * 2 + gx - 1) / gx;
        int64_t maxb = ceil_div(M, (int64_t)lanes * 4); if (want > maxb) want = maxb; if (want < 1) want = 1;
        dim3 grid(gx, (unsigned)want);
        B2_LAUNCH((k_bn_bwd_reduce_atomic<bf16>), grid, 256, 0, s, (const bf16 *)x, (const bf16 *)y, (const bf16 *)dy, w.mean, w.invstd,
                  relu, M, C, opb, bsum);
    }
    int64_t total = M * C;
    B2_CHECK(total < ((int64_t)1 << 32), B2_ERR_UNSUPPORTED, "bn fused: tensor of %lld elements (32-bit indexing)", (long long)total);
    int64_t blocks = ceil_div(total / 8, 256); if (blocks > kNumSMs * 2) blocks = kNumSMs * 2;
    B2_LAUNCH((k_bn_bwd_apply_fused<bf16, 8>), (int)blocks, 256, 3 * C * sizeof(float), s, (const bf16 *)x, (const bf16 *)y,
              (const bf16 *)dy, (bf16 *)dx, (bf16 *)dres, w.mean, w.invstd, bsum, gamma, dgamma, dbeta, M, relu, acc_dx, acc_dres,
              total, C);
    return 0;
}

}  // namespace b2

// =====================================================================================
// input pipeline / eval kernels
// =====================================================================================
namespace b2 {

__global__ void __launch_bounds__(256)
k_u8_to_normalized(const uint8_t *__restrict__ x, float *__restrict__ y, int N, int H, int W, int C, float m0, float m1, float m2,
                   float s0, float s1, float s2) {
    pdl_sync(); B2_TRACE_BEGIN();
    const int64_t total = (int64_t)N * C * H * W;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int w = i % W; int64_t r = i / W;
        int h = r % H; r /= H;
        int c = r % C; int n = (int)(r / C);
        float v = (float)x[(((int64_t)n * H + h) * W + w) * C + c] / 255.0f;      // ToTensor
        float m = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
        y[i] = (v - m) / sd;                                                       // Normalize
    }
}
int op_u8_to_normalized(const uint8_t *x, float *y, int N, int H, int W, int C, const float *mean3, const float *std3, cudaStream_t s) {
    B2_CHECK(C == 3, B2_ERR_UNSUPPORTED, "u8 preprocess: 3-channel images only");
    int64_t total = (int64_t)N * C * H * W;
    int blocks = (int)ceil_div(total, 256); if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    B2_LAUNCH(k_u8_to_normalized, blocks, 256, 0, s, x, y, N, H, W, C, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
    return 0;
}

__global__ void __launch_bounds__(256)
k_argmax_count(const float *__restrict__ logits, const int64_t *__restrict__ labels, int B, int C, long long *counts) {
    pdl_sync(); B2_TRACE_BEGIN();
    int correct = 0;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const float *row = logits + (int64_t)b * C;
        int best = 0; float bv = row[0];
        for (int j = 1; j < C; ++j) if (row[j] > bv) { bv = row[j]; best = j; }
        correct += (best == (int)labels[b]);
    }
    correct = __reduce_add_sync(0xffffffffu, correct);
    if ((threadIdx.x & 31) == 0 && correct) atomicAdd((unsigned long long *)counts, (unsigned long long)correct);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd((unsigned long long *)counts + 1, (unsigned long long)B);
}
int op_argmax_count(const float *logits, const int64_t *labels, int B, int C, long long *counts, cudaStream_t s) {
    B2_CHECK(B > 0 && C > 0, B2_ERR_ARG, "argmax_count: empty batch");
    int blocks = (int)ceil_div(B, 256); if (blocks > 64) blocks = 64;
    B2_LAUNCH(k_argmax_count, blocks, 256, 0, s, logits, labels, B, C, counts);
    return 0;
}

}  // namespace b2

namespace b2 { B2_TRACE_INSTALLER(trace_install_elem) }
