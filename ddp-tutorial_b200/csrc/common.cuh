// Shared helpers for the b2ddp kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/b2ddp.h"

namespace b2 {

typedef __nv_bfloat16 bf16;

// ---- error plumbing -------------------------------------------------------
void set_error(const char *fmt, ...);
extern thread_local char g_err[1024];
extern int64_t g_launches;

#define B2_CUDA(expr)                                                              \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            b2::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,            \
                          cudaGetErrorString(_e));                                 \
            return B2_ERR_CUDA;                                                    \
        }                                                                          \
    } while (0)

#define B2_CHECK(cond, code, ...)                                                  \
    do {                                                                           \
        if (!(cond)) {                                                             \
            b2::set_error(__VA_ARGS__);                                            \
            return (code);                                                         \
        }                                                                          \
    } while (0)

#define B2_TRY(expr)                                                               \
    do {                                                                           \
        int _r = (expr);                                                           \
        if (_r != 0) return _r;                                                    \
    } while (0)

// every launch goes through this so b2_launch_count() is honest
#define B2_LAUNCH(kernel, grid, block, smem, stream, ...)                          \
    do {                                                                           \
        kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);  \
        ++b2::g_launches;                                                          \
        B2_CUDA(cudaGetLastError());                                               \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- scalar conversion ----------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int kNumSMs = 148;
// fused-BatchNorm sum buffers are replicated to spread same-address atomics: layout [copies][2][C]
constexpr int kBnSumCopies = 8;

// ---- conv geometry shared by every implementation ---------------------------
struct ConvGeom {
    int N, Hi, Wi, Ci;   // input  NHWC
    int Ho, Wo, Co;      // output NHWC
    int kh, kw, stride, pad;
    int K() const { return kh * kw * Ci; }
    int64_t M() const { return (int64_t)N * Ho * Wo; }
};

}  // namespace b2
