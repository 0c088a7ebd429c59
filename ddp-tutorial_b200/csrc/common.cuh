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

// Programmatic dependent launch: every kernel is launched with the stream-serialization
// attribute and begins with pdl_sync() (launch_dependents + wait), so the NEXT kernel's launch
// latency and prologue overlap this kernel's execution while all of its global-memory traffic
// still happens after this grid has completed and flushed.
extern bool g_pdl;
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_trigger(); pdl_wait(); }

// ---- in-graph kernel trace (diagnostic; off unless b2_trace_start() installed a buffer) ----------------
// Block 0 / thread 0 of a kernel appends (tag, %globaltimer) records: tag = TU id << 24 | source line << 4 | kind
// (0 = past griddepcontrol.wait, 1 = end of block 0).  Works inside replayed CUDA graphs and on every stream, which
// is what cudaEvents / ncu cannot show: what overlaps what in one real training step (tools/step_trace.py).
// The symbols are per translation unit (no -rdc); B2_TRACE_TU defines that TU's id and its installer.
static __device__ unsigned long long *d_trace_buf;
static __device__ unsigned int *d_trace_cnt;
static __device__ unsigned int d_trace_cap;
#ifndef B2_TRACE_TU_ID
#define B2_TRACE_TU_ID 0
#endif
__device__ __forceinline__ void trace_mark(int line, int kind) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) {
        unsigned long long *buf = d_trace_buf;
        if (buf) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            const unsigned int i = atomicAdd(d_trace_cnt, 1u);
            if (i < d_trace_cap) {
                buf[2 * i] = ((unsigned long long)B2_TRACE_TU_ID << 24) | ((unsigned long long)line << 4) | (unsigned)kind;
                buf[2 * i + 1] = t;
            }
        }
    }
}
#define B2_TRACE_BEGIN() b2::trace_mark(__LINE__, 0)
#define B2_TRACE_END() b2::trace_mark(__LINE__, 1)
#define B2_TRACE_INSTALLER(name)                                                                     \
    int name(unsigned long long *buf, unsigned int *cnt, unsigned int cap) {                           \
        if (cudaMemcpyToSymbol(d_trace_cnt, &cnt, sizeof(cnt)) != cudaSuccess) return B2_ERR_CUDA;     \
        if (cudaMemcpyToSymbol(d_trace_cap, &cap, sizeof(cap)) != cudaSuccess) return B2_ERR_CUDA;     \
        if (cudaMemcpyToSymbol(d_trace_buf, &buf, sizeof(buf)) != cudaSuccess) return B2_ERR_CUDA;     \
        return 0;                                                                                      \
    }
int trace_install_comm(unsigned long long *buf, unsigned int *cnt, unsigned int cap);
int trace_install_simt(unsigned long long *buf, unsigned int *cnt, unsigned int cap);
int trace_install_umma(unsigned long long *buf, unsigned int *cnt, unsigned int cap);
int trace_install_elem(unsigned long long *buf, unsigned int *cnt, unsigned int cap);

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                                unsigned cluster_x, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (g_pdl) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster_x > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = at; cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                        Args &&...args) {
    return launch_kernel_cluster(kernel, grid, block, smem, s, 1u, static_cast<Args &&>(args)...);
}

// every launch goes through this so b2_launch_count() is honest
#define B2_LAUNCH(kernel, grid, block, smem, stream, ...)                          \
    do {                                                                           \
        cudaError_t _le = b2::launch_kernel(kernel, dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(stream), __VA_ARGS__); \
        ++b2::g_launches;                                                          \
        if (_le != cudaSuccess) {                                                  \
            b2::set_error("%s:%d: launch of %s failed: %s", __FILE__, __LINE__, #kernel, cudaGetErrorString(_le)); \
            return B2_ERR_CUDA;                                                    \
        }                                                                          \
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
