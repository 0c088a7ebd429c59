// Peer-memory communicator: the DDP gradient exchange without NCCL.
//
// Replaces, for /root/reference/cifar_example_ddp.py:
//   :57-58  init_process_group('nccl') + barrier()         -> IPC peer mapping + flag barrier
//   :83,106 c10d::Reducer bucket all-reduce (reducer.hpp:524, div_factor_ :497-500)
//   :107    optimizer.step() (sgd.py)                        -> fused into the same kernel
//
// Every rank maps every peer's gradient / parameter / flag buffers (cudaIpc over NVLink 5 /
// NVSwitch: uniform bandwidth to every peer).  One kernel per bucket does
//   one-shot : every rank reads all W gradient copies, sums them in rank order, applies SGD
//              locally (small buckets: one NVLink round trip, one barrier pair)
//   two-shot : rank r reduces slice r from all W peers (reduce-scatter), applies the
//              SGD-momentum update to its slice, then stores the updated *parameters* of the
//              slice into all W replicas (all-gather of params, not grads)
// so neither a separate all-reduce nor a separate elementwise optimizer kernel exists.
// Summation order is fixed (rank 0..W-1) => replicas stay bit-identical.
#include <stdlib.h>
#include <string.h>
#include <vector>

#define B2_TRACE_TU_ID 1
#include "ops.cuh"

namespace b2 {

constexpr int kMaxWorld = 8;
constexpr int kCommBlocks = 148;      // most blocks a comm kernel launches (flag slots; same on every rank)
constexpr int kCommThreadsMax = 512;
// Launch shape of the exchange kernels (B2_COMM_THREADS / B2_COMM_BLOCKS override; same value on every rank).
// Measured at W=2 on ResNet-18 (profiles/r02_comm_shape_sweep.txt): 96 x 512 beats 148 x 128 and 96/148 x 256 although
// only the small CTAs fit beside a persistent conv CTA (352 threads x 120 registers): the exchange slows the step mostly
// through the memory system (latency-bound kernels next to it run 2-3x longer, tools/step_trace.py), so the shortest
// exchange wins, not the most polite one.
static int comm_threads() {
    static const int v = [] { const char *e = getenv("B2_COMM_THREADS"); int t = e ? atoi(e) : 0;
                              return (t >= 32 && t <= kCommThreadsMax && t % 32 == 0) ? t : 512; }();
    return v;
}
static int comm_blocks_env() {          // 0 = not set
    static const int v = [] { const char *e = getenv("B2_COMM_BLOCKS"); int t = e ? atoi(e) : 0;
                              return (t > 0 && t <= kCommBlocks) ? t : 0; }();
    return v;
}
// Grid of one exchange kernel.  A stand-alone all-reduce (MODE 0) wants the whole machine: 96 CTAs.  The fused
// all-reduce + SGD of a gradient bucket (MODE 1) runs UNDER backward with hundreds of microseconds to spare, and every
// SM that hosts one of its CTAs serves the latency-bound kernels beside it 2-10x slower (their loads queue behind
// ~2000 outstanding NVLink requests per comm CTA; tools/step_trace.py showed single BatchNorm launches of 68 us instead
// of 6.5 us).  Fewer CTAs = fewer slow SMs: measured step time at W=8 (NVLS, ResNet-18): 96 CTAs 0.996 ms, 64 0.995,
// 48 0.984, 32 0.972, 24 0.960; at W=2: 96 0.977, 48 0.951 (profiles/r02_comm_shape_sweep.txt).  The slice a rank owns
// shrinks with W, so the grid does too.
static int comm_blocks(int mode, int world, bool nvls) {
    if (comm_blocks_env()) return comm_blocks_env();
    if (mode == 0 || !nvls) return 96;
    return world <= 2 ? 48 : (world <= 4 ? 32 : 24);
}
// flag layout (uint32 words): [phase 0/1][src rank][block]
constexpr int kFlagStride = kCommBlocks;
constexpr int kOneShotPerThread = 4;  // plain one-shot all-reduce keeps its sums in registers across the barrier
constexpr int kMcBatch = 4;           // multimem.ld_reduce in flight per thread

struct CommPtrs {
    float *grads[kMaxWorld];
    float *params[kMaxWorld];
    uint32_t *flags[kMaxWorld];
};

struct CommState {
    bool inited = false;
    int rank = 0, world = 1, device = 0;
    bool registered = false;
    CommPtrs ptrs{};
    int64_t n_elems = 0;
    float *mc_grads = nullptr, *mc_params = nullptr;   // NVSwitch multicast addresses of the same buffers (NVLS), or null
    bool use_nvls = false;
    uint32_t *epoch = nullptr;     // local device memory: per-block barrier epoch [2][kCommBlocks]
    int *err = nullptr;            // local device word: set by a kernel whose flag wait timed out
    std::vector<void *> allocs;
};
static CommState g_comm;

// ---- system-scope flag primitives ------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float *p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer_f4(float *p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// All W ranks launch the same grid; block b of every rank meets block b of every other rank.
// phase selects one of two flag banks so that back-to-back barriers never alias.
// FENCE: publish this CTA's earlier (peer) stores system-wide before signalling; not needed when
// the barrier only says "I have arrived / I am done reading".
template <int W, bool FENCE = true>
__device__ __forceinline__ void peer_barrier(const CommPtrs &P, int rank, uint32_t *epoch_mem, int phase, int *err) {
    __syncthreads();
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) {
        uint32_t e = epoch_mem[phase * kCommBlocks + blockIdx.x] + 1;
        epoch_mem[phase * kCommBlocks + blockIdx.x] = e;
        s_epoch = e;
    }
    __syncthreads();                        // every thread's earlier (peer) stores are ordered before the signalling threads' release
    const uint32_t e = s_epoch;
    if (threadIdx.x < W) {
        const int peer = threadIdx.x;
        // one cumulative system-scope fence per signalling thread publishes the whole CTA's stores (bar.sync above
        // orders them before it) -- a fence in all 512 threads costs several microseconds per barrier
        if (FENCE) __threadfence_system();
        uint32_t *remote = P.flags[peer] + ((phase * kMaxWorld + rank) * kFlagStride + blockIdx.x);
        st_release_sys(remote, e);
        const uint32_t *mine = P.flags[rank] + ((phase * kMaxWorld + peer) * kFlagStride + blockIdx.x);
        long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
            // ~2 min without the peer: a rank died or diverged (a rank that is merely late -- host-side work between
            // steps -- must not trip this).  Continuing would reduce stale data and let the replicas drift apart
            // silently, so record the reason and kill the context (every later CUDA call on this rank fails loudly;
            // b2_comm_check() reports the flag when the host polls first).
            if (clock64() - t0 > 230000000000LL) { *err = 1; __threadfence_system(); __trap(); }
            __nanosleep(64);        // poll gently: the flag lives in local memory next to a training step's traffic
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void st_bf16x4(bf16 *dst, float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk; pk.x = *(uint32_t *)&lo; pk.y = *(uint32_t *)&hi;
    *(uint2 *)dst = pk;
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }

__device__ __forceinline__ void sgd_apply(float4 &p, float4 &m, const float4 g, float lr, float mom, int first) {
    if (first) m = g;
    else {
        m.x = __fadd_rn(__fmul_rn(m.x, mom), g.x); m.y = __fadd_rn(__fmul_rn(m.y, mom), g.y);
        m.z = __fadd_rn(__fmul_rn(m.z, mom), g.z); m.w = __fadd_rn(__fmul_rn(m.w, mom), g.w);
    }
    p.x = fmaf(-lr, m.x, p.x); p.y = fmaf(-lr, m.y, p.y); p.z = fmaf(-lr, m.z, p.z); p.w = fmaf(-lr, m.w, p.w);
}

// MODE 0: plain all-reduce (result * scale written to every rank's grad buffer)
// MODE 1: fused all-reduce/W + SGD-momentum (params updated on every rank)
template <int W, int MODE, bool ONESHOT>
__global__ void __launch_bounds__(kCommThreadsMax)
k_allreduce(const CommPtrs P, int rank, float *__restrict__ mom_buf, int64_t off, int64_t n4 /* float4 count */,
            float scale, float lr, float momentum, int first, bf16 *__restrict__ mirror, uint32_t *epoch_mem, int *err) {
    pdl_sync(); B2_TRACE_BEGIN();
    peer_barrier<W, false>(P, rank, epoch_mem, 0, err);      // earlier kernels' stores are already visible
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ONESHOT) {
        if (MODE == 1) {
            for (int64_t i = tid; i < n4; i += stride) {
                const int64_t e = off + i * 4;
                float4 acc = ld_peer_f4(P.grads[0] + e);
#pragma unroll
                for (int r = 1; r < W; ++r) acc = f4_add(acc, ld_peer_f4(P.grads[r] + e));
                acc = f4_scale(acc, scale);
                float4 p = *(float4 *)(P.params[rank] + e), m = first ? acc : *(float4 *)(mom_buf + e);
                sgd_apply(p, m, acc, lr, momentum, first);
                *(float4 *)(mom_buf + e) = m;
                *(float4 *)(P.params[rank] + e) = p;
                if (mirror) st_bf16x4(mirror + e, p);
            }
            peer_barrier<W, false>(P, rank, epoch_mem, 1, err);      // peers are done reading my grads
        } else {
            // at most kOneShotPerThread float4 per thread (host guarantees n4 <= that * grid * block): read all, barrier, write own
            float4 acc[kOneShotPerThread];
#pragma unroll
            for (int u = 0; u < kOneShotPerThread; ++u) {
                const int64_t i = tid + u * stride;
                acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < n4) {
                    const int64_t e = off + i * 4;
                    acc[u] = ld_peer_f4(P.grads[0] + e);
#pragma unroll
                    for (int r = 1; r < W; ++r) acc[u] = f4_add(acc[u], ld_peer_f4(P.grads[r] + e));
                    acc[u] = f4_scale(acc[u], scale);
                }
            }
            peer_barrier<W, false>(P, rank, epoch_mem, 1, err);
#pragma unroll
            for (int u = 0; u < kOneShotPerThread; ++u) {
                const int64_t i = tid + u * stride;
                if (i < n4) *(float4 *)(P.grads[rank] + off + i * 4) = acc[u];
            }
        }
    } else {
        // slice owned by this rank, in float4 units
        const int64_t per = (n4 + W - 1) / W;
        const int64_t lo = (int64_t)rank * per;
        int64_t hi = lo + per; if (hi > n4) hi = n4;
        constexpr int U = W >= 8 ? 1 : (W == 4 ? 2 : 4);       // U * W peer loads in flight per thread (the asm loads are volatile: batch them by hand)
        for (int64_t i0 = lo + tid; i0 < hi; i0 += stride * U) {
            float4 part[U][W];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + u * stride;
                if (i < hi) {
#pragma unroll
                    for (int r = 0; r < W; ++r) part[u][r] = ld_peer_f4(P.grads[r] + off + i * 4);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + u * stride;
                if (i >= hi) break;
                const int64_t e = off + i * 4;
                float4 acc = part[u][0];
#pragma unroll
                for (int r = 1; r < W; ++r) acc = f4_add(acc, part[u][r]);
                acc = f4_scale(acc, scale);
                if (MODE == 1) {
                    float4 p = *(float4 *)(P.params[rank] + e), m = first ? acc : *(float4 *)(mom_buf + e);
                    sgd_apply(p, m, acc, lr, momentum, first);
                    *(float4 *)(mom_buf + e) = m;
#pragma unroll
                    for (int r = 0; r < W; ++r) st_peer_f4(P.params[r] + e, p);
                } else {
#pragma unroll
                    for (int r = 0; r < W; ++r) st_peer_f4(P.grads[r] + e, acc);
                }
            }
        }
        peer_barrier<W>(P, rank, epoch_mem, 1, err);          // all slices have landed everywhere
        if (MODE == 1 && mirror) {
            // every replica now holds the full updated fp32 range: refresh the local bf16 mirror
            for (int64_t i0 = tid; i0 < n4; i0 += stride * 4) {          // 4 loads in flight (the asm loads are volatile)
                float4 pv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int64_t i = i0 + u * stride; if (i < n4) pv[u] = ld_peer_f4(P.params[rank] + off + i * 4); }
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int64_t i = i0 + u * stride; if (i < n4) st_bf16x4(mirror + off + i * 4, pv[u]); }
            }
        }
    }
    __syncthreads();
    B2_TRACE_END();
}


// ---- NVLS: in-switch reduction / broadcast through the multicast mapping (comm_vmm.cu) ---------------
__device__ __forceinline__ float4 mc_ld_reduce_f4(const float *mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_f4(float *mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Same contract as k_allreduce, but the W-way sum is ONE multimem.ld_reduce per float4 (the NVSwitch reads the W
// replicas and adds them in fp32) and the result / updated parameters go out with ONE multimem.st (the switch
// writes all W replicas): each GPU moves 1/W of the bytes over its own links instead of (W-1)/W.
template <int W, int MODE, bool ONESHOT>
__global__ void __launch_bounds__(kCommThreadsMax)
k_allreduce_nvls(const CommPtrs P, int rank, float *__restrict__ mc_grads, float *__restrict__ mc_params, float *__restrict__ mom_buf,
                 int64_t off, int64_t n4, float scale, float lr, float momentum, int first, bf16 *__restrict__ mirror,
                 uint32_t *epoch_mem, int *err) {
    pdl_sync(); B2_TRACE_BEGIN();
    peer_barrier<W, false>(P, rank, epoch_mem, 0, err);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ONESHOT) {
        // small buckets: every rank asks the switch for the whole sum and updates its own replica
        if (MODE == 1) {
            for (int64_t i0 = tid; i0 < n4; i0 += stride * kMcBatch) {
                float4 red[kMcBatch];
#pragma unroll
                for (int u = 0; u < kMcBatch; ++u) {
                    const int64_t i = i0 + u * stride;
                    if (i < n4) red[u] = mc_ld_reduce_f4(mc_grads + off + i * 4);
                }
#pragma unroll
                for (int u = 0; u < kMcBatch; ++u) {
                    const int64_t i = i0 + u * stride;
                    if (i >= n4) break;
                    const int64_t e = off + i * 4;
                    float4 acc = f4_scale(red[u], scale);
                    float4 p = *(float4 *)(P.params[rank] + e), m = first ? acc : *(float4 *)(mom_buf + e);
                    sgd_apply(p, m, acc, lr, momentum, first);
                    *(float4 *)(mom_buf + e) = m;
                    *(float4 *)(P.params[rank] + e) = p;
                    if (mirror) st_bf16x4(mirror + e, p);
                }
            }
            peer_barrier<W, false>(P, rank, epoch_mem, 1, err);      // peers are done reading my grads
        } else {
            float4 acc[kOneShotPerThread];
#pragma unroll
            for (int u = 0; u < kOneShotPerThread; ++u) {
                const int64_t i = tid + u * stride;
                acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < n4) acc[u] = f4_scale(mc_ld_reduce_f4(mc_grads + off + i * 4), scale);
            }
            peer_barrier<W, false>(P, rank, epoch_mem, 1, err);
#pragma unroll
            for (int u = 0; u < kOneShotPerThread; ++u) {
                const int64_t i = tid + u * stride;
                if (i < n4) *(float4 *)(P.grads[rank] + off + i * 4) = acc[u];
            }
        }
    } else {
        const int64_t per = (n4 + W - 1) / W;
        const int64_t lo = (int64_t)rank * per;
        int64_t hi = lo + per; if (hi > n4) hi = n4;
        for (int64_t i0 = lo + tid; i0 < hi; i0 += stride * kMcBatch) {
            float4 red[kMcBatch];                               // kMcBatch switch reductions in flight per thread
#pragma unroll
            for (int u = 0; u < kMcBatch; ++u) {
                const int64_t i = i0 + u * stride;
                if (i < hi) red[u] = mc_ld_reduce_f4(mc_grads + off + i * 4);
            }
#pragma unroll
            for (int u = 0; u < kMcBatch; ++u) {
                const int64_t i = i0 + u * stride;
                if (i >= hi) break;
                const int64_t e = off + i * 4;
                float4 acc = f4_scale(red[u], scale);
                if (MODE == 1) {
                    float4 p = *(float4 *)(P.params[rank] + e), m = first ? acc : *(float4 *)(mom_buf + e);
                    sgd_apply(p, m, acc, lr, momentum, first);
                    *(float4 *)(mom_buf + e) = m;
                    mc_st_f4(mc_params + e, p);                     // all W replicas, mine included
                } else {
                    mc_st_f4(mc_grads + e, acc);
                }
            }
        }
        peer_barrier<W>(P, rank, epoch_mem, 1, err);          // all slices have landed everywhere
        if (MODE == 1 && mirror) {
            for (int64_t i0 = tid; i0 < n4; i0 += stride * 4) {          // 4 loads in flight (the asm loads are volatile)
                float4 pv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int64_t i = i0 + u * stride; if (i < n4) pv[u] = ld_peer_f4(P.params[rank] + off + i * 4); }
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int64_t i = i0 + u * stride; if (i < n4) st_bf16x4(mirror + off + i * 4, pv[u]); }
            }
        }
    }
    __syncthreads();
    B2_TRACE_END();
}

template <int W>
__global__ void __launch_bounds__(kCommThreadsMax)
k_barrier_only(const CommPtrs P, int rank, uint32_t *epoch_mem, int *err) {
    pdl_sync(); B2_TRACE_BEGIN();
    peer_barrier<W>(P, rank, epoch_mem, 0, err);
    peer_barrier<W>(P, rank, epoch_mem, 1, err);
}

template <int MODE>
static int launch_allreduce(int64_t off, int64_t n, float *m, float scale, float lr, float momentum, int first,
                            int algo, bf16 *mirror, cudaStream_t s) {
    CommState &c = g_comm;
    B2_CHECK(c.inited && c.registered, B2_ERR_STATE, "comm: not initialised / no buffers registered");
    B2_CHECK(off >= 0 && n >= 0 && off + n <= c.n_elems, B2_ERR_ARG, "comm: range [%lld,+%lld) outside registered %lld",
             (long long)off, (long long)n, (long long)c.n_elems);
    B2_CHECK(off % 4 == 0, B2_ERR_ARG, "comm: offset must be a multiple of 4 elements");
    if (n == 0) return 0;
    const int64_t n4 = (n + 3) / 4;            // buffers are padded to 4 elements by the allocator
    B2_CHECK(off + n4 * 4 <= (c.n_elems + 3) / 4 * 4, B2_ERR_ARG, "comm: padded range exceeds buffer");
    const int threads = comm_threads();
    const bool nvls = c.use_nvls && c.mc_grads && (MODE == 0 || c.mc_params) && c.world > 1;
    const int env_blocks = comm_blocks(MODE, c.world, nvls);
    bool oneshot;
    if (algo == 1) oneshot = true;
    else if (algo == 2) oneshot = false;
    else oneshot = (n * 4 * (int64_t)c.world <= ((int64_t)8 << 20));   // bytes each rank reads in one-shot mode
    if (MODE == 0 && oneshot && n4 > (int64_t)kOneShotPerThread * env_blocks * threads) oneshot = false;
    if (c.world == 1) oneshot = true;
    // small messages: only as many blocks as there is data (same on every rank: a function of n)
    int blocks = env_blocks;            // two-shot kernels stride over their slice, so any grid <= kCommBlocks works (same on every rank)
    if (oneshot) { int64_t need = (n4 + threads - 1) / threads; if (need < blocks) blocks = need < 1 ? 1 : (int)need; }
#define LAUNCH_W(Wv)                                                                                              \
    if (nvls && oneshot) { B2_LAUNCH((k_allreduce_nvls<Wv, MODE, true>), blocks, threads, 0, s, c.ptrs, c.rank, c.mc_grads, c.mc_params, m, \
                                     off, n4, scale, lr, momentum, first, mirror, c.epoch, c.err); }                  \
    else if (nvls) { B2_LAUNCH((k_allreduce_nvls<Wv, MODE, false>), blocks, threads, 0, s, c.ptrs, c.rank, c.mc_grads, c.mc_params, m, \
                               off, n4, scale, lr, momentum, first, mirror, c.epoch, c.err); }                        \
    else if (oneshot) { B2_LAUNCH((k_allreduce<Wv, MODE, true>), blocks, threads, 0, s, c.ptrs, c.rank, m, off, n4, \
                             scale, lr, momentum, first, mirror, c.epoch, c.err); }                                   \
    else { B2_LAUNCH((k_allreduce<Wv, MODE, false>), blocks, threads, 0, s, c.ptrs, c.rank, m, off, n4, scale,   \
                     lr, momentum, first, mirror, c.epoch, c.err); }
    switch (c.world) {
    case 1: LAUNCH_W(1); break;
    case 2: LAUNCH_W(2); break;
    case 4: LAUNCH_W(4); break;
    case 8: LAUNCH_W(8); break;
    default: set_error("comm: world size %d not supported (1,2,4,8)", c.world); return B2_ERR_UNSUPPORTED;
    }
#undef LAUNCH_W
    return 0;
}

// entry for the model's bucketed backward (model.cu)
int comm_allreduce_sgd_range(float *m, int64_t off, int64_t n, float lr, float momentum, int first, int algo, bf16 *mirror,
                             cudaStream_t s) {
    B2_CHECK(g_comm.inited && g_comm.ptrs.params[0] != nullptr, B2_ERR_STATE, "allreduce_sgd: communicator has no parameter buffers");
    return launch_allreduce<1>(off, n, m, 1.f / (float)g_comm.world, lr, momentum, first, algo, mirror, s);
}
bool comm_ready() { return g_comm.inited && g_comm.registered; }

}  // namespace b2

using namespace b2;

extern "C" {

int b2_comm_init(int rank, int world, int device) {
    CommState &c = g_comm;
    B2_CHECK(!c.inited, B2_ERR_STATE, "comm: already initialised");
    B2_CHECK(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, B2_ERR_ARG, "comm: bad rank %d / world %d", rank, world);
    B2_CUDA(cudaSetDevice(device));
    c.rank = rank; c.world = world; c.device = device;
    B2_CUDA(cudaMalloc(&c.epoch, 2 * kCommBlocks * sizeof(uint32_t)));
    B2_CUDA(cudaMemset(c.epoch, 0, 2 * kCommBlocks * sizeof(uint32_t)));
    B2_CUDA(cudaMalloc(&c.err, sizeof(int)));
    B2_CUDA(cudaMemset(c.err, 0, sizeof(int)));
    c.inited = true;
    return 0;
}

int b2_comm_finalize(void) {
    CommState &c = g_comm;
    if (!c.inited) return 0;
    cudaDeviceSynchronize();
    for (void *p : c.allocs) cudaFree(p);
    c.allocs.clear();
    if (c.epoch) cudaFree(c.epoch);
    if (c.err) cudaFree(c.err);
    c = CommState();
    return 0;
}

int b2_comm_alloc(int64_t bytes, void **ptr) {
    B2_CHECK(ptr && bytes > 0, B2_ERR_ARG, "comm_alloc: bad arguments");
    bytes = (bytes + 255) / 256 * 256;
    void *p = nullptr;
    B2_CUDA(cudaMalloc(&p, bytes));
    B2_CUDA(cudaMemset(p, 0, bytes));
    B2_CUDA(cudaDeviceSynchronize());
    g_comm.allocs.push_back(p);
    *ptr = p;
    return 0;
}
int b2_comm_free(void *ptr) {
    CommState &c = g_comm;
    for (size_t i = 0; i < c.allocs.size(); ++i)
        if (c.allocs[i] == ptr) { cudaFree(ptr); c.allocs.erase(c.allocs.begin() + i); return 0; }
    set_error("comm_free: pointer not owned by the library");
    return B2_ERR_ARG;
}
int b2_comm_export(void *ptr, void *handle_out) {
    B2_CHECK(ptr && handle_out, B2_ERR_ARG, "comm_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == B2_IPC_HANDLE_BYTES, "ipc handle size");
    cudaIpcMemHandle_t h;
    B2_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}
int b2_comm_import(const void *handle, void **peer_ptr) {
    B2_CHECK(handle && peer_ptr, B2_ERR_ARG, "comm_import: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    B2_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int b2_comm_unimport(void *peer_ptr) {
    if (peer_ptr) B2_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return 0;
}

int b2_comm_register(void *const *grads, void *const *params, void *const *flags, int64_t n_elems) {
    CommState &c = g_comm;
    B2_CHECK(c.inited, B2_ERR_STATE, "comm: not initialised");
    B2_CHECK(grads && flags && n_elems > 0, B2_ERR_ARG, "comm_register: bad arguments");
    for (int r = 0; r < c.world; ++r) {
        B2_CHECK(grads[r] && flags[r], B2_ERR_ARG, "comm_register: NULL buffer for rank %d", r);
        c.ptrs.grads[r] = (float *)grads[r];
        c.ptrs.params[r] = params ? (float *)params[r] : nullptr;
        c.ptrs.flags[r] = (uint32_t *)flags[r];
    }
    c.n_elems = n_elems;
    c.registered = true;
    c.mc_grads = c.mc_params = nullptr; c.use_nvls = false;      // a new registration starts without multicast
    return 0;
}

/* NVLS: multicast addresses of the SAME gradient / parameter buffers registered above (comm_vmm.cu); NULL disables */
int b2_comm_register_multicast(void *mc_grads, void *mc_params) {
    CommState &c = g_comm;
    B2_CHECK(c.inited && c.registered, B2_ERR_STATE, "comm: register the peer buffers first");
    c.mc_grads = (float *)mc_grads; c.mc_params = (float *)mc_params;
    c.use_nvls = mc_grads != nullptr;
    return 0;
}
/* 1: all-reduce through the switch (multimem) when multicast addresses are registered; 0: peer loads/stores */
int b2_comm_set_nvls(int enable) { g_comm.use_nvls = enable != 0 && g_comm.mc_grads != nullptr; return 0; }
int b2_comm_nvls_active(void) { return g_comm.use_nvls ? 1 : 0; }

int b2_comm_flag_bytes(void) { return 2 * kMaxWorld * kFlagStride * (int)sizeof(uint32_t); }

int b2_comm_check(void) {
    // host-side: did any comm kernel time out on a flag wait?  (synchronises the device)
    CommState &c = g_comm;
    B2_CHECK(c.inited, B2_ERR_STATE, "comm: not initialised");
    int e = 0;
    B2_CUDA(cudaMemcpy(&e, c.err, sizeof(int), cudaMemcpyDeviceToHost));
    B2_CHECK(e == 0, B2_ERR_TIMEOUT, "comm: a peer flag wait timed out (a rank died or diverged)");
    return 0;
}

int b2_comm_barrier(void *stream) {
    CommState &c = g_comm;
    B2_CHECK(c.inited && c.registered, B2_ERR_STATE, "comm: not initialised / no buffers registered");
    switch (c.world) {
    case 1: return 0;
    case 2: B2_LAUNCH((k_barrier_only<2>), comm_blocks(0, c.world, false), comm_threads(), 0, stream, c.ptrs, c.rank, c.epoch, c.err); break;
    case 4: B2_LAUNCH((k_barrier_only<4>), comm_blocks(0, c.world, false), comm_threads(), 0, stream, c.ptrs, c.rank, c.epoch, c.err); break;
    case 8: B2_LAUNCH((k_barrier_only<8>), comm_blocks(0, c.world, false), comm_threads(), 0, stream, c.ptrs, c.rank, c.epoch, c.err); break;
    default: set_error("comm: world size %d not supported", c.world); return B2_ERR_UNSUPPORTED;
    }
    return 0;
}

int b2_allreduce(int64_t n_elems, float scale, int algo, void *stream) {
    return launch_allreduce<0>(0, n_elems, nullptr, scale, 0.f, 0.f, 0, algo, nullptr, (cudaStream_t)stream);
}
int b2_allreduce_range(int64_t offset, int64_t n_elems, float scale, int algo, void *stream) {
    return launch_allreduce<0>(offset, n_elems, nullptr, scale, 0.f, 0.f, 0, algo, nullptr, (cudaStream_t)stream);
}

int b2_allreduce_sgd(float *m, int64_t n_elems, float lr, float momentum, int first_step, int algo, void *p_bf16,
                     void *stream) {
    B2_CHECK(m, B2_ERR_ARG, "allreduce_sgd: NULL momentum buffer");
    return comm_allreduce_sgd_range(m, 0, n_elems, lr, momentum, first_step, algo, (bf16 *)p_bf16, (cudaStream_t)stream);
}
int b2_allreduce_sgd_range(float *m, int64_t offset, int64_t n_elems, float lr, float momentum, int first_step, int algo,
                           void *p_bf16, void *stream) {
    B2_CHECK(m, B2_ERR_ARG, "allreduce_sgd: NULL momentum buffer");
    return comm_allreduce_sgd_range(m, offset, n_elems, lr, momentum, first_step, algo, (bf16 *)p_bf16, (cudaStream_t)stream);
}

}  // extern "C"

namespace b2 { B2_TRACE_INSTALLER(trace_install_comm) }
