// Microbenchmark: per-SM TMA (cp.async.bulk.tensor) fill rate for the box shapes the conv kernels use.
// Question it answers: is the ~30 B/clk/SM the conv main loops see a property of
// 4-D boxes with 128-byte rows, of the L2, or of a single issuing thread?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_bench.bin tools/tma_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { while (!mbar_try(bar, parity)) {} }
__device__ __forceinline__ void tma2(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma4(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

struct Params {
    int mode;        // 0: A = 2-D box {64, 128 rows};  1: A = 4-D box {64, BW, BH, BN} with tap shifts
    int with_b;      // also load an 8 KB 2-D weight box per stage
    int iters;       // stages to stream per CTA
    int m_tiles;     // distinct A tiles (CTA b starts at tile b % m_tiles)
    int BN, h_blocks, BH;   // 4-D: images per tile, row blocks per image
    int c_chunks;    // channel chunks of 64
    int producers;   // 1 or 2 issuing threads (A by warp 0, B by warp 2 when 2)
};

template <int STAGES>
__global__ void __launch_bounds__(128) k_bench(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap b_map,
                                               const Params p, long long *cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    constexpr int kA = 16384, kB = 8192;
    uint8_t *sA = smem, *sB = smem + STAGES * kA;
    uint64_t *full = (uint64_t *)(sB + STAGES * kB), *empty = full + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], (p.with_b && p.producers == 2) ? 2 : 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int tile = blockIdx.x % p.m_tiles;
    if (warp == 0 && lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(&empty[stage], phase ^ 1);
            const bool both = p.with_b && p.producers == 1;
            mbar_expect(&full[stage], kA + (both ? kB : 0));
            const int tap = it % 9, cc = (it / 9) % p.c_chunks;
            if (p.mode == 0) tma2(sA + stage * kA, &a_map, &full[stage], (it % (9 * p.c_chunks)) * 64, tile * 128);
            else {
                const int n_blk = tile / p.h_blocks, h_blk = tile % p.h_blocks;
                tma4(sA + stage * kA, &a_map, &full[stage], cc * 64, tap % 3 - 1, h_blk * p.BH + tap / 3 - 1, n_blk * p.BN);
            }
            if (both) tma2(sB + stage * kB, &b_map, &full[stage], (it % (9 * p.c_chunks)) * 64, 0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 2 && lane == 0 && p.with_b && p.producers == 2) {
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect(&full[stage], kB);
            tma2(sB + stage * kB, &b_map, &full[stage], (it % (9 * p.c_chunks)) * 64, 0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1 && lane == 0) {
        int stage = 0; uint32_t phase = 0;
        long long t0 = 0;
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(&full[stage], phase);
            if (it == STAGES) t0 = clock64();          // steady state: skip the pipeline fill
            mbar_arrive(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn g_enc;

static void map2d(CUtensorMap *m, void *base, uint64_t rows, uint64_t cols, uint32_t box_rows, CUtensorMapL2promotion pr) {
    cuuint64_t dims[2] = {cols, rows}; cuuint64_t strides[1] = {cols * 2}; cuuint32_t box[2] = {64, box_rows}; cuuint32_t es[2] = {1, 1};
    CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode2d failed %d\n", (int)r); exit(1); }
}
static void map4d(CUtensorMap *m, void *base, int N, int H, int W, int C, int bw, int bh, int bn, CUtensorMapL2promotion pr) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn}; cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode4d failed %d\n", (int)r); exit(1); }
}


// ---- variant 2: P producer threads, each issuing 1/P of the A tile (a sub-box) and 1/P of the B tile
struct Params2 { int iters, m_tiles, BN, BH, h_blocks, c_chunks, P, f, bn_sub, bh_sub, same_warp, with_b; };

template <int STAGES>
__global__ void __launch_bounds__(576) k_bench2(const __grid_constant__ CUtensorMap a_sub, const __grid_constant__ CUtensorMap b_sub,
                                                const Params2 p, long long *cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    constexpr int kA = 16384, kB = 8192;
    uint8_t *sA = smem, *sB = smem + STAGES * kA;
    uint64_t *full = (uint64_t *)(sB + STAGES * kB), *empty = full + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], p.P); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int tile = blockIdx.x % p.m_tiles;
    // producer id: warps 1..P (lane 0) or lanes 0..P-1 of warp 1
    int pid = -1;
    if (p.same_warp) { if (warp == 1 && lane < p.P) pid = lane; }
    else if (warp >= 1 && warp <= p.P && lane == 0) pid = warp - 1;
    if (pid >= 0) {
        const int pn = pid / p.f, ph = pid % p.f;
        const int abytes = kA / p.P, bbytes = p.with_b ? kB / p.P : 0;
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect(&full[stage], abytes + bbytes);
            const int tap = it % 9, cc = (it / 9) % p.c_chunks;
            const int n_blk = tile / p.h_blocks, h_blk = tile % p.h_blocks;
            tma4(sA + stage * kA + pid * abytes, &a_sub, &full[stage], cc * 64, tap % 3 - 1, h_blk * p.BH + ph * p.bh_sub + tap / 3 - 1,
                 n_blk * p.BN + pn * p.bn_sub);
            if (p.with_b) tma2(sB + stage * kB + pid * bbytes, &b_sub, &full[stage], (it % (9 * p.c_chunks)) * 64, pid * (64 / p.P));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 0 && lane == 0) {
        int stage = 0; uint32_t phase = 0;
        long long t0 = 0;
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(&full[stage], phase);
            if (it == STAGES) t0 = clock64();
            mbar_arrive(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

template <int STAGES>
static void run2(const char *name, void *act, void *w, int N, int H, int W, int C, int BW, int BH, int BN, int P, int same_warp, int with_b,
                 int grid, long long *d_cyc) {
    Params2 p{};
    p.iters = 400; p.BN = BN; p.BH = BH; p.h_blocks = H / BH; p.m_tiles = (N / BN) * p.h_blocks; p.c_chunks = C / 64; p.P = P;
    p.same_warp = same_warp; p.with_b = with_b;
    p.bn_sub = BN / P > 0 ? BN / P : 1;
    p.f = P / (BN / p.bn_sub);
    p.bh_sub = BH / p.f;
    if (p.bh_sub < 1) { printf("%s: cannot split\n", name); return; }
    CUtensorMap a_sub, b_sub;
    map4d(&a_sub, act, N, H, W, C, BW, p.bh_sub, p.bn_sub, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    map2d(&b_sub, w, 256, 2304, 64 / P, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    const int smem = STAGES * (16384 + 8192) + 1024 + 256;
    CK(cudaFuncSetAttribute(k_bench2<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int rep = 0; rep < 2; ++rep) { k_bench2<STAGES><<<grid, 576, smem>>>(a_sub, b_sub, p, d_cyc); CK(cudaDeviceSynchronize()); }
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0; double avg = 0;
    for (long long c : h) { if (c > mx) mx = c; avg += (double)c; }
    avg /= grid;
    const double bytes = (double)(p.iters - STAGES) * (16384 + (with_b ? 8192 : 0));
    printf("%-22s P=%2d %s B=%d stages=%d grid=%3d : %6.1f B/clk/SM (slowest %6.1f), %5.0f clk/stage\n", name, P, same_warp ? "lanes" : "warps", with_b, STAGES,
           grid, bytes / avg, bytes / (double)mx, avg / (p.iters - STAGES));
}

template <int STAGES>
static void run(const char *name, const CUtensorMap &a, const CUtensorMap &b, Params p, int grid, long long *d_cyc) {
    const int smem = STAGES * (16384 + 8192) + 1024 + 256;
    CK(cudaFuncSetAttribute(k_bench<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int rep = 0; rep < 2; ++rep) {
        k_bench<STAGES><<<grid, 128, smem>>>(a, b, p, d_cyc);
        CK(cudaDeviceSynchronize());
    }
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0; double avg = 0;
    for (long long c : h) { if (c > mx) mx = c; avg += (double)c; }
    avg /= grid;
    const double bytes = (double)(p.iters - STAGES) * (16384 + (p.with_b ? 8192 : 0));
    printf("%-34s stages=%d grid=%3d prod=%d B=%d : %6.1f B/clk/SM (slowest CTA %6.1f), %5.0f clk/stage\n", name, STAGES, grid, p.producers, p.with_b,
           bytes / avg, bytes / (double)mx, avg / (p.iters - STAGES));
}

int main() {
    CK(cudaSetDevice(0));
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    g_enc = (EncodeFn)fn;
    // activations: 512 images; layer1-like [512,8,8,64], layer2-like [512,4,4,128], layer3-like [512,2,2,256]; the same
    // bytes viewed as a [32768][576]-style GEMM operand for the 2-D case
    const size_t act_elems = (size_t)512 * 8 * 8 * 64;
    __nv_bfloat16 *act, *gemm_a, *w;
    CK(cudaMalloc(&act, act_elems * 2)); CK(cudaMemset(act, 0, act_elems * 2));
    CK(cudaMalloc(&gemm_a, (size_t)32768 * 576 * 2)); CK(cudaMemset(gemm_a, 0, (size_t)32768 * 576 * 2));
    CK(cudaMalloc(&w, (size_t)256 * 2304 * 2)); CK(cudaMemset(w, 0, (size_t)256 * 2304 * 2));
    long long *d_cyc; CK(cudaMalloc(&d_cyc, 1024 * sizeof(long long)));
    (void)gemm_a;
    for (int g : {148, 64}) {
        for (int with_b : {0, 1}) {
            for (int P : {1, 2, 4, 8, 16}) {
                run2<8>("layer1 {64,8,8,2}", act, w, 512, 8, 8, 64, 8, 8, 2, P, 0, with_b, g, d_cyc);
                if (P <= 8) run2<8>("layer1 {64,8,8,2}", act, w, 512, 8, 8, 64, 8, 8, 2, P, 1, with_b, g, d_cyc);
            }
            for (int P : {1, 4, 8}) {
                run2<8>("layer2 {64,4,4,8}", act, w, 1024, 4, 4, 128, 4, 4, 8, P, 0, with_b, g, d_cyc);
                run2<8>("layer3 {64,2,2,32}", act, w, 2048, 2, 2, 256, 2, 2, 32, P, 0, with_b, g, d_cyc);
            }
        }
    }
    run2<4>("layer1 {64,8,8,2}", act, w, 512, 8, 8, 64, 8, 8, 2, 4, 0, 1, 148, d_cyc);
    run2<4>("layer1 {64,8,8,2}", act, w, 512, 8, 8, 64, 8, 8, 2, 8, 0, 1, 148, d_cyc);
    printf("done\n");
    return 0;
}
