// Microbenchmark 3: do TMA loads issued by different warps through different tensor-map objects overlap?
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
// producer p: kind 0 = none, 1 = 4-D activation box (map m), 2 = 2-D weight box (map m); n_off/h_off/row_off select the sub-box
struct Prod { int kind, bytes, smem_off, n_off, h_off, row_off; };
struct P3 { int iters, nprod, stage_bytes; Prod pr[4]; };
constexpr int STAGES = 6;
constexpr int kStage = 32768;
__global__ void __launch_bounds__(192) k3(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1,
                                          const __grid_constant__ CUtensorMap m2, const __grid_constant__ CUtensorMap m3, const P3 p, long long *cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full = (uint64_t *)(smem + STAGES * kStage), *empty = full + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], p.nprod); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int tile = blockIdx.x;
    if (warp >= 1 && warp <= p.nprod && lane == 0) {
        const Prod q = p.pr[warp - 1];
        const CUtensorMap *m = warp == 1 ? &m0 : warp == 2 ? &m1 : warp == 3 ? &m2 : &m3;
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect(&full[stage], q.bytes);
            const int tap = it % 9;
            if (q.kind == 1) tma4(smem + stage * kStage + q.smem_off, m, &full[stage], 0, tap % 3 - 1, q.h_off + tap / 3 - 1, tile * 2 + q.n_off);
            else tma2(smem + stage * kStage + q.smem_off, m, &full[stage], tap * 64, q.row_off);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 0 && lane == 0) {
        int stage = 0; uint32_t phase = 0; long long t0 = 0;
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(&full[stage], phase);
            if (it == STAGES) t0 = clock64();
            mbar_arrive(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn g_enc;
static CUtensorMap map2d(void *base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    CUtensorMap m; cuuint64_t dims[2] = {cols, rows}; cuuint64_t strides[1] = {cols * 2}; cuuint32_t box[2] = {64, box_rows}; cuuint32_t es[2] = {1, 1};
    if (g_enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("enc2d\n"); exit(1); }
    return m;
}
static CUtensorMap map4d(void *base, int N, int H, int W, int C, int bw, int bh, int bn) {
    CUtensorMap m; cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn}; cuuint32_t es[4] = {1, 1, 1, 1};
    if (g_enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("enc4d\n"); exit(1); }
    return m;
}
static void run(const char *name, CUtensorMap m0, CUtensorMap m1, CUtensorMap m2, CUtensorMap m3, P3 p, int grid, long long *d) {
    const int smem = STAGES * kStage + 1024 + 256;
    CK(cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    p.iters = 300;
    for (int r = 0; r < 2; ++r) { k3<<<grid, 192, smem>>>(m0, m1, m2, m3, p, d); CK(cudaDeviceSynchronize()); }
    std::vector<long long> h(grid); CK(cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (long long c : h) avg += (double)c; avg /= grid;
    int bytes = 0; for (int i = 0; i < p.nprod; ++i) bytes += p.pr[i].bytes;
    printf("%-58s grid=%3d: %5.0f clk/stage, %5.1f B/clk/SM (%d B/stage)\n", name, grid, avg / (p.iters - STAGES), bytes * (double)(p.iters - STAGES) / avg, bytes);
}
int main() {
    CK(cudaSetDevice(0));
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q)); g_enc = (EncodeFn)fn;
    __nv_bfloat16 *act, *act2, *w, *w2;
    CK(cudaMalloc(&act, (size_t)512 * 64 * 64 * 2)); CK(cudaMalloc(&act2, (size_t)512 * 64 * 64 * 2));
    CK(cudaMalloc(&w, (size_t)256 * 2304 * 2)); CK(cudaMalloc(&w2, (size_t)256 * 2304 * 2));
    CK(cudaMemset(act, 0, (size_t)512 * 64 * 64 * 2)); CK(cudaMemset(act2, 0, (size_t)512 * 64 * 64 * 2));
    CK(cudaMemset(w, 0, (size_t)256 * 2304 * 2)); CK(cudaMemset(w2, 0, (size_t)256 * 2304 * 2));
    long long *d; CK(cudaMalloc(&d, 1024 * 8));
    CUtensorMap A = map4d(act, 512, 8, 8, 64, 8, 8, 2), Ah = map4d(act, 512, 8, 8, 64, 8, 8, 1), Ah2 = map4d(act, 512, 8, 8, 64, 8, 8, 1);
    CUtensorMap A2 = map4d(act2, 512, 8, 8, 64, 8, 8, 2), Aq = map4d(act, 512, 8, 8, 64, 8, 4, 1);
    CUtensorMap B64 = map2d(w, 256, 2304, 64), B128 = map2d(w, 256, 2304, 128), B64b = map2d(w, 256, 2304, 64), B64c = map2d(w2, 256, 2304, 64);
    for (int g : {148, 32}) {
        P3 p{};
        p.nprod = 1; p.pr[0] = {1, 16384, 0, 0, 0, 0}; run("A128", A, A, A, A, p, g, d);
        p.nprod = 1; p.pr[0] = {2, 16384, 0, 0, 0, 0}; run("B128", B128, A, A, A, p, g, d);
        p.nprod = 2; p.pr[0] = {1, 16384, 0, 0, 0, 0}; p.pr[1] = {2, 8192, 16384, 0, 0, 0}; run("A128 | B64", A, B64, A, A, p, g, d);
        p.pr[1] = {2, 16384, 16384, 0, 0, 0}; run("A128 | B128", A, B128, A, A, p, g, d);
        p.nprod = 2; p.pr[0] = {1, 8192, 0, 0, 0, 0}; p.pr[1] = {1, 8192, 8192, 1, 0, 0}; run("A64(map Ah) | A64(same map Ah)", Ah, Ah, A, A, p, g, d);
        run("A64(map Ah) | A64(map copy Ah2)", Ah, Ah2, A, A, p, g, d);
        p.nprod = 2; p.pr[0] = {1, 16384, 0, 0, 0, 0}; p.pr[1] = {1, 16384, 16384, 0, 0, 0}; run("A128 | A128 (same map, same tile)", A, A, A, A, p, g, d);
        run("A128 | A128 (other tensor act2)", A, A2, A, A, p, g, d);
        p.nprod = 2; p.pr[0] = {2, 8192, 0, 0, 0, 0}; p.pr[1] = {2, 8192, 8192, 0, 0, 64}; run("B64 | B64 (same map)", B64, B64, A, A, p, g, d);
        run("B64 | B64 (map copy)", B64, B64b, A, A, p, g, d);
        run("B64 | B64 (other tensor w2)", B64, B64c, A, A, p, g, d);
        p.nprod = 4; p.pr[0] = {1, 8192, 0, 0, 0, 0}; p.pr[1] = {1, 8192, 8192, 1, 0, 0}; p.pr[2] = {2, 8192, 16384, 0, 0, 0}; p.pr[3] = {2, 8192, 24576, 0, 0, 64};
        run("A64 | A64 | B64 | B64 (4 warps, 2 maps)", Ah, Ah, B64, B64, p, g, d);
        run("A64 | A64' | B64 | B64' (4 warps, 4 map copies)", Ah, Ah2, B64, B64b, p, g, d);
        p.nprod = 3; p.pr[0] = {1, 16384, 0, 0, 0, 0}; p.pr[1] = {2, 8192, 16384, 0, 0, 0}; p.pr[2] = {2, 8192, 24576, 0, 0, 64};
        run("A128 | B64 | B64 (3 warps)", A, B64, B64b, A, p, g, d);
    }
    printf("done\n");
    return 0;
}
