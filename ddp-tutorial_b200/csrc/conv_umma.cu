// tcgen05 + TMA implicit-GEMM convolution kernels (sm_100a).
//
// Replaces what torch dispatches to cuDNN for conv2d forward / backward under
// /root/reference/cifar_example_ddp.py:104,106 on the ResNet bodies BASELINE.json names.
//
// Data layout: activations NHWC bf16, so one output pixel's 64-channel slice is one 128-byte
// row.  A 128-pixel tile (BN images x BH rows x BW cols) of the *shifted* input for filter tap
// (kh,kw) is ONE 4-D TMA box {64 ch, BW, BH, BN} at coordinates {c0, kw-pad, h0+kh-pad, n0}
// (element strides {1,s,s,1} for a strided conv); out-of-bounds pixels are zero-filled by the
// TMA unit, which is exactly the convolution's zero padding.  No im2col buffer ever exists.
//
//   fprop / dgrad  (k_tapgemm): D[128 pix][64 ch] += A_tap[128 pix][64 k] * W_tap[64 ch][64 k]^T
//       A, B K-major SWIZZLE_128B tiles, UMMA 128x64x16, fp32 accumulators in TMEM,
//       warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue (TMEM -> bf16 -> HBM)
//   wgrad (k_wgrad): dW[co][k] += dY[pix][co]^T * X_tap[pix][k]: both operands MN-major
//       (the contraction runs over pixels, i.e. over tile rows), split over pixel tiles.
#include <mutex>
#include <algorithm>
#include <stdlib.h>
#include "conv_umma.cuh"
#include "umma.cuh"

namespace b2 {
using namespace umma;

// =================================================================================
// host: tensor maps
// =================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

int umma_init() {
    if (g_encode) return 0;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    B2_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    B2_CHECK(fn && qres == cudaDriverEntryPointSuccess, B2_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
    g_encode = (EncodeTiledFn)fn;
    return 0;
}

// NHWC bf16 tensor [N][H][W][C]: box {64, bw*es, bh*es, bn}, element strides {1, es, es, 1}
static int make_map_nhwc(CUtensorMap *m, const void *base, int N, int H, int W, int C, int bw, int bh, int bn, int es) {
    B2_TRY(umma_init());
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(bw * es), (cuuint32_t)(bh * es), (cuuint32_t)bn};
    cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B2_CHECK(r == CUDA_SUCCESS, B2_ERR_CUDA, "cuTensorMapEncodeTiled(4D %dx%dx%dx%d box %dx%dx%d es %d) failed: %d", N, H, W, C,
             bw, bh, bn, es, (int)r);
    return 0;
}
// row-major bf16 matrix [rows][cols]: box {64 cols, box_rows}
static int make_map_2d(CUtensorMap *m, const void *base, int64_t rows, int64_t cols, int box_rows) {
    B2_TRY(umma_init());
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B2_CHECK(r == CUDA_SUCCESS, B2_ERR_CUDA, "cuTensorMapEncodeTiled(2D %lldx%lld) failed: %d", (long long)rows, (long long)cols, (int)r);
    return 0;
}

// 128-pixel tile geometry over a box space of N x Hb x Wb pixels
static bool tile_geometry(int N, int Hb, int Wb, int *BW, int *BH, int *BN) {
    if (Wb < 1 || Wb > 128 || (128 % Wb) != 0) return false;
    int bw = Wb;
    int bh = 128 / bw; if (bh > Hb) bh = Hb;
    if (Hb % bh != 0 || (128 % (bw * bh)) != 0) return false;
    int bn = 128 / (bw * bh);
    if (bn > 1 && bh != Hb) return false;
    if (N % bn != 0) return false;
    *BW = bw; *BH = bh; *BN = bn;
    return true;
}

bool umma_conv_eligible(const ConvGeom &g) {
    if (g.Ci % 64 || g.Co % 64) return false;
    if (g.kh != g.kw || g.kh * g.kw > kMaxTaps) return false;
    if (g.stride != 1 && g.stride != 2) return false;
    if (g.stride == 2 && ((g.Hi & 1) || (g.Wi & 1) || g.Ho * 2 != g.Hi || g.Wo * 2 != g.Wi)) return false;
    if (g.stride == 1 && (g.Ho != g.Hi || g.Wo != g.Wi)) return false;
    int a, b, c;
    if (!tile_geometry(g.N, g.Ho, g.Wo, &a, &b, &c)) return false;
    if (!tile_geometry(g.N, g.Hi, g.Wi, &a, &b, &c)) return false;
    return true;
}

// does a tap at source offset d (source = box*step + d) ever touch [0, src) for box in [0, nbox)?
static bool tap_alive(int d, int step, int nbox, int src) {
    for (int b = 0; b < nbox; ++b) { int v = b * step + d; if (v >= 0 && v < src) return true; }
    return false;
}

// =================================================================================
// device: fprop / dgrad
// =================================================================================
constexpr int kATileBytes = 128 * 128;       // 128 pixel rows x 64 bf16
constexpr int kBTileBytes = 64 * 128;        // 64 channel rows x 64 bf16
// Pipeline depth: the per-SM load rate is bytes-in-flight / L2 latency, so a grid that gives each SM
// one CTA needs a deep ring (8 stages = 192 KB in flight); grids with >= 2 CTAs per SM use 4 stages
// (96 KB each) so that two CTAs co-reside and overlap one's epilogue with the other's main loop.
// SPLIT: a cluster of two CTAs shares one output tile, each running half of the K steps (doubles the
// number of SMs pulling operands for layers whose grid is smaller than the chip); CTA 1 hands its
// fp32 accumulators to CTA 0 through distributed shared memory, CTA 0 runs the epilogue.
constexpr int kSplitStageBytes = 128 * 68 * 4;      // [128 rows][64 + 4 pad] fp32
constexpr int tap_smem(int stages, bool split = false) {
    return stages * (kATileBytes + kBTileBytes) + (split ? kSplitStageBytes : 0) + 1024 /*align*/ + 256 /*barriers*/;
}

template <int kStagesTap, bool SPLIT>
__global__ void __launch_bounds__(192)
k_tapgemm(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap b_map,
          const __grid_constant__ TapGemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = smem + kStagesTap * kATileBytes;
    float *sX = (float *)(sB + kStagesTap * kBTileBytes);                 // SPLIT: partner's accumulators land here
    uint64_t *bars = (uint64_t *)((uint8_t *)sX + (SPLIT ? kSplitStageBytes : 0));
    uint64_t *full = bars, *empty = bars + kStagesTap, *tmem_full = bars + 2 * kStagesTap;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * kStagesTap + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TapClass &cls = p.cls[blockIdx.z];
    const uint32_t crank = SPLIT ? cluster_ctarank() : 0u;
    const int tile_x = SPLIT ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_blk = tile_x / p.h_blocks, h_blk = tile_x % p.h_blocks;
    const int ch0 = blockIdx.y * 64;
    const int ksteps_all = cls.ntaps * p.c_chunks;
    const int ks_begin = SPLIT ? (crank == 0 ? 0 : (ksteps_all + 1) / 2) : 0;
    const int ks_end = SPLIT ? (crank == 0 ? (ksteps_all + 1) / 2 : ksteps_all) : ksteps_all;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&a_map); prefetch_tmap(&b_map);
        for (int i = 0; i < kStagesTap; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 64); tmem_relinquish(); }
    pdl_trigger();                 // the next kernel may begin its own prologue
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                    // everything below touches global memory of earlier kernels

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int ks = ks_begin; ks < ks_end; ++ks) {
                const int tap = ks / p.c_chunks, cc = ks - tap * p.c_chunks;
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], kATileBytes + kBTileBytes);
                tma_load_4d(sA + stage * kATileBytes, &a_map, &full[stage], cc * 64, cls.dw[tap],
                            h_blk * p.BH * p.a_hstep + cls.dh[tap], n_blk * p.BN);
                if (p.b_mn) tma_load_2d(sB + stage * kBTileBytes, &b_map, &full[stage], cls.kofs[tap] + ch0, cc * 64);
                else tma_load_2d(sB + stage * kBTileBytes, &b_map, &full[stage], cls.kofs[tap] + cc * 64, ch0);
                if (++stage == kStagesTap) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, 64, 0, p.b_mn);
            const uint32_t bstep = p.b_mn ? 2048 : 32;     // 16 K-rows of 128 B vs 16 K-elements
            int stage = 0; uint32_t phase = 0;
            for (int ks = ks_begin; ks < ks_end; ++ks) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA + stage * kATileBytes), b0 = smem_u32(sB + stage * kBTileBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    mma_bf16_ss(tmem_base, make_desc_sw128(a0 + k * 32, 16, 1024),
                                make_desc_sw128(b0 + k * bstep, kBTileBytes, 1024), idesc, (ks > ks_begin || k > 0) ? 1u : 0u);
                }
                mma_commit(&empty[stage]);
                if (++stage == kStagesTap) { stage = 0; phase ^= 1; }
            }
            mma_commit(tmem_full);
        }
    } else {
        // ---- epilogue: TMEM lane quarter (warp % 4) -> rows
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int w_l = r % p.BW; const int t = r / p.BW;
        const int h_l = t % p.BH; const int n_l = t / p.BH;
        const int n = n_blk * p.BN + n_l;
        const int h = (h_blk * p.BH + h_l) * p.os + cls.oh;
        const int w = w_l * p.os + cls.ow;
        bf16 *dst = p.out + (((int64_t)n * p.out_H + h) * p.out_W + w) * p.out_C + ch0;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        uint32_t v[32];
        if (SPLIT && crank == 1) {
            // hand this CTA's partial accumulators to CTA 0 (its sX), then leave through the cluster barrier
            const uint32_t remote = map_to_cta(smem_u32(sX), 0) + (uint32_t)r * 68 * 4;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + half * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    st_cluster_f4(remote + (half * 32 + j * 4) * 4, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                  __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            }
            cluster_arrive();
            cluster_wait();
        } else {
        if (SPLIT) { cluster_arrive(); cluster_wait(); }       // partner's accumulators are in sX now
        float *sT = (float *)sA;                       // [128][65] fp32 staging for per-channel sums
        float *sT2 = sT + 128 * 65;                    // second plane (spills into the B stages: 66.5 KB < 96 KB)
        const bool want_fwd = p.stats != nullptr, want_bwd = p.bn_bsum != nullptr;
        const int64_t pix_off = (((int64_t)n * p.out_H + h) * p.out_W + w) * p.out_C + ch0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + half * 32, v);
            tmem_ld_wait();
            uint4 *d4 = (uint4 *)(dst + half * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j * 8 + e]);
                if (SPLIT) {
                    const float4 x0 = *(const float4 *)(sX + r * 68 + half * 32 + j * 8);
                    const float4 x1 = *(const float4 *)(sX + r * 68 + half * 32 + j * 8 + 4);
                    f[0] += x0.x; f[1] += x0.y; f[2] += x0.z; f[3] += x0.w; f[4] += x1.x; f[5] += x1.y; f[6] += x1.z; f[7] += x1.w;
                }
                if (p.acc) {
                    uint4 old = d4[j];
                    const __nv_bfloat162 *o2 = (const __nv_bfloat162 *)&old;
#pragma unroll
                    for (int e = 0; e < 4; ++e) { float2 of = __bfloat1622float2(o2[e]); f[2 * e] += of.x; f[2 * e + 1] += of.y; }
                }
                uint4 pk;
                __nv_bfloat162 *p2 = (__nv_bfloat162 *)&pk;
#pragma unroll
                for (int e = 0; e < 4; ++e) p2[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                d4[j] = pk;
                const int col = half * 32 + j * 8;
                if (want_fwd) {                        // statistics of the values BatchNorm will read (bf16-rounded)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float2 rf = __bfloat1622float2(p2[e]);
                        sT[r * 65 + col + 2 * e] = rf.x;
                        sT[r * 65 + col + 2 * e + 1] = rf.y;
                    }
                }
                if (want_bwd) {                        // g = dY*[y>0],  g*xhat  of the BatchNorm that produced this tensor
                    uint4 yr = *(const uint4 *)(p.bn_y + pix_off + col), xr = *(const uint4 *)(p.bn_x + pix_off + col);
                    const __nv_bfloat162 *y2 = (const __nv_bfloat162 *)&yr, *x2 = (const __nv_bfloat162 *)&xr;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float2 g = __bfloat1622float2(p2[e]), yv = __bfloat1622float2(y2[e]), xv = __bfloat1622float2(x2[e]);
                        const int c = ch0 + col + 2 * e;
                        if (p.bn_relu) { if (!(yv.x > 0.f)) g.x = 0.f; if (!(yv.y > 0.f)) g.y = 0.f; }
                        sT[r * 65 + col + 2 * e] = g.x;
                        sT[r * 65 + col + 2 * e + 1] = g.y;
                        sT2[r * 65 + col + 2 * e] = g.x * ((xv.x - __ldg(p.bn_mean + c)) * __ldg(p.bn_invstd + c));
                        sT2[r * 65 + col + 2 * e + 1] = g.y * ((xv.y - __ldg(p.bn_mean + c + 1)) * __ldg(p.bn_invstd + c + 1));
                    }
                }
            }
        }
        const int e = (warp - 2) * 32 + lane;
        if (want_fwd || want_bwd) {
            asm volatile("bar.sync 1, 128;" ::: "memory");      // the four epilogue warps only
            const int c = e & 63, r0 = (e >> 6) * 64;
            float s0 = 0.f, s1 = 0.f;
            if (want_fwd) {
#pragma unroll 8
                for (int rr = 0; rr < 64; ++rr) { float x = sT[(r0 + rr) * 65 + c]; s0 += x; s1 = fmaf(x, x, s1); }
            } else {
#pragma unroll 8
                for (int rr = 0; rr < 64; ++rr) { s0 += sT[(r0 + rr) * 65 + c]; s1 += sT2[(r0 + rr) * 65 + c]; }
            }
            float *st = (want_fwd ? p.stats : p.bn_bsum) + (size_t)((blockIdx.x + blockIdx.z) % kBnSumCopies) * 2 * p.out_C;
            atomicAdd(st + ch0 + c, s0);
            atomicAdd(st + p.out_C + ch0 + c, s1);
        }
        if (p.fuse_bn) {
            // ---- grid-wide barrier: every CTA of this launch has contributed its tile's sums
            __threadfence();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (e == 0) {
                atomicAdd(p.fb_counter, 1u);
                const uint32_t total = gridDim.x * gridDim.y * gridDim.z;
                const long long t0 = clock64();
                while (ld_acquire_gpu_u32(p.fb_counter) < total)
                    if (clock64() - t0 > 6000000000LL) { *p.fb_err = 1; break; }      // ~3 s: fail, do not hang
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            float *sS = sT;                                   // [0,64) scale, [64,128) shift of this CTA's channels
            if (e < 64) {
                const int c = ch0 + e;
                float t0s = 0.f, t1s = 0.f;
#pragma unroll
                for (int k = 0; k < kBnSumCopies; ++k) { t0s += __ldcg(p.stats + (2 * k) * p.out_C + c); t1s += __ldcg(p.stats + (2 * k + 1) * p.out_C + c); }
                const double mu = (double)t0s / (double)p.fb_count;
                double var = (double)t1s / (double)p.fb_count - mu * mu;
                if (var < 0.0) var = 0.0;
                const float is = (float)(1.0 / sqrt(var + (double)p.fb_eps));
                const float g = p.fb_gamma[c];
                sS[e] = is * g;
                sS[64 + e] = p.fb_beta[c] - (float)mu * is * g;
                if (blockIdx.x == 0 && blockIdx.z == 0) {
                    p.fb_save_mean[c] = (float)mu; p.fb_save_invstd[c] = is;
                    if (p.fb_running_mean) {
                        const double unbiased = p.fb_count > 1.f ? var * (double)p.fb_count / ((double)p.fb_count - 1.0) : var;
                        p.fb_running_mean[c] = (1.f - p.fb_momentum) * p.fb_running_mean[c] + p.fb_momentum * (float)mu;
                        p.fb_running_var[c] = (1.f - p.fb_momentum) * p.fb_running_var[c] + p.fb_momentum * (float)unbiased;
                    }
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            bf16 *ydst = p.fb_y + pix_off;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + half * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int col = half * 32 + j * 8;
                    uint4 rr = make_uint4(0, 0, 0, 0);
                    if (p.fb_res) rr = *(const uint4 *)(p.fb_res + pix_off + col);
                    const __nv_bfloat162 *r2 = (const __nv_bfloat162 *)&rr;
                    uint4 pk;
                    __nv_bfloat162 *p2 = (__nv_bfloat162 *)&pk;
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) {
                        // normalise the bf16-rounded conv output (what is stored and what backward reads)
                        float2 xr = __bfloat1622float2(__floats2bfloat162_rn(__uint_as_float(v[j * 8 + 2 * k2]), __uint_as_float(v[j * 8 + 2 * k2 + 1])));
                        float a = fmaf(xr.x, sS[col + 2 * k2], sS[64 + col + 2 * k2]);
                        float b = fmaf(xr.y, sS[col + 2 * k2 + 1], sS[64 + col + 2 * k2 + 1]);
                        if (p.fb_res) { float2 rf = __bfloat1622float2(r2[k2]); a += rf.x; b += rf.y; }
                        if (p.fb_relu) { a = a > 0.f ? a : 0.f; b = b > 0.f ? b : 0.f; }
                        p2[k2] = __floats2bfloat162_rn(a, b);
                    }
                    *(uint4 *)(ydst + col) = pk;
                }
            }
        }
        }   // !(SPLIT && crank == 1)
        tc_fence_before();
    }
    if (SPLIT && warp < 2) { cluster_arrive(); cluster_wait(); }      // every thread of both CTAs joins the barrier once
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

// =================================================================================
// device: wgrad
// =================================================================================
constexpr int kStagesWg = 3;
constexpr int kBoxBytes = 128 * 128;         // [128 pixel rows][64 bf16]
constexpr int kWgStageBytes = 4 * kBoxBytes; // up to 2 dY boxes + 2 X boxes
constexpr int kWgSmem = kStagesWg * kWgStageBytes + 1024 + 256;

__global__ void __launch_bounds__(192, 1)
k_wgrad(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap dy_map,
        const __grid_constant__ WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(smem + kStagesWg * kWgStageBytes);
    uint64_t *full = bars, *empty = bars + kStagesWg, *tmem_full = bars + 2 * kStagesWg;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * kStagesWg + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // N tile: up to two 64-wide K chunks (chunk index over live taps x channel chunks)
    const int total_chunks = p.ntaps_alive * p.c_chunks;
    const int chunk0 = blockIdx.x * 2;
    const int nb = (chunk0 + 1 < total_chunks) ? 2 : 1;
    const int mb = p.m_boxes;
    const int co0 = blockIdx.y * 64 * mb;
    const int pt0 = blockIdx.z * p.tiles_per_split;
    int pt1 = pt0 + p.tiles_per_split; if (pt1 > p.pix_tiles) pt1 = p.pix_tiles;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&x_map); prefetch_tmap(&dy_map);
        for (int i = 0; i < kStagesWg; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 128); tmem_relinquish(); }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int tap[2], cc[2];
            for (int j = 0; j < nb; ++j) { int c = chunk0 + j; tap[j] = c / p.c_chunks; cc[j] = c - tap[j] * p.c_chunks; }
            int stage = 0; uint32_t phase = 0;
            for (int pt = pt0; pt < pt1; ++pt) {
                const int n_blk = pt / p.h_blocks, h_blk = pt - n_blk * p.h_blocks;
                uint8_t *st = smem + stage * kWgStageBytes;
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], (mb + nb) * kBoxBytes);
                for (int j = 0; j < mb; ++j)
                    tma_load_2d(st + j * kBoxBytes, &dy_map, &full[stage], co0 + j * 64, pt * 128);
                for (int j = 0; j < nb; ++j)
                    tma_load_4d(st + (2 + j) * kBoxBytes, &x_map, &full[stage], cc[j] * 64, p.dw[tap[j]],
                                h_blk * p.BH * p.a_hstep + p.dh[tap[j]], n_blk * p.BN);
                if (++stage == kStagesWg) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(64 * mb, 64 * nb, 1, 1);
            int stage = 0; uint32_t phase = 0;
            for (int pt = pt0; pt < pt1; ++pt) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + stage * kWgStageBytes), b0 = a0 + 2 * kBoxBytes;
#pragma unroll
                for (int k = 0; k < 8; ++k) {     // 16 pixel rows (= UMMA K) per instruction
                    mma_bf16_ss(tmem_base, make_desc_sw128(a0 + k * 2048, kBoxBytes, 1024),
                                make_desc_sw128(b0 + k * 2048, kBoxBytes, 1024), idesc, (pt > pt0 || k > 0) ? 1u : 0u);
                }
                mma_commit(&empty[stage]);
                if (++stage == kStagesWg) { stage = 0; phase ^= 1; }
            }
            mma_commit(tmem_full);
        }
    } else {
        const int q = warp & 3;
        // M = 128: lane i <-> row i.  M = 64: rows 16q..16q+15 live in lanes 32q..32q+15.
        int row; bool valid;
        if (mb == 2) { row = q * 32 + lane; valid = true; }
        else { row = q * 16 + lane; valid = lane < 16; }
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        float *dst_row = p.out + ((int64_t)blockIdx.z * p.Co + co0 + row) * p.out_pitch;
        uint32_t v[32];
        for (int j = 0; j < nb; ++j) {
            const int c = chunk0 + j;
            const int tap = c / p.c_chunks, cc = c - tap * p.c_chunks;
            float *dst = dst_row + (p.compact ? (int64_t)c * 64 : (int64_t)p.tap_id[tap] * p.Ci + cc * 64);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + j * 64 + half * 32, v);
                tmem_ld_wait();
                if (valid) {
                    float4 *d4 = (float4 *)(dst + half * 32);
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        d4[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]),
                                            __uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3]));
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

// =================================================================================
// host: plans and launches
// =================================================================================
static std::once_flag g_attr_once;
static void set_attrs() {
    cudaFuncSetAttribute(k_tapgemm<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tap_smem(4));
    cudaFuncSetAttribute(k_tapgemm<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tap_smem(8));
    cudaFuncSetAttribute(k_tapgemm<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tap_smem(6, true));
    cudaFuncSetAttribute(k_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
}

int umma_plan_fprop(const ConvGeom &g, const bf16 *x, const bf16 *w, bf16 *y, UmmaConvPlan *plan) {
    plan->valid = false;
    B2_CHECK(umma_conv_eligible(g), B2_ERR_UNSUPPORTED, "umma fprop: shape not eligible");
    TapGemmParams &p = plan->p;
    memset(&p, 0, sizeof(p));
    B2_CHECK(tile_geometry(g.N, g.Ho, g.Wo, &p.BW, &p.BH, &p.BN), B2_ERR_UNSUPPORTED, "umma fprop: tile geometry");
    p.h_blocks = g.Ho / p.BH; p.a_hstep = g.stride; p.c_chunks = g.Ci / 64;
    p.out_H = g.Ho; p.out_W = g.Wo; p.out_C = g.Co; p.os = 1; p.out = y; p.acc = 0; p.stats = nullptr;
    p.nclasses = 1;
    TapClass &c = p.cls[0];
    c.ntaps = 0; c.oh = c.ow = 0;
    for (int kh = 0; kh < g.kh; ++kh)
        for (int kw = 0; kw < g.kw; ++kw) {
            int dh = kh - g.pad, dw = kw - g.pad;
            if (!tap_alive(dh, g.stride, g.Ho, g.Hi) || !tap_alive(dw, g.stride, g.Wo, g.Wi)) continue;
            c.dh[c.ntaps] = dh; c.dw[c.ntaps] = dw; c.kofs[c.ntaps] = (kh * g.kw + kw) * g.Ci; ++c.ntaps;
        }
    B2_TRY(make_map_nhwc(&plan->a_map, x, g.N, g.Hi, g.Wi, g.Ci, p.BW, p.BH, p.BN, g.stride));
    B2_TRY(make_map_2d(&plan->b_map, w, g.Co, (int64_t)g.kh * g.kw * g.Ci, 64));
    plan->grid = dim3((unsigned)(g.N / p.BN * p.h_blocks), (unsigned)(g.Co / 64), 1);
    plan->flops = 2.0 * g.M() * g.Co * (double)c.ntaps * g.Ci;
    plan->valid = true;
    return 0;
}

int umma_plan_dgrad(const ConvGeom &g, const bf16 *dy, const bf16 *w, bf16 *dx, bool acc, UmmaConvPlan *plan) {
    plan->valid = false;
    B2_CHECK(umma_conv_eligible(g), B2_ERR_UNSUPPORTED, "umma dgrad: shape not eligible");
    TapGemmParams &p = plan->p;
    memset(&p, 0, sizeof(p));
    // box space = dY's pixel grid for stride 2 (one class per output parity), dX's grid for stride 1
    B2_CHECK(tile_geometry(g.N, g.Ho, g.Wo, &p.BW, &p.BH, &p.BN), B2_ERR_UNSUPPORTED, "umma dgrad: tile geometry");
    p.h_blocks = g.Ho / p.BH; p.a_hstep = 1; p.c_chunks = g.Co / 64;
    p.out_H = g.Hi; p.out_W = g.Wi; p.out_C = g.Ci; p.os = g.stride; p.out = dx; p.acc = acc ? 1 : 0; p.stats = nullptr;
    p.b_mn = 1;
    const int s = g.stride;
    p.nclasses = 0;
    double taps_total = 0;
    for (int ph = 0; ph < s; ++ph)
        for (int pw = 0; pw < s; ++pw) {
            TapClass &c = p.cls[p.nclasses];
            c.ntaps = 0; c.oh = ph; c.ow = pw;
            for (int kh = 0; kh < g.kh; ++kh)
                for (int kw = 0; kw < g.kw; ++kw) {
                    // hi = s*a + ph ; ho = (hi + pad - kh)/s must be integral: a + (ph + pad - kh)/s
                    int th = ph + g.pad - kh, tw = pw + g.pad - kw;
                    if (((th % s) + s) % s != 0 || ((tw % s) + s) % s != 0) continue;
                    int dh = th >= 0 ? th / s : -((-th) / s), dw = tw >= 0 ? tw / s : -((-tw) / s);
                    if (!tap_alive(dh, 1, g.Ho, g.Ho) || !tap_alive(dw, 1, g.Wo, g.Wo)) continue;
                    c.dh[c.ntaps] = dh; c.dw[c.ntaps] = dw; c.kofs[c.ntaps] = (kh * g.kw + kw) * g.Ci; ++c.ntaps;
                }
            taps_total += c.ntaps;
            if (c.ntaps > 0) ++p.nclasses;      // a class without taps contributes nothing
            else B2_CHECK(acc || s == 1, B2_ERR_UNSUPPORTED, "umma dgrad: empty parity class needs acc (caller zero-fills)");
        }
    B2_TRY(make_map_nhwc(&plan->a_map, dy, g.N, g.Ho, g.Wo, g.Co, p.BW, p.BH, p.BN, 1));
    B2_TRY(make_map_2d(&plan->b_map, w, g.Co, (int64_t)g.kh * g.kw * g.Ci, 64));
    plan->grid = dim3((unsigned)(g.N / p.BN * p.h_blocks), (unsigned)(g.Ci / 64), (unsigned)p.nclasses);
    plan->flops = 2.0 * (double)g.N * g.Ho * g.Wo * g.Ci * taps_total * g.Co;
    plan->valid = true;
    return 0;
}

int umma_conv_launch(const UmmaConvPlan &plan, cudaStream_t s) {
    B2_CHECK(plan.valid, B2_ERR_STATE, "umma conv: plan not valid");
    std::call_once(g_attr_once, set_attrs);
    const unsigned ctas = plan.grid.x * plan.grid.y * plan.grid.z;
    int min_ksteps = 1 << 30;
    for (int c = 0; c < plan.p.nclasses; ++c) min_ksteps = std::min(min_ksteps, plan.p.cls[c].ntaps * plan.p.c_chunks);
    // measured (r01): parity-green but 4% slower end to end (1.051 vs 1.007 ms/step) -- the extra
    // cluster barrier / DSMEM hand-off and the competition with the concurrent wgrad stream outweigh
    // the additional SMs.  Kept behind B2_SPLITK=1 for layer-level experiments.
    static const bool allow_split = [] { const char *e = getenv("B2_SPLITK"); return e && e[0] == '1'; }();
    if (allow_split && ctas <= (unsigned)kNumSMs && min_ksteps >= 4) {
        // split-K cluster pair: twice the CTAs (and SMs) per output tile
        dim3 g(plan.grid.x * 2, plan.grid.y, plan.grid.z);
        cudaError_t e = launch_kernel_cluster(k_tapgemm<6, true>, g, dim3(192), (size_t)tap_smem(6, true), s, 2u, plan.a_map, plan.b_map, plan.p);
        ++g_launches;
        B2_CHECK(e == cudaSuccess, B2_ERR_CUDA, "k_tapgemm<split> launch failed: %s", cudaGetErrorString(e));
    } else if (ctas <= (unsigned)kNumSMs) { B2_LAUNCH((k_tapgemm<8, false>), plan.grid, 192, tap_smem(8), s, plan.a_map, plan.b_map, plan.p); }
    else { B2_LAUNCH((k_tapgemm<4, false>), plan.grid, 192, tap_smem(4), s, plan.a_map, plan.b_map, plan.p); }
    return 0;
}

static void wgrad_shape(const ConvGeom &g, int *ntaps_alive, int *m_boxes, int *k_tiles, int *pix_tiles, int *splits, int *per) {
    int alive = 0;
    for (int kh = 0; kh < g.kh; ++kh)
        for (int kw = 0; kw < g.kw; ++kw)
            if (tap_alive(kh - g.pad, g.stride, g.Ho, g.Hi) && tap_alive(kw - g.pad, g.stride, g.Wo, g.Wi)) ++alive;
    *ntaps_alive = alive;
    *m_boxes = g.Co >= 128 ? 2 : 1;
    int chunks = alive * (g.Ci / 64);
    *k_tiles = (chunks + 1) / 2;
    *pix_tiles = (int)(g.M() / 128);
    int tiles = *k_tiles * (g.Co / (64 * *m_boxes));
    int want = (2 * kNumSMs + tiles - 1) / tiles;
    int maxs = (*pix_tiles + 1) / 2; if (maxs < 1) maxs = 1;
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    *per = (*pix_tiles + want - 1) / want;
    *splits = (*pix_tiles + *per - 1) / *per;
}

size_t umma_wgrad_workspace(const ConvGeom &g) {
    int a, mb, kt, pt, sp, per;
    wgrad_shape(g, &a, &mb, &kt, &pt, &sp, &per);
    return sp > 1 ? (size_t)sp * g.Co * a * g.Ci : 0;
}

int umma_plan_wgrad(const ConvGeom &g, const bf16 *x, const bf16 *dy, float *dw, float *partial, size_t partial_floats,
                    UmmaWgradPlan *plan) {
    plan->valid = false;
    B2_CHECK(umma_conv_eligible(g), B2_ERR_UNSUPPORTED, "umma wgrad: shape not eligible");
    WgradParams &p = plan->p;
    memset(&p, 0, sizeof(p));
    int kt;
    wgrad_shape(g, &p.ntaps_alive, &p.m_boxes, &kt, &p.pix_tiles, &p.splits, &p.tiles_per_split);
    B2_CHECK(tile_geometry(g.N, g.Ho, g.Wo, &p.BW, &p.BH, &p.BN), B2_ERR_UNSUPPORTED, "umma wgrad: tile geometry");
    p.h_blocks = g.Ho / p.BH; p.a_hstep = g.stride; p.c_chunks = g.Ci / 64; p.Ci = g.Ci; p.Co = g.Co; p.K = g.K();
    int n = 0;
    for (int kh = 0; kh < g.kh; ++kh)
        for (int kw = 0; kw < g.kw; ++kw)
            if (tap_alive(kh - g.pad, g.stride, g.Ho, g.Hi) && tap_alive(kw - g.pad, g.stride, g.Wo, g.Wi)) {
                p.tap_id[n] = kh * g.kw + kw; p.dh[n] = kh - g.pad; p.dw[n] = kw - g.pad; ++n;
            }
    plan->need_zero = n < g.kh * g.kw;
    plan->dw = dw;
    if (p.splits > 1) {
        B2_CHECK((size_t)p.splits * g.Co * n * g.Ci <= partial_floats, B2_ERR_ARG, "umma wgrad: workspace too small");
        p.out = partial; p.out_pitch = n * g.Ci; p.compact = 1;
    } else { p.out = dw; p.out_pitch = g.K(); p.compact = 0; }
    B2_TRY(make_map_nhwc(&plan->x_map, x, g.N, g.Hi, g.Wi, g.Ci, p.BW, p.BH, p.BN, g.stride));
    B2_TRY(make_map_2d(&plan->dy_map, dy, g.M(), g.Co, 128));
    plan->grid = dim3((unsigned)kt, (unsigned)(g.Co / (64 * p.m_boxes)), (unsigned)p.splits);
    plan->flops = 2.0 * g.M() * g.Co * (double)n * g.Ci;
    plan->valid = true;
    return 0;
}

// dW[co][tap_id*Ci + c] = sum_z partial[z][co][a*Ci + c]   (a = live-tap index), float4 lanes
__global__ void k_wgrad_reduce(const float *__restrict__ part, float *__restrict__ dw, WgradParams p) {
    pdl_sync();
    const int akc = p.out_pitch;                 // live taps * Ci
    const int64_t per_split = (int64_t)p.Co * akc;
    const int64_t n4 = per_split >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 s = ((const float4 *)part)[i];
        for (int z = 1; z < p.splits; ++z) {
            float4 v = ((const float4 *)(part + (int64_t)z * per_split))[i];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        const int64_t e = i << 2;
        const int co = (int)(e / akc); const int col = (int)(e - (int64_t)co * akc);
        const int a = col / p.Ci; const int c = col - a * p.Ci;
        *(float4 *)(dw + (int64_t)co * p.K + (int64_t)p.tap_id[a] * p.Ci + c) = s;
    }
}

int umma_wgrad_zero_dead(const UmmaWgradPlan &plan, cudaStream_t s) {
    if (plan.need_zero) B2_CUDA(cudaMemsetAsync(plan.dw, 0, (size_t)plan.p.Co * plan.p.K * 4, s));
    return 0;
}

int umma_wgrad_launch(const UmmaWgradPlan &plan, cudaStream_t s) {
    B2_CHECK(plan.valid, B2_ERR_STATE, "umma wgrad: plan not valid");
    std::call_once(g_attr_once, set_attrs);
    const WgradParams &p = plan.p;
    B2_LAUNCH(k_wgrad, plan.grid, 192, kWgSmem, s, plan.x_map, plan.dy_map, plan.p);
    if (p.splits > 1) {
        int64_t n4 = (int64_t)p.Co * p.out_pitch / 4;
        int blocks = (int)ceil_div(n4, 256); if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        B2_LAUNCH(k_wgrad_reduce, blocks, 256, 0, s, p.out, plan.dw, p);
    }
    return 0;
}

}  // namespace b2
