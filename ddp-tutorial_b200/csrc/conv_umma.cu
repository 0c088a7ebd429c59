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
//   fprop / dgrad  (k_conv<BN_TILE>): D[128 pix][BN_TILE ch] += A_tap[128 pix][64 k] * W_tap[BN_TILE ch][64 k]^T
//       persistent: grid <= #SMs, CTA b walks tiles b, b+grid, ...; accumulators double-buffered in TMEM so the
//       epilogue of tile i overlaps the main loop of tile i+1.  Warp roles: 0 = TMA producer of A,
//       1 = TMA producer of B (measured, tools/tma_bench3.cu: the TMA unit processes the boxes of ONE issuing
//       warp back to back at ~2 clk per 128-byte row, but streams from two warps overlap: 16 KB A + 16 KB B per
//       k-step in ~260-330 clk instead of ~540), 2 = MMA issuer, 3-6 = epilogue.
//       Epilogue: TMEM -> bf16 -> 128B-swizzled staging tile in shared memory -> row-coalesced 16-byte global
//       stores; the same pass accumulates per-channel BatchNorm sums (forward statistics, or the backward sums
//       sum(g), sum(g*xhat) of the BatchNorm that produced the tensor whose gradient this dgrad finalises).
//   wgrad (k_wgrad): dW[co][k] += dY[pix][co]^T * X_tap[pix][k]: both operands MN-major
//       (the contraction runs over pixels, i.e. over tile rows), split over pixel tiles; the splits add into
//       dW with red.global.add.v4.f32 (or, deterministic mode, through partials + a reduce kernel).
#include <mutex>
#include <algorithm>
#include <stdlib.h>
#define B2_TRACE_TU_ID 3
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
static int g_num_sms = 0;

int umma_init() {
    if (g_encode) return 0;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    B2_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    B2_CHECK(fn && qres == cudaDriverEntryPointSuccess, B2_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
    int dev = 0;
    B2_CUDA(cudaGetDevice(&dev));
    B2_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    g_encode = (EncodeTiledFn)fn;
    return 0;
}
int umma_num_sms() { return g_num_sms > 0 ? g_num_sms : kNumSMs; }

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
static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

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

bool umma_stem_supported(const ConvGeom &g, const ConvGeom &v, int kpad) {
    // the gather code is written for exactly this geometry (torchvision's stem on CIFAR-sized input)
    return g.Ci == 3 && g.Hi == 32 && g.Wi == 32 && g.kh == 7 && g.kw == 7 && g.stride == 2 && g.pad == 3 && g.Ho == 16 && g.Wo == 16 &&
           g.Co == 64 && kpad == 192 && v.Ci == 192 && v.kh == 1 && v.Ho == 16 && v.Wo == 16;
}

// does a tap at source offset d (source = box*step + d) ever touch [0, src) for box in [0, nbox)?
static bool tap_alive(int d, int step, int nbox, int src) {
    for (int b = 0; b < nbox; ++b) { int v = b * step + d; if (v >= 0 && v < src) return true; }
    return false;
}

// =================================================================================
// device: fprop / dgrad
// =================================================================================
constexpr int kStemXsElemsC = 3 * 21 * 40;
template <int BN_TILE, int STAGES = (BN_TILE == 128 ? 5 : 6)>
struct ConvCfg {
    static constexpr int kStages = STAGES;
    static constexpr int kABytes = 128 * 128;              // 128 pixel rows x 64 bf16
    static constexpr int kBBytes = BN_TILE * 128;          // BN_TILE channel rows x 64 bf16 (or 64 k-rows x BN_TILE)
    static constexpr int kStagingBytes = (BN_TILE / 64) * 16384;   // [half][128 rows][128 B], 128B-swizzled
    static constexpr int kRedFloats = 8 * 2 * BN_TILE;     // per epilogue warp: [2][BN_TILE] partial sums
    static constexpr int kSmem = kStages * (kABytes + kBBytes) + kStagingBytes + (kRedFloats + 2 * BN_TILE) * 4 + 256 /*barriers*/ +
                                 2 * kStemXsElemsC * 2 /*stem input windows*/ + 1024 /*align*/;
    static constexpr int kTmemCols = 2 * BN_TILE;          // two accumulator buffers
    static constexpr int kCPR = BN_TILE / 8;               // 16-byte chunks per output row
};
constexpr int kEpiWarps = 8, kEpiThreads = kEpiWarps * 32;
constexpr int kConvThreads = (3 + kEpiWarps) * 32;     // warps: 0 A-TMA, 1 B-TMA, 2 MMA, 3..10 epilogue
constexpr int kStemExtraThreads = 128; // stem variants: 4 more warps gather im2col rows into the A stages

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---- stem gather (shared by k_conv<64, true> and k_wgrad<true>) ----------------------------------------------------
// Geometry is fixed: 3 x 32 x 32 fp32 image, 7x7 / stride 2 / pad 3, 16x16 outputs, tile = 8 output rows x 16 columns
// of one image.  The 21 input rows a tile needs are staged as bf16 [3][21][40] (3 zero columns left, 5 right), so the
// 8 consecutive input pixels of im2col chunk (c, kh) at output column ow start at element 2*ow of row 2*oh_l + kh.
constexpr int kStemXsRow = 40, kStemXsRows = 21, kStemXsElems = 3 * kStemXsRows * kStemXsRow;
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void stem_bar() { asm volatile("bar.sync 2, 128;" ::: "memory"); }
struct StemRaw { float4 v[4]; };
// issue the global loads of the 3 x 21 input rows of (image n, row half h_blk); 504 float4, up to 4 per thread
__device__ __forceinline__ void stem_load(const float *__restrict__ x, int n, int h_blk, int gt, StemRaw &raw) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int idx = gt + u * 128;
        raw.v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < 3 * kStemXsRows * 8) {
            const int c = idx / (kStemXsRows * 8), rem = idx - c * (kStemXsRows * 8), rr = rem >> 3, q4 = rem & 7;
            const int ih = h_blk * 16 - 3 + rr;
            if (ih >= 0 && ih < 32) raw.v[u] = __ldg((const float4 *)(x + (((int64_t)n * 3 + c) * 32 + ih) * 32) + q4);
        }
    }
}
__device__ __forceinline__ void stem_store(bf16 *xs, int gt, const StemRaw &raw) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int idx = gt + u * 128;
        if (idx < 3 * kStemXsRows * 8) {
            const int c = idx / (kStemXsRows * 8), rem = idx - c * (kStemXsRows * 8), rr = rem >> 3, q4 = rem & 7;
            bf16 *d = xs + (c * kStemXsRows + rr) * kStemXsRow + 3 + q4 * 4;
            d[0] = __float2bfloat16_rn(raw.v[u].x); d[1] = __float2bfloat16_rn(raw.v[u].y);
            d[2] = __float2bfloat16_rn(raw.v[u].z); d[3] = __float2bfloat16_rn(raw.v[u].w);
        }
    }
}
// im2col k-chunk kc (64 columns = 8 (c, kh) groups of 8 pixels) of tile row r = oh_l * 16 + ow -> 128B-swizzled stage row
__device__ __forceinline__ void stem_build_chunk(const bf16 *xs, int r, int kc, uint8_t *stage) {
    const int oh_l = r >> 4, ow = r & 15;
    uint8_t *row = stage + r * 128;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        const int j = kc * 8 + jj;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (j < 21) {
            const int c = j / 7, kh = j - c * 7;
            const uint32_t *src = (const uint32_t *)(xs + (c * kStemXsRows + 2 * oh_l + kh) * kStemXsRow + 2 * ow);
            v = make_uint4(src[0], src[1], src[2], src[3]);
        }
        *(uint4 *)(row + ((jj ^ (r & 7)) << 4)) = v;
    }
}

struct TileCoord { int z, n, n_blk, h_blk, ch0, m; };
// tile t = (class * n_tiles + n) * m_tiles + m
template <int BN_TILE>
__device__ __forceinline__ TileCoord decode_tile(const ConvParams &p, int t) {
    TileCoord c;
    c.m = t % p.m_tiles;
    const int r = t / p.m_tiles;
    c.n = r % p.n_tiles; c.z = r / p.n_tiles;
    c.n_blk = c.m / p.h_blocks; c.h_blk = c.m - c.n_blk * p.h_blocks;
    c.ch0 = c.n * BN_TILE;
    return c;
}
// global element offset of tile row `row`, 8-channel chunk `chunk`
__device__ __forceinline__ int64_t row_offset(const ConvParams &p, const TileCoord &tc, int row, int chunk, int oh, int ow) {
    const int w_l = row & (p.BW - 1);
    const int tt = row >> p.lgBW;
    const int h_l = tt & (p.BH - 1), n_l = tt >> p.lgBH;
    const int n = tc.n_blk * p.BN + n_l;
    const int h = (tc.h_blk * p.BH + h_l) * p.os + oh;
    const int w = w_l * p.os + ow;
    return (((int64_t)n * p.out_H + h) * p.out_W + w) * p.out_C + tc.ch0 + chunk * 8;
}
__device__ __forceinline__ uint4 ldg_nc_u4(const bf16 *p) {       // read-only path: may be hoisted above the stores of the loop
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int BN_TILE, bool STEM, int STAGES = (BN_TILE == 128 ? 5 : 6)>
__global__ void __launch_bounds__(kConvThreads + (STEM ? kStemExtraThreads : 0), STAGES == 3 ? 2 : 1)
k_conv(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap b_map,
       const __grid_constant__ CUtensorMap out_map, const __grid_constant__ ConvParams p) {
    using Cfg = ConvCfg<BN_TILE, STAGES>;
    constexpr int S = Cfg::kStages, kA = Cfg::kABytes, kB = Cfg::kBBytes, CPR = Cfg::kCPR;
    constexpr int RP = kEpiThreads / CPR;                            // tile rows covered by one pass of the epilogue threads
    constexpr int NI = 128 / RP;                                     // 16-byte items per epilogue thread and tile
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = sA + S * kA;
    uint8_t *sStage = sB + S * kB;                                   // epilogue staging tile
    float *sRed = (float *)(sStage + Cfg::kStagingBytes);            // [kEpiWarps][2][BN_TILE]
    float *sSS = sRed + Cfg::kRedFloats;                             // scale[BN_TILE], shift[BN_TILE]
    uint64_t *bars = (uint64_t *)(sSS + 2 * BN_TILE);
    uint64_t *full = bars, *empty = bars + S, *tfull = bars + 2 * S, *tempty = bars + 2 * S + 2;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * S + 4);
    bf16 *sXs = (bf16 *)(bars + 2 * S + 6);                          // STEM: two staged input windows

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long *tl = (p.timeline && blockIdx.x == 0) ? p.timeline : nullptr;
#define TL(i) do { if (tl) tl[i] = clock64(); } while (0)
    if (threadIdx.x == 0) {
        TL(0);
        prefetch_tmap(&a_map); prefetch_tmap(&b_map);
        if (p.tma_store) prefetch_tmap(&out_map);
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
    pdl_trigger();                 // the next kernel may begin its own prologue
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TL(1);
    pdl_wait(); B2_TRACE_BEGIN();                    // everything below touches global memory of earlier kernels
    if (threadIdx.x == 0) TL(2);

    if (STEM && warp >= 3 + kEpiWarps) {
        // ---------------- stem: gather im2col rows from the fp32 image into the A stages (no TMA, no HBM im2col buffer)
        const int gt = threadIdx.x - kConvThreads;
        for (int i = gt; i < 2 * kStemXsElems; i += kStemExtraThreads) sXs[i] = __float2bfloat16_rn(0.f);   // pad columns stay zero
        stem_bar();
        StemRaw raw;
        int stage = 0; uint32_t phase = 0, it = 0;
        if ((int)blockIdx.x < p.total_tiles) {
            const TileCoord t0 = decode_tile<BN_TILE>(p, blockIdx.x);
            stem_load(p.stem_x, t0.n_blk, t0.h_blk, gt, raw);
            stem_store(sXs, gt, raw);
        }
        stem_bar();
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
            const bf16 *xs = sXs + (it & 1) * kStemXsElems;
            const bool more = t + (int)gridDim.x < p.total_tiles;
            if (more) {                                     // next tile's rows: loads in flight while this tile is built
                const TileCoord tn = decode_tile<BN_TILE>(p, t + gridDim.x);
                stem_load(p.stem_x, tn.n_blk, tn.h_blk, gt, raw);
            }
#pragma unroll
            for (int kc = 0; kc < 3; ++kc) {
                mbar_wait(&empty[stage], phase ^ 1);
                stem_build_chunk(xs, gt, kc, sA + stage * kA);
                fence_proxy_async_smem();                   // generic-proxy stores -> visible to the tensor core's async reads
                stem_bar();
                if (gt == 0) mbar_arrive(&full[stage]);
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
            if (more) stem_store(sXs + ((it + 1) & 1) * kStemXsElems, gt, raw);
            stem_bar();
        }
    } else if (warp == 0) {
        // ---------------- TMA producer: activation (A) tiles
        if (!STEM && lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile<BN_TILE>(p, t);
                const TapClass &cls = p.cls[tc.z];
                const int ksteps = cls.ntaps * p.c_chunks;
                int tap = 0, cc = 0;
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], kA);
                    tma_load_4d(sA + stage * kA, &a_map, &full[stage], cc * 64, cls.dw[tap],
                                tc.h_blk * p.BH * p.a_hstep + cls.dh[tap], tc.n_blk * p.BN);
                    if (++cc == p.c_chunks) { cc = 0; ++tap; }
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- TMA producer: weight (B) tiles
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile<BN_TILE>(p, t);
                const TapClass &cls = p.cls[tc.z];
                const int ksteps = cls.ntaps * p.c_chunks;
                int tap = 0, cc = 0;
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], kB);
                    uint8_t *dst = sB + stage * kB;
                    if (p.b_mn) {
#pragma unroll
                        for (int j = 0; j < BN_TILE / 64; ++j)
                            tma_load_2d(dst + j * 8192, &b_map, &full[stage], cls.kofs[tap] + tc.ch0 + j * 64, cc * 64);
                    } else {
                        tma_load_2d(dst, &b_map, &full[stage], cls.kofs[tap] + cc * 64, tc.ch0);
                    }
                    if (++cc == p.c_chunks) { cc = 0; ++tap; }
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        // ---------------- MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, BN_TILE, 0, p.b_mn);
            const uint32_t bstep = p.b_mn ? 2048 : 32;     // 16 K-rows of 128 B vs 16 K-elements
            const uint32_t blbo = p.b_mn ? 8192 : kB;      // MN-major: distance between 64-channel atoms
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
                const TileCoord tc = decode_tile<BN_TILE>(p, t);
                const int ksteps = p.cls[tc.z].ntaps * p.c_chunks;
                const int buf = it & 1;
                mbar_wait(&tempty[buf], ((it >> 1) & 1) ^ 1);          // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN_TILE;
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(&full[stage], phase);
                    if (it == 0 && ks == 0) TL(3);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + stage * kA), b0 = smem_u32(sB + stage * kB);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_bf16_ss(d_tmem, make_desc_sw128(a0 + k * 32, 16, 1024), make_desc_sw128(b0 + k * bstep, blbo, 1024), idesc,
                                    (ks > 0 || k > 0) ? 1u : 0u);
                    mma_commit(&empty[stage]);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                mma_commit(&tfull[buf]);
                if (it == 0) TL(4);
            }
        }
    } else {
        // ---------------- epilogue: 8 warps.  TMEM lanes 32*(warp % 4).. are the only ones a warp may read, so two warps
        // share each lane quarter and split the tile's columns; phases B / 2 use all 256 threads as (row, 16-byte chunk) owners
        const int ew = warp - 3;                         // 0..7
        const int q = warp & 3;                          // TMEM lane quarter
        const int hh = ew >> 2;                          // column half of that quarter's rows
        const int r = q * 32 + lane;                     // phase A: tile row
        const int et = ew * 32 + lane;                   // phases B / 2: linear id 0..255
        const int chunk = et % CPR, rsub = et / CPR;     // fixed 8-channel chunk; rows rsub + i * RP
        const int half = chunk >> 3, c8 = chunk & 7;
        const bool want_fwd = p.stats != nullptr, want_bwd = p.bn_bsum != nullptr;
        bf16 *__restrict__ outp = p.out;
        int it = 0, last_tile = -1;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
            const TileCoord tc = decode_tile<BN_TILE>(p, t);
            const TapClass &cls = p.cls[tc.z];
            const int buf = it & 1;
            mbar_wait(&tfull[buf], (it >> 1) & 1);
            if (it == 0 && et == 0) TL(5);
            tc_fence_after();
            // ---- phase A: TMEM -> bf16 -> swizzled staging rows (this warp's half of the columns)
            {
                const int sw = r & 7;
#pragma unroll
                for (int jj = 0; jj < BN_TILE / 64; ++jj) {
                    const int col0 = hh * (BN_TILE / 2) + jj * 32;
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + buf * BN_TILE + ((uint32_t)(q * 32) << 16) + col0, v);
                    tmem_ld_wait();
                    uint8_t *hb = sStage + (col0 >> 6) * 16384 + r * 128;
                    const int cbase = (col0 & 63) >> 3;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        uint4 pk;
                        __nv_bfloat162 *p2 = (__nv_bfloat162 *)&pk;
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            p2[e] = __floats2bfloat162_rn(__uint_as_float(v[k4 * 8 + 2 * e]), __uint_as_float(v[k4 * 8 + 2 * e + 1]));
                        *(uint4 *)(hb + (((cbase + k4) ^ sw) << 4)) = pk;
                    }
                }
            }
            tc_fence_before();
            if (p.tma_store) fence_proxy_async();            // this thread's staging stores -> visible to the TMA engine
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);        // the MMA warp may start tile it+2 into this buffer
            epi_bar();
            if (p.tma_store && et == 0) {
                // the whole 128-pixel tile leaves with one TMA tile store per 64-channel half (box = the A-load box shape)
#pragma unroll
                for (int hf = 0; hf < BN_TILE / 64; ++hf)
                    tma_store_4d(&out_map, sStage + hf * 16384, tc.ch0 + hf * 64, 0, tc.h_blk * p.BH, tc.n_blk * p.BN);
                tma_store_commit();
            }
            last_tile = t;
            if (it == 0 && et == 0) TL(6);
            // ---- phase B: row-coalesced stores (+ read-modify-write, + BatchNorm sums)
            float s0[8], s1[8], mu[8], is[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { s0[e] = 0.f; s1[e] = 0.f; }
            if (want_bwd) {
#pragma unroll
                for (int e = 0; e < 8; ++e) { mu[e] = __ldg(p.bn_mean + tc.ch0 + chunk * 8 + e); is[e] = __ldg(p.bn_invstd + tc.ch0 + chunk * 8 + e); }
            }
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const int row = rsub + i * RP;
                const int64_t off = row_offset(p, tc, row, chunk, cls.oh, cls.ow);
                uint4 pk = *(const uint4 *)(sStage + half * 16384 + row * 128 + ((c8 ^ (row & 7)) << 4));
                __nv_bfloat162 *p2 = (__nv_bfloat162 *)&pk;
                if (p.acc) {
                    const uint4 old = *(const uint4 *)(outp + off);
                    const __nv_bfloat162 *o2 = (const __nv_bfloat162 *)&old;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 a = __bfloat1622float2(p2[e]), b = __bfloat1622float2(o2[e]);
                        p2[e] = __floats2bfloat162_rn(a.x + b.x, a.y + b.y);
                    }
                }
                if (!p.tma_store) *(uint4 *)(outp + off) = pk;
                if (want_fwd) {                        // statistics of the values BatchNorm will read (bf16-rounded)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = __bfloat1622float2(p2[e]);
                        s0[2 * e] += f.x; s1[2 * e] = fmaf(f.x, f.x, s1[2 * e]);
                        s0[2 * e + 1] += f.y; s1[2 * e + 1] = fmaf(f.y, f.y, s1[2 * e + 1]);
                    }
                } else if (want_bwd) {                 // g = dY*[y>0],  g*xhat  of the BatchNorm that produced this tensor
                    const uint4 yr = ldg_nc_u4(p.bn_y + off), xr = ldg_nc_u4(p.bn_x + off);
                    const __nv_bfloat162 *y2 = (const __nv_bfloat162 *)&yr, *x2 = (const __nv_bfloat162 *)&xr;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float2 g = __bfloat1622float2(p2[e]);
                        const float2 yv = __bfloat1622float2(y2[e]), xv = __bfloat1622float2(x2[e]);
                        if (p.bn_relu) { if (!(yv.x > 0.f)) g.x = 0.f; if (!(yv.y > 0.f)) g.y = 0.f; }
                        s0[2 * e] += g.x; s1[2 * e] = fmaf(g.x, (xv.x - mu[2 * e]) * is[2 * e], s1[2 * e]);
                        s0[2 * e + 1] += g.y; s1[2 * e + 1] = fmaf(g.y, (xv.y - mu[2 * e + 1]) * is[2 * e + 1], s1[2 * e + 1]);
                    }
                }
            }
            if (it == 0 && et == 0) TL(7);
            if (want_fwd || want_bwd) {
                // lanes with equal (lane % CPR) own the same channels: butterfly over the others, then over the 8 warps
#pragma unroll
                for (int o = CPR; o < 32; o <<= 1) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) { s0[e] += __shfl_xor_sync(0xffffffffu, s0[e], o); s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], o); }
                }
                float *mine = sRed + ew * 2 * BN_TILE;
                if (lane < CPR) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) { mine[lane * 8 + e] = s0[e]; mine[BN_TILE + lane * 8 + e] = s1[e]; }
                }
                epi_bar();
                float *dst = (want_fwd ? p.stats : p.bn_bsum) + (size_t)(blockIdx.x % kBnSumCopies) * 2 * p.out_C;
                for (int j = et; j < 2 * BN_TILE; j += kEpiThreads) {
                    float v = 0.f;
#pragma unroll
                    for (int w8 = 0; w8 < kEpiWarps; ++w8) v += sRed[w8 * 2 * BN_TILE + j];
                    const int which = j / BN_TILE, c = j - which * BN_TILE;
                    atomicAdd(dst + which * p.out_C + tc.ch0 + c, v);
                }
            }
            // the staging tile is rewritten by the next tile's phase A: the TMA engine must have read it
            if (p.tma_store && et == 0 && t + (int)gridDim.x < p.total_tiles) tma_store_wait_read();
            epi_bar();                                     // staging tile and sRed are free again
            if (it == 0 && et == 0) TL(8);
        }
        if (p.tma_store && et == 0) tma_store_wait_all();  // every tile store of this CTA is performed (pass 2 / the next kernel read them)
        if (p.fuse_bn) {
            // ---- grid-wide barrier: every CTA of this launch has stored its tiles and contributed its sums.
            // One thread fences (cumulative over the CTA's writes ordered before it by the bar.sync) and arrives.
            if (et == 0) TL(9);
            if (et == 0) {
                __threadfence();
                const uint32_t gen = ld_acquire_gpu_u32(p.fb_sync + 1);
                if (atomicAdd(p.fb_sync, 1u) == gridDim.x - 1) {
                    *(volatile uint32_t *)p.fb_sync = 0u;               // self-resetting: ready for the next launch
                    __threadfence();
                    st_release_gpu_u32(p.fb_sync + 1, gen + 1);
                } else {
                    const long long t0 = clock64();
                    while (ld_acquire_gpu_u32(p.fb_sync + 1) == gen)
                        if (clock64() - t0 > 8000000000LL) { *p.fb_err = 1; __threadfence_system(); __trap(); }   // ~4 s: fail loudly
                }
                __threadfence();
            }
            epi_bar();
            if (et == 0) TL(10);
            // ---- pass 2: y = relu(bn(z) + res) over this CTA's own tiles: the last one is still in the staging
            // tile (shared memory), earlier ones are re-read from L2
            int cur_n = -1;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile<BN_TILE>(p, t);
                if (tc.n != cur_n) {
                    cur_n = tc.n;
                    epi_bar();
                    for (int c = et; c < BN_TILE; c += kEpiThreads) {
                        const int ch = tc.ch0 + c;
                        float t0s = 0.f, t1s = 0.f;
#pragma unroll
                        for (int k = 0; k < kBnSumCopies; ++k) { t0s += __ldcg(p.stats + (2 * k) * p.out_C + ch); t1s += __ldcg(p.stats + (2 * k + 1) * p.out_C + ch); }
                        // mean / variance in double (E[x^2] - mean^2 cancels), 1/sqrt in fp32 like torch's fp32 kernels
                        const double mean_d = (double)t0s * p.fb_inv_count;
                        double var_d = fma(-mean_d, mean_d, (double)t1s * p.fb_inv_count);
                        if (var_d < 0.0) var_d = 0.0;
                        const float mean = (float)mean_d, var = (float)var_d;
                        const float istd = rsqrtf(var + p.fb_eps);
                        const float g = p.fb_gamma[ch];
                        sSS[c] = istd * g;
                        sSS[BN_TILE + c] = fmaf(-mean, istd * g, p.fb_beta[ch]);
                        if (tc.m == 0 && tc.z == 0) {                 // exactly one tile per channel block records the statistics
                            p.fb_save_mean[ch] = mean; p.fb_save_invstd[ch] = istd;
                            if (p.fb_running_mean) {
                                const float unbiased = p.fb_count > 1.f ? var * p.fb_count / (p.fb_count - 1.f) : var;
                                p.fb_running_mean[ch] = (1.f - p.fb_momentum) * p.fb_running_mean[ch] + p.fb_momentum * mean;
                                p.fb_running_var[ch] = (1.f - p.fb_momentum) * p.fb_running_var[ch] + p.fb_momentum * unbiased;
                            }
                        }
                    }
                    epi_bar();
                }
                float sc[8], sh[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { sc[e] = sSS[chunk * 8 + e]; sh[e] = sSS[BN_TILE + chunk * 8 + e]; }
                const bool staged = t == last_tile;
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const int row = rsub + i * RP;
                    const int64_t off = row_offset(p, tc, row, chunk, 0, 0);
                    uint4 zr, rr = make_uint4(0, 0, 0, 0);
                    if (staged) zr = *(const uint4 *)(sStage + half * 16384 + row * 128 + ((c8 ^ (row & 7)) << 4));
                    else zr = __ldcg((const uint4 *)(outp + off));
                    if (p.fb_res) rr = ldg_nc_u4(p.fb_res + off);
                    const __nv_bfloat162 *z2 = (const __nv_bfloat162 *)&zr, *r2 = (const __nv_bfloat162 *)&rr;
                    uint4 pk;
                    __nv_bfloat162 *p2 = (__nv_bfloat162 *)&pk;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 zf = __bfloat1622float2(z2[e]);
                        float a = fmaf(zf.x, sc[2 * e], sh[2 * e]), b = fmaf(zf.y, sc[2 * e + 1], sh[2 * e + 1]);
                        if (p.fb_res) { const float2 rf = __bfloat1622float2(r2[e]); a += rf.x; b += rf.y; }
                        if (p.fb_relu) { a = a > 0.f ? a : 0.f; b = b > 0.f ? b : 0.f; }
                        p2[e] = __floats2bfloat162_rn(a, b);
                    }
                    *(uint4 *)(p.fb_y + off) = pk;
                }
            }
            if (et == 0) TL(11);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (threadIdx.x == 0) TL(12);
    B2_TRACE_END();
#undef TL
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::kTmemCols); }
}

// =================================================================================
// device: wgrad
// =================================================================================
constexpr int kStagesWg = 3;
constexpr int kBoxBytes = 128 * 128;         // [128 pixel rows][64 bf16]
constexpr int kWgStageBytes = 4 * kBoxBytes; // up to 2 dY boxes + 2 X boxes
constexpr int kWgSmem = kStagesWg * kWgStageBytes + 1024 + 256 + 2 * kStemXsElemsC * 2;
constexpr int kWgThreads = 224;              // warps: 0 dY-TMA, 1 MMA, 2..5 epilogue, 6 X-TMA  (+ 7..10 gather, stem variant)

template <bool STEM>
__global__ void __launch_bounds__(kWgThreads + (STEM ? kStemExtraThreads : 0), 1)
k_wgrad(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap dy_map,
        const __grid_constant__ WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(smem + kStagesWg * kWgStageBytes);
    uint64_t *full = bars, *empty = bars + kStagesWg, *tmem_full = bars + 2 * kStagesWg;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * kStagesWg + 1);
    bf16 *sXs = (bf16 *)(bars + 2 * kStagesWg + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // N tile: up to cpt 64-wide K chunks (chunk index over live taps x channel chunks); a stage holds mb dY boxes
    // followed by nb X boxes, mb + nb <= 4: cpt = 3 for 64-channel layers (one dY box), 2 otherwise
    const int total_chunks = p.ntaps_alive * p.c_chunks;
    const int chunk0 = blockIdx.x * p.cpt;
    const int nb = total_chunks - chunk0 < p.cpt ? total_chunks - chunk0 : p.cpt;
    const int mb = p.m_boxes;
    const int co0 = blockIdx.y * 64 * mb;
    const int pt0 = blockIdx.z * p.tiles_per_split;
    int pt1 = pt0 + p.tiles_per_split; if (pt1 > p.pix_tiles) pt1 = p.pix_tiles;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&x_map); prefetch_tmap(&dy_map);
        for (int i = 0; i < kStagesWg; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait(); B2_TRACE_BEGIN();

    if (warp == 0) {
        // ---- TMA producer: dY boxes (the two producers' streams overlap in the TMA unit)
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int pt = pt0; pt < pt1; ++pt) {
                uint8_t *st = smem + stage * kWgStageBytes;
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], mb * kBoxBytes);
                for (int j = 0; j < mb; ++j)
                    tma_load_2d(st + j * kBoxBytes, &dy_map, &full[stage], co0 + j * 64, pt * 128);
                if (++stage == kStagesWg) { stage = 0; phase ^= 1; }
            }
        }
    } else if (STEM && warp >= 7) {
        // ---- stem: the "shifted input" operand is the im2col tile, gathered from the fp32 image (k-chunks chunk0..)
        const int gt = threadIdx.x - kWgThreads;
        for (int i = gt; i < 2 * kStemXsElems; i += kStemExtraThreads) sXs[i] = __float2bfloat16_rn(0.f);
        stem_bar();
        StemRaw raw;
        int stage = 0; uint32_t phase = 0, it = 0;
        if (pt0 < pt1) { stem_load(p.stem_x, pt0 / p.h_blocks, pt0 % p.h_blocks, gt, raw); stem_store(sXs, gt, raw); }
        stem_bar();
        for (int pt = pt0; pt < pt1; ++pt, ++it) {
            const bf16 *xs = sXs + (it & 1) * kStemXsElems;
            const bool more = pt + 1 < pt1;
            if (more) stem_load(p.stem_x, (pt + 1) / p.h_blocks, (pt + 1) % p.h_blocks, gt, raw);
            uint8_t *st = smem + stage * kWgStageBytes;
            mbar_wait(&empty[stage], phase ^ 1);
            for (int j = 0; j < nb; ++j) stem_build_chunk(xs, gt, chunk0 + j, st + (mb + j) * kBoxBytes);
            fence_proxy_async_smem();
            stem_bar();
            if (gt == 0) mbar_arrive(&full[stage]);
            if (++stage == kStagesWg) { stage = 0; phase ^= 1; }
            if (more) stem_store(sXs + ((it + 1) & 1) * kStemXsElems, gt, raw);
            stem_bar();
        }
    } else if (warp == 6) {
        // ---- TMA producer: shifted input boxes
        if (!STEM && lane == 0) {
            int tap[3], cc[3];
            for (int j = 0; j < nb; ++j) { int c = chunk0 + j; tap[j] = c / p.c_chunks; cc[j] = c - tap[j] * p.c_chunks; }
            int stage = 0; uint32_t phase = 0;
            for (int pt = pt0; pt < pt1; ++pt) {
                const int n_blk = pt / p.h_blocks, h_blk = pt - n_blk * p.h_blocks;
                uint8_t *st = smem + stage * kWgStageBytes;
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], nb * kBoxBytes);
                for (int j = 0; j < nb; ++j)
                    tma_load_4d(st + (mb + j) * kBoxBytes, &x_map, &full[stage], cc[j] * 64, p.dw[tap[j]],
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
                const uint32_t a0 = smem_u32(smem + stage * kWgStageBytes), b0 = a0 + mb * kBoxBytes;
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
        float *dst_row = p.out + ((int64_t)(p.atomic ? 0 : blockIdx.z) * p.Co + co0 + row) * p.out_pitch;
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
                    if (p.atomic) {
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            red_add_f4((float *)(d4 + e), __uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]),
                                       __uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            d4[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]),
                                                __uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3]));
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    B2_TRACE_END();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// =================================================================================
// host: plans and launches
// =================================================================================
static std::once_flag g_attr_once;
static void set_attrs() {
    cudaFuncSetAttribute(k_conv<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<64>::kSmem);
    cudaFuncSetAttribute(k_conv<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<128>::kSmem);
    cudaFuncSetAttribute(k_conv<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<64>::kSmem);
    cudaFuncSetAttribute(k_conv<64, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<64, 3>::kSmem);
    cudaFuncSetAttribute(k_wgrad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
    cudaFuncSetAttribute(k_wgrad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
}

// Output-channel tile: 128 when the 64-wide tiling would need more than one wave of CTAs (halves the activation
// re-reads and the shared-memory traffic per FLOP); 64 otherwise (more CTAs, shorter epilogues).
// (Measured and removed in round 2: split-K over an fp32 workspace with red.global.add -- 42 us against 22 us for
//  the unsplit layer3 launch: the L2 atomics and the read-back cost more than the shorter main loop saves.)
static int pick_bn_tile(int64_t m_tiles, int channels, int classes) {
    if (channels % 128 != 0) return 64;
    return m_tiles * (channels / 64) * classes > (int64_t)umma_num_sms() ? 128 : 64;
}

static void finish_plan(UmmaConvPlan *plan, int64_t m_tiles, int channels) {
    ConvParams &p = plan->p;
    p.lgBW = ilog2(p.BW); p.lgBH = ilog2(p.BH);
    p.m_tiles = (int)m_tiles; p.n_tiles = channels / plan->bn_tile;
    p.total_tiles = p.m_tiles * p.n_tiles * p.nclasses;
    plan->stem = false;
    static const bool shallow = [] { const char *e = getenv("B2_CONV_SHALLOW"); return e && e[0] == '1'; }();
    plan->shallow = shallow;
    plan->grid = std::min(p.total_tiles, umma_num_sms());
}

int umma_plan_fprop(const ConvGeom &g, const bf16 *x, const bf16 *w, bf16 *y, UmmaConvPlan *plan) {
    plan->valid = false;
    B2_CHECK(umma_conv_eligible(g), B2_ERR_UNSUPPORTED, "umma fprop: shape not eligible");
    B2_TRY(umma_init());
    ConvParams &p = plan->p;
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
    const int64_t m_tiles = (int64_t)(g.N / p.BN) * p.h_blocks;
    plan->bn_tile = pick_bn_tile(m_tiles, g.Co, 1);
    B2_TRY(make_map_nhwc(&plan->a_map, x, g.N, g.Hi, g.Wi, g.Ci, p.BW, p.BH, p.BN, g.stride));
    B2_TRY(make_map_2d(&plan->b_map, w, g.Co, (int64_t)g.kh * g.kw * g.Ci, plan->bn_tile));
    finish_plan(plan, m_tiles, g.Co);
    // output tile store: the conv output has the tile's own pixel geometry (dense lattice): one box per 64 channels
    static const bool tma_st = [] { const char *e = getenv("B2_TMA_STORE"); return !(e && e[0] == '0'); }();
    B2_TRY(make_map_nhwc(&plan->out_map, y, g.N, g.Ho, g.Wo, g.Co, p.BW, p.BH, p.BN, 1));
    p.tma_store = tma_st ? 1 : 0;
    plan->flops = 2.0 * g.M() * g.Co * (double)c.ntaps * g.Ci;
    plan->valid = true;
    return 0;
}

int umma_plan_dgrad(const ConvGeom &g, const bf16 *dy, const bf16 *w, bf16 *dx, bool acc, UmmaConvPlan *plan) {
    plan->valid = false;
    B2_CHECK(umma_conv_eligible(g), B2_ERR_UNSUPPORTED, "umma dgrad: shape not eligible");
    B2_TRY(umma_init());
    ConvParams &p = plan->p;
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
    const int64_t m_tiles = (int64_t)(g.N / p.BN) * p.h_blocks;
    plan->bn_tile = pick_bn_tile(m_tiles, g.Ci, p.nclasses);
    B2_TRY(make_map_nhwc(&plan->a_map, dy, g.N, g.Ho, g.Wo, g.Co, p.BW, p.BH, p.BN, 1));
    B2_TRY(make_map_2d(&plan->b_map, w, g.Co, (int64_t)g.kh * g.kw * g.Ci, 64));
    finish_plan(plan, m_tiles, g.Ci);
    plan->out_map = plan->a_map;      // unused (tma_store = 0: strided lattices / read-modify-write go through st.global)
    plan->flops = 2.0 * (double)g.N * g.Ho * g.Wo * g.Ci * taps_total * g.Co;
    plan->valid = true;
    return 0;
}

int umma_conv_check_coresident(const UmmaConvPlan &plan) {
    std::call_once(g_attr_once, set_attrs);
    int per_sm = 0;
    if (plan.stem) B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_conv<64, true>, kConvThreads + kStemExtraThreads, ConvCfg<64>::kSmem));
    else if (plan.bn_tile == 64 && plan.shallow) { B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_conv<64, false, 3>, kConvThreads, (ConvCfg<64, 3>::kSmem))); per_sm = per_sm > 1 ? 1 : per_sm; }
    else if (plan.bn_tile == 128) B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_conv<128, false>, kConvThreads, ConvCfg<128>::kSmem));
    else B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_conv<64, false>, kConvThreads, ConvCfg<64>::kSmem));
    B2_CHECK(per_sm >= 1 && plan.grid <= per_sm * umma_num_sms(), B2_ERR_UNSUPPORTED,
             "fused conv+BatchNorm: grid of %d CTAs is not co-resident (%d per SM x %d SMs)", plan.grid, per_sm, umma_num_sms());
    return 0;
}

int umma_conv_launch(const UmmaConvPlan &plan, cudaStream_t s) {
    B2_CHECK(plan.valid, B2_ERR_STATE, "umma conv: plan not valid");
    std::call_once(g_attr_once, set_attrs);
    if (plan.stem) {
        B2_CHECK(plan.p.stem_x && plan.bn_tile == 64, B2_ERR_STATE, "umma conv: stem plan without an input image / unexpected tiling");
        B2_LAUNCH((k_conv<64, true>), plan.grid, kConvThreads + kStemExtraThreads, ConvCfg<64>::kSmem, s, plan.a_map, plan.b_map, plan.out_map, plan.p);
    } else if (plan.bn_tile == 64 && plan.shallow) {
        // 3-stage ring, 95 KB of shared memory: two CTAs fit on an SM, so the NEXT launch's CTAs (programmatic dependent
        // launch) are resident and past their prologue while this grid is still in its barrier / pass 2
        B2_LAUNCH((k_conv<64, false, 3>), plan.grid, kConvThreads, (ConvCfg<64, 3>::kSmem), s, plan.a_map, plan.b_map, plan.out_map, plan.p);
    } else if (plan.bn_tile == 128) { B2_LAUNCH((k_conv<128, false>), plan.grid, kConvThreads, ConvCfg<128>::kSmem, s, plan.a_map, plan.b_map, plan.out_map, plan.p); }
    else { B2_LAUNCH((k_conv<64, false>), plan.grid, kConvThreads, ConvCfg<64>::kSmem, s, plan.a_map, plan.b_map, plan.out_map, plan.p); }
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
    const int cpt = 2;        // K chunks per CTA (3 for the 64-channel layers was measured slower: 15.5-17.4 vs 13.3-13.9 us)
    *k_tiles = (chunks + cpt - 1) / cpt;
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
                    bool deterministic, UmmaWgradPlan *plan) {
    plan->valid = false;
    B2_CHECK(umma_conv_eligible(g), B2_ERR_UNSUPPORTED, "umma wgrad: shape not eligible");
    WgradParams &p = plan->p;
    memset(&p, 0, sizeof(p));
    int kt;
    wgrad_shape(g, &p.ntaps_alive, &p.m_boxes, &kt, &p.pix_tiles, &p.splits, &p.tiles_per_split);
    B2_CHECK(tile_geometry(g.N, g.Ho, g.Wo, &p.BW, &p.BH, &p.BN), B2_ERR_UNSUPPORTED, "umma wgrad: tile geometry");
    p.h_blocks = g.Ho / p.BH; p.a_hstep = g.stride; p.c_chunks = g.Ci / 64; p.Ci = g.Ci; p.Co = g.Co; p.K = g.K();
    p.cpt = 2;
    int n = 0;
    for (int kh = 0; kh < g.kh; ++kh)
        for (int kw = 0; kw < g.kw; ++kw)
            if (tap_alive(kh - g.pad, g.stride, g.Ho, g.Hi) && tap_alive(kw - g.pad, g.stride, g.Wo, g.Wi)) {
                p.tap_id[n] = kh * g.kw + kw; p.dh[n] = kh - g.pad; p.dw[n] = kw - g.pad; ++n;
            }
    plan->need_zero = n < g.kh * g.kw;
    plan->dw = dw;
    plan->stem = false;
    if (p.splits > 1 && deterministic) {
        B2_CHECK((size_t)p.splits * g.Co * n * g.Ci <= partial_floats, B2_ERR_ARG, "umma wgrad: workspace too small");
        p.out = partial; p.out_pitch = n * g.Ci; p.compact = 1; p.atomic = 0;
    } else { p.out = dw; p.out_pitch = g.K(); p.compact = 0; p.atomic = p.splits > 1 ? 1 : 0; }
    B2_TRY(make_map_nhwc(&plan->x_map, x, g.N, g.Hi, g.Wi, g.Ci, p.BW, p.BH, p.BN, g.stride));
    B2_TRY(make_map_2d(&plan->dy_map, dy, g.M(), g.Co, 128));
    plan->grid = dim3((unsigned)kt, (unsigned)(g.Co / (64 * p.m_boxes)), (unsigned)p.splits);
    plan->flops = 2.0 * g.M() * g.Co * (double)n * g.Ci;
    plan->valid = true;
    return 0;
}

// dW[co][tap_id*Ci + c] = sum_z partial[z][co][a*Ci + c]   (a = live-tap index), float4 lanes
__global__ void k_wgrad_reduce(const float *__restrict__ part, float *__restrict__ dw, WgradParams p) {
    pdl_sync(); B2_TRACE_BEGIN();
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
    if (plan.stem) {
        B2_CHECK(plan.p.stem_x, B2_ERR_STATE, "umma wgrad: stem plan without an input image");
        B2_LAUNCH(k_wgrad<true>, plan.grid, kWgThreads + kStemExtraThreads, kWgSmem, s, plan.x_map, plan.dy_map, plan.p);
    } else {
        B2_LAUNCH(k_wgrad<false>, plan.grid, kWgThreads, kWgSmem, s, plan.x_map, plan.dy_map, plan.p);
    }
    if (p.splits > 1 && !p.atomic) {
        int64_t n4 = (int64_t)p.Co * p.out_pitch / 4;
        int blocks = (int)ceil_div(n4, 256); if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        B2_LAUNCH(k_wgrad_reduce, blocks, 256, 0, s, p.out, plan.dw, p);
    }
    return 0;
}

}  // namespace b2

namespace b2 { B2_TRACE_INSTALLER(trace_install_umma) }
