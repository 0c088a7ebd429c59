// CUDA-core implicit-GEMM convolution: fprop, dgrad, wgrad in fp32 FMA.
//
// This is the arithmetic "parity mode" (fp32 storage) and the any-shape path (Net's 5x5
// convs with Ci=3/6, the 3-channel ResNet stem, the N=10 classifier) -- shapes whose
// K or N are too small / ragged for tcgen05 tiles.  Replaces what torch dispatches for
// conv2d / linear forward+backward under cifar_example_ddp.py:104,106.
//
// All three are C[M][N] (+)= sum_k A(m,k) * B(k,n) with 64x64x16 tiles, 256 threads,
// 4x4 micro-tiles, operands gathered on the fly (no materialised im2col).
#define B2_TRACE_TU_ID 2
#include "ops.cuh"

namespace b2 {

constexpr int BM = 64, BN = 64, BK = 16;

enum { MODE_FPROP = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

struct SimtArgs {
    ConvGeom g;
    const void *a;      // FPROP: x      DGRAD: dy     WGRAD: dy
    const void *b;      // FPROP: w      DGRAD: w      WGRAD: x
    void *c;            // FPROP: y      DGRAD: dx     WGRAD: partial/dw (fp32)
    const float *bias;
    int wpitch;
    int relu;
    int acc;
    int64_t M; int N; int64_t K;        // GEMM extents (K = full contraction length)
    int64_t k_per_split;                // WGRAD: contraction chunk per blockIdx.z
};

// TA: activation type (x, dy, y, dx); TW: weight type
template <int MODE, typename TA, typename TW, typename TC>
__global__ void __launch_bounds__(256)
k_conv_simt(const SimtArgs p) {
    pdl_sync(); B2_TRACE_BEGIN();
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const ConvGeom &g = p.g;
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int taps = g.kh * g.kw;

    int64_t kbeg = 0, kend = p.K;
    if (MODE == MODE_WGRAD) {
        kbeg = (int64_t)blockIdx.z * p.k_per_split;
        kend = kbeg + p.k_per_split; if (kend > p.K) kend = p.K;
    }

    // ---- per-thread load assignments ------------------------------------------
    // "k-fast" operand (contiguous along k in memory): thread loads rows r = tid/16 + 16*i, col tid%16
    // "mn-fast" operand (contiguous along m or n):     thread loads col   = tid%64, k rows tid/64 + 4*i
    const int lk = tid & 15, lr = tid >> 4;     // k-fast
    const int lc = tid & 63, lkk = tid >> 6;    // mn-fast

    // FPROP: A k-fast (x gather: c fastest), B k-fast (w[n][k])
    // DGRAD: A k-fast (dy gather: co fastest), B n-fast (w[(co*taps+tap)*Ci + n])
    // WGRAD: A m-fast (dy[r][co]), B n-fast (x gather: c fastest along n)
    int a_n[4], a_h[4], a_w[4];   // decomposed rows for gather operands (FPROP/DGRAD A rows)
    bool a_ok[4];
    if (MODE == MODE_FPROP || MODE == MODE_DGRAD) {
        const int H = MODE == MODE_FPROP ? g.Ho : g.Hi, W = MODE == MODE_FPROP ? g.Wo : g.Wi;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int64_t m = m0 + lr + 16 * i;
            a_ok[i] = m < p.M;
            int64_t mm = a_ok[i] ? m : 0;
            a_w[i] = mm % W; mm /= W;
            a_h[i] = mm % H; a_n[i] = (int)(mm / H);
        }
    }
    // WGRAD B columns: n -> (tap, c) fixed per thread
    int b_tap = 0, b_c = 0; bool b_ok = false;
    if (MODE == MODE_WGRAD) {
        int n = n0 + lc;
        b_ok = n < p.N;
        int nn = b_ok ? n : 0;
        b_tap = nn / g.Ci; b_c = nn - b_tap * g.Ci;
    }

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;
    const TA *A = (const TA *)p.a;

    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
        // ---------------- load A tile -> As[k][m] ----------------
        if (MODE == MODE_FPROP) {
            int64_t k = k0 + lk;
            bool kok = k < kend;
            int tap = kok ? (int)(k / g.Ci) : 0; int c = kok ? (int)(k - (int64_t)tap * g.Ci) : 0;
            int fh = tap / g.kw, fw = tap - fh * g.kw;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int hi = a_h[i] * g.stride - g.pad + fh, wi = a_w[i] * g.stride - g.pad + fw;
                float v = 0.f;
                if (kok && a_ok[i] && hi >= 0 && hi < g.Hi && wi >= 0 && wi < g.Wi)
                    v = to_f<TA>(A[(((int64_t)a_n[i] * g.Hi + hi) * g.Wi + wi) * g.Ci + c]);
                As[lk][lr + 16 * i] = v;
            }
        } else if (MODE == MODE_DGRAD) {
            int64_t k = k0 + lk;
            bool kok = k < kend;
            int tap = kok ? (int)(k / g.Co) : 0; int co = kok ? (int)(k - (int64_t)tap * g.Co) : 0;
            int fh = tap / g.kw, fw = tap - fh * g.kw;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int th = a_h[i] + g.pad - fh, tw = a_w[i] + g.pad - fw;
                float v = 0.f;
                if (kok && a_ok[i] && th >= 0 && tw >= 0 && (th % g.stride) == 0 && (tw % g.stride) == 0) {
                    int ho = th / g.stride, wo = tw / g.stride;
                    if (ho < g.Ho && wo < g.Wo)
                        v = to_f<TA>(A[(((int64_t)a_n[i] * g.Ho + ho) * g.Wo + wo) * g.Co + co]);
                }
                As[lk][lr + 16 * i] = v;
            }
        } else {  // WGRAD: A(m=co, r) = dy[r*Co + co]
            int co = (int)m0 + lc;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int64_t r = k0 + lkk + 4 * i;
                float v = 0.f;
                if (r < kend && co < p.M) v = to_f<TA>(A[r * g.Co + co]);
                As[lkk + 4 * i][lc] = v;
            }
        }
        // ---------------- load B tile -> Bs[k][n] ----------------
        if (MODE == MODE_FPROP) {
            const TW *Wt = (const TW *)p.b;
            int64_t k = k0 + lk;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int n = n0 + lr + 16 * i;
                float v = 0.f;
                if (k < kend && n < p.N) v = to_f<TW>(Wt[(int64_t)n * p.wpitch + k]);
                Bs[lk][lr + 16 * i] = v;
            }
        } else if (MODE == MODE_DGRAD) {
            const TW *Wt = (const TW *)p.b;
            int n = n0 + lc;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int64_t k = k0 + lkk + 4 * i;
                float v = 0.f;
                if (k < kend && n < p.N) {
                    int tap = (int)(k / g.Co); int co = (int)(k - (int64_t)tap * g.Co);
                    v = to_f<TW>(Wt[(int64_t)co * p.wpitch + (int64_t)tap * g.Ci + n]);
                }
                Bs[lkk + 4 * i][lc] = v;
            }
        } else {  // WGRAD: B(r, n) = x gather at output pixel r, tap/channel of n
            const TA *X = (const TA *)p.b;
            int fh = b_tap / g.kw, fw = b_tap - fh * g.kw;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int64_t r = k0 + lkk + 4 * i;
                float v = 0.f;
                if (r < kend && b_ok) {
                    int wo = r % g.Wo; int64_t t = r / g.Wo;
                    int ho = t % g.Ho; int n = (int)(t / g.Ho);
                    int hi = ho * g.stride - g.pad + fh, wi = wo * g.stride - g.pad + fw;
                    if (hi >= 0 && hi < g.Hi && wi >= 0 && wi < g.Wi)
                        v = to_f<TA>(X[(((int64_t)n * g.Hi + hi) * g.Wi + wi) * g.Ci + b_c]);
                }
                Bs[lkk + 4 * i][lc] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 av = *(const float4 *)&As[kk][tm];
            float4 bv = *(const float4 *)&Bs[kk][tn];
            float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
        }
        __syncthreads();
    }
    (void)taps;
    // ---------------- epilogue ----------------
    TC *C = (TC *)p.c;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t m = m0 + tm + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tn + j;
            if (n >= p.N) continue;
            float v = acc[i][j];
            if (MODE == MODE_FPROP) {
                if (p.bias) v += p.bias[n];
                if (p.relu) v = v > 0.f ? v : 0.f;
                C[m * p.N + n] = from_f<TC>(v);
            } else if (MODE == MODE_DGRAD) {
                if (p.acc) v += to_f<TC>(C[m * p.N + n]);
                C[m * p.N + n] = from_f<TC>(v);
            } else {
                // partial[z][co][k]
                C[((int64_t)blockIdx.z * p.M + m) * p.N + n] = from_f<TC>(v);
            }
        }
    }
}

__global__ void k_reduce_partials(const float *__restrict__ part, float *__restrict__ out, int64_t n, int splits) {
    pdl_sync(); B2_TRACE_BEGIN();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += part[(int64_t)z * n + i];
        out[i] = s;
    }
}

#define DISPATCH2(adt, wdt, TA, TW, ...)                                        \
    if ((adt) == DT_F32 && (wdt) == DT_F32) { typedef float TA; typedef float TW; __VA_ARGS__; }       \
    else if ((adt) == DT_BF16 && (wdt) == DT_BF16) { typedef bf16 TA; typedef bf16 TW; __VA_ARGS__; }  \
    else if ((adt) == DT_BF16 && (wdt) == DT_F32) { typedef bf16 TA; typedef float TW; __VA_ARGS__; }  \
    else { typedef float TA; typedef bf16 TW; __VA_ARGS__; }

int conv_simt_fprop(const ConvGeom &g, const void *x, int dt, const void *w, int wdt, int wpitch, const float *bias,
                    bool relu, void *y, int ydt, cudaStream_t s) {
    SimtArgs p{};
    p.g = g; p.a = x; p.b = w; p.c = y; p.bias = bias; p.wpitch = wpitch; p.relu = relu; p.acc = 0;
    p.M = g.M(); p.N = g.Co; p.K = g.K(); p.k_per_split = p.K;
    dim3 grid((unsigned)ceil_div(p.M, BM), (unsigned)ceil_div(p.N, BN), 1);
    if (ydt == dt) {
        DISPATCH2(dt, wdt, TA, TW, B2_LAUNCH((k_conv_simt<MODE_FPROP, TA, TW, TA>), grid, 256, 0, s, p));
    } else {
        B2_CHECK(ydt == DT_F32, B2_ERR_ARG, "conv fprop: output type must match input or be fp32");
        DISPATCH2(dt, wdt, TA, TW, B2_LAUNCH((k_conv_simt<MODE_FPROP, TA, TW, float>), grid, 256, 0, s, p));
    }
    return 0;
}

int conv_simt_dgrad(const ConvGeom &g, const void *dy, int dt, const void *w, int wdt, int wpitch, void *dx, bool acc,
                    cudaStream_t s) {
    SimtArgs p{};
    p.g = g; p.a = dy; p.b = w; p.c = dx; p.bias = nullptr; p.wpitch = wpitch; p.relu = 0; p.acc = acc;
    p.M = (int64_t)g.N * g.Hi * g.Wi; p.N = g.Ci; p.K = (int64_t)g.kh * g.kw * g.Co; p.k_per_split = p.K;
    dim3 grid((unsigned)ceil_div(p.M, BM), (unsigned)ceil_div(p.N, BN), 1);
    DISPATCH2(dt, wdt, TA, TW, B2_LAUNCH((k_conv_simt<MODE_DGRAD, TA, TW, TA>), grid, 256, 0, s, p));
    return 0;
}

static int wgrad_splits(const ConvGeom &g) {
    int64_t tiles = ceil_div(g.Co, BM) * ceil_div(g.K(), BN);
    int64_t R = g.M();
    int64_t want = ceil_div((int64_t)kNumSMs * 3, tiles);
    int64_t maxs = ceil_div(R, 256);
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    if (want > 512) want = 512;
    return (int)want;
}
size_t conv_simt_wgrad_workspace(const ConvGeom &g) {
    int sp = wgrad_splits(g);
    return sp > 1 ? (size_t)sp * g.Co * g.K() : 0;
}
int conv_simt_wgrad(const ConvGeom &g, const void *x, const void *dy, int dt, float *dw, float *partial,
                    size_t partial_floats, cudaStream_t s) {
    SimtArgs p{};
    int sp = wgrad_splits(g);
    p.g = g; p.a = dy; p.b = x; p.bias = nullptr; p.wpitch = g.K(); p.relu = 0; p.acc = 0;
    p.M = g.Co; p.N = g.K(); p.K = g.M();
    int64_t per = ceil_div(p.K, sp);
    per = ceil_div(per, BK) * BK;
    sp = (int)ceil_div(p.K, per);
    p.k_per_split = per;
    if (sp > 1) {
        B2_CHECK((size_t)sp * g.Co * g.K() <= partial_floats, B2_ERR_ARG, "wgrad: workspace too small");
        p.c = partial;
    } else {
        p.c = dw;
    }
    dim3 grid((unsigned)ceil_div(p.M, BM), (unsigned)ceil_div(p.N, BN), sp);
    if (dt == DT_F32) { B2_LAUNCH((k_conv_simt<MODE_WGRAD, float, float, float>), grid, 256, 0, s, p); }
    else { B2_LAUNCH((k_conv_simt<MODE_WGRAD, bf16, float, float>), grid, 256, 0, s, p); }
    if (sp > 1) {
        int64_t n = (int64_t)g.Co * g.K();
        int blocks = (int)ceil_div(n, 256); if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        B2_LAUNCH(k_reduce_partials, blocks, 256, 0, s, partial, dw, n, sp);
    }
    return 0;
}

}  // namespace b2

namespace b2 { B2_TRACE_INSTALLER(trace_install_simt) }
