// tcgen05 / TMA implicit-GEMM convolution for NHWC bf16 activations (fprop, dgrad, wgrad).
#pragma once
#include <cuda.h>
#include "ops.cuh"

namespace b2 {

constexpr int kMaxTaps = 9;

// One group of filter taps sharing an output-pixel lattice: fprop / stride-1 dgrad have a
// single class (all taps, dense outputs); a stride-2 dgrad has one class per output parity.
struct TapClass {
    int ntaps;
    int oh, ow;                 // output lattice offset (pixel = box_pixel * os + (oh, ow))
    int dh[kMaxTaps], dw[kMaxTaps];   // source-box coordinate offsets of each tap
    int kofs[kMaxTaps];         // element offset of the tap's 64-wide K slices in the B matrix
};

// Parameters of the persistent fprop / dgrad kernel k_conv<BN_TILE> (conv_umma.cu).
struct ConvParams {
    int nclasses;
    TapClass cls[4];
    int c_chunks;               // source channels / 64
    int BW, BH, BN, h_blocks;   // 128-pixel tile = BN images x BH rows x BW cols (box space), all powers of two
    int lgBW, lgBH;
    int a_hstep;                // source rows per box row (conv stride for fprop, else 1)
    int out_H, out_W, out_C, os;
    bf16 *out;
    int acc;                    // 1: out += result (bf16 read-modify-write)
    int tma_store;              // 1: the staged bf16 tile goes to `out` with one TMA tile store per 64-channel half (dense lattice, no acc)
    int b_mn;                   // 0: B tile rows = output channel, cols = K (fprop)
                                // 1: B tile rows = K (co), cols = output channel (dgrad reads the
                                //    same [Co][taps*Ci] weight copy as an MN-major operand)
    int m_tiles, n_tiles, total_tiles;   // tile t = (class * n_tiles + n) * m_tiles + m
    float *stats;               // optional [copies][2][out_C] per-channel sum / sum of squares (atomics); may be null
    // optional (dgrad that produces the FINAL gradient of a fused BatchNorm's output tensor): also
    // accumulate that BatchNorm's backward sums  sum(g), sum(g*xhat),  g = dY * [y > 0]
    float *bn_bsum;             // [copies][2][out_C]; null = off
    const bf16 *bn_x, *bn_y;    // the BatchNorm's input / output (same NHWC shape as `out`)
    const float *bn_mean, *bn_invstd;
    int bn_relu;
    // optional (fprop): apply the consumer BatchNorm (+residual, ReLU) in this kernel.  After a CTA has
    // stored its tiles (bf16 conv output `out`, which backward needs anyway) and added their per-channel
    // sums to `stats`, it joins a grid-wide barrier (the grid is persistent: at most one CTA per SM, all
    // co-resident) and then normalises its own tiles -- re-read from L2 -- into the BatchNorm output `fb_y`.
    // One launch replaces conv + BN-statistics + BN-apply.
    int fuse_bn;
    bf16 *fb_y; const bf16 *fb_res;
    const float *fb_gamma, *fb_beta;
    float *fb_save_mean, *fb_save_invstd, *fb_running_mean, *fb_running_var;
    unsigned int *fb_sync;      // [0] arrival counter, [1] generation (self-resetting sense barrier)
    int *fb_err;
    float fb_count, fb_eps, fb_momentum; int fb_relu;
    double fb_inv_count;
    // stem mode (k_conv<64, true>): the A tiles are not TMA boxes of an NHWC tensor but im2col rows gathered in shared
    // memory by four extra warps straight from the fp32 NCHW image (K order (c, kh, kw padded to 8), 3*7*8 -> 192):
    // the 7x7 / stride-2 / 3-channel stem never materialises an im2col buffer in HBM.
    const float *stem_x;        // [N][3][32][32] fp32; null = not the stem
    long long *timeline;        // optional (tests/tools): CTA 0 records clock64() at 13 points of its first tile; null in the product
};

struct UmmaConvPlan {
    CUtensorMap a_map, b_map, out_map;
    ConvParams p;
    int bn_tile;                // 64 or 128 output channels per tile
    bool stem;                  // launch the gather variant (p.stem_x must be set before every launch)
    bool shallow;               // 3-stage ring / 2 CTAs per SM variant (experiment: B2_CONV_SHALLOW=1)
    int grid;                   // persistent CTAs (<= SM count)
    double flops;               // executed FLOPs (dead taps skipped)
    bool valid;
};

struct WgradParams {
    int ntaps_alive;
    int tap_id[kMaxTaps];       // index of each live tap in the (kh,kw) grid -> K offset tap_id*Ci
    int dh[kMaxTaps], dw[kMaxTaps];
    int c_chunks, Ci, Co, K;    // K = taps*Ci (row pitch of dW)
    int BW, BH, BN, h_blocks, a_hstep;
    int pix_tiles, tiles_per_split, splits;
    int m_boxes;                // Co tile = 64*m_boxes (1 or 2)
    int cpt;                    // 64-wide K chunks per CTA: 3 when m_boxes == 1, else 2 (a stage holds 4 boxes)
    float *out;                 // dW (atomic / single-split mode) or compact partials [splits][Co][alive taps*Ci]
    int out_pitch;              // row pitch of `out` in floats
    int compact;                // 1: columns are (alive tap, ci); 0: columns are (tap id, ci) as in dW
    int atomic;                 // 1: red.global.add straight into dW (dW zeroed once per step by the caller)
    const float *stem_x;        // stem mode (k_wgrad<true>): gather the im2col operand from the fp32 NCHW image; null otherwise
};

struct UmmaWgradPlan {
    CUtensorMap x_map, dy_map;
    WgradParams p;
    dim3 grid;
    float *dw;
    bool stem;                  // gather variant (p.stem_x must be set before every launch)
    bool need_zero;             // dead taps exist: dW must be zeroed first
    double flops;
    bool valid;
};

bool umma_conv_eligible(const ConvGeom &g);
int umma_init();   // resolves the driver entry point for tensor-map encoding
int umma_num_sms();
int umma_plan_fprop(const ConvGeom &g, const bf16 *x, const bf16 *w, bf16 *y, UmmaConvPlan *plan);
// the 7x7 / stride 2 / pad 3 / 3-channel stem on 32x32 images as a 1x1 conv over virtual im2col rows of kpad = 192 columns
bool umma_stem_supported(const ConvGeom &stem, const ConvGeom &virt, int kpad);
int umma_plan_dgrad(const ConvGeom &g, const bf16 *dy, const bf16 *w, bf16 *dx, bool acc, UmmaConvPlan *plan);
// checks (once per plan) that the persistent grid is co-resident; required before enabling fuse_bn
int umma_conv_check_coresident(const UmmaConvPlan &plan);
int umma_conv_launch(const UmmaConvPlan &plan, cudaStream_t s);
// deterministic: split-K partials + a reduce kernel (fixed summation order); otherwise the splits add
// into dW with red.global.add (dW must have been zeroed: umma_wgrad_needs_zero()).
size_t umma_wgrad_workspace(const ConvGeom &g);
int umma_plan_wgrad(const ConvGeom &g, const bf16 *x, const bf16 *dy, float *dw, float *partial, size_t partial_floats,
                    bool deterministic, UmmaWgradPlan *plan);
int umma_wgrad_launch(const UmmaWgradPlan &plan, cudaStream_t s);
// zero the dW columns of taps that never touch the input (call once; they are never written)
int umma_wgrad_zero_dead(const UmmaWgradPlan &plan, cudaStream_t s);

}  // namespace b2
