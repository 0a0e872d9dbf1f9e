// salun_gemm.cuh -- host-side interface of the tcgen05 GEMM kernels (salun_gemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "salun_act.cuh"
#include "salun_common.cuh"

namespace salun {

// Optional epilogue fusion for the dgrad GEMMs: the tile just computed is dX, the gradient w.r.t. an activation
// act = relu(BN(y) [+ ...]).  The BatchNorm backward of that BN needs  sum_pixels dZ  and  sum_pixels dZ * xhat  with
// dZ = dX * (act > 0), xhat = (y - mean) * invstd : they are reduced here from the fp32 accumulators (per (tile, warp)
// column partials, same layout as the forward statistics) instead of in a separate pass over dX, act and y.
struct BnBwdFuse {
  const act_t *act;           // halo-padded activation (ReLU mask), nullptr = fusion off (bf16 build only)
  const act_t *y;             // raw conv output [M][N] feeding that BN
  const float *mean, *invstd; // saved statistics of that BN
  float *partials;            // [tiles*4][2][N]
};

// D[M][N] = sum_k A[m][k] * B[n][k]   (bf16 x bf16 -> fp32 accumulate in TMEM)
// A comes either from a plain 2-D matrix or, for stride-1 convolutions, directly from the
// halo-padded NHWC activation through a 4-D tensor map (implicit GEMM: k-block = (tap, 64 channels)).
struct ConvGemmArgs {
  int mode_a;        // 0: A is [M][K] row-major; 1: A is padded NHWC [batch][H+2][W+2][C], taps walk it
  int num_k_blocks;  // K / 64 (K in activation elements; launch_conv_gemm rescales for the split build)
  int k_wrap;        // set by launch_conv_gemm.  > 0: the A operand restarts at k-block 0 when the k loop reaches this
                     // block (split build: pass 2 multiplies the same activation tile with the lo half of the weights)
  int cin_blocks;    // mode 1: 64-channel blocks per tap
  int kw;            // mode 1: taps per kernel row (3 for 3x3, 1 for 1x1)
  int tap_y0, tap_x0;  // mode 1: padded coordinate of tap (0,0) for output pixel (0,0): 0 for 3x3/pad1, 1 for 1x1/pad0
  int H, W;          // mode 1: output spatial size (== input spatial size, stride 1)
  int M, N;          // valid rows / columns of D
  act_t *out_bf16;          // activation-typed output [M][ld_out] or nullptr (bf16, or (hi, lo) pairs in the split build)
  float *out_f32;           // [M][ld_out] or nullptr
  int ld_out;
  float *stat_sum, *stat_sq;  // per-(m tile, warp) column partial sums [gridDim.x * 4][N], or nullptr
  const act_t *addend;          // optional [M][ld_out]: D += addend before the store (residual-gradient merge)
  long long *dbg;               // optional per-CTA role timing [gridDim][8] (bring-up / profiling only)
  BnBwdFuse f1, f2;             // up to two consumer BatchNorms of the produced gradient (bn2 + projection-shortcut BN)
  int fH, fW;                   // image size of the output pixels (padded-offset arithmetic of f1/f2.act and of out_pad)
  // ---- U-Net path (persistent kernel only) ----
  const float *bias;            // optional [N] fp32: D += bias[col]            (conv bias, DDPM/models/diffusion.py:104-119)
  const float *rowbias;         // optional fp32 [rows >> rb_shift][rb_ld]: D += rowbias[(row >> rb_shift) * rb_ld + col]
  int rb_shift, rb_ld;          //   (the per-sample temb/cemb projection added after conv1, diffusion.py:131-132)
  int out_pad;                  // 1: rows of out_bf16 / addend are the pixels of a halo-padded NHWC tensor
                                //    [n][fH+2][fW+2][ld_out] (row m -> interior pixel), 0: flat [M][ld_out]
  int batch_rows_a, batch_rows_b;  // batched GEMM (one B matrix per group of batch_rows_a rows of A): the B tile of
                                   // output tile (m, n) starts at row (m*128 / batch_rows_a) * batch_rows_b + n*BN.  0 = off
  float *splitk_ws;                // optional scratch: when the launch has too few output tiles to fill the GPU and K is
  long long splitk_ws_floats;      //   deep, launch_conv_gemm splits the k loop over several CTAs per tile (fp32 partial tiles
                                   //   in the scratch, then k_splitk_epilogue adds them in a fixed order and applies bias /
                                   //   rowbias / addend / the output conversion).  Needs stat_sum == nullptr and no f1/f2.
  int ksplit;                      // set by launch_conv_gemm (kernel side): k-slices per tile (0/1 = off) ...
  long long split_stride;          //   ... and the float distance between the partial slabs in out_f32
  int pair;                        // 1: launch_conv_gemm routes to the CTA-pair kernel (k_gemm2, 256 x bn tiles): tmB must
                                   // have been encoded with bn/2 box rows; bn in {128, 256}; batch_rows_a % 256 == 0
};

// dW[co][b*64 + j] += sum_p dY[p][co] * X_b[p][j]   (both operands MN-major: pixels are the K dimension)
// X_b is either a 64-column slab of a 2-D matrix Col[M][Kc] or the (tap, 64-channel) slab of the padded
// NHWC activation.  Split-K over pixel ranges, fp32 red.global.add into dW.
struct WgradArgs {
  int mode_a;         // 0: dY is [M][Cout] row-major; 1: dY is halo-padded NHWC (interior is read)
  int mode_b;         // 0: 2-D Col[M][Kc]; 1: 4-D padded NHWC taps
  int kb_total;       // M / 64 (k-blocks of 64 pixels)
  int kb_per_split;
  int cin_blocks, kw, tap_y0, tap_x0, H, W;
  int n_blocks;       // 64-wide column blocks per CTA (1..4)
  int total_blocks;   // column blocks overall (= Kc / 64)
  int Cout;           // valid rows of dW
  int ldw;            // row stride of dW in floats
  int kvalid;         // valid columns of dW (<= total_blocks * 64)
  float *dw;
  long long split_stride;  // 0: every split red.add's into dw (caller zeroes dw).  > 0: split s STORES its partial
                           // tile into dw + s*split_stride (a workspace slab of [Cout][ldw]); a deterministic
                           // reduction over the slabs follows (launch_wgrad_reduce) -- no atomics, no zero-fill
  int swap_lbo_sbo;   // debug knob for descriptor bring-up (0 in production)
  long long *dbg;     // optional per-CTA role timing [gridDim][8]
};

// persistent stride-1 3x3 convolution with smem-resident weights (k_conv_rw): layer1 / layer2 shapes
struct ConvRwArgs {
  int H, W;             // image size (stride 1: in == out)
  int cin_blocks;       // Cin / 64
  int num_tiles;        // ceil(M / 128)
  int M, N;             // valid rows / columns (N = Cout)
  act_t *out_bf16;
  int ld_out;
  float *stat_sum, *stat_sq;
  const act_t *addend;
  long long *dbg;
  BnBwdFuse f1, f2;
};

struct TmapBox4 {
  int c, w, h, n;
};

int make_tmap_2d_bf16(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                      uint32_t box_cols);
// Typed forms used by the runtimes (sizes in ACTIVATION / WEIGHT elements; the split build widens them):
//   act : an activation matrix [rows][cols]              -> bf16 [rows][cols * kActK], 64-wide boxes
//   wop : a prepared weight operand [rows][K] (K padded)  -> bf16 [rows][K * kWopK]
static inline int make_tmap_2d_act(CUtensorMap *m, const act_t *base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  return make_tmap_2d_bf16(m, base, rows, cols * kActK, box_rows, 64);
}
static inline int make_tmap_2d_wop(CUtensorMap *m, const wop_t *base, uint64_t rows, uint64_t K, uint32_t box_rows) {
  return make_tmap_2d_bf16(m, base, rows, K * kWopK, box_rows, 64);
}
// same with a row stride of `ld` elements (ld >= cols, ld * 2 bytes a multiple of 16); columns >= cols read as zero
int make_tmap_2d_bf16_ld(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                         uint32_t box_cols);
int make_tmap_4d_bf16(CUtensorMap *m, const void *base, uint64_t C, uint64_t Wp, uint64_t Hp, uint64_t N,
                      TmapBox4 box);
static inline int make_tmap_4d_act(CUtensorMap *m, const act_t *base, uint64_t C, uint64_t Wp, uint64_t Hp, uint64_t N,
                                   TmapBox4 box) {
  return make_tmap_4d_bf16(m, base, C * kActK, Wp, Hp, N, box);
}


// box decomposition of `pixels` consecutive output pixels of an H x W image batch (power-of-two sizes)
int conv_box(int H, int W, int pixels, TmapBox4 *box);

int launch_conv_gemm(const CUtensorMap &tmA, const CUtensorMap &tmB, const ConvGemmArgs &a, int bn, cudaStream_t st);
// fused attention forward for head widths <= 64 (salun_attn.cu); operands are the packed per-head buffers of salun_sd_attention
bool flash_attn_supported(int d);
int launch_flash_attn(const act_t *Qh, const wop_t *Kh, const wop_t *Vt, act_t *out, int n, int Tq, int Tqp, int Tk, int Tkp, int heads,
                      int d, cudaStream_t st);
int launch_wgrad(const CUtensorMap &tmA, const CUtensorMap &tmB, const WgradArgs &a, int co_tiles, int col_groups,
                 int splits, cudaStream_t st);
// CTA-pair (cta_group::2) GEMM, 256 x bn tiles; tmB must be encoded with bn/2 box rows
int launch_gemm2(const CUtensorMap &tmA, const CUtensorMap &tmB, const ConvGemmArgs &a, int bn, cudaStream_t st);
int wgrad_pick_blocks(int total_blocks);
// Tile decomposition of one weight-gradient GEMM dW[cout][kcp] (sizes in weight elements, kcp % 64 == 0); the split
// build computes the 2 x 2 partial products of every (co, k) pair: twice the rows and columns, and a 4x workspace.
struct WgradGeom {
  int co_tiles, groups, n_blocks, total_blocks;
};
WgradGeom wgrad_geometry(int cout, int kcp);
static inline size_t wgrad_ws_elems(int cout, int kc) { return (size_t)cout * kc * kActK * kActK; }
// dst[off + i] = sum_s ws[s*count + i], one table entry per convolution, ONE launch for the whole network
struct WgReduceEntry {
  const float *ws;
  long long dst_off;
  long long count;     // cout * kc weight elements
  int splits;
  int kc;              // row length of the weight (split build: the workspace row is 2 * kc wide)
};
void launch_wgrad_reduce(const WgReduceEntry *table_dev, int n_entries, float *grads, cudaStream_t st);
bool conv_rw_supported(int W, int cin, int cout);
// tmA: 4-D map of the padded activation with box {64, W, 128/W + 2, 1}; tmB: 2-D weight map with box {64, 64}
int launch_conv_rw(const CUtensorMap &tmA, const CUtensorMap &tmB, const ConvRwArgs &a, int num_sms, cudaStream_t st);

}  // namespace salun
