// salun_elem.cuh -- launchers of the HBM-bound elementwise / reduction kernels of the ResNet path
// (BatchNorm forward/backward around the tcgen05 GEMMs, ReLU, residual add, stride-2 im2col/col2im,
// stem input normalisation, average pool + FC + cross-entropy head, weight re-layout).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "salun_act.cuh"

namespace salun {

constexpr int kStatSlices = 8;      // row slices of the forward BN statistics reduction
constexpr int kBwdPartialRows = 296;  // CTAs (= partial rows) of the backward BN reduction

struct BnFwd {            // one BatchNorm applied to a raw conv output y[M][C]
  const act_t *y;
  const double *slices;   // [kStatSlices][2][C] (sum, sumsq), train mode
  const float *gamma, *beta;
  float *running_mean, *running_var;  // updated in train mode (momentum, unbiased var)
  float *saved_mean, *saved_invstd;   // written for the backward pass
  const double *count_dev;            // optional: the element count of the statistics (sync-BN: the GLOBAL pixel count);
                                      // nullptr = the M of this launch
};

// partial column sums written by the GEMM epilogue -> kStatSlices double slices
void launch_bn_stats_reduce(const float *stat_sum, const float *stat_sq, int rows, int C, double *slices,
                            cudaStream_t st);

// out = relu?( bn_a(y_a) [+ bn_b(y_b)] [+ resid] ), written into the halo-padded NHWC activation
// relu_mask_out (optional): 1 bit per output element, [M][C/8] bytes, bit i of byte (m, c/8) = out[m][c + i] > 0 --
// the backward kernels read it instead of the 16x larger activation
void launch_bn_apply(const BnFwd &a, const BnFwd *b, const act_t *resid_padded, act_t *out_padded,
                     uint8_t *relu_mask_out, int n_img, int H, int W, int C, int relu, int train, float eps,
                     float momentum, cudaStream_t st);

// sums over pixels of dZ and dZ*xhat, dZ = dout * (out > 0)  [relu_mask == nullptr: no ReLU in front]
void launch_bn_bwd_reduce(const act_t *dout, const uint8_t *relu_mask, const act_t *y,
                          const float *saved_mean, const float *saved_invstd, float *partials, int n_img, int H, int W,
                          int C, cudaStream_t st);
// partials -> dgamma, dbeta (into the grad arena) and the per-channel coefficients k1, m1, m2 of the apply pass
// rows > 0: number of partial rows (fused dgrad-epilogue partials); rows == 0: the rows launch_bn_bwd_reduce wrote
// xchg_out (optional, sync-BN): the two sums in double, [2][C], for the cross-rank exchange that then owns coef[C..3C)
void launch_bn_bwd_finalize(const float *partials, int rows, int C, const float *gamma, const float *saved_invstd,
                            float count, int train, float *dgamma, float *dbeta, float *coef, cudaStream_t st,
                            double *xchg_out = nullptr);

// ---- sync-BN: statistics of a batch sharded over ranks, exchanged through NVLink peer memory ----
// Every rank owns a symmetric slot [2][C] + 2 doubles per BatchNorm and direction.  One CTA: publish the local sums,
// cross-rank barrier on per-peer epoch flags (release / acquire at system scope), sum the peers' slots in RANK order
// (identical result everywhere).
//   mode 0 (forward) : local = the kStatSlices slices of this rank (sum x, sum x^2) and its pixel count;
//                      out   = slices[0] <- global sums, slices[1..] <- 0, count_out <- global count
//   mode 1 (backward): local = xchg slot already written by launch_bn_bwd_finalize; out = coef[C..3C) <- global sums / count
struct SyncBnPeers {
  double *slot[8];                 // this BatchNorm's slot on every rank (slot[rank] is the local one)
  unsigned long long *flags[8];    // flags[r][s]: epoch last published by rank s, in rank r's memory
  int rank, world;
};
void launch_syncbn_exchange(const SyncBnPeers &p, unsigned long long epoch, int mode, int C, double *slices,
                            double local_count, double *count_out, float *coef, cudaStream_t st);
// dY = k1 * (dZ - m1 - xhat * m2); written padded (for the 4-D TMA consumers) or flat [M][C]; optionally dZ too
void launch_bn_bwd_apply(const act_t *dout, const uint8_t *relu_mask, const act_t *y,
                         const float *saved_mean, const float *saved_invstd, const float *coef, act_t *dy,
                         int dy_padded, act_t *dz_flat, int n_img, int H, int W, int C, cudaStream_t st);

// stem: x fp32 NCHW [n][3][H][W] -> ((x-mean)/std) -> 3x3/pad-1 patches, col[M][64] bf16 (27 valid, tap-major)
void launch_stem_im2col(const float *x, act_t *col, int n_img, int H, int W, const float *mean3,
                        const float *inv_std3, cudaStream_t st);
// stride-2 patches of the padded NHWC activation: col[Mout][ks*ks*C]
void launch_im2col_s2(const act_t *in_padded, act_t *col, int n_img, int Hin, int Win, int C, int ks,
                      cudaStream_t st);
// dX[Min][C] = col2im(dcol3 [Mout][9C]) (+ dcol1 [Mout][C] at even pixels) (+ addend)
void launch_col2im_s2(const act_t *dcol3, const act_t *dcol1, act_t *dx, int n_img, int Hin,
                      int Win, int C, cudaStream_t st);

// weights: fp32 native [Cout][taps][Cin] -> bf16 operands; one launch for the whole network
struct WPrepEntry {
  long long w_off;         // offset of the fp32 master weight in the parameter arena
  wop_t *w_fwd;    // [cout][kcp * kWopK]  (zero padded beyond kc)
  wop_t *w_dgrad;  // dgrad operand or nullptr
  int cout, cin, kc, kcp;
  int dgrad_mode;          // 0: none, 1: stride-1 flipped [ci][tap'][co], 2: transposed [kc][co]
  int ldo;                 // row stride (in co) of the dgrad operand; 0 = cout (conv_out of the U-Net pads 3 -> 64)
};
void launch_prep_w_all(const WPrepEntry *table_dev, int n_convs, const float *params, int need_dgrad, cudaStream_t st);

// ---- generic (any spatial size / stride / padding) kernels on FLAT NHWC activations [n*H*W][C]: Bottleneck nets ----
// col[Mout][ks*ks*C] (tap-major) from a flat activation; out-of-image taps are zero
void launch_im2col_flat(const act_t *in_flat, act_t *col, int n_img, int Hin, int Win, int C, int ks,
                        int stride, int pad, int Hout, int Wout, cudaStream_t st);
// stem of any geometry: x fp32 NCHW -> normalise -> col[Mout][kcp] (kc = ks*ks*3 valid, tap-major then channel)
void launch_stem_im2col_generic(const float *x, act_t *col, int n_img, int Hin, int Win, int ks, int stride,
                                int pad, int Hout, int Wout, int kcp, const float *mean3, const float *inv_std3,
                                cudaStream_t st);
// dx[Min][C] = col2im(dcol[Mout][ks*ks*C]) (+ addend[Min][C]); gather form, deterministic
void launch_col2im_flat(const act_t *dcol, const act_t *addend, act_t *dx, int n_img, int Hin,
                        int Win, int C, int ks, int stride, int pad, int Hout, int Wout, cudaStream_t st);
// out = relu?(bn_a(y_a) [+ bn_b(y_b)] [+ resid_flat]) written FLAT
void launch_bn_apply_flat(const BnFwd &a, const BnFwd *b, const act_t *resid_flat, act_t *out_flat,
                          uint8_t *relu_mask_out, int M, int C, int relu, int train, float eps, float momentum,
                          cudaStream_t st);
// 3x3 / stride 2 / pad 1 max pooling (nn.MaxPool2d(3, 2, 1), ResNet.py:229) and its backward (argmax offsets kept)
void launch_maxpool_fwd(const act_t *in_flat, act_t *out_flat, uint8_t *argmax, int n_img, int Hin,
                        int Win, int C, cudaStream_t st);
void launch_maxpool_bwd(const act_t *dout_flat, const uint8_t *argmax, act_t *dx_flat, int n_img, int Hin,
                        int Win, int C, cudaStream_t st);
void launch_avgpool_flat(const act_t *act_flat, float *pooled, int n_img, int pix, int C, cudaStream_t st);

// head
void launch_avgpool(const act_t *act_padded, float *pooled, int n_img, int H, int W, int C, cudaStream_t st);
void launch_fc_ce(const float *pooled, const float *w, const float *b, const int64_t *labels, float *logits,
                  float *dlogits, float *loss_per_sample, int n_img, int C, int K, float sign, cudaStream_t st);
void launch_loss_sum(const float *loss_per_sample, int n_img, float sign, float *loss_out, cudaStream_t st);
void launch_ce_rows(const float *logits, const int64_t *labels, float *dlogits, float *loss_per_sample, int n_img, int K,
                    float sign, cudaStream_t st);
void launch_pool_grad_bcast(const float *dpooled, act_t *dact_flat, int n_img, int C, int pix, cudaStream_t st);
void launch_fc_bwd(const float *pooled, const float *dlogits, const float *w, float *dw, float *db,
                   act_t *dact_flat, int n_img, int C, int K, int pix, cudaStream_t st);

}  // namespace salun
