// salun_unet_elem.cuh -- launchers of the HBM-bound kernels of the DDPM U-Net path (salun_unet.cu): GroupNorm + swish
// (+ dropout) forward / backward, per-sample column reductions (bias and temb-projection gradients), skip concat /
// split, nearest upsample, strided-conv patches, softmax, batched transposes, the timestep / class embedding MLPs.
// Replaces the ATen elementwise chain of DDPM/models/diffusion.py:38-192,357-413 and its autograd backward.
//
// Activation layouts (bf16, square power-of-two images of side H):
//   padded ("P"): [n][H+2][H+2][C], zero halo, written once -- what the 4-D-TMA convolution kernels read
//   flat   ("F"): [n*H*H][C]
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "salun_act.cuh"

namespace salun {

constexpr int kGnGroups = 32;  // Normalize() = GroupNorm(32, C, eps 1e-6)   diffusion.py:43-46

// slices of one image's H*H pixels handled by separate CTAs in the per-sample kernels
int unet_slices(int H);

// ---- GroupNorm forward: partial[n][S][2][C] (sum x, sum x^2 per slice) -> stats[n][32][2] (mean, rstd; kept for the
// backward pass) -> out = dropout(act(gamma * (x - mean) * rstd + beta)); act = swish or identity; out padded or flat ----
// ep_sum / ep_sq (optional): per-(32-row group) column partials [n*H*H/32][C] of x that the producing GEMM's epilogue
// wrote (ConvGemmArgs::stat_sum / stat_sq) -- the statistics pass over x is then skipped.
void launch_gn_forward(const act_t *x_pad, const float *ep_sum, const float *ep_sq, float *partial, float *stats,
                       const float *gamma, const float *beta, act_t *out, int out_flat, int swish, float drop_p,
                       uint32_t drop_seed, int n, int H, int C, float eps, cudaStream_t st);
// epilogue partials of a skip concatenation: row r of the result = [row r of a | row r of b]
void launch_concat_stats(const float *a_sum, const float *a_sq, int Ca, const float *b_sum, const float *b_sq, int Cb,
                         float *o_sum, float *o_sq, long long rows, cudaStream_t st);
// ---- GroupNorm backward.  dout: gradient w.r.t. the GN(+swish+dropout) output, FLAT [n*H*H][C] ----
//   partial[n][S][2][C] = per-sample, per-slice sums over pixels of dyh and dyh * xhat (dyh = dout * act'(yh) * dropmask)
//   persample[n][2][C]  = the same summed over the slices; the caller sums it over n into dbeta / dgamma (SumEntry)
//   dx = rstd * (gamma*dyh - coefA - xhat*coefB), written padded; accumulate != 0: added to what dx already holds
//   dout_flat is OVERWRITTEN with dyh (it has no other consumer)
void launch_gn_backward(act_t *dout_flat, const act_t *x_pad, const float *stats, const float *gamma,
                        const float *beta, int swish, float drop_p, uint32_t drop_seed, float *partial, float *persample,
                        act_t *dx_pad, int accumulate, int n, int H, int C, cudaStream_t st);

// ---- cross-sample sums, all in one launch at the end of the backward pass ----
//   gbase[d0 + c] (and gbase[d0b + c]) = sum_rows src[row*ld + c];  K == 2: gbase[d1 + c] = sum_rows src[row*ld + C + c]
struct SumEntry {
  const float *src;
  long long ld;
  int rows, K, C;
  long long d0, d0b, d1;  // offsets into the gradient arena, -1 = unused
};
void launch_sum_rows_table(const SumEntry *table_dev, int n_entries, float *grad_base, cudaStream_t st);

// ---- bias gradients: per-sample, per-slice column sums of dY (padded or flat): partial[n][S][C]; returns n*S.
//   rowsum (optional) [n][rowsum_ld] at column rowsum_col0: per-sample sums (gradient of the temb/cemb projection output)
int launch_bias_partial(const act_t *dy, int dy_flat, float *partial, float *rowsum, int rowsum_ld,
                        int rowsum_col0, int n, int H, int C, cudaStream_t st);
int unet_slices_for(int H, int n);  // slices launched for a batch of n (<= unet_slices(H))

// ---- copies on padded tensors ----
void launch_concat(const act_t *a_pad, int Ca, const act_t *b_pad, int Cb, act_t *out_pad, int n,
                   int H, cudaStream_t st);
// da (=|+=) dcat[..., :Ca] ; db (=|+=) dcat[..., Ca:]
void launch_split(const act_t *dcat_pad, act_t *da_pad, int Ca, int acc_a, act_t *db_pad, int Cb,
                  int acc_b, int n, int H, cudaStream_t st);
// dst (=|+=) src over `count` bf16 elements (count % 8 == 0)
void launch_add_into(const act_t *src, act_t *dst, int accumulate, long long count, cudaStream_t st);
// nearest x2 (F.interpolate(scale_factor=2, mode="nearest"), diffusion.py:59): in side H -> out side 2H
void launch_upsample2(const act_t *in_pad, act_t *out_pad, int n, int H, int C, cudaStream_t st);
void launch_upsample2_bwd(const act_t *dout_pad, act_t *din_pad, int accumulate, int n, int H, int C,
                          cudaStream_t st);
// Downsample (diffusion.py:75-79): pad (0,1,0,1) then 3x3 / stride 2 / no padding.  col[n*Ho*Ho][9*C], tap-major
void launch_down_im2col(const act_t *in_pad, act_t *col, int n, int H, int C, cudaStream_t st);
void launch_down_col2im(const act_t *dcol, act_t *din_pad, int accumulate, int n, int H, int C,
                        cudaStream_t st);
// eps[n][3][H][H] fp32 = y[m][0..2] + bias   (y: fp32 [n*H*H][64] GEMM output of conv_out padded to 64 channels)
void launch_eps_out(const float *y, const float *bias3, float *eps_nchw, int n, int H, cudaStream_t st);
// dy_pad[n][H+2][H+2][64] bf16 (channels 0..2) = deps[n][3][H][H]; dbias_partial[n][3] = per-sample sums over pixels
void launch_eps_in(const float *deps_nchw, act_t *dy_pad, float *dbias_partial, int n, int H, cudaStream_t st);

// ---- attention (diffusion.py:167-192): rows of S fp32 [M][Te] -> P bf16; block-diagonal mask of block T inside Te ----
void launch_softmax(const float *S, act_t *P, int M, int Te, int T, float scale, cudaStream_t st);
// dS = scale * P * (dP - sum_j dP*P)
void launch_softmax_bwd(const float *dP, const act_t *P, act_t *dS, int M, int Te, float scale,
                        cudaStream_t st);
// out[g][c][r] = in[g][r][c]; in rows have stride ld_in, out rows stride R; G groups of R rows; R, Cc multiples of 32.
// The result is an activation matrix (the A operand of the next GEMM) ...
void launch_transpose(const act_t *in, int ld_in, act_t *out, int R, int Cc, int G, cudaStream_t st);
// ... or a prepared WEIGHT-side operand (the B operand: rows of logical length R in the wop layout of salun_act.cuh)
void launch_transpose_wop(const act_t *in, int ld_in, wop_t *out, int R, int Cc, int G, cudaStream_t st);
// B operand of a GEMM whose B matrix is itself an activation (K of S = Q K^T, V of dP = dO V^T): wop rows of length K
// from act rows.  bf16 build: the layouts coincide, nothing is launched and `in` is returned.
const wop_t *launch_pack_wop(const act_t *in, wop_t *out, long long rows, int K, cudaStream_t st);

// ---- embedding MLPs (fp32, CUDA cores; tiny) ----
// sincos[n][ch] (get_timestep_embedding, diffusion.py:17-35) and ce[n][ch] = drop[n] ? null_emb : class_emb[c[n]]
void launch_emb_inputs(const float *t, const int64_t *c, const uint8_t *drop, const float *class_emb,
                       const float *null_emb, float *sincos, float *ce, int n, int ch, int n_classes, cudaStream_t st);
// C[i][j] (=|+=) sum_k A[i*sai + k*sak] * B[k*sbk + j*sbj] (+ bias[j])
void launch_sgemm(const float *A, long long sai, long long sak, const float *B, long long sbk, long long sbj, float *C,
                  int ldc, int M, int N, int K, const float *bias, int accumulate, cudaStream_t st);
void launch_swish_f32(const float *in, float *out, long long count, cudaStream_t st);
// out = dy * swish'(pre)
void launch_dswish_f32(const float *dy, const float *pre, float *out, long long count, cudaStream_t st);
// out[j] = sum_i a[i*ld + j]
void launch_colsum_f32(const float *a, int ld, int rows, int cols, float *out, cudaStream_t st);
// gradient of the embedding lookup: d_class_emb[k][:] = sum_{i: c[i]==k, !drop[i]} dce[i][:] ; d_null = sum_{drop[i]} dce[i]
void launch_emb_scatter(const float *dce, const int64_t *c, const uint8_t *drop, float *d_class_emb, float *d_null,
                        int n, int ch, int n_classes, cudaStream_t st);


// ---- the temb/cemb projections of all ResnetBlocks as one tensor-core GEMM: operand staging ----
// wcat[r][k] = act(params[row_w[r] + k]), bcat[r] = params[row_b[r]]  for r < rows, k < K (K % 4 == 0)
void launch_gather_proj(const float *params, const long long *row_w, const long long *row_b, act_t *wcat,
                        float *bcat, int rows, int K, cudaStream_t st);
// dst[row_w[r] + k] = src[r][k]
void launch_scatter_rows(const float *src, const long long *row_w, float *dst, int rows, int K, cudaStream_t st);
// out[i*ld_out + j] = act(in[i*ld_in + j]); cols % 4 == 0, 16-byte aligned rows
void launch_f32_to_bf16(const float *in, int ld_in, act_t *out, int ld_out, int rows, int cols, cudaStream_t st);


// ---- q-sample and eps-prediction losses (functions/losses.py:21-37, runners/diffusion.py:533-572) ----
void launch_q_sample(const float *x01, const float *e, const int64_t *t, const float *sqrt_abar, const float *sqrt_1m_abar,
                     int num_t, int rescale, int n, int chw, float *xt, cudaStream_t st);
// ss[n] = sum_chw (eps - target)^2 ; d_eps = 2 w[n] (eps - target) ; loss = sum_n w[n] ss[n]
void launch_eps_loss_grad(const float *eps, const float *target, const float *w, int n, int chw, float *d_eps, float *ss,
                          float *loss, cudaStream_t st);

}  // namespace salun
