/*
 * salun.h -- C ABI of libsalun.so, the B200-native (sm_100a) SalUn engine.
 *
 * The reference (OPTML-Group/Unlearn-Saliency) has no FFI: its hot path is Python calling
 * ATen.  This header is the boundary a maintainer binds instead (ctypes stub in
 * INTEGRATION.md): every entry point names the reference lines it replaces.  Citations
 * are relative to the reference checkout.
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch types.  `stream` is a cudaStream_t passed as
 *     void* (0 = legacy default stream).  Every call only ENQUEUES work on `stream` unless
 *     stated otherwise; buffers are owned by the caller (PyTorch) and must stay alive until
 *     the stream reaches the enqueued work.
 *   - pointers are DEVICE pointers unless the parameter name ends in `_host`.
 *   - return 0 on success, negative salun_status otherwise; salun_last_error() returns a
 *     thread-local message for the last failure on the calling thread.
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns
 *     SALUN_ERR_CUDA.
 *   - mask_bits: packed 1-bit mask, element i is bit (i & 31) of word (i >> 5); the buffer
 *     holds (n + 31) / 32 words and padding bits are zero.  The on-disk format stays the
 *     reference's dict{name: int64 0/1 tensor} (Classification/generate_mask.py:76-82).
 */
#ifndef SALUN_H_
#define SALUN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum salun_status {
  SALUN_OK = 0,
  SALUN_ERR_INVALID = -1, /* bad argument (null pointer, negative size, misaligned buffer) */
  SALUN_ERR_CUDA = -2,    /* a CUDA runtime / driver call failed, see salun_last_error() */
  SALUN_ERR_STATE = -3,   /* object used in the wrong state (e.g. backward before forward) */
  SALUN_ERR_UNSUPPORTED = -4
} salun_status;

typedef struct salun_ctx salun_ctx;

/* ABI version: major*1000 + minor */
int salun_version(void);
/* number of kernels this library has launched since it was loaded (bench.py: gpu_launches) */
long long salun_launch_count(void);
/* Per-launch CUDA-event timing of the two tensor-core kernels (category 0 = conv forward/dgrad GEMM,
 * 1 = wgrad GEMM).  begin() arms it; end() synchronises the device and returns, per category, summed
 * milliseconds, launch count and algorithmic FLOPs.  Arrays of 2; any may be NULL. */
int salun_profile_begin(void);
/* bring-up aid: attach a device buffer of >= 8 * 1024 int64; every tensor-core launch then records per-CTA role
 * timings (cycles) into it: [cta][0] producer wait-empty, [1] producer total, [2] MMA wait-full, [3] MMA wait-tmem-empty,
 * [4] MMA total, [5] epilogue wait-tmem-full, [6] epilogue total.  NULL detaches. */
int salun_debug_role_timing(long long *buf_dev);
int salun_profile_end(double *ms_by_cat, int64_t *launches_by_cat, double *flops_by_cat);
const char *salun_last_error(void);

/* One context per (thread, device).  Owns the small device workspaces (radix-select
 * histograms, reduction partials) and a pinned host mailbox.  Not thread-safe: use one
 * context per thread/stream. */
int salun_ctx_create(int device, salun_ctx **out);
int salun_ctx_destroy(salun_ctx *ctx);

/* ---------------------------------------------------------------------------------------
 * (i) saliency mask generation tail
 * ------------------------------------------------------------------------------------- */

/* accum_flat[off_t + j] += scale * grads[t][j]   for every tensor t, j < numels[t]
 * replaces  gradients[name] += param.grad.data        Classification/generate_mask.py:41-44
 *                                                      DDPM/runners/diffusion.py:992-996 (no .cpu() round trip)
 *                                                      SD/train-scripts/generate_mask.py:66-69
 * grads_host / numels_host: HOST arrays of n_tensors device pointers / element counts, laid
 * out back to back in accum_flat in the given order (= model.named_parameters() order).
 * scale_dev: optional device scalar (e.g. the clip coefficient of runners/diffusion.py:985-990);
 * NULL means 1. */
int salun_saliency_accumulate(salun_ctx *ctx, const float *const *grads_host,
                              const int64_t *numels_host, int n_tensors, float *accum_flat,
                              const float *scale_dev, void *stream);

/* accum[i] += scale * grad[i] on one flat arena (the engine's own gradient layout). */
int salun_saliency_accumulate_flat(salun_ctx *ctx, const float *grad, float *accum, int64_t n,
                                   const float *scale_dev, void *stream);

/* a[i] = |a[i]|       replaces torch.abs_  generate_mask.py:46-48 */
int salun_abs_inplace(salun_ctx *ctx, float *a, int64_t n, void *stream);

typedef struct salun_topk_info {
  uint32_t thr_key;  /* key(|g|) of the k-th largest element; key = bits(|g|)+1, NaN -> 0 */
  float thr_value;   /* |g| of the k-th largest element */
  int64_t n_greater; /* elements strictly above the threshold */
  int64_t n_equal;   /* elements equal to the threshold (ties); k - n_greater of them are selected */
} salun_topk_info;

/* Global top-k saliency mask: element i gets 1 iff its rank in descending |accum| is < k.
 * replaces  all_elements = -cat(...); argsort; argsort; ranks < k   generate_mask.py:57-80
 *           DDPM/runners/diffusion.py:1006-1037, SD/train-scripts/generate_mask.py:78-106
 * with a 3-pass radix select (no sort).  `accum` may be signed: |.| is applied on the fly.
 * Ties at the threshold are resolved in flat order (stable-argsort semantics); NaN ranks last.
 * k is computed by the caller as int(n * ratio) in double arithmetic (generate_mask.py:60).
 * mask_i64 / mask_bits: either may be NULL.  info_host: optional HOST struct; when non-NULL
 * the call synchronises `stream` before returning. */
int salun_topk_mask(salun_ctx *ctx, const float *accum, int64_t n, int64_t k, int64_t *mask_i64,
                    uint32_t *mask_bits, salun_topk_info *info_host, void *stream);

/* The reference's sweep over threshold_list = [0.1 .. 1.0] (generate_mask.py:50-82) as ONE call: the three histogram passes,
 * the tie count and the mask write read `accum` once for all n_ratios <= 16 ratios.  ks_host[r] = int(n * ratio_r);
 * mask_i64_host / mask_bits_host: HOST arrays of n_ratios device pointers (either array, or single entries, may be NULL as
 * long as every ratio has an output); infos_host: optional HOST array of n_ratios structs (synchronises `stream`).
 * Output r is bit-identical to salun_topk_mask(accum, n, ks_host[r], ...). */
int salun_topk_mask_multi(salun_ctx *ctx, const float *accum, int64_t n, const int64_t *ks_host, int n_ratios,
                          int64_t *const *mask_i64_host, uint32_t *const *mask_bits_host, salun_topk_info *infos_host,
                          void *stream);

/* int64 {0,1} <-> packed bits (loading / saving the reference's on-disk mask) */
int salun_pack_mask(salun_ctx *ctx, const int64_t *mask_i64, int64_t n, uint32_t *mask_bits,
                    void *stream);
int salun_unpack_mask(salun_ctx *ctx, const uint32_t *mask_bits, int64_t n, int64_t *mask_i64,
                      void *stream);

/* ---------------------------------------------------------------------------------------
 * (ii) masked unlearning step tail
 * ------------------------------------------------------------------------------------- */

/* g[i] *= m[i]     replaces _apply_mask_to_grads  Classification/unlearn/RL.py:11-14
 *                  DDPM/runners/diffusion.py:589-592 (incl. its 309 MB per-step H2D), SD/train-scripts/train-esd.py:318-321 */
int salun_apply_mask(salun_ctx *ctx, float *g, const uint32_t *mask_bits, int64_t n,
                     void *stream);

/* Fused  mask(.)grad -> SGD(momentum, weight decay) -> restore  on a flat arena:
 *   m=1: g' = g + wd*p ; v = momentum*v + g' ; p = p - lr*v        m=0: p untouched, v = 0
 * replaces  _apply_mask_to_grads + optimizer.step() + _restore_masked_params
 *           Classification/unlearn/RL.py:134-140 (same in GA.py:119-125, FT.py:137-142)
 *           with torch.optim.SGD built at Classification/unlearn/impl.py:68-73.
 * v must be zero-initialised before the first step (torch's "buf = clone(g')" first step is
 * then reproduced exactly).  mask_bits == NULL: plain SGD on every coordinate. */
int salun_masked_sgd_step(salun_ctx *ctx, float *p, const float *g, float *v,
                          const uint32_t *mask_bits, int64_t n, float lr, float momentum,
                          float wd, void *stream);

/* Data-parallel form of the step above, ONE kernel over NVLink peer memory instead of
 * all-reduce + scale + salun_masked_sgd_step:
 *   shard [lo, hi) = salun_dp_shard(n, rank, world) of the flat arena is owned by `rank`;
 *   g = (sum over ranks r = 0..world-1 of grad_peers[r][i]) / world   (peer loads, fixed order -> identical replicas)
 *   masked SGD on p[i], v_shard[i - lo] exactly as salun_masked_sgd_step
 *   p written to param_peers[r][i] for every r                          (peer stores)
 * param_peers_host / grad_peers_host: HOST arrays of `world` device pointers, entry r = rank r's arena mapped into
 * this process (e.g. torch symmetric memory buffer_ptrs); world <= 8.  v_shard: hi - lo floats, zero-initialised.
 * The caller orders the ranks: a cross-rank barrier on `stream` before (all gradients written) and after (all weights
 * delivered) the call.  Semantics are those of DDP-averaged gradients followed by RL.py:134-140. */
int salun_dp_shard(int64_t n, int rank, int world, int64_t *lo, int64_t *hi);
int salun_dp_masked_sgd_step(salun_ctx *ctx, float *const *param_peers_host, const float *const *grad_peers_host,
                             float *v_shard, const uint32_t *mask_bits, int64_t n, int rank, int world, float lr,
                             float momentum, float wd, void *stream);

/* Data-parallel form of clip + mask + Adam (DDPM/runners/diffusion.py:582-593 under DDP-averaged gradients), two
 * kernels over NVLink peer memory around ONE cross-rank barrier, instead of all-reduce + grad_sumsq + clip_coef +
 * masked_adam_step.  The clip needs the norm of the AVERAGED gradient before masking, so:
 *   salun_dp_grad_reduce_sumsq : g_shard[i - lo] = (sum_r grad_peers[r][i]) / world for the shard [lo, hi) =
 *                                salun_dp_shard(n, rank, world) (peer loads, fixed order); norm_slot[0] = sum g_shard^2
 *                                (double).  norm_slot must live in memory every peer can read (symmetric memory).
 *   -- cross-rank barrier on the stream (all norm slots written) --
 *   salun_dp_masked_adam_step  : coef = min(1, max_norm / (sqrt(sum_r *norm_peers[r]) + 1e-6)) (max_norm <= 0: no clip);
 *                                Adam exactly as salun_masked_adam_step on p[lo:hi) with g_shard and the shard-sized
 *                                moments m1_shard / m2_shard (zero-initialised); the new weights are stored into
 *                                param_peers[r] for every r (peer stores).  coef_norm_dev (optional, 2 floats): the clip
 *                                coefficient and the pre-clip global gradient norm.
 *   -- cross-rank barrier (all weights delivered) --
 * Optimizer state exists only for the owned shard (1/world of the arena per GPU). */
int salun_dp_grad_reduce_sumsq(salun_ctx *ctx, const float *const *grad_peers_host, float *g_shard, double *norm_slot,
                               int64_t n, int rank, int world, void *stream);
int salun_dp_masked_adam_step(salun_ctx *ctx, float *const *param_peers_host, const double *const *norm_peers_host,
                              const float *g_shard, float *m1_shard, float *m2_shard, const uint32_t *mask_bits, int64_t n,
                              int rank, int world, float lr, float beta1, float beta2, float eps, float wd, int64_t step,
                              float max_norm, float *coef_norm_dev, void *stream);

/* sumsq_dev[0] = sum_i g[i]^2 in double (deterministic two-stage reduction).
 * replaces the norm inside clip_grad_norm_   DDPM/runners/diffusion.py:582-587,985-990 */
int salun_grad_sumsq(salun_ctx *ctx, const float *g, int64_t n, double *sumsq_dev, void *stream);

/* g[i] += alpha * sign(p[i]);  l1_dev[0] = sum_i |p[i]| in double: the gradient and the value of the FT_l1 penalty
 * alpha * torch.linalg.norm(cat(params), ord=1).
 * replaces l1_regularization + its autograd   Classification/unlearn/FT.py:13-17,133-134 */
int salun_l1_penalty_grad(salun_ctx *ctx, const float *p, float *g, int64_t n, float alpha, double *l1_dev, void *stream);

/* coef_dev[0] = min(1, max_norm / (sqrt(sumsq_dev[0]) + 1e-6))   (torch clip_grad_norm_) */
int salun_clip_coef(salun_ctx *ctx, const double *sumsq_dev, float max_norm, float *coef_dev,
                    void *stream);

/* Fused  clip -> mask(.)grad -> Adam  on a flat arena (torch.optim.Adam arithmetic, amsgrad off):
 *   g = coef*g*m ; m1 += (g-m1)(1-b1) ; m2 = b2*m2 + (1-b2) g^2 ;
 *   p -= lr/(1-b1^t) * m1 / (sqrt(m2)/sqrt(1-b2^t) + eps)
 * replaces  clip_grad_norm_ + mask multiply + optimizer.step()
 *           DDPM/runners/diffusion.py:582-593 with Adam from DDPM/functions/__init__.py:9-18;
 *           SD/train-scripts/train-esd.py:318-323, random_label.py:132-139 (coef_dev = NULL: no clip).
 * step is 1-based.  mask_bits == NULL: no mask. */
int salun_masked_adam_step(salun_ctx *ctx, float *p, const float *g, float *m1, float *m2,
                           const uint32_t *mask_bits, int64_t n, float lr, float beta1,
                           float beta2, float eps, float wd, int64_t step,
                           const float *coef_dev, void *stream);

/* ---------------------------------------------------------------------------------------
 * tcgen05 tensor-core building blocks (bf16 operands, fp32 accumulation in TMEM, TMA-fed).
 * These replace the cuDNN/cuBLAS calls behind Classification/models/ResNet.py:58-74,108-124
 * (conv forward) and behind loss.backward() (Classification/unlearn/RL.py:132: dgrad, wgrad).
 * Activation layout: halo-padded NHWC bf16 [batch][H+2][W+2][C] with a zero halo -- the
 * convolution reads its taps straight out of it with shifted 4-D TMA boxes (no im2col buffer).
 * Weight layout: bf16 [Cout][kh*kw*Cin], tap-major then input channel.
 * ------------------------------------------------------------------------------------- */

/* D[M][N] = A[M][K] . B[N][K]^T ; A, B bf16 row-major ; out_f32 and/or out_bf16 [M][N].
 * Requires K % 64 == 0 and N % 64 == 0. */
int salun_gemm_bf16_tn(salun_ctx *ctx, const void *A, const void *B, float *out_f32, void *out_bf16,
                       int64_t M, int64_t N, int64_t K, void *stream);

/* Same contract through the CTA-pair kernel (tcgen05 cta_group::2, 256 x {128,256} tiles shared by two SMs);
 * N % 128 == 0. */
int salun_gemm2_bf16_tn(salun_ctx *ctx, const void *A, const void *B, float *out_f32, void *out_bf16,
                        int64_t M, int64_t N, int64_t K, void *stream);

/* Stride-1 convolution forward, ksize 3 (pad 1) or 1 (pad 0):
 *   y[batch*H*W][Cout] = conv(xpad, wk)        (y_bf16 and/or y_f32; either may be NULL, not both)
 *   stat_sum / stat_sq: optional fp32 [(batch*H*W/128)*4][Cout] per-tile column sums of y and y^2
 *   (the BatchNorm batch statistics, reduced by the caller).  Cin, Cout multiples of 64,
 *   batch*H*W a multiple of 128, H and W powers of two. */
int salun_conv_fwd_bf16(salun_ctx *ctx, const void *xpad, const void *wk, void *y_bf16, float *y_f32,
                        float *stat_sum, float *stat_sq, int batch, int H, int W, int Cin, int Cout,
                        int ksize, void *stream);

/* salun_conv_fwd_bf16 for 3x3 convolutions on 32x32 / 16x16 images with Cin in {64,128}, through the persistent
 * kernel that keeps its weight slice resident in shared memory (k_conv_rw). */
int salun_conv_rw_fwd_bf16(salun_ctx *ctx, const void *xpad, const void *wk, void *y_bf16, float *stat_sum,
                           float *stat_sq, int batch, int H, int W, int Cin, int Cout, void *stream);

/* Stride-1 convolution weight gradient: dw[Cout][ksize*ksize*Cin] (fp32, tap-major) +=
 *   sum over pixels of dy[p][Cout]^T . x_tap[p][Cin].  dw must be zeroed by the caller (split-K
 *   accumulates with fp32 red.global.add).  splits <= 0 picks a split count that fills the GPU.
 *   swap_lbo_sbo is a bring-up knob and must be 0. */
int salun_conv_wgrad_bf16(salun_ctx *ctx, const void *dy, const void *xpad, float *dw, int batch, int H,
                          int W, int Cin, int Cout, int ksize, int splits, int swap_lbo_sbo, void *stream);

/* ---------------------------------------------------------------------------------------
 * ResNet (BasicBlock, CIFAR stem) forward + backward engine.
 * replaces  output = model(image); loss = +/-criterion(output, target); loss.backward()
 *           Classification/generate_mask.py:35-39  (model.eval(): BN uses running statistics, loss_sign = -1)
 *           Classification/unlearn/RL.py:128-132, GA.py:113-117, FT.py:128-135  (model.train())
 * for the architectures of Classification/models/ResNet.py:180-322: resnet18 / resnet34 (BasicBlock :77-124, CIFAR
 * stem, power-of-two images: implicit-GEMM path on halo-padded activations) and resnet50 / 101 / 152 (Bottleneck
 * :127-177, CIFAR or ImageNet stem, any image size: BASELINE config 4).  bf16 tensor-core operands, fp32 accumulation,
 * fp32 master weights / gradients / BatchNorm statistics.
 *
 * Arena layout (caller-owned device buffers, fp32):
 *   params / grads : tensors back to back in model.named_parameters() order; every conv weight
 *                    is stored [Cout][kh][kw][Cin] (the reference's [Cout][Cin][kh][kw] permuted
 *                    (0,2,3,1)); BN weight/bias and fc weight/bias as in PyTorch.
 *   running_mean / running_var : BatchNorm buffers back to back in module order
 *                    (bn1, layer1.0.bn1, layer1.0.bn2, [layer*.0.downsample.1], ...).
 * ------------------------------------------------------------------------------------- */
typedef struct salun_resnet salun_resnet;
typedef struct salun_resnet_cfg {
  int depth;       /* 18 / 34 (BasicBlock, CIFAR stem) ; 50 / 101 / 152 (Bottleneck, either stem) */
  int num_classes; /* arg_parser.py --num_classes */
  int image_size;  /* BasicBlock nets: 32 or 64 ; Bottleneck nets: any (e.g. 224) */
  int max_batch;   /* largest batch a call will pass */
  float mean[3];   /* NormalizeByChannelMeanStd, ResNet.py:7-28, values replaced per dataset in utils.py:115-117 */
  float std[3];
  float bn_eps;      /* 1e-5 */
  float bn_momentum; /* 0.1 */
  int imagenet_stem; /* Bottleneck nets: 1 = 7x7/2 conv + 3x3/2 max pool (ResNet.py:224-230, imagenet=True), 0 = 3x3/1 */
} salun_resnet_cfg;

int64_t salun_resnet_param_count(const salun_resnet_cfg *cfg); /* elements of params / grads */
int64_t salun_resnet_bn_channels(const salun_resnet_cfg *cfg); /* elements of running_mean / running_var */
int salun_resnet_create(salun_ctx *ctx, const salun_resnet_cfg *cfg, float *params, float *grads,
                        float *running_mean, float *running_var, salun_resnet **out);
int salun_resnet_destroy(salun_resnet *net);

/* One mini-batch: logits = net(x); loss = loss_sign * mean CE(logits, labels); grads = dloss/dparams
 * (grads is overwritten: the zero_grad() of the reference loops is implied).
 *   x      : fp32 NCHW [n][3][S][S] in [0,1] (what image.cuda() hands the reference model)
 *   labels : int64 [n]
 *   train  : 1 = batch statistics + running-stat update (model.train()), 0 = running statistics (model.eval())
 *   loss_dev (1 float) / logits_dev ([n][num_classes] fp32): optional device outputs. */
int salun_resnet_forward_backward(salun_resnet *net, const float *x, const int64_t *labels, int n,
                                  int train, float loss_sign, float *loss_dev, float *logits_dev,
                                  void *stream);
/* Sync-BN for a mini-batch sharded over ranks: train-mode BatchNorm statistics (sum x, sum x^2,
 * pixel count) and the two batch means of the BatchNorm backward are exchanged through NVLink peer-mapped memory, so the
 * sharded step computes the single-process step of the concatenated batch (nn.BatchNorm2d over the global batch:
 * Classification/models/ResNet.py:111-119 under the reference's single-GPU loop).
 *   peer_sums_host[r]  : rank r's arena of salun_resnet_syncbn_doubles(cfg) doubles (zero-initialised, peer-mapped)
 *   peer_flags_host[r] : rank r's `world` 64-bit epoch flags (zero-initialised, peer-mapped)
 * Every rank must enqueue the same sequence of forward_backward calls (each BatchNorm exchange is a cross-rank barrier
 * of one CTA; a missing rank traps after ~10 s instead of hanging).  resnet18 / resnet34, world <= 8. */
int64_t salun_resnet_syncbn_doubles(const salun_resnet_cfg *cfg);
int salun_resnet_enable_syncbn(salun_resnet *net, double *const *peer_sums_host, unsigned long long *const *peer_flags_host,
                               int rank, int world);

/* eval-mode inference (trainer/val.py:6-72 validate): logits only */
int salun_resnet_forward(salun_resnet *net, const float *x, int n, float *logits_dev, void *stream);


/* ---------------------------------------------------------------------------------------
 * Class-conditional DDPM U-Net forward + backward engine.
 * replaces  model(x_t, t.float(), c, mode=..., ...)  and  loss.backward()  of
 *           DDPM/runners/diffusion.py:974-983  (generate_mask: cond + null pass, eval mode)
 *           DDPM/runners/diffusion.py:533-580  (saliency_unlearn: remain / forget / pseudo-label passes, train mode)
 *           DDPM/functions/losses.py:21-37     (the model call inside noise_estimation_loss_conditional)
 * for the architecture of DDPM/models/diffusion.py:195-413 (Conditional_Model) with the keys of
 * DDPM/configs/cifar10_saliency_unlearn.yml:14-27.  bf16 tensor-core operands and activations, fp32 accumulation,
 * GroupNorm statistics, embeddings, master weights and gradients.
 *
 * Arena layout (caller-owned device buffers, fp32): tensors back to back in model.named_parameters() order
 * (null_classes_emb, temb.dense.*, classes_emb.weight, cemb.dense.*, conv_in, down.*, mid.*, up.0 ... up.L-1, norm_out,
 * conv_out); every conv weight is stored [Cout][kh][kw][Cin] (the reference's [Cout][Cin][kh][kw] permuted (0,2,3,1)),
 * everything else as in PyTorch.
 *
 * GroupNorm does not couple samples, so the caller may run several of the reference's model calls as ONE batch
 * (e.g. remain + forget mini-batches, or the conditional + null passes of classifier-free guidance via drop[]) and
 * hand back per-sample dL/d(eps).
 * ------------------------------------------------------------------------------------- */
typedef struct salun_unet salun_unet;
typedef struct salun_unet_cfg {
  int ch;             /* model.ch (128: the reference's ResnetBlock hard-codes cemb_channels = 512 = 4 * 128) */
  int n_levels;       /* len(model.ch_mult) */
  int ch_mult[8];     /* model.ch_mult */
  int num_res_blocks; /* model.num_res_blocks */
  int n_attn_res;     /* len(model.attn_resolutions) */
  int attn_res[8];    /* model.attn_resolutions (<= 16) */
  int image_size;     /* data.image_size: power of two, smallest level >= 4x4 */
  int in_channels;    /* 3 */
  int out_ch;         /* 3 */
  int n_classes;      /* data.n_classes */
  int max_batch;      /* largest batch a call will pass */
  float dropout;      /* model.dropout; applied when train != 0 (counter-based generator, see salun_unet_forward) */
} salun_unet_cfg;

int64_t salun_unet_param_count(const salun_unet_cfg *cfg);
int salun_unet_create(salun_ctx *ctx, const salun_unet_cfg *cfg, float *params, float *grads, salun_unet **out);
int salun_unet_destroy(salun_unet *net);

/* eps = model._forward(x, t, c) for n samples.
 *   x    : fp32 NCHW [n][3][S][S] (x_t)          t : fp32 [n] (the reference passes t.float())       c : int64 [n]
 *          (labels in [0, n_classes); out-of-range labels are clamped instead of reading outside the embedding table)
 *   drop : optional uint8 [n]; 1 = the class embedding of that sample is replaced by null_classes_emb
 *          (diffusion.py:372-376; the caller draws the cond_drop_prob decisions, or passes all-ones for the null pass
 *          of _forward_with_cond_scale :340-355).  NULL = keep every class embedding.
 *   train: != 0 applies dropout(cfg.dropout) after norm2 + swish of every ResnetBlock with a counter-based generator
 *          keyed by `seed` (the mask is regenerated, not stored, in the backward pass).  It is a different stream of
 *          random numbers than torch's Philox: parity runs use dropout 0.
 *   save_for_backward: != 0 keeps the activations for salun_unet_backward (a later forward overwrites them).
 *   eps_out : fp32 NCHW [n][3][S][S]. */
int salun_unet_forward(salun_unet *net, const float *x, const float *t, const int64_t *c, const uint8_t *drop, int n,
                       int train, uint64_t seed, int save_for_backward, float *eps_out, void *stream);
/* grads (=|+=) d loss / d params for the last saved forward, given d_eps = d loss / d eps_out (fp32 NCHW [n][3][S][S]).
 * accumulate == 0 overwrites the gradient arena (the zero_grad() of the reference loop is implied). */
int salun_unet_backward(salun_unet *net, const float *d_eps, int accumulate, void *stream);

/* x_t = x0 * sqrt_abar[t] + e * sqrt_1m_abar[t], x0 = rescale ? 2 * x01 - 1 : x01          (n samples of chw floats)
 * replaces  data_transform (DDPM/datasets/__init__.py:241-255) and the q-sample of DDPM/functions/losses.py:31-32 /
 *           runners/diffusion.py:558-559, :971-973.  sqrt_abar / sqrt_1m_abar: device tables over the T timesteps
 *           ((1 - betas).cumprod(0).sqrt() and (1 - cumprod).sqrt(), fp32; num_timesteps entries, t is clamped into
 *           them).  Same rounding sequence as the torch statements: bit-identical. */
int salun_ddpm_q_sample(salun_ctx *ctx, const float *x01, const float *e, const int64_t *t, const float *sqrt_abar,
                        const float *sqrt_1m_abar, int num_timesteps, int rescale, int n, int chw, float *x_t,
                        void *stream);
/* The eps-prediction losses of an iteration and their gradient w.r.t. eps in one pass:
 *   sumsq_ps[i] = sum_chw (eps[i] - target[i])^2 ;  d_eps[i] = 2 * w[i] * (eps[i] - target[i]) ;  loss = sum_i w[i] * sumsq_ps[i]
 * replaces  noise_estimation_loss_conditional (losses.py:33-37: target = e, w = alpha / n for the remain batch, -1 / n for
 *           method ga), MSELoss(pseudo, output) (runners/diffusion.py:570: target = pseudo, w = 1 / (n * chw)), their sum
 *           (:572) and the first step of loss.backward() (:579-580).  w: device fp32 [n] per-sample weights. */
int salun_ddpm_eps_loss_grad(salun_ctx *ctx, const float *eps, const float *target, const float *w, int n, int chw,
                             float *d_eps, float *sumsq_ps, float *loss_dev, void *stream);

/* bring-up / parity-test accessors: the tape's activation tensors, exported as fp32 NCHW [n][C][H][H]
 * (which = 0: value, 1: gradient of the last backward). */
int salun_unet_num_tensors(const salun_unet *net);
int salun_unet_tensor_info(const salun_unet *net, int idx, char *name_buf, int name_cap, int *C, int *H);
int salun_unet_export_tensor(salun_unet *net, int idx, int which, float *out_nchw, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * The steps either side of the hot path, csrc/salun_data.cu
 * --------------------------------------------------------------------------------------------------------------- */

/* out_nchw[i] (fp32 [n][3][H][W] in [0,1]) = ToTensor(HorizontalFlip_{flip[i]}(Crop_{crop_xy[i]}(ZeroPad_{pad}(
 * images_hwc[index[i]])))) from a uint8 [n_images][H][W][3] dataset resident in device memory.
 *   index   int64 [n] or NULL (identity);  crop_xy int32 [n][2] = (left, top) of the crop window inside the padded image,
 *   0 .. 2*pad, or NULL (centre = no crop);  flip uint8 [n] or NULL (no flip).  Bit-identical to the torchvision ops.
 * replaces the DataLoader gather + RandomCrop(32, padding=4) + RandomHorizontalFlip + ToTensor
 * Classification/dataset.py:549-555, main_forget.py:42-48 */
int salun_augment_batch(salun_ctx *ctx, const uint8_t *images_hwc, int64_t n_images, const int64_t *index,
                        const int32_t *crop_xy, const uint8_t *flip, int n, int H, int W, int pad, float *out_nchw,
                        void *stream);

/* One batch of logits [n][K]:  loss_sum_dev[0] += sum_i CE(logits_i, labels_i)  (double),  correct_dev[0] += #{argmax ==
 * label},  probs (optional, [n][K]) = softmax(logits).  labels may be NULL when only probs are wanted.  Accumulators live
 * on the device: a whole validation pass needs one read-back.
 * replaces criterion + utils.accuracy + the two .item() per batch of  Classification/trainer/val.py:44-61  and
 * F.softmax of  Classification/evaluation/SVC_MIA.py:44-46 */
int salun_eval_logits(salun_ctx *ctx, const float *logits, const int64_t *labels, int n, int K, float *probs,
                      double *loss_sum_dev, int64_t *correct_dev, void *stream);

/* One step of the class-conditional generalized (DDIM) sampler after the two U-Net passes of classifier-free guidance:
 *   et = (1 + s) eps_cond - s eps_null ;  x0_t = (xt - et sqrt(1 - at)) / sqrt(at) ;
 *   c1 = eta sqrt((1 - at/at_next)(1 - at_next)/(1 - at)) ;  c2 = sqrt((1 - at_next) - c1^2) ;
 *   x_next = sqrt(at_next) x0_t + c1 noise + c2 et          (at, at_next: fp32 [n] = compute_alpha of t and next_t)
 * eps_null NULL: no guidance (et = eps_cond); noise may be NULL when eta == 0; x0_out optional.
 * replaces the update statements of generalized_steps_conditional   DDPM/functions/denoising.py:72-95 */
int salun_ddim_step(salun_ctx *ctx, const float *eps_cond, const float *eps_null, const float *xt, const float *noise,
                    const float *at, const float *at_next, float cond_scale, float eta, int n, int chw, float *x_next,
                    float *x0_out, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Op-level entry points (csrc/salun_ops.cu): the tcgen05 convolution / GEMM and the elementwise kernels one layer at a
 * time, plus the layers the Stable-Diffusion U-Net adds (LayerNorm, GEGLU, multi-head self / cross attention, cos|sin
 * timestep embedding, GroupNorm of any width).  The SD U-Net forward (SD/ldm/modules/diffusionmodules/openaimodel.py:814-846)
 * is composed from them by unlearn_saliency_b200/sd/engine.py and replayed from a CUDA graph.
 * Activations ("act") are bf16 in libsalun.so and (hi, lo) bf16 pairs in libsalun_split.so: salun_act_bytes() per element.
 * Layouts: padded NHWC [n][H+2][W+2][C] (zero halo, written once by the caller's memset) or flat [rows][C].
 * --------------------------------------------------------------------------------------------------------------- */
int salun_act_bytes(void);   /* 2 (bf16) or 4 (bf16 hi/lo pair) */
int salun_wop_k(void);       /* bf16 elements a prepared weight operand spends per weight element: 1 or 4 */
int salun_op_f32_to_act(salun_ctx *ctx, const float *src, int64_t ld_src, void *dst, int64_t ld_dst, int64_t rows, int64_t cols,
                        void *stream);
int salun_op_act_to_f32(salun_ctx *ctx, const void *src, int64_t ld_src, float *dst, int64_t ld_dst, int64_t rows, int64_t cols,
                        void *stream);
/* x fp32 NCHW [n][C][H][W] -> padded NHWC act with Cp >= C channels (the extra ones zero) and back */
int salun_op_nchw_to_padded(salun_ctx *ctx, const float *x, void *out_pad, int n, int C, int Cp, int H, int W, void *stream);
int salun_op_padded_to_nchw(salun_ctx *ctx, const void *in_pad, float *out, int n, int C, int H, int W, void *stream);
/* out NCHW [n][C][H][W] = y[pixel][c] (+ bias[c]); y: fp32 rows of leading dimension ld >= C (a GEMM with padded N) */
int salun_op_rows_to_nchw(salun_ctx *ctx, const float *y, int ld, const float *bias, float *out, int n, int C, int H, int W,
                          void *stream);
/* PyTorch weight (Conv2d OIHW fp32, or Linear [out][in] with ks = 1) -> tensor-core operand rows [cout_pad][ks*ks*cin_pad]
 * (tap-major then channel, zero padded); salun_wop_k() * 2 bytes per element */
int salun_op_prep_weight(salun_ctx *ctx, const float *w, void *wop, int cout, int cin, int ks, int cout_pad, int cin_pad,
                         void *stream);
/* Stride-1 convolution (ks 3 / pad 1 or ks 1) or Linear over token rows (in_flat, ks 1) with the epilogue fused:
 *   out = conv(in, w) + bias[col] + rowbias[sample][col] + addend        (bias / rowbias / addend optional)
 * in: padded act [n][H+2][W+2][cin] or flat [n*H*W][cin]; out: act flat or padded (out_pad) and / or fp32 flat (out_f32);
 * addend has the layout of out.  cin, cout multiples of 64 (pad the operand).  rowbias: fp32 [n][rb_ld].
 * replaces nn.Conv2d / nn.Linear + the adds of ResBlock._forward (openaimodel.py:268-288), CrossAttention.to_q / to_k / to_v
 * / to_out, FeedForward (attention.py:37-66,168-192) */
int salun_op_conv(salun_ctx *ctx, const void *in, int in_flat, const void *wop, const float *bias, const float *rowbias, int rb_ld,
                  const void *addend, void *out, int out_pad, float *out_f32, int n, int H, int W, int cin, int cout, int ks,
                  void *stream);
/* Downsample: conv 3x3 / stride 2 / padding 1 (openaimodel.py:131-160).  col_scratch: act [n*(Hin/2)*(Win/2)][9*cin] */
int salun_op_conv_s2(salun_ctx *ctx, const void *in_pad, void *col_scratch, const void *wop, const float *bias, void *out_padded,
                     int n, int Hin, int Win, int cin, int cout, void *stream);
/* GroupNorm(32, C, eps) (+ SiLU): padded in -> padded or flat out (util.py:217-224 normalization).  stats_ws: scratch of
 * salun_op_groupnorm_ws_floats(n) floats, 8-byte aligned (per-(sample, group, pixel-slice) fp64 partial sums) */
int64_t salun_op_groupnorm_ws_floats(int n);
/* Optional caller-owned scratch (16-byte aligned; 32 MiB covers SD v1.4 at batch 2) that lets salun_op_conv / salun_op_conv_s2
 * split the k loop of small-M, deep-K GEMMs over several CTAs per output tile (fp32 partial tiles, fixed-order sum: results
 * do not depend on the grid).  NULL / 0 = off.  One stream per context at a time while it is set. */
int salun_op_set_scratch(salun_ctx *ctx, void *scratch, int64_t bytes);
int salun_op_groupnorm(salun_ctx *ctx, const void *in_pad, const float *gamma, const float *beta, float *stats_ws, void *out,
                       int out_flat, int n, int H, int W, int C, float eps, int swish, void *stream);
int salun_op_upsample2(salun_ctx *ctx, const void *in_pad, void *out_pad, int n, int H, int C, void *stream);
int salun_op_concat(salun_ctx *ctx, const void *a_pad, int Ca, const void *b_pad, int Cb, void *out_pad, int n, int H, void *stream);
/* out[n][N] = (silu_in ? silu(x) : x)[n][K] . w[N][K]^T + b  in fp32 (time_embed / emb_layers, openaimodel.py:556-560,222-229) */
int salun_op_linear_f32(salun_ctx *ctx, const float *x, const float *w, const float *b, float *out, float *tmp, int n, int K, int N,
                        int silu_in, void *stream);
/* timestep_embedding(t, dim, max_period): [cos | sin]   (SD/ldm/modules/diffusionmodules/util.py:173-197) */
int salun_sd_timestep_embedding(salun_ctx *ctx, const float *t, float *out, int n, int dim, float max_period, void *stream);
/* nn.LayerNorm(C) over token rows (attention.py:221-223) */
int salun_sd_layernorm(salun_ctx *ctx, const void *x, const float *gamma, const float *beta, void *out, int64_t rows, int C,
                       float eps, void *stream);
/* GEGLU: out[r][c] = proj[r][c] * gelu(proj[r][Ci + c])   (attention.py:37-44) */
int salun_sd_geglu(salun_ctx *ctx, const void *proj, void *out, int64_t rows, int Ci, void *stream);
/* Multi-head attention: out = merge_heads(softmax(q_h k_h^T / sqrt(d)) v_h); q [n*Tq][heads*d], k / v [n*Tk][heads*d] (act);
 * two batched tcgen05 GEMMs per call over (sample, head) units, fp32 scores.  ws: salun_sd_attention_ws_bytes(...) bytes.
 * replaces CrossAttention.forward after the projections   SD/ldm/modules/attention.py:177-191 */
int64_t salun_sd_attention_ws_bytes(int n, int Tq, int Tk, int heads, int d);
int salun_sd_attention(salun_ctx *ctx, void *ws, int64_t ws_bytes, const void *q, const void *k, const void *v, void *out, int n,
                       int Tq, int Tk, int heads, int d, void *stream);
/* the same with row strides (elements) for q, k, v: column slices of a fused q | k | v (self-attention) or k | v (cross-
 * attention) projection output, so one GEMM serves the three / two Linears of attention.py:177-179 */
int salun_sd_attention_ld(salun_ctx *ctx, void *ws, int64_t ws_bytes, const void *q, int ldq, const void *k, int ldk, const void *v,
                          int ldv, void *out, int n, int Tq, int Tk, int heads, int d, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SALUN_H_ */
