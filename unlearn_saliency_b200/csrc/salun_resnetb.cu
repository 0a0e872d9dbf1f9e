// salun_resnetb.cu -- runtime of the Bottleneck ResNets (resnet50/101/152, Classification/models/ResNet.py:127-177,
// 358-390) with either stem (imagenet=True: 7x7/2 conv + 3x3/2 max pool, ResNet.py:224-230; imagenet=False: 3x3/1 conv).
//
// BASELINE.json config 4 (ResNet-50 / ImageNet shape).  Spatial sizes here are 56/28/14/7 (not powers of two), so this
// path keeps activations FLAT ([n*H*W][C] bf16) and feeds the tcgen05 GEMM kernels in their plain 2-D mode:
//   1x1 / stride 1 convs (2/3 of the layers, ~half of the FLOPs): A is the activation itself -- no copy at all;
//   3x3, strided 1x1 and the stem: explicit bf16 patch matrix, one buffer per conv (kept for the wgrad);
//   dgrad of 1x1/s1: one GEMM straight into the input gradient (+ residual addend); others: dcol GEMM + col2im gather.
// Everything else (BatchNorm kernels, 1-bit ReLU masks, split-K workspace + deterministic reduce, fused masked SGD)
// is shared with the BasicBlock runtime (salun_resnet.cu).
#include <stdlib.h>

#include <map>
#include <vector>

#include "salun_elem.cuh"
#include "salun_gemm.cuh"
#include "salun_resnetb.cuh"
#include "salun_unet_elem.cuh"  // launch_sgemm, launch_colsum_f32

namespace salun {


struct FConv {
  int cin, cout, ks, stride, pad, hin, win, hout, wout;
  bool stem;
  int kc, kcp;
  int64_t w_off, g_off, b_off;
  int rs_off, in_act;
  bool direct;  // 1x1 / stride 1: the input activation IS the GEMM operand
  wop_t *w_fwd, *w_dg;
  act_t *col;  // own patch matrix [n*hout*wout][kcp] of a 3x3 / strided / stem conv: built in the forward pass, read again by
               // the weight gradient (a shared scratch meant rebuilding all 20 of them per step: 7.6 ms of 40 at batch 256)
  act_t *y, *dy;
  float *stat_sum, *stat_sq, *saved_mean, *saved_invstd, *coef, *bwd_partials, *wg_ws;
  double *slices;
  int wg_splits_max;
};
struct FAct {
  int C, H, W;
  act_t *p, *dout, *dz;
  uint8_t *rmask;
};
struct FBlock {
  int c1, c2, c3, cd, in_act, mid1, mid2, out_act;
};
struct FMaps {
  CUtensorMap fwdA, fwdB, dgA, dgB, wgA, wgB;
};

struct FlatNet {
  salun_ctx *ctx;
  salun_resnet_cfg cfg;
  float *params, *grads, *rmean, *rvar;
  int64_t n_params;
  int n_bn;
  std::vector<FConv> convs;
  std::vector<FAct> acts;
  std::vector<FBlock> blocks;
  int stem_act, pool_act;  // pool_act == stem_act when there is no max pool
  uint8_t *pool_argmax;
  int64_t fc_w_off, fc_b_off;
  int feat;
  float *pooled, *logits, *dlogits, *loss_ps, *dpooled;
  act_t *scratch_dcol;
  WPrepEntry *wprep_table;
  WgReduceEntry *wgred_table, *wgred_host;
  std::vector<int> wg_splits;
  std::vector<void *> allocs;
  std::map<int, std::vector<FMaps>> plans;
  int last_n, last_train;
  bool fwd_done;
};

static int bottleneck_blocks(int depth, int s) {
  static const int r50[4] = {3, 4, 6, 3}, r101[4] = {3, 4, 23, 3}, r152[4] = {3, 8, 36, 3};
  return depth == 152 ? r152[s] : (depth == 101 ? r101[s] : r50[s]);
}

static int fbuild(const salun_resnet_cfg &c, FlatNet *net) {
  int64_t off = 0;
  int rs = 0;
  auto add_act = [&](int C, int H, int W) {
    FAct a{};
    a.C = C; a.H = H; a.W = W;
    net->acts.push_back(a);
    return (int)net->acts.size() - 1;
  };
  auto add_conv = [&](int cin, int cout, int ks, int stride, int pad, int hin, int win, bool stem, int in_act) {
    FConv L{};
    L.cin = cin; L.cout = cout; L.ks = ks; L.stride = stride; L.pad = pad; L.hin = hin; L.win = win;
    L.hout = (hin + 2 * pad - ks) / stride + 1;
    L.wout = (win + 2 * pad - ks) / stride + 1;
    L.stem = stem;
    L.kc = ks * ks * cin;
    L.kcp = (L.kc + 63) / 64 * 64;
    L.w_off = off; off += (int64_t)cout * L.kc;
    L.g_off = off; off += cout;
    L.b_off = off; off += cout;
    L.rs_off = rs; rs += cout;
    L.in_act = in_act;
    L.direct = !stem && ks == 1 && stride == 1;
    net->convs.push_back(L);
    return (int)net->convs.size() - 1;
  };
  int H = c.image_size, W = c.image_size;
  if (c.imagenet_stem) {
    const int s = add_conv(3, 64, 7, 2, 3, H, W, true, -1);
    H = net->convs[s].hout; W = net->convs[s].wout;
    net->stem_act = add_act(64, H, W);
    H = (H + 2 - 3) / 2 + 1; W = (W + 2 - 3) / 2 + 1;
    net->pool_act = add_act(64, H, W);
  } else {
    add_conv(3, 64, 3, 1, 1, H, W, true, -1);
    net->stem_act = net->pool_act = add_act(64, H, W);
  }
  int cur = net->pool_act, inpl = 64;
  const int planes[4] = {64, 128, 256, 512};
  for (int s = 0; s < 4; ++s)
    for (int b = 0; b < bottleneck_blocks(c.depth, s); ++b) {
      const int stride = (b == 0 && s > 0) ? 2 : 1;
      const int width = planes[s], outc = planes[s] * 4;
      FBlock B{};
      B.in_act = cur;
      // named_parameters order: conv1, bn1, conv2, bn2, conv3, bn3, downsample.0, downsample.1   (ResNet.py:148-156)
      B.c1 = add_conv(inpl, width, 1, 1, 0, H, W, false, cur);
      B.mid1 = add_act(width, H, W);
      B.c2 = add_conv(width, width, 3, stride, 1, H, W, false, B.mid1);
      const int Ho = net->convs[B.c2].hout, Wo = net->convs[B.c2].wout;
      B.mid2 = add_act(width, Ho, Wo);
      B.c3 = add_conv(width, outc, 1, 1, 0, Ho, Wo, false, B.mid2);
      B.cd = -1;
      if (stride != 1 || inpl != outc) B.cd = add_conv(inpl, outc, 1, stride, 0, H, W, false, cur);
      B.out_act = add_act(outc, Ho, Wo);
      net->blocks.push_back(B);
      cur = B.out_act;
      inpl = outc;
      H = Ho; W = Wo;
    }
  net->feat = inpl;
  net->fc_w_off = off; off += (int64_t)c.num_classes * inpl;
  net->fc_b_off = off; off += c.num_classes;
  net->n_params = off;
  net->n_bn = rs;
  return SALUN_OK;
}

#define TRY(expr)        \
  do {                   \
    int _rc = (expr);    \
    if (_rc) return _rc; \
  } while (0)

static int fpick_bn(int N, int64_t M) {
  if (N % 256 == 0 && ((M + 127) / 128) * (N / 256) >= 96) return 256;
  return N % 128 == 0 ? 128 : 64;
}

template <typename T>
static int fmalloc(FlatNet *net, T **p, size_t count, bool zero) {
  void *q = nullptr;
  SALUN_CUDA_OK(cudaMalloc(&q, count * sizeof(T)));
  if (zero) SALUN_CUDA_OK(cudaMemset(q, 0, count * sizeof(T)));
  net->allocs.push_back(q);
  *p = (T *)q;
  return SALUN_OK;
}

static const act_t *conv_operand(FlatNet *net, const FConv &L) { return L.direct ? net->acts[L.in_act].p : L.col; }

static int fplan(FlatNet *net, int n, std::vector<FMaps> **out) {
  auto it = net->plans.find(n);
  if (it != net->plans.end()) {
    *out = &it->second;
    return SALUN_OK;
  }
  std::vector<FMaps> maps(net->convs.size());
  for (size_t i = 0; i < net->convs.size(); ++i) {
    const FConv &L = net->convs[i];
    FMaps &m = maps[i];
    const int64_t Mo = (int64_t)n * L.hout * L.wout;
    const act_t *A = conv_operand(net, L);
    TRY(make_tmap_2d_act(&m.fwdA, A, Mo, L.kcp, 128));
    TRY(make_tmap_2d_wop(&m.fwdB, L.w_fwd, L.cout, L.kcp, fpick_bn(L.cout, Mo)));
    TRY(make_tmap_2d_act(&m.wgA, L.dy, Mo, L.cout, 64));
    TRY(make_tmap_2d_act(&m.wgB, A, Mo, L.kcp, 64));
    if (!L.stem) {
      TRY(make_tmap_2d_act(&m.dgA, L.dy, Mo, L.cout, 128));
      TRY(make_tmap_2d_wop(&m.dgB, L.w_dg, L.kc, L.cout, fpick_bn(L.kc, Mo)));
    }
  }
  auto res = net->plans.emplace(n, std::move(maps));
  *out = &res.first->second;
  return SALUN_OK;
}

static void build_col(FlatNet *net, const FConv &L, const float *x, int n, cudaStream_t st) {
  if (L.direct) return;
  const salun_resnet_cfg &c = net->cfg;
  if (L.stem) {
    const float inv_std[3] = {1.f / c.std[0], 1.f / c.std[1], 1.f / c.std[2]};
    launch_stem_im2col_generic(x, L.col, n, L.hin, L.win, L.ks, L.stride, L.pad, L.hout, L.wout, L.kcp, c.mean,
                               inv_std, st);
  } else {
    launch_im2col_flat(net->acts[L.in_act].p, L.col, n, L.hin, L.win, L.cin, L.ks, L.stride, L.pad, L.hout,
                       L.wout, st);
  }
}

static int fconv_forward(FlatNet *net, int ci, const FMaps &m, const float *x, int n, int train, cudaStream_t st) {
  const FConv &L = net->convs[ci];
  const int M = n * L.hout * L.wout;
  build_col(net, L, x, n, st);
  ConvGemmArgs a{};
  a.mode_a = 0;
  a.num_k_blocks = L.kcp / 64;
  a.M = M;
  a.N = L.cout;
  a.out_bf16 = L.y;
  a.ld_out = L.cout;
  if (train) {
    a.stat_sum = L.stat_sum;
    a.stat_sq = L.stat_sq;
  }
  TRY(launch_conv_gemm(m.fwdA, m.fwdB, a, fpick_bn(L.cout, M), st));
  if (train) launch_bn_stats_reduce(L.stat_sum, L.stat_sq, (M + 127) / 128 * 4, L.cout, L.slices, st);
  return SALUN_OK;
}

// the one-CTA-per-sample head kernels re-read the whole weight per sample: fine for 10 x 2048, 2.7 ms for 1000 x 2048 at batch 256
static bool wide_head(const FlatNet *net) { return (int64_t)net->cfg.num_classes * net->feat >= (1 << 18); }

static BnFwd fbn_of(FlatNet *net, const FConv &L) {
  BnFwd b{};
  b.y = L.y;
  b.slices = L.slices;
  b.gamma = net->params + L.g_off;
  b.beta = net->params + L.b_off;
  b.running_mean = net->rmean + L.rs_off;
  b.running_var = net->rvar + L.rs_off;
  b.saved_mean = L.saved_mean;
  b.saved_invstd = L.saved_invstd;
  return b;
}

static int fforward(FlatNet *net, const float *x, const int64_t *labels, int n, int train, float sign, float *loss_dev,
                    float *logits_out, bool need_bwd, cudaStream_t st) {
  std::vector<FMaps> *plan;
  TRY(fplan(net, n, &plan));
  launch_prep_w_all(net->wprep_table, (int)net->convs.size(), net->params, need_bwd ? 1 : 0, st);
  const salun_resnet_cfg &c = net->cfg;
  auto apply = [&](const FConv &L, const BnFwd *second, const act_t *resid, FAct &out, int relu) {
    BnFwd b = fbn_of(net, L);
    launch_bn_apply_flat(b, second, resid, out.p, need_bwd ? out.rmask : nullptr, n * L.hout * L.wout, L.cout, relu, train,
                         c.bn_eps, c.bn_momentum, st);
  };
  TRY(fconv_forward(net, 0, (*plan)[0], x, n, train, st));
  apply(net->convs[0], nullptr, nullptr, net->acts[net->stem_act], 1);
  if (net->pool_act != net->stem_act) {
    const FAct &s = net->acts[net->stem_act];
    launch_maxpool_fwd(s.p, net->acts[net->pool_act].p, net->pool_argmax, n, s.H, s.W, s.C, st);
  }
  for (const FBlock &B : net->blocks) {
    TRY(fconv_forward(net, B.c1, (*plan)[B.c1], x, n, train, st));
    apply(net->convs[B.c1], nullptr, nullptr, net->acts[B.mid1], 1);
    TRY(fconv_forward(net, B.c2, (*plan)[B.c2], x, n, train, st));
    apply(net->convs[B.c2], nullptr, nullptr, net->acts[B.mid2], 1);
    TRY(fconv_forward(net, B.c3, (*plan)[B.c3], x, n, train, st));
    if (B.cd >= 0) {
      TRY(fconv_forward(net, B.cd, (*plan)[B.cd], x, n, train, st));
      BnFwd bd = fbn_of(net, net->convs[B.cd]);
      apply(net->convs[B.c3], &bd, nullptr, net->acts[B.out_act], 1);
    } else {
      apply(net->convs[B.c3], nullptr, net->acts[B.in_act].p, net->acts[B.out_act], 1);
    }
  }
  const FAct &last = net->acts[net->blocks.back().out_act];
  launch_avgpool_flat(last.p, net->pooled, n, last.H * last.W, last.C, st);
  if (wide_head(net)) {  // logits[n][K] = pooled[n][C] . W[K][C]^T + b as an fp32 GEMM, then the per-sample cross entropy
    float *lg = logits_out ? logits_out : net->logits;
    launch_sgemm(net->pooled, net->feat, 1, net->params + net->fc_w_off, 1, net->feat, lg, c.num_classes, n, c.num_classes,
                 net->feat, net->params + net->fc_b_off, 0, st);
    if (labels) launch_ce_rows(lg, labels, net->dlogits, net->loss_ps, n, c.num_classes, sign, st);
  } else {
    launch_fc_ce(net->pooled, net->params + net->fc_w_off, net->params + net->fc_b_off, labels,
                 logits_out ? logits_out : net->logits, net->dlogits, net->loss_ps, n, net->feat, c.num_classes, sign, st);
  }
  if (labels && loss_dev) launch_loss_sum(net->loss_ps, n, sign, loss_dev, st);
  SALUN_CUDA_OK(cudaGetLastError());
  net->last_n = n;
  net->last_train = train;
  net->fwd_done = true;
  return SALUN_OK;
}

static void fbn_backward(FlatNet *net, const FConv &L, const act_t *dout, const uint8_t *rmask, act_t *dz, int n, int train,
                         cudaStream_t st) {
  const int M = n * L.hout * L.wout;
  // the reduce / apply kernels only use n*H*W as a row count when dY is flat: pass it as (n = M, H = W = 1)
  launch_bn_bwd_reduce(dout, rmask, L.y, L.saved_mean, L.saved_invstd, L.bwd_partials, M, 1, 1, L.cout, st);
  launch_bn_bwd_finalize(L.bwd_partials, 0, L.cout, net->params + L.g_off, L.saved_invstd, (float)M, train,
                         net->grads + L.g_off, net->grads + L.b_off, L.coef, st);
  launch_bn_bwd_apply(dout, rmask, L.y, L.saved_mean, L.saved_invstd, L.coef, L.dy, 0, dz, M, 1, 1, L.cout, st);
}

static int fwgrad(FlatNet *net, int ci, const FMaps &m, const float *x, int n, cudaStream_t st) {
  const FConv &L = net->convs[ci];
  const int64_t M = (int64_t)n * L.hout * L.wout;
  (void)x;  // the patch matrix of the forward pass is still in L.col
  WgradArgs a{};
  a.mode_a = 0;
  a.mode_b = 0;
  a.kb_total = (int)((M + 63) / 64);
  a.cin_blocks = 1;
  a.kw = 1;
  a.H = a.W = 1;
  const WgradGeom geo = wgrad_geometry(L.cout, L.kcp);
  a.total_blocks = geo.total_blocks;
  a.n_blocks = geo.n_blocks;
  a.Cout = L.cout;
  a.ldw = L.kc;
  a.kvalid = L.kc;
  const int co_tiles = geo.co_tiles, groups = geo.groups;
  int splits = net->ctx->num_sms / (co_tiles * groups);
  if (splits < 1) splits = 1;
  if (splits > L.wg_splits_max) splits = L.wg_splits_max;
  if (splits > a.kb_total) splits = a.kb_total;
  a.kb_per_split = (a.kb_total + splits - 1) / splits;
  splits = (a.kb_total + a.kb_per_split - 1) / a.kb_per_split;
  a.dw = L.wg_ws;
  a.split_stride = (long long)L.cout * L.kc;
  net->wg_splits[ci] = splits;
  return launch_wgrad(m.wgA, m.wgB, a, co_tiles, groups, splits, st);
}

// gradient w.r.t. the conv's input activation (+ addend), written to acts[L.in_act].dout
static int fdgrad(FlatNet *net, int ci, const FMaps &m, const act_t *addend, int n, cudaStream_t st) {
  const FConv &L = net->convs[ci];
  const int Mo = n * L.hout * L.wout;
  FAct &in = net->acts[L.in_act];
  ConvGemmArgs a{};
  a.mode_a = 0;
  a.num_k_blocks = L.cout / 64;
  a.M = Mo;
  a.N = L.kc;
  a.ld_out = L.kc;
  if (L.direct) {
    a.out_bf16 = in.dout;
    a.addend = addend;
    return launch_conv_gemm(m.dgA, m.dgB, a, fpick_bn(L.kc, Mo), st);
  }
  a.out_bf16 = net->scratch_dcol;
  TRY(launch_conv_gemm(m.dgA, m.dgB, a, fpick_bn(L.kc, Mo), st));
  launch_col2im_flat(net->scratch_dcol, addend, in.dout, n, L.hin, L.win, L.cin, L.ks, L.stride, L.pad, L.hout, L.wout, st);
  return SALUN_OK;
}

static int fbackward(FlatNet *net, const float *x, cudaStream_t st) {
  if (!net->fwd_done) {
    set_error("salun_resnet backward called before forward");
    return SALUN_ERR_STATE;
  }
  const int n = net->last_n, train = net->last_train;
  std::vector<FMaps> *plan;
  TRY(fplan(net, n, &plan));
  const FAct &last = net->acts[net->blocks.back().out_act];
  if (wide_head(net)) {
    const int K = net->cfg.num_classes, C = net->feat;
    // dW[K][C] = dlogits^T . pooled ; db = column sums of dlogits ; dpooled[n][C] = dlogits . W   (fp32 GEMMs, fixed order)
    launch_sgemm(net->dlogits, 1, K, net->pooled, C, 1, net->grads + net->fc_w_off, C, K, C, n, nullptr, 0, st);
    launch_colsum_f32(net->dlogits, K, n, K, net->grads + net->fc_b_off, st);
    launch_sgemm(net->dlogits, K, 1, net->params + net->fc_w_off, C, 1, net->dpooled, C, n, C, K, nullptr, 0, st);
    launch_pool_grad_bcast(net->dpooled, last.dout, n, C, last.H * last.W, st);
  } else {
    launch_fc_bwd(net->pooled, net->dlogits, net->params + net->fc_w_off, net->grads + net->fc_w_off,
                  net->grads + net->fc_b_off, last.dout, n, net->feat, net->cfg.num_classes, last.H * last.W, st);
  }
  for (int bi = (int)net->blocks.size() - 1; bi >= 0; --bi) {
    const FBlock &B = net->blocks[bi];
    FAct &in = net->acts[B.in_act], &m1 = net->acts[B.mid1], &m2 = net->acts[B.mid2], &out = net->acts[B.out_act];
    const bool identity = B.cd < 0;
    // out = relu(bn3(y3) + shortcut)
    fbn_backward(net, net->convs[B.c3], out.dout, out.rmask, identity ? out.dz : nullptr, n, train, st);
    if (!identity) fbn_backward(net, net->convs[B.cd], out.dout, out.rmask, nullptr, n, train, st);
    TRY(fdgrad(net, B.c3, (*plan)[B.c3], nullptr, n, st));  // -> mid2.dout
    TRY(fwgrad(net, B.c3, (*plan)[B.c3], x, n, st));
    fbn_backward(net, net->convs[B.c2], m2.dout, m2.rmask, nullptr, n, train, st);
    TRY(fdgrad(net, B.c2, (*plan)[B.c2], nullptr, n, st));  // -> mid1.dout
    TRY(fwgrad(net, B.c2, (*plan)[B.c2], x, n, st));
    fbn_backward(net, net->convs[B.c1], m1.dout, m1.rmask, nullptr, n, train, st);
    if (identity) {
      TRY(fdgrad(net, B.c1, (*plan)[B.c1], out.dz, n, st));  // in.dout = dgrad(conv1) + identity-shortcut gradient
    } else {
      TRY(fdgrad(net, B.cd, (*plan)[B.cd], nullptr, n, st));  // in.dout = dgrad(projection shortcut)
      TRY(fwgrad(net, B.cd, (*plan)[B.cd], x, n, st));
      TRY(fdgrad(net, B.c1, (*plan)[B.c1], in.dout, n, st));  //         += dgrad(conv1)   (element-wise read-modify-write)
    }
    TRY(fwgrad(net, B.c1, (*plan)[B.c1], x, n, st));
  }
  FAct &sa = net->acts[net->stem_act];
  if (net->pool_act != net->stem_act)
    launch_maxpool_bwd(net->acts[net->pool_act].dout, net->pool_argmax, sa.dout, n, sa.H, sa.W, sa.C, st);
  fbn_backward(net, net->convs[0], sa.dout, sa.rmask, nullptr, n, train, st);
  TRY(fwgrad(net, 0, (*plan)[0], x, n, st));
  for (size_t i = 0; i < net->convs.size(); ++i) {
    const FConv &L = net->convs[i];
    WgReduceEntry &e = net->wgred_host[i];
    e.ws = L.wg_ws;
    e.dst_off = L.w_off;
    e.count = (long long)L.cout * L.kc;
    e.splits = net->wg_splits[i];
    e.kc = L.kc;
  }
  SALUN_CUDA_OK(cudaMemcpyAsync(net->wgred_table, net->wgred_host, net->convs.size() * sizeof(WgReduceEntry),
                                cudaMemcpyHostToDevice, st));
  launch_wgrad_reduce(net->wgred_table, (int)net->convs.size(), net->grads, st);
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

// ------------------------------------------------------------------------------------------------ public (internal) API
int64_t flatnet_param_count(const salun_resnet_cfg *cfg, int64_t *n_bn) {
  FlatNet tmp{};
  if (fbuild(*cfg, &tmp)) return -1;
  if (n_bn) *n_bn = tmp.n_bn;
  return tmp.n_params;
}

void flatnet_destroy(FlatNet *net) {
  if (!net) return;
  cudaSetDevice(net->ctx->device);
  for (void *p : net->allocs) cudaFree(p);
  if (net->wgred_host) cudaFreeHost(net->wgred_host);
  delete net;
}

int flatnet_create(salun_ctx *ctx, const salun_resnet_cfg *cfg, float *params, float *grads, float *rmean, float *rvar,
                   FlatNet **out) {
  FlatNet *net = new FlatNet();
  net->ctx = ctx;
  net->cfg = *cfg;
  net->params = params;
  net->grads = grads;
  net->rmean = rmean;
  net->rvar = rvar;
  net->fwd_done = false;
  net->wgred_host = nullptr;
  net->pool_argmax = nullptr;
  int rc = fbuild(*cfg, net);
  if (rc) {
    delete net;
    return rc;
  }
  const int64_t nb = cfg->max_batch;
#define A(expr)              \
  do {                       \
    int _rc = (expr);        \
    if (_rc) {               \
      flatnet_destroy(net);  \
      return _rc;            \
    }                        \
  } while (0)
  for (FAct &a : net->acts) {
    const size_t e = (size_t)nb * a.H * a.W * a.C;
    A(fmalloc(net, &a.p, e, true));
    A(fmalloc(net, &a.dout, e, false));
    A(fmalloc(net, &a.dz, e, false));
    A(fmalloc(net, &a.rmask, e / 8, true));
  }
  if (net->pool_act != net->stem_act) {
    const FAct &p = net->acts[net->pool_act];
    A(fmalloc(net, &net->pool_argmax, (size_t)nb * p.H * p.W * p.C, true));
  }
  size_t dcol_max = 0;
  for (FConv &L : net->convs) {
    const size_t Mo = (size_t)nb * L.hout * L.wout;
    A(fmalloc(net, &L.w_fwd, (size_t)L.cout * L.kcp * kWopK, true));
    if (!L.stem) A(fmalloc(net, &L.w_dg, (size_t)L.cout * L.kc * kWopK, true));
    A(fmalloc(net, &L.y, Mo * L.cout, false));
    A(fmalloc(net, &L.dy, Mo * L.cout, false));
    if (!L.direct) {
      A(fmalloc(net, &L.col, Mo * L.kcp + 64, true));
      if (!L.stem && Mo * L.kc > dcol_max) dcol_max = Mo * L.kc;
    }
    const size_t rows = (Mo + 127) / 128 * 4;
    A(fmalloc(net, &L.stat_sum, rows * L.cout, true));
    A(fmalloc(net, &L.stat_sq, rows * L.cout, true));
    A(fmalloc(net, &L.slices, (size_t)kStatSlices * 2 * L.cout, true));
    A(fmalloc(net, &L.saved_mean, (size_t)L.cout, true));
    A(fmalloc(net, &L.saved_invstd, (size_t)L.cout, true));
    A(fmalloc(net, &L.coef, (size_t)3 * L.cout, true));
    A(fmalloc(net, &L.bwd_partials, (size_t)kBwdPartialRows * 2 * L.cout, true));
    const WgradGeom geo = wgrad_geometry(L.cout, L.kcp);
    const int tiles = geo.co_tiles * geo.groups;
    L.wg_splits_max = ctx->num_sms / tiles;
    if (L.wg_splits_max < 1) L.wg_splits_max = 1;
    A(fmalloc(net, &L.wg_ws, (size_t)L.wg_splits_max * wgrad_ws_elems(L.cout, L.kc), false));
  }
  A(fmalloc(net, &net->scratch_dcol, dcol_max + 64, true));
  {
    std::vector<WPrepEntry> tab;
    for (const FConv &L : net->convs) {
      WPrepEntry e{};
      e.w_off = L.w_off;
      e.w_fwd = L.w_fwd;
      e.w_dgrad = L.stem ? nullptr : L.w_dg;
      e.cout = L.cout;
      e.cin = L.cin;
      e.kc = L.kc;
      e.kcp = L.kcp;
      e.dgrad_mode = L.stem ? 0 : 2;  // transposed [kc][cout] for every conv of this runtime
      tab.push_back(e);
    }
    A(fmalloc(net, &net->wprep_table, tab.size(), false));
    if (cudaMemcpy(net->wprep_table, tab.data(), tab.size() * sizeof(WPrepEntry), cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("cudaMemcpy(wprep_table) failed");
      flatnet_destroy(net);
      return SALUN_ERR_CUDA;
    }
  }
  net->wg_splits.assign(net->convs.size(), 1);
  A(fmalloc(net, &net->wgred_table, net->convs.size(), false));
  if (cudaMallocHost(&net->wgred_host, net->convs.size() * sizeof(WgReduceEntry)) != cudaSuccess) {
    set_error("cudaMallocHost(wgred_host) failed");
    flatnet_destroy(net);
    return SALUN_ERR_CUDA;
  }
  A(fmalloc(net, &net->pooled, (size_t)nb * net->feat, true));
  A(fmalloc(net, &net->logits, (size_t)nb * cfg->num_classes, true));
  A(fmalloc(net, &net->dlogits, (size_t)nb * cfg->num_classes, true));
  A(fmalloc(net, &net->loss_ps, (size_t)nb, true));
  A(fmalloc(net, &net->dpooled, (size_t)nb * net->feat, true));
#undef A
  *out = net;
  return SALUN_OK;
}

int flatnet_forward_backward(FlatNet *net, const float *x, const int64_t *labels, int n, int train, float sign,
                             float *loss_dev, float *logits_dev, cudaStream_t st) {
  TRY(fforward(net, x, labels, n, train, sign, loss_dev, logits_dev, true, st));
  return fbackward(net, x, st);
}

int flatnet_forward(FlatNet *net, const float *x, int n, float *logits_dev, cudaStream_t st) {
  int rc = fforward(net, x, nullptr, n, 0, 1.f, nullptr, logits_dev, false, st);
  net->fwd_done = false;
  return rc;
}

}  // namespace salun
