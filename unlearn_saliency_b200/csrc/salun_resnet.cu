// salun_resnet.cu -- runtime of the BasicBlock ResNet (CIFAR stem) forward + backward on sm_100a.
//
// Replaces, for the SalUn hot path, model(image) and loss.backward() of
//   Classification/generate_mask.py:35-39   (eval-mode BN, loss = -CE)        and
//   Classification/unlearn/RL.py:128-132    (train-mode BN, CE)  (same in GA.py:113-117, FT.py:128-135)
// on the architecture of Classification/models/ResNet.py:180-322 (resnet18/34, imagenet=False stem).
//
// Data layout in HBM
//   parameters / gradients : one flat fp32 arena each, tensors in named_parameters() order; conv weights are kept
//                            "native" = [Cout][kh][kw][Cin] (OHWI) so that the bf16 GEMM operand is a plain cast and
//                            the wgrad kernel's red.add rows are contiguous.  The host mirror permutes at the boundary.
//   activations            : bf16 NHWC with a zero halo, [n][H+2][W+2][C]; the tcgen05 conv kernels read their 3x3 taps
//                            out of it with shifted 4-D TMA boxes (no im2col buffer for stride-1 convs).
//   raw conv outputs       : bf16 [n*H*W][C]; BatchNorm batch statistics come out of the GEMM epilogue (fp32
//                            accumulators) as per-tile partials and are reduced in double.
// Every kernel launch below is enqueued on the caller's stream; nothing synchronises.
#include <stdlib.h>

#include <map>
#include <vector>

#include "salun_elem.cuh"
#include "salun_gemm.cuh"
#include "salun_resnetb.cuh"

namespace salun {

struct ConvL {
  int cin, cout, ks, stride, hin, hout;
  bool stem;
  int kc, kcp;           // valid / padded reduction length (ks*ks*cin)
  int64_t w_off, g_off, b_off;  // arena offsets: weight, BN gamma, BN beta
  int rs_off;            // offset into the running-stat arenas
  int in_act;            // index of the padded input activation (-1: network input)
  wop_t *w_fwd, *w_dgrad;
  act_t *col, *y, *dy, *dcol;
  float *stat_sum, *stat_sq, *saved_mean, *saved_invstd, *coef, *bwd_partials;
  float *wg_ws;          // split-K workspace of the weight gradient: [wg_splits_max][cout][kc]
  int wg_splits_max;
  double *slices;
  bool dy_padded;
};
struct Act {
  int C, H;
  act_t *p;      // padded [n][H+2][H+2][C]
  act_t *dout;   // gradient w.r.t. this activation, flat [n*H*H][C]
  act_t *dz;     // dout * (act > 0), flat (identity-shortcut blocks only)
  uint8_t *rmask;  // 1-bit ReLU mask of this activation, [n*H*H][C/8] (read by the BatchNorm backward instead of p)
};
struct Block {
  int c1, c2, cd;          // conv indices (cd = -1: identity shortcut)
  int in_act, mid_act, out_act;
};
struct ConvMaps {
  CUtensorMap fwdA, fwdB, dgA, dgB, wgA, wgB;
  CUtensorMap rwA, rwB, rwdA, rwdB;  // k_conv_rw (resident weights) forward / dgrad maps
  bool rw_fwd, rw_dgrad;
  int pair_fwd, pair_dg;   // forward / dgrad through the CTA-pair kernel (fwdB / dgB encoded with bn/2 box rows)
};

}  // namespace salun

using namespace salun;

struct salun_resnet {
  salun::FlatNet *flat;  // non-null: Bottleneck runtime (salun_resnetb.cu); everything below is unused then
  salun_ctx *ctx;
  salun_resnet_cfg cfg;
  float *params, *grads, *rmean, *rvar;
  int64_t n_params;
  int n_bn_channels;
  std::vector<ConvL> convs;
  std::vector<Act> acts;
  std::vector<Block> blocks;
  int64_t fc_w_off, fc_b_off;
  int feat;  // channels of the last stage
  float *pooled, *logits, *dlogits, *loss_ps;
  WPrepEntry *wprep_table;
  WgReduceEntry *wgred_table;          // device copy, refreshed per backward (split counts depend on the batch size)
  WgReduceEntry *wgred_host;           // pinned staging
  // weight-gradient GEMMs only feed the final gradient: they run on a side stream, under the HBM-bound BatchNorm
  // backward kernels of the main chain (tensor-bound + memory-bound work co-resident on the SMs)
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_join;
  int use_side;
  std::vector<int> wg_splits_host;
  std::vector<int> bwd_fused_rows;     // per conv: partial rows written by a fused dgrad epilogue in this backward (0 = none)
  int use_bwd_fuse;
  int use_conv_rw;  // SALUN_CONV_RW: 0 = k_conv_gemm everywhere, 1 = k_conv_rw where supported, 2 = only 32x32 layers
  std::vector<void *> allocs;
  std::map<int, std::vector<ConvMaps>> plans;
  int last_n, last_train;
  bool fwd_done;
  // sync-BN (salun_resnet_enable_syncbn): per-BatchNorm statistics slots in NVLink peer-mapped memory
  bool syncbn;
  double *sb_sums[8];                 // every rank's slot arena
  unsigned long long *sb_flags[8];    // every rank's epoch flags [world]
  int sb_rank, sb_world;
  unsigned long long sb_epoch;
  std::vector<int64_t> sb_off;        // per conv: offset of its forward slot; the backward slot sits sb_half further
  int64_t sb_half;
  double *sb_count;                   // device [n_convs]: global pixel count of each BatchNorm (written by the exchange)
};

namespace salun {

static int stage_blocks(int depth, int s) {
  static const int r18[4] = {2, 2, 2, 2}, r34[4] = {3, 4, 6, 3};
  return depth == 34 ? r34[s] : r18[s];
}

// shared by create / param_count: walks the architecture in named_parameters() order
static int build_arch(const salun_resnet_cfg &c, std::vector<ConvL> *convs, std::vector<Act> *acts,
                      std::vector<Block> *blocks, int64_t *n_params, int *n_bn, int64_t *fcw, int64_t *fcb, int *feat) {
  if (!(c.depth == 18 || c.depth == 34)) {
    set_error("salun_resnet: depth %d not supported (BasicBlock nets 18/34)", c.depth);
    return SALUN_ERR_UNSUPPORTED;
  }
  if (c.image_size != 32 && c.image_size != 64) {
    set_error("salun_resnet: image_size %d not supported (power-of-two CIFAR-style stem: 32 or 64)", c.image_size);
    return SALUN_ERR_UNSUPPORTED;
  }
  int64_t off = 0;
  int rs = 0;
  auto add_conv = [&](int cin, int cout, int ks, int stride, int hin, bool stem, int in_act) {
    ConvL L{};
    L.cin = cin; L.cout = cout; L.ks = ks; L.stride = stride; L.hin = hin; L.hout = hin / stride; L.stem = stem;
    L.kc = ks * ks * cin;
    L.kcp = (L.kc + 63) / 64 * 64;
    L.w_off = off; off += (int64_t)cout * L.kc;
    L.g_off = off; off += cout;
    L.b_off = off; off += cout;
    L.rs_off = rs; rs += cout;
    L.in_act = in_act;
    L.dy_padded = (!stem && stride == 1);
    convs->push_back(L);
    return (int)convs->size() - 1;
  };
  auto add_act = [&](int C, int H) {
    Act a{};
    a.C = C; a.H = H;
    acts->push_back(a);
    return (int)acts->size() - 1;
  };
  int H = c.image_size;
  add_conv(3, 64, 3, 1, H, true, -1);
  int cur = add_act(64, H);
  int inpl = 64;
  const int planes[4] = {64, 128, 256, 512};
  for (int s = 0; s < 4; ++s) {
    for (int b = 0; b < stage_blocks(c.depth, s); ++b) {
      const int stride = (b == 0 && s > 0) ? 2 : 1;
      Block B{};
      B.in_act = cur;
      // named_parameters order inside a BasicBlock: conv1, bn1, conv2, bn2, downsample.0, downsample.1
      B.c1 = add_conv(inpl, planes[s], 3, stride, H, false, cur);
      const int Ho = H / stride;
      B.mid_act = add_act(planes[s], Ho);
      B.c2 = add_conv(planes[s], planes[s], 3, 1, Ho, false, B.mid_act);
      B.cd = -1;
      if (stride != 1 || inpl != planes[s]) B.cd = add_conv(inpl, planes[s], 1, stride, H, false, cur);
      B.out_act = add_act(planes[s], Ho);
      blocks->push_back(B);
      cur = B.out_act;
      inpl = planes[s];
      H = Ho;
    }
  }
  *feat = inpl;
  *fcw = off; off += (int64_t)c.num_classes * inpl;
  *fcb = off; off += c.num_classes;
  *n_params = off;
  *n_bn = rs;
  return SALUN_OK;
}

// N-tile of the conv GEMMs: the L2->SM port (64 B/clk) bounds the 128-wide tiles at ~50% of the tensor pipe; 256-wide
// tiles move 25% fewer operand bytes per MMA cycle and are used whenever they still give ~100+ CTAs
static int pick_bn(int N, int64_t M) {
  static int allow256 = -1;
  if (allow256 < 0) {
    const char *e = getenv("SALUN_BN256");
    allow256 = e ? atoi(e) : 1;
  }
  if (allow256 && N % 256 == 0 && ((M + 127) / 128) * (N / 256) >= 96) return 256;
  return N % 128 == 0 ? 128 : 64;
}

// CTA pairs (k_gemm2: 256 x bn tiles, each CTA pulls half of B) for the GEMMs that still give every pair a tile.
// SALUN_RESNET_PAIR: 0 = off (default), 1 = when >= 74 pair-tiles, 2 = when >= 37
static bool rn_use_pair(int64_t M, int N, int bn) {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("SALUN_RESNET_PAIR");
    on = e ? atoi(e) : 0;
  }
  if (!on || bn < 128) return false;
  const int64_t tiles = ((M + 255) / 256) * ((N + bn - 1) / bn);
  return tiles >= (on == 2 ? 37 : 74);
}

template <typename T>
static int dmalloc(salun_resnet *net, T **p, size_t count, bool zero) {
  void *q = nullptr;
  SALUN_CUDA_OK(cudaMalloc(&q, count * sizeof(T)));
  if (zero) SALUN_CUDA_OK(cudaMemset(q, 0, count * sizeof(T)));
  net->allocs.push_back(q);
  *p = (T *)q;
  return SALUN_OK;
}
#define TRY(expr)              \
  do {                         \
    int _rc = (expr);          \
    if (_rc) return _rc;       \
  } while (0)

static int build_plan(salun_resnet *net, int n, std::vector<ConvMaps> **out) {
  auto it = net->plans.find(n);
  if (it != net->plans.end()) {
    *out = &it->second;
    return SALUN_OK;
  }
  std::vector<ConvMaps> maps(net->convs.size());
  for (size_t i = 0; i < net->convs.size(); ++i) {
    const ConvL &L = net->convs[i];
    ConvMaps &m = maps[i];
    m.rw_fwd = m.rw_dgrad = false;
    const int64_t Mout = (int64_t)n * L.hout * L.hout;
    const int bn = pick_bn(L.cout, Mout);
    m.pair_fwd = m.pair_dg = 0;
    const bool rw_f = L.dy_padded && net->use_conv_rw && (net->use_conv_rw == 1 || L.hin == 32) && L.ks == 3 && conv_rw_supported(L.hin, L.cin, L.cout);
    if (!rw_f && rn_use_pair(Mout, L.cout, bn)) m.pair_fwd = 1;
    TRY(make_tmap_2d_wop(&m.fwdB, L.w_fwd, L.cout, L.kcp, m.pair_fwd ? bn / 2 : bn));
    TmapBox4 bx128, bx64;
    if (L.dy_padded) {
      // stride-1 3x3: forward A and wgrad B read the padded input activation, dgrad A / wgrad A the padded dY
      const Act &in = net->acts[L.in_act];
      TRY(conv_box(L.hin, L.hin, 128, &bx128));
      TRY(conv_box(L.hin, L.hin, 64, &bx64));
      TRY(make_tmap_4d_act(&m.fwdA, in.p, L.cin, L.hin + 2, L.hin + 2, n, bx128));
      TRY(make_tmap_4d_act(&m.dgA, L.dy, L.cout, L.hout + 2, L.hout + 2, n, bx128));
      const int bnd = pick_bn(L.cin, Mout);
      const bool rw_d = net->use_conv_rw && (net->use_conv_rw == 1 || L.hout == 32) && L.ks == 3 && conv_rw_supported(L.hout, L.cout, L.cin);
      if (!rw_d && !net->use_bwd_fuse && rn_use_pair(Mout, L.cin, bnd)) m.pair_dg = 1;
      TRY(make_tmap_2d_wop(&m.dgB, L.w_dgrad, L.cin, (uint64_t)L.ks * L.ks * L.cout, m.pair_dg ? bnd / 2 : bnd));
      TRY(make_tmap_4d_act(&m.wgA, L.dy, L.cout, L.hout + 2, L.hout + 2, n, bx64));
      TRY(make_tmap_4d_act(&m.wgB, in.p, L.cin, L.hin + 2, L.hin + 2, n, bx64));
      m.rw_fwd = net->use_conv_rw && (net->use_conv_rw == 1 || L.hin == 32) && L.ks == 3 && conv_rw_supported(L.hin, L.cin, L.cout);
      m.rw_dgrad = net->use_conv_rw && (net->use_conv_rw == 1 || L.hout == 32) && L.ks == 3 && conv_rw_supported(L.hout, L.cout, L.cin);
      TmapBox4 bxr{64, L.hin, 128 / L.hin + 2, 1};
      if (m.rw_fwd) {
        TRY(make_tmap_4d_act(&m.rwA, in.p, L.cin, L.hin + 2, L.hin + 2, n, bxr));
        TRY(make_tmap_2d_wop(&m.rwB, L.w_fwd, L.cout, L.kcp, 64));
      }
      if (m.rw_dgrad) {
        TRY(make_tmap_4d_act(&m.rwdA, L.dy, L.cout, L.hout + 2, L.hout + 2, n, bxr));
        TRY(make_tmap_2d_wop(&m.rwdB, L.w_dgrad, L.cin, (uint64_t)9 * L.cout, 64));
      }
    } else {
      // stem / stride-2: explicit patch matrix col[Mout][kcp]
      TRY(make_tmap_2d_act(&m.fwdA, L.col, Mout, L.kcp, 128));
      TRY(make_tmap_2d_act(&m.wgA, L.dy, Mout, L.cout, 64));
      TRY(make_tmap_2d_act(&m.wgB, L.col, Mout, L.kcp, 64));
      if (!L.stem) {
        // dgrad: dcol[Mout][kc] = dY[Mout][Cout] . Wt[kc][Cout]^T
        TRY(make_tmap_2d_act(&m.dgA, L.dy, Mout, L.cout, 128));
        const int bnd = pick_bn(L.kc, Mout);
        TRY(make_tmap_2d_wop(&m.dgB, L.w_dgrad, L.kc, L.cout, bnd));
      }
    }
  }
  auto res = net->plans.emplace(n, std::move(maps));
  *out = &res.first->second;
  return SALUN_OK;
}

static SyncBnPeers sb_peers(const salun_resnet *net, int ci, bool backward) {
  SyncBnPeers p{};
  p.rank = net->sb_rank;
  p.world = net->sb_world;
  for (int r = 0; r < net->sb_world; ++r) {
    p.slot[r] = net->sb_sums[r] + net->sb_off[ci] + (backward ? net->sb_half : 0);
    p.flags[r] = net->sb_flags[r];
  }
  return p;
}
// batch statistics of a sharded batch: sum the ranks' (sum x, sum x^2, count) before the BatchNorm apply reads them
static void sync_stats(salun_resnet *net, const ConvL &L, int M, cudaStream_t st) {
  if (!net->syncbn) return;
  const int ci = (int)(&L - net->convs.data());
  launch_syncbn_exchange(sb_peers(net, ci, false), ++net->sb_epoch, 0, L.cout, L.slices, (double)M, net->sb_count + ci, nullptr, st);
}

static int conv_forward(salun_resnet *net, const ConvL &L, const ConvMaps &m, int n, int train, cudaStream_t st) {
  const int M = n * L.hout * L.hout;
  ConvGemmArgs a{};
  a.M = M;
  a.N = L.cout;
  a.out_bf16 = L.y;
  a.ld_out = L.cout;
  if (train) {
    a.stat_sum = L.stat_sum;
    a.stat_sq = L.stat_sq;
  }
  const int bn = pick_bn(L.cout, M);
  if (m.rw_fwd) {
    ConvRwArgs r{};
    r.H = r.W = L.hout;
    r.cin_blocks = L.cin / 64;
    r.num_tiles = (M + 127) / 128;
    r.M = M;
    r.N = L.cout;
    r.out_bf16 = L.y;
    r.ld_out = L.cout;
    r.stat_sum = a.stat_sum;
    r.stat_sq = a.stat_sq;
    TRY(launch_conv_rw(m.rwA, m.rwB, r, net->ctx->num_sms, st));
    if (train) {
      launch_bn_stats_reduce(L.stat_sum, L.stat_sq, (M + 127) / 128 * 4, L.cout, L.slices, st);
      sync_stats(net, L, M, st);
    }
    return SALUN_OK;
  }
  if (L.dy_padded) {
    a.mode_a = 1;
    a.cin_blocks = L.cin / 64;
    a.num_k_blocks = L.ks * L.ks * a.cin_blocks;
    a.kw = L.ks;
    a.tap_y0 = a.tap_x0 = L.ks == 3 ? 0 : 1;
    a.H = a.W = L.hout;
  } else {
    a.mode_a = 0;
    a.num_k_blocks = L.kcp / 64;
  }
  a.pair = m.pair_fwd;
  TRY(launch_conv_gemm(m.fwdA, m.fwdB, a, bn, st));
  if (train) {
    launch_bn_stats_reduce(L.stat_sum, L.stat_sq, (M + 127) / 128 * 4, L.cout, L.slices, st);
    sync_stats(net, L, M, st);
  }
  return SALUN_OK;
}

static BnFwd bn_of(salun_resnet *net, const ConvL &L) {
  BnFwd b{};
  b.count_dev = net->syncbn ? net->sb_count + (&L - net->convs.data()) : nullptr;
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

static int wgrad_conv(salun_resnet *net, const ConvL &L, const ConvMaps &m, int n, cudaStream_t main_st) {
  cudaStream_t st = main_st;
  if (net->use_side) {  // fork: everything enqueued on the main stream so far (dY, activations) precedes this wgrad
    SALUN_CUDA_OK(cudaEventRecord(net->ev_fork, main_st));
    SALUN_CUDA_OK(cudaStreamWaitEvent(net->side, net->ev_fork, 0));
    st = net->side;
  }
  const int64_t M = (int64_t)n * L.hout * L.hout;
  WgradArgs a{};
  a.mode_a = L.dy_padded ? 1 : 0;
  a.mode_b = L.dy_padded ? 1 : 0;
  a.kb_total = (int)((M + 63) / 64);
  a.cin_blocks = L.cin >= 64 ? L.cin / 64 : 1;
  a.kw = L.ks;
  a.tap_y0 = a.tap_x0 = L.ks == 3 ? 0 : 1;
  a.H = a.W = L.hout;
  const WgradGeom geo = wgrad_geometry(L.cout, L.kcp);
  a.total_blocks = geo.total_blocks;
  a.n_blocks = geo.n_blocks;
  a.Cout = L.cout;
  a.ldw = L.kc;
  a.kvalid = L.kc;
  const int co_tiles = geo.co_tiles, groups = geo.groups;
  // one wave: as many pixel splits as fit on the SMs next to the (co tile, tap group) decomposition
  int splits = net->ctx->num_sms / (co_tiles * groups);
  if (splits < 1) splits = 1;
  if (splits > L.wg_splits_max) splits = L.wg_splits_max;
  if (splits > a.kb_total) splits = a.kb_total;
  a.kb_per_split = (a.kb_total + splits - 1) / splits;
  splits = (a.kb_total + a.kb_per_split - 1) / a.kb_per_split;
  a.dw = L.wg_ws;                                   // split s stores its partial into slab s ...
  a.split_stride = (long long)L.cout * L.kc;
  net->wg_splits_host[&L - net->convs.data()] = splits;  // ... and launch_wgrad_reduce sums the slabs into the grad arena
  return launch_wgrad(m.wgA, m.wgB, a, co_tiles, groups, splits, st);
}

// BN backward of one conv's BatchNorm: dout (flat) [* relu mask of out_act] -> L.dy (+ dz)
static void bn_backward(salun_resnet *net, const ConvL &L, const act_t *dout, const uint8_t *relu_act, act_t *dz, int n,
                        int train, cudaStream_t st) {
  const int H = L.hout;
  const int ci = (int)(&L - net->convs.data());
  const int fused_rows = net->bwd_fused_rows[ci];  // > 0: the dgrad GEMM that produced `dout` already reduced dZ, dZ*xhat
  if (fused_rows == 0)
    launch_bn_bwd_reduce(dout, relu_act, L.y, L.saved_mean, L.saved_invstd, L.bwd_partials, n, H, H, L.cout, st);
  const bool sync = net->syncbn && train;
  SyncBnPeers sp{};
  if (sync) sp = sb_peers(net, ci, true);
  launch_bn_bwd_finalize(L.bwd_partials, fused_rows, L.cout, net->params + L.g_off, L.saved_invstd, (float)(n * H * H),
                         train, net->grads + L.g_off, net->grads + L.b_off, L.coef, st, sync ? sp.slot[sp.rank] : nullptr);
  // sync-BN: dgamma / dbeta stay this rank's sums (the data-parallel gradient average combines them); the batch means
  // inside dX are over the GLOBAL batch
  if (sync) launch_syncbn_exchange(sp, ++net->sb_epoch, 1, L.cout, nullptr, (double)(n * H * H), nullptr, L.coef, st);
  launch_bn_bwd_apply(dout, relu_act, L.y, L.saved_mean, L.saved_invstd, L.coef, L.dy, L.dy_padded ? 1 : 0, dz, n, H, H,
                      L.cout, st);
}

static int prep_weights(salun_resnet *net, bool need_dgrad, cudaStream_t st) {
  launch_prep_w_all(net->wprep_table, (int)net->convs.size(), net->params, need_dgrad ? 1 : 0, st);
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

static int forward_impl(salun_resnet *net, const float *x, const int64_t *labels, int n, int train, float sign,
                        float *loss_dev, float *logits_out, bool need_bwd, cudaStream_t st) {
  std::vector<ConvMaps> *plan;
  TRY(build_plan(net, n, &plan));
  TRY(prep_weights(net, need_bwd, st));
  const salun_resnet_cfg &c = net->cfg;
  const float inv_std[3] = {1.f / c.std[0], 1.f / c.std[1], 1.f / c.std[2]};
  // stem
  {
    const ConvL &L = net->convs[0];
    launch_stem_im2col(x, L.col, n, L.hin, L.hin, c.mean, inv_std, st);
    TRY(conv_forward(net, L, (*plan)[0], n, train, st));
    BnFwd b = bn_of(net, L);
    launch_bn_apply(b, nullptr, nullptr, net->acts[0].p, need_bwd ? net->acts[0].rmask : nullptr, n, L.hout, L.hout, L.cout, 1, train, c.bn_eps, c.bn_momentum,
                    st);
  }
  for (const Block &B : net->blocks) {
    const ConvL &L1 = net->convs[B.c1], &L2 = net->convs[B.c2];
    const Act &in = net->acts[B.in_act], &mid = net->acts[B.mid_act], &out = net->acts[B.out_act];
    if (!L1.dy_padded) launch_im2col_s2(in.p, L1.col, n, L1.hin, L1.hin, L1.cin, 3, st);
    TRY(conv_forward(net, L1, (*plan)[B.c1], n, train, st));
    BnFwd b1 = bn_of(net, L1);
    launch_bn_apply(b1, nullptr, nullptr, mid.p, need_bwd ? mid.rmask : nullptr, n, L1.hout, L1.hout, L1.cout, 1, train, c.bn_eps, c.bn_momentum, st);
    TRY(conv_forward(net, L2, (*plan)[B.c2], n, train, st));
    BnFwd b2 = bn_of(net, L2);
    if (B.cd >= 0) {
      const ConvL &Ld = net->convs[B.cd];
      launch_im2col_s2(in.p, Ld.col, n, Ld.hin, Ld.hin, Ld.cin, 1, st);
      TRY(conv_forward(net, Ld, (*plan)[B.cd], n, train, st));
      BnFwd bd = bn_of(net, Ld);
      launch_bn_apply(b2, &bd, nullptr, out.p, need_bwd ? out.rmask : nullptr, n, L2.hout, L2.hout, L2.cout, 1, train, c.bn_eps, c.bn_momentum, st);
    } else {
      launch_bn_apply(b2, nullptr, in.p, out.p, need_bwd ? out.rmask : nullptr, n, L2.hout, L2.hout, L2.cout, 1, train, c.bn_eps, c.bn_momentum, st);
    }
  }
  const Act &last = net->acts[net->blocks.back().out_act];
  launch_avgpool(last.p, net->pooled, n, last.H, last.H, last.C, st);
  launch_fc_ce(net->pooled, net->params + net->fc_w_off, net->params + net->fc_b_off, labels,
               logits_out ? logits_out : net->logits, net->dlogits, net->loss_ps, n, net->feat, c.num_classes, sign, st);
  if (labels && loss_dev) launch_loss_sum(net->loss_ps, n, sign, loss_dev, st);
  SALUN_CUDA_OK(cudaGetLastError());
  net->last_n = n;
  net->last_train = train;
  net->fwd_done = true;
  return SALUN_OK;
}

static int backward_impl(salun_resnet *net, cudaStream_t st) {
  if (!net->fwd_done) {
    set_error("salun_resnet backward called before forward");
    return SALUN_ERR_STATE;
  }
  const int n = net->last_n, train = net->last_train;
  std::vector<ConvMaps> *plan;
  TRY(build_plan(net, n, &plan));
  net->bwd_fused_rows.assign(net->convs.size(), 0);
  // consumer BatchNorm(s) of the gradient a dgrad GEMM produces: fills f1/f2 and marks them as reduced
  auto fuse_for = [&](int conv_idx, const Act &act, BnBwdFuse *f, int rows) {
    if (!net->use_bwd_fuse) return;
    ConvL &C = net->convs[conv_idx];
    f->act = act.p;
    f->y = C.y;
    f->mean = C.saved_mean;
    f->invstd = C.saved_invstd;
    f->partials = C.bwd_partials;
    net->bwd_fused_rows[conv_idx] = rows;
  };
  // every gradient element has exactly one writer (BN finalize, FC backward, wgrad reduce): no zero-fill needed
  const Act &last = net->acts[net->blocks.back().out_act];
  launch_fc_bwd(net->pooled, net->dlogits, net->params + net->fc_w_off, net->grads + net->fc_w_off,
                net->grads + net->fc_b_off, last.dout, n, net->feat, net->cfg.num_classes, last.H * last.H, st);
  for (int bi = (int)net->blocks.size() - 1; bi >= 0; --bi) {
    const Block &B = net->blocks[bi];
    const ConvL &L1 = net->convs[B.c1], &L2 = net->convs[B.c2];
    const Act &in = net->acts[B.in_act], &mid = net->acts[B.mid_act], &out = net->acts[B.out_act];
    const bool identity = B.cd < 0;
    // out = relu(bn2(y2) + shortcut):  dZ = dOut * (out > 0)
    bn_backward(net, L2, out.dout, out.rmask, identity ? out.dz : nullptr, n, train, st);
    if (!identity) bn_backward(net, net->convs[B.cd], out.dout, out.rmask, nullptr, n, train, st);
    // conv2: dgrad -> d(mid), wgrad
    if ((*plan)[B.c2].rw_dgrad) {
      ConvRwArgs r{};
      r.H = r.W = L2.hout;
      r.cin_blocks = L2.cout / 64;
      r.M = n * L2.hout * L2.hout;
      r.num_tiles = (r.M + 127) / 128;
      r.N = L2.cin;
      r.out_bf16 = mid.dout;
      r.ld_out = L2.cin;
      fuse_for(B.c1, mid, &r.f1, r.num_tiles * 4);  // mid = relu(bn1(y1))
      TRY(launch_conv_rw((*plan)[B.c2].rwdA, (*plan)[B.c2].rwdB, r, net->ctx->num_sms, st));
      TRY(wgrad_conv(net, L2, (*plan)[B.c2], n, st));
    } else {
      ConvGemmArgs a{};
      a.mode_a = 1;
      a.cin_blocks = L2.cout / 64;
      a.num_k_blocks = 9 * a.cin_blocks;
      a.kw = 3;
      a.H = a.W = L2.hout;
      a.M = n * L2.hout * L2.hout;
      a.N = L2.cin;
      a.out_bf16 = mid.dout;
      a.ld_out = L2.cin;
      a.fH = a.fW = L2.hout;
      fuse_for(B.c1, mid, &a.f1, (a.M + 127) / 128 * 4);
      a.pair = (*plan)[B.c2].pair_dg;
      TRY(launch_conv_gemm((*plan)[B.c2].dgA, (*plan)[B.c2].dgB, a, pick_bn(L2.cin, a.M), st));
      TRY(wgrad_conv(net, L2, (*plan)[B.c2], n, st));
    }
    // mid = relu(bn1(y1))
    bn_backward(net, L1, mid.dout, mid.rmask, nullptr, n, train, st);
    if (L1.dy_padded && (*plan)[B.c1].rw_dgrad) {
      ConvRwArgs r{};
      r.H = r.W = L1.hout;
      r.cin_blocks = L1.cout / 64;
      r.M = n * L1.hout * L1.hout;
      r.num_tiles = (r.M + 127) / 128;
      r.N = L1.cin;
      r.out_bf16 = in.dout;
      r.ld_out = L1.cin;
      r.addend = identity ? out.dz : nullptr;
      if (bi > 0) {  // in = relu(bn2(y2) + shortcut) of the previous block
        const Block &P = net->blocks[bi - 1];
        fuse_for(P.c2, in, &r.f1, r.num_tiles * 4);
        if (P.cd >= 0) fuse_for(P.cd, in, &r.f2, r.num_tiles * 4);
      } else {
        fuse_for(0, in, &r.f1, r.num_tiles * 4);  // acts[0] = relu(bn(stem conv))
      }
      TRY(launch_conv_rw((*plan)[B.c1].rwdA, (*plan)[B.c1].rwdB, r, net->ctx->num_sms, st));
      if (!identity) {
        set_error("salun_resnet: stride-1 block with projection shortcut is not supported");
        return SALUN_ERR_UNSUPPORTED;
      }
    } else if (L1.dy_padded) {
      ConvGemmArgs a{};
      a.mode_a = 1;
      a.cin_blocks = L1.cout / 64;
      a.num_k_blocks = 9 * a.cin_blocks;
      a.kw = 3;
      a.H = a.W = L1.hout;
      a.M = n * L1.hout * L1.hout;
      a.N = L1.cin;
      a.out_bf16 = in.dout;
      a.ld_out = L1.cin;
      a.addend = identity ? out.dz : nullptr;  // gradient of the identity shortcut
      a.fH = a.fW = L1.hout;
      {
        const int rows = (a.M + 127) / 128 * 4;
        if (bi > 0) {
          const Block &P = net->blocks[bi - 1];
          fuse_for(P.c2, in, &a.f1, rows);
          if (P.cd >= 0) fuse_for(P.cd, in, &a.f2, rows);
        } else {
          fuse_for(0, in, &a.f1, rows);
        }
      }
      a.pair = (*plan)[B.c1].pair_dg;
      TRY(launch_conv_gemm((*plan)[B.c1].dgA, (*plan)[B.c1].dgB, a, pick_bn(L1.cin, a.M), st));
      if (!identity) {
        set_error("salun_resnet: stride-1 block with projection shortcut is not supported");
        return SALUN_ERR_UNSUPPORTED;
      }
    } else {
      // stride 2: dcol = dY . W ; then col2im gathers the 3x3 taps and the 1x1 projection shortcut
      const ConvL &Ld = net->convs[B.cd];
      const int Mo = n * L1.hout * L1.hout;
      ConvGemmArgs a{};
      a.mode_a = 0;
      a.num_k_blocks = L1.cout / 64;
      a.M = Mo;
      a.N = L1.kc;
      a.out_bf16 = L1.dcol;
      a.ld_out = L1.kc;
      TRY(launch_conv_gemm((*plan)[B.c1].dgA, (*plan)[B.c1].dgB, a, pick_bn(L1.kc, a.M), st));
      ConvGemmArgs d{};
      d.mode_a = 0;
      d.num_k_blocks = Ld.cout / 64;
      d.M = Mo;
      d.N = Ld.kc;
      d.out_bf16 = Ld.dcol;
      d.ld_out = Ld.kc;
      TRY(launch_conv_gemm((*plan)[B.cd].dgA, (*plan)[B.cd].dgB, d, pick_bn(Ld.kc, d.M), st));
      launch_col2im_s2(L1.dcol, Ld.dcol, in.dout, n, L1.hin, L1.hin, L1.cin, st);
      TRY(wgrad_conv(net, Ld, (*plan)[B.cd], n, st));
    }
    TRY(wgrad_conv(net, L1, (*plan)[B.c1], n, st));
  }
  // stem: act0 = relu(bn(y0)); no dgrad (the input needs no gradient)
  {
    const ConvL &L = net->convs[0];
    bn_backward(net, L, net->acts[0].dout, net->acts[0].rmask, nullptr, n, train, st);
    TRY(wgrad_conv(net, L, (*plan)[0], n, st));
  }
  if (net->use_side) {  // join: the reduction below consumes every wgrad workspace
    SALUN_CUDA_OK(cudaEventRecord(net->ev_join, net->side));
    SALUN_CUDA_OK(cudaStreamWaitEvent(st, net->ev_join, 0));
  }
  // deterministic split-K reduction of every conv weight gradient into the flat gradient arena (one launch)
  for (size_t i = 0; i < net->convs.size(); ++i) {
    const ConvL &L = net->convs[i];
    WgReduceEntry &e = net->wgred_host[i];
    e.ws = L.wg_ws;
    e.dst_off = L.w_off;
    e.count = (long long)L.cout * L.kc;
    e.splits = net->wg_splits_host[i];
    e.kc = L.kc;
  }
  SALUN_CUDA_OK(cudaMemcpyAsync(net->wgred_table, net->wgred_host, net->convs.size() * sizeof(WgReduceEntry),
                                cudaMemcpyHostToDevice, st));
  launch_wgrad_reduce(net->wgred_table, (int)net->convs.size(), net->grads, st);
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

}  // namespace salun

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int64_t salun_resnet_param_count(const salun_resnet_cfg *cfg) {
  if (!cfg) return -1;
  if (cfg->depth >= 50) return flatnet_param_count(cfg, nullptr);
  std::vector<ConvL> c;
  std::vector<Act> a;
  std::vector<Block> b;
  int64_t n, fw, fb;
  int nb, feat;
  if (build_arch(*cfg, &c, &a, &b, &n, &nb, &fw, &fb, &feat)) return -1;
  return n;
}

int64_t salun_resnet_bn_channels(const salun_resnet_cfg *cfg) {
  if (!cfg) return -1;
  if (cfg->depth >= 50) {
    int64_t nbn = -1;
    return flatnet_param_count(cfg, &nbn) < 0 ? -1 : nbn;
  }
  std::vector<ConvL> c;
  std::vector<Act> a;
  std::vector<Block> b;
  int64_t n, fw, fb;
  int nb, feat;
  if (build_arch(*cfg, &c, &a, &b, &n, &nb, &fw, &fb, &feat)) return -1;
  return nb;
}

int salun_resnet_create(salun_ctx *ctx, const salun_resnet_cfg *cfg, float *params, float *grads, float *running_mean,
                        float *running_var, salun_resnet **out) {
  SALUN_REQUIRE(ctx && cfg && params && grads && running_mean && running_var && out, "NULL argument");
  SALUN_REQUIRE(cfg->max_batch > 0 && cfg->num_classes > 0, "max_batch and num_classes must be positive");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  salun_resnet *net = new salun_resnet();
  net->flat = nullptr;
  if (cfg->depth >= 50) {
    if (!(cfg->depth == 50 || cfg->depth == 101 || cfg->depth == 152)) {
      delete net;
      set_error("salun_resnet: depth %d not supported", cfg->depth);
      return SALUN_ERR_UNSUPPORTED;
    }
    net->ctx = ctx;
    net->cfg = *cfg;
    int rcf = flatnet_create(ctx, cfg, params, grads, running_mean, running_var, &net->flat);
    if (rcf) {
      delete net;
      return rcf;
    }
    *out = net;
    return SALUN_OK;
  }
  net->ctx = ctx;
  net->cfg = *cfg;
  net->params = params;
  net->grads = grads;
  net->rmean = running_mean;
  net->rvar = running_var;
  net->fwd_done = false;
  net->syncbn = false;
  net->sb_epoch = 0;
  net->sb_count = nullptr;
  net->wgred_host = nullptr;
  net->side = nullptr;
  net->ev_fork = net->ev_join = nullptr;
  {
    const char *e = getenv("SALUN_WGRAD_SIDE_STREAM");
    net->use_side = e ? atoi(e) : 1;
    if (net->use_side) {
      if (cudaStreamCreateWithFlags(&net->side, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&net->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&net->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        set_error("salun_resnet: could not create the side stream / events");
        delete net;
        return SALUN_ERR_CUDA;
      }
    }
  }
  {
    // measured on B200 (profiles/README.md): the dgrad epilogues are already the slower side of their kernels, so
    // moving the BatchNorm-backward reduction into them costs more (-6% steps/s) than the 16 reduce launches it saves
    const char *e = getenv("SALUN_BN_BWD_FUSE");
    net->use_bwd_fuse = (e && !kSplit) ? atoi(e) : 0;
  }
  {
    const char *e = getenv("SALUN_CONV_RW");
    net->use_conv_rw = e ? atoi(e) : 1;
  }
  int rc = build_arch(*cfg, &net->convs, &net->acts, &net->blocks, &net->n_params, &net->n_bn_channels, &net->fc_w_off,
                      &net->fc_b_off, &net->feat);
  if (rc) {
    delete net;
    return rc;
  }
  const int64_t nb = (cfg->max_batch + 7) / 8 * 8;  // TMA boxes cover up to 8 images of the 4x4 stage
#define A(expr)                     \
  do {                              \
    int _rc = (expr);               \
    if (_rc) {                      \
      salun_resnet_destroy(net);    \
      return _rc;                   \
    }                               \
  } while (0)
  for (Act &a : net->acts) {
    const size_t padded = (size_t)nb * (a.H + 2) * (a.H + 2) * a.C;
    A(dmalloc(net, &a.p, padded, true));
    A(dmalloc(net, &a.dout, (size_t)nb * a.H * a.H * a.C, false));
    A(dmalloc(net, &a.dz, (size_t)nb * a.H * a.H * a.C, false));
    A(dmalloc(net, &a.rmask, (size_t)nb * a.H * a.H * a.C / 8, true));
  }
  for (ConvL &L : net->convs) {
    const size_t Mo = (size_t)nb * L.hout * L.hout;
    A(dmalloc(net, &L.w_fwd, (size_t)L.cout * L.kcp * kWopK, true));
    if (!L.stem) A(dmalloc(net, &L.w_dgrad, (size_t)L.cout * L.kc * kWopK, true));
    A(dmalloc(net, &L.y, Mo * L.cout, false));
    if (L.dy_padded) {
      A(dmalloc(net, &L.dy, (size_t)nb * (L.hout + 2) * (L.hout + 2) * L.cout, true));
    } else {
      A(dmalloc(net, &L.dy, Mo * L.cout, false));
      A(dmalloc(net, &L.col, Mo * L.kcp, true));
      if (!L.stem) A(dmalloc(net, &L.dcol, Mo * L.kc, false));
    }
    const size_t rows = (Mo + 255) / 256 * 8;   // whole CTA pairs: the pair kernel's second CTA owns a (possibly empty) tile
    A(dmalloc(net, &L.stat_sum, rows * L.cout, true));
    A(dmalloc(net, &L.stat_sq, rows * L.cout, true));
    A(dmalloc(net, &L.slices, (size_t)kStatSlices * 2 * L.cout, true));
    A(dmalloc(net, &L.saved_mean, (size_t)L.cout, true));
    A(dmalloc(net, &L.saved_invstd, (size_t)L.cout, true));
    A(dmalloc(net, &L.coef, (size_t)3 * L.cout, true));
    {
      size_t prow = (Mo + 127) / 128 * 4;  // fused dgrad epilogues write one partial row per (tile, warp)
      if (prow < (size_t)kBwdPartialRows) prow = kBwdPartialRows;
      A(dmalloc(net, &L.bwd_partials, prow * 2 * L.cout, true));
    }
    {
      const WgradGeom geo = wgrad_geometry(L.cout, L.kcp);
      const int tiles = geo.co_tiles * geo.groups;
      L.wg_splits_max = ctx->num_sms / tiles;
      if (L.wg_splits_max < 1) L.wg_splits_max = 1;
      A(dmalloc(net, &L.wg_ws, (size_t)L.wg_splits_max * wgrad_ws_elems(L.cout, L.kc), false));
    }
  }
  net->wg_splits_host.assign(net->convs.size(), 1);
  A(dmalloc(net, &net->wgred_table, net->convs.size(), false));
  if (cudaMallocHost(&net->wgred_host, net->convs.size() * sizeof(WgReduceEntry)) != cudaSuccess) {
    set_error("cudaMallocHost(wgred_host) failed");
    salun_resnet_destroy(net);
    return SALUN_ERR_CUDA;
  }
  {
    std::vector<WPrepEntry> tab;
    for (const ConvL &L : net->convs) {
      WPrepEntry e{};
      e.w_off = L.w_off;
      e.w_fwd = L.w_fwd;
      e.w_dgrad = L.stem ? nullptr : L.w_dgrad;
      e.cout = L.cout;
      e.cin = L.cin;
      e.kc = L.kc;
      e.kcp = L.kcp;
      e.dgrad_mode = L.stem ? 0 : (L.dy_padded ? 1 : 2);
      tab.push_back(e);
    }
    A(dmalloc(net, &net->wprep_table, tab.size(), false));
    cudaError_t ce = cudaMemcpy(net->wprep_table, tab.data(), tab.size() * sizeof(WPrepEntry), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
      set_error("cudaMemcpy(wprep_table) failed: %s", cudaGetErrorString(ce));
      salun_resnet_destroy(net);
      return SALUN_ERR_CUDA;
    }
  }
  A(dmalloc(net, &net->pooled, (size_t)nb * net->feat, true));
  A(dmalloc(net, &net->logits, (size_t)nb * cfg->num_classes, true));
  A(dmalloc(net, &net->dlogits, (size_t)nb * cfg->num_classes, true));
  A(dmalloc(net, &net->loss_ps, (size_t)nb, true));
#undef A
  *out = net;
  return SALUN_OK;
}

static int64_t syncbn_layout(const std::vector<ConvL> &convs, std::vector<int64_t> *off) {
  int64_t o = 0;
  for (const ConvL &L : convs) {
    if (off) off->push_back(o);
    o += 2 * (int64_t)L.cout + 2;
  }
  return o;  // one direction; the arena holds forward and backward slots: 2 * o doubles
}

int64_t salun_resnet_syncbn_doubles(const salun_resnet_cfg *cfg) {
  if (!cfg || cfg->depth >= 50) return -1;
  std::vector<ConvL> c;
  std::vector<Act> a;
  std::vector<Block> b;
  int64_t n, fw, fb;
  int nb, feat;
  if (build_arch(*cfg, &c, &a, &b, &n, &nb, &fw, &fb, &feat)) return -1;
  return 2 * syncbn_layout(c, nullptr);
}

int salun_resnet_enable_syncbn(salun_resnet *net, double *const *peer_sums_host, unsigned long long *const *peer_flags_host,
                               int rank, int world) {
  SALUN_REQUIRE(net && peer_sums_host && peer_flags_host, "NULL argument");
  SALUN_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "rank / world out of range (<= 8 ranks)");
  if (net->flat) {
    set_error("salun_resnet_enable_syncbn: served for the BasicBlock runtime (resnet18 / 34)");
    return SALUN_ERR_UNSUPPORTED;
  }
  SALUN_CUDA_OK(cudaSetDevice(net->ctx->device));
  net->sb_off.clear();
  net->sb_half = syncbn_layout(net->convs, &net->sb_off);
  for (int r = 0; r < world; ++r) {
    SALUN_REQUIRE(peer_sums_host[r] && peer_flags_host[r], "NULL peer buffer");
    net->sb_sums[r] = peer_sums_host[r];
    net->sb_flags[r] = peer_flags_host[r];
  }
  net->sb_rank = rank;
  net->sb_world = world;
  if (!net->sb_count) {
    void *q = nullptr;
    SALUN_CUDA_OK(cudaMalloc(&q, net->convs.size() * sizeof(double)));
    SALUN_CUDA_OK(cudaMemset(q, 0, net->convs.size() * sizeof(double)));
    net->allocs.push_back(q);
    net->sb_count = (double *)q;
  }
  net->sb_epoch = 0;
  net->syncbn = true;
  return SALUN_OK;
}

int salun_resnet_destroy(salun_resnet *net) {
  if (!net) return SALUN_OK;
  if (net->flat) {
    flatnet_destroy(net->flat);
    delete net;
    return SALUN_OK;
  }
  cudaSetDevice(net->ctx->device);
  for (void *p : net->allocs) cudaFree(p);
  if (net->wgred_host) cudaFreeHost(net->wgred_host);
  if (net->side) cudaStreamDestroy(net->side);
  if (net->ev_fork) cudaEventDestroy(net->ev_fork);
  if (net->ev_join) cudaEventDestroy(net->ev_join);
  delete net;
  return SALUN_OK;
}

int salun_resnet_forward_backward(salun_resnet *net, const float *x, const int64_t *labels, int n, int train,
                                  float loss_sign, float *loss_dev, float *logits_dev, void *stream) {
  SALUN_REQUIRE(net && x && labels, "NULL argument");
  SALUN_REQUIRE(n > 0 && n <= net->cfg.max_batch, "batch size out of range");
  SALUN_CUDA_OK(cudaSetDevice(net->ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (net->flat) return flatnet_forward_backward(net->flat, x, labels, n, train, loss_sign, loss_dev, logits_dev, st);
  TRY(forward_impl(net, x, labels, n, train, loss_sign, loss_dev, logits_dev, true, st));
  return backward_impl(net, st);
}

int salun_resnet_forward(salun_resnet *net, const float *x, int n, float *logits_dev, void *stream) {
  SALUN_REQUIRE(net && x && logits_dev, "NULL argument");
  SALUN_REQUIRE(n > 0 && n <= net->cfg.max_batch, "batch size out of range");
  SALUN_CUDA_OK(cudaSetDevice(net->ctx->device));
  if (net->flat) return flatnet_forward(net->flat, x, n, logits_dev, (cudaStream_t)stream);
  int rc = forward_impl(net, x, nullptr, n, 0, 1.f, nullptr, logits_dev, false, (cudaStream_t)stream);
  net->fwd_done = false;
  return rc;
}

}  // extern "C"
