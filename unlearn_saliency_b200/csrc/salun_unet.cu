// salun_unet.cu -- runtime of the class-conditional DDPM U-Net forward + backward on sm_100a.
//
// Replaces, for the DDPM SalUn hot path, the model(x_t, t, c, ...) calls and loss.backward() of
//   DDPM/runners/diffusion.py:959-996 (generate_mask: two eval-mode passes, CFG-combined)   and
//   DDPM/runners/diffusion.py:519-593 (saliency_unlearn: remain / forget / pseudo-label passes)
// on the architecture of DDPM/models/diffusion.py:195-413 (Conditional_Model; ResnetBlock :82-145, AttnBlock :148-192,
// Downsample / Upsample :49-79, get_timestep_embedding :17-35), config DDPM/configs/cifar10_saliency_unlearn.yml.
//
// The network is compiled once into a tape of ops over activation tensors; forward walks the tape, backward walks it in
// reverse.  GroupNorm has no batch coupling, so the caller may concatenate the remain / forget batches (and the two CFG
// passes of mask generation) into ONE batch with per-sample dL/d(eps).
//
// Data layout in HBM
//   parameters / gradients : one flat fp32 arena each, tensors in named_parameters() order (null_classes_emb first, up.0
//                            before up.3); conv weights are kept [Cout][kh][kw][Cin] (OHWI), Linear weights as in PyTorch.
//   activations            : bf16 NHWC with a zero halo [n][H+2][H+2][C] ("P"); the 3x3 and 1x1 convolutions read them
//                            with shifted 4-D TMA boxes.  GroupNorm outputs feeding attention, q / k / v, the attention
//                            output and all GroupNorm-output gradients are flat [n*H*H][C] ("F").
//   convolution            : k_conv_gemm_p (tcgen05, TMEM, TMA) with bias / temb-projection / residual fused in its
//                            epilogue; dgrad is the same kernel on the padded dY; wgrad = k_wgrad on a side stream.
//   attention              : S = Q K^T, P = softmax, O = P V and their five backward GEMMs as BATCHED k_conv_gemm_p
//                            launches (one B matrix per group of Te rows); 4x4 / 8x8 images share a 128-row group
//                            through a block-diagonal softmax mask.
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "salun_elem.cuh"
#include "salun_gemm.cuh"
#include "salun_unet_elem.cuh"

namespace salun {

typedef act_t bf16;  // activation element of this build (salun_act.cuh)

struct UTensor {
  std::string name;
  int C, H;
  bool vflat, gflat;   // layout of the value / of the gradient
  bool need_g;
  bf16 *v, *g;
  bool g_live;         // backward: the gradient buffer already holds a contribution
  bool want_stats;     // a GroupNorm reads this tensor (directly or through a concat): its producer's GEMM epilogue
  float *st_sum, *st_sq;  // emits the per-(32-row group) column sums [nb*H*H/32][C] the GroupNorm statistics need
};
enum ConvKind { CK_S1 = 0, CK_IN = 1, CK_DOWN = 2, CK_OUT = 3 };
struct UConv {
  int kind, cin, cout, ks, H;      // H: output side
  int64_t w_off, b_off, pb_off;    // pb_off: bias of the temb_cemb_proj Linear that shares this conv's bias gradient (-1: none)
  int in, out, addend, rb_col;     // tensor ids; rb_col: column of this block's projection in RB (-1: none)
  int kc, kcp, cout_p;
  wop_t *w_fwd, *w_dgrad;
  bf16 *col, *dcol;
  float *yf;                       // CK_OUT: fp32 [M][64] GEMM output
  bf16 *dy64;                      // CK_OUT: padded 64-channel dY
  float *wg_ws;
  int wg_splits_max;
  float *bias_partial;             // [rows][cout_p] per-sample, per-slice column sums of dY (CK_OUT: [nb][3])
};
struct UGn {
  int in, out, C, H, swish, dropout;
  int64_t g_off, b_off;
  float *stats;      // [nb][32][2]
  float *persample;  // [nb][2][C] backward: per-sample sums of dyh, dyh*xhat
  uint32_t id;
};
struct UAttn {
  int q, k, v, o, T, Te, C;
  bf16 *P;  // [nb*T][Te] softmax probabilities (kept for the backward pass)
};
struct RbParams {   // one ResnetBlock's parameter offsets
  int cin, cout;
  int64_t n1w, n1b, c1w, c1b, pw, pb, n2w, n2b, c2w, c2b, ninw, ninb;
  int rb_col;
};
struct AtParams {
  int C;
  int64_t nw, nb, qw, qb, kw, kb, vw, vb, pw, pb;
};
enum OpType { OP_CONV, OP_GN, OP_CONCAT, OP_UP, OP_ATTN };
struct UOp {
  int type, idx;       // idx into convs / gns / attns; CONCAT: a, b, out tensors; UP: in, out
  int a, b, c;
};
struct UConvMaps {
  CUtensorMap fwdA, fwdB, dgA, dgB, wgA, wgB;
  int pair_fwd, pair_dg;  // forward / dgrad GEMM through the CTA-pair kernel (fwdB / dgB encoded with bn/2 box rows)
};
struct UAttnMaps {
  CUtensorMap q, k, v, p, xt_b, tt_a, og, ds;  // see build_plan
  int bn_s, bn_o;
};
struct UPlan {
  std::vector<UConvMaps> conv;
  std::vector<UAttnMaps> attn;
  CUtensorMap e_a, wcat_b, drb_a, wcatT_b, drbt_a, et_b;  // projection GEMMs (forward, dE, dW)
  int bn_rb, bn_de, bn_dw;
  SumEntry *sum_table;  // device: the cross-sample sums of one backward pass (GroupNorm dgamma/dbeta, conv biases)
  int n_sums;
};

}  // namespace salun

using namespace salun;

struct salun_unet {
  salun_ctx *ctx;
  salun_unet_cfg cfg;
  float *params, *grads, *gscratch;
  int64_t n_params;
  int nb;  // allocated batch (max_batch rounded up to 8)
  std::vector<UTensor> ts;
  std::vector<UConv> convs;
  std::vector<UGn> gns;
  std::vector<UAttn> attns;
  std::vector<UOp> ops;
  std::vector<RbParams> rbs;
  // embedding path
  int64_t null_off, t0w, t0b, t1w, t1b, cew, c0w, c0b, c1w, c1b;
  int emb;      // 4 * ch
  int ld_rb;    // sum of the ResnetBlock output widths
  float *sincos, *ce, *pre_t, *h_t, *pre_c, *h_c, *cat, *E, *RB, *dRB, *dE, *dcat, *dh, *dpre, *dce;
  // all temb_cemb_proj Linears as one tensor-core GEMM (operands staged per step)
  bf16 *wcat_act, *E_bf, *dRB_bf, *dRBt_bf;   // activation-side (A) operands; wcat_act: the gathered weights
  wop_t *wcat, *wcatT, *Et_bf;                // weight-side (B) operands (bf16 build: wcat aliases wcat_act)
  float *bcat, *dWcat;
  long long *row_w, *row_b;
  int nb32;
  float *t_dev;
  int64_t *c_dev;
  uint8_t *drop_dev;
  bool have_drop;
  // scratch
  float *gn_partial, *S_f32;
  bf16 *tt, *dS;
  wop_t *xt1, *wpk;   // B operands of the attention GEMMs: transposes, and (split build) the packed K / V matrix
  WPrepEntry *wprep_table;
  WgReduceEntry *wgred_table, *wgred_host;
  std::vector<int> wg_splits;
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_join;
  std::vector<void *> allocs;
  std::map<int, UPlan> plans;
  int last_n;
  bool fwd_saved;
  float drop_p;
  uint32_t seed;
};

namespace salun {

#define TRY(expr)              \
  do {                         \
    int _rc = (expr);          \
    if (_rc) return _rc;       \
  } while (0)

static int ilog2(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return s;
}
static int u_pick_bn(int N, int64_t M) {
  if (N % 256 == 0 && ((M + 127) / 128) * (N / 256) >= 96) return 256;
  return N % 128 == 0 ? 128 : 64;
}
// CTA pairs (cta_group::2) halve the B-operand bytes each SM pulls through its L2 port -- the measured bound of the
// 128 x bn tiles -- and are used when a GEMM still gives every pair at least two 256-row tiles.  SALUN_UNET_PAIR=0: off, 2: always.
static bool use_pair(int64_t M, int N, int bn) {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("SALUN_UNET_PAIR");
    on = e ? atoi(e) : 1;
  }
  if (!on || bn < 128) return false;
  if (on == 2) return true;  // tests: force the pair kernel on small shapes too
  return ((M + 255) / 256) * ((N + bn - 1) / bn) >= 148;
}
static bool has_res(const salun_unet_cfg &c, int res) {
  for (int i = 0; i < c.n_attn_res; ++i)
    if (c.attn_res[i] == res) return true;
  return false;
}

// ------------------------------------------------------------------------------------------------------------------
// architecture -> parameter offsets (named_parameters() order) + tape (execution order)
// ------------------------------------------------------------------------------------------------------------------
struct ArchBuilder {
  salun_unet *net;
  int add_tensor(const std::string &name, int C, int H, bool vflat, bool gflat, bool need_g = true) {
    UTensor t{};
    t.name = name; t.C = C; t.H = H; t.vflat = vflat; t.gflat = gflat; t.need_g = need_g;
    net->ts.push_back(t);
    return (int)net->ts.size() - 1;
  }
  int add_conv(int kind, int cin, int cout, int ks, int H, int64_t w_off, int64_t b_off, int in, int out, int addend = -1,
               int rb_col = -1, int64_t pb_off = -1) {
    UConv L{};
    L.kind = kind; L.cin = cin; L.cout = cout; L.ks = ks; L.H = H; L.w_off = w_off; L.b_off = b_off; L.pb_off = pb_off;
    L.in = in; L.out = out; L.addend = addend; L.rb_col = rb_col;
    L.kc = ks * ks * cin;
    L.kcp = (L.kc + 63) / 64 * 64;
    L.cout_p = (cout + 63) / 64 * 64;
    net->convs.push_back(L);
    net->ops.push_back(UOp{OP_CONV, (int)net->convs.size() - 1, 0, 0, 0});
    return (int)net->convs.size() - 1;
  }
  void add_gn(int in, int out, int C, int H, int swish, int dropout, int64_t g_off, int64_t b_off) {
    UGn g{};
    g.in = in; g.out = out; g.C = C; g.H = H; g.swish = swish; g.dropout = dropout; g.g_off = g_off; g.b_off = b_off;
    g.id = (uint32_t)net->gns.size();
    net->gns.push_back(g);
    net->ops.push_back(UOp{OP_GN, (int)net->gns.size() - 1, 0, 0, 0});
  }
  // ResnetBlock (diffusion.py:124-145)
  int resblock(const std::string &nm, int x, const RbParams &p, int H) {
    const int a1 = add_tensor(nm + ".a1", p.cin, H, false, true);
    add_gn(x, a1, p.cin, H, 1, 0, p.n1w, p.n1b);
    const int h1 = add_tensor(nm + ".h1", p.cout, H, false, false);
    add_conv(CK_S1, p.cin, p.cout, 3, H, p.c1w, p.c1b, a1, h1, -1, p.rb_col, p.pb);
    const int a2 = add_tensor(nm + ".a2", p.cout, H, false, true);
    add_gn(h1, a2, p.cout, H, 1, 1, p.n2w, p.n2b);
    int sc = x;
    if (p.cin != p.cout) {
      sc = add_tensor(nm + ".sc", p.cout, H, false, false);
      add_conv(CK_S1, p.cin, p.cout, 1, H, p.ninw, p.ninb, x, sc);
    }
    const int out = add_tensor(nm + ".out", p.cout, H, false, false);
    add_conv(CK_S1, p.cout, p.cout, 3, H, p.c2w, p.c2b, a2, out, sc);
    return out;
  }
  // AttnBlock (diffusion.py:167-192)
  int attn(const std::string &nm, int x, const AtParams &p, int H) {
    const int C = p.C;
    const int xn = add_tensor(nm + ".xn", C, H, true, true);
    add_gn(x, xn, C, H, 0, 0, p.nw, p.nb);
    const int q = add_tensor(nm + ".q", C, H, true, true), k = add_tensor(nm + ".k", C, H, true, true),
              v = add_tensor(nm + ".v", C, H, true, true);
    add_conv(CK_S1, C, C, 1, H, p.qw, p.qb, xn, q);
    add_conv(CK_S1, C, C, 1, H, p.kw, p.kb, xn, k);
    add_conv(CK_S1, C, C, 1, H, p.vw, p.vb, xn, v);
    const int o = add_tensor(nm + ".o", C, H, true, true);
    UAttn A{};
    A.q = q; A.k = k; A.v = v; A.o = o; A.C = C; A.T = H * H; A.Te = A.T < 128 ? 128 : A.T;
    net->attns.push_back(A);
    net->ops.push_back(UOp{OP_ATTN, (int)net->attns.size() - 1, 0, 0, 0});
    const int out = add_tensor(nm + ".out", C, H, false, false);
    add_conv(CK_S1, C, C, 1, H, p.pw, p.pb, o, out, x);
    return out;
  }
};

static int build_arch(salun_unet *net) {
  const salun_unet_cfg &c = net->cfg;
  if (c.ch != 128) {
    set_error("salun_unet: ch = %d; the reference's ResnetBlock hard-codes cemb_channels = 512 (diffusion.py:93), so only ch = 128 is consistent", c.ch);
    return SALUN_ERR_UNSUPPORTED;
  }
  if (c.n_levels < 1 || c.n_levels > 8 || c.num_res_blocks < 1 || c.in_channels != 3 || c.out_ch != 3) {
    set_error("salun_unet: unsupported config (levels %d, res blocks %d, in %d, out %d)", c.n_levels, c.num_res_blocks,
              c.in_channels, c.out_ch);
    return SALUN_ERR_UNSUPPORTED;
  }
  const int S = c.image_size;
  if (S < 8 || S > 64 || (S & (S - 1)) || (S >> (c.n_levels - 1)) < 4) {
    set_error("salun_unet: image_size %d with %d levels not supported (power of two, smallest level >= 4x4)", S, c.n_levels);
    return SALUN_ERR_UNSUPPORTED;
  }
  for (int i = 0; i < c.n_attn_res; ++i)
    if (c.attn_res[i] > 16) {
      set_error("salun_unet: attention at %dx%d not supported (<= 16x16)", c.attn_res[i], c.attn_res[i]);
      return SALUN_ERR_UNSUPPORTED;
    }
  const int ch = c.ch, L = c.n_levels, nrb = c.num_res_blocks;
  net->emb = 4 * ch;
  auto mult = [&](int l) { return c.ch_mult[l]; };
  // ---- channel plan in EXECUTION order ----
  struct Lvl { std::vector<RbParams> blocks; std::vector<AtParams> attns; int res; int64_t rs_w, rs_b; int rs_c; };
  std::vector<Lvl> down(L), up(L);
  RbParams mid1{}, mid2{};
  AtParams mida{};
  int res = S, cur = ch;
  std::vector<int> skip_ch;
  skip_ch.push_back(ch);
  for (int l = 0; l < L; ++l) {
    down[l].res = res;
    for (int i = 0; i < nrb; ++i) {
      RbParams p{};
      p.cin = cur; p.cout = ch * mult(l);
      cur = p.cout;
      down[l].blocks.push_back(p);
      if (has_res(c, res)) { AtParams a{}; a.C = cur; down[l].attns.push_back(a); }
      skip_ch.push_back(cur);
    }
    if (l != L - 1) { down[l].rs_c = cur; skip_ch.push_back(cur); res /= 2; }
  }
  mid1.cin = mid1.cout = mid2.cin = mid2.cout = cur;
  mida.C = cur;
  for (int l = L - 1; l >= 0; --l) {
    up[l].res = res;
    for (int i = 0; i < nrb + 1; ++i) {
      RbParams p{};
      const int sk = skip_ch.back();
      skip_ch.pop_back();
      p.cin = cur + sk; p.cout = ch * mult(l);
      cur = p.cout;
      up[l].blocks.push_back(p);
      if (has_res(c, res)) { AtParams a{}; a.C = cur; up[l].attns.push_back(a); }
    }
    if (l != 0) { up[l].rs_c = cur; res *= 2; }
  }
  const int c_last = cur;
  // ---- parameter offsets in NAMING order ----
  int64_t off = 0;
  int rb_col = 0;
  auto take = [&](int64_t n) { int64_t o = off; off += n; return o; };
  auto name_rb = [&](RbParams &p) {
    p.n1w = take(p.cin); p.n1b = take(p.cin);
    p.c1w = take((int64_t)p.cout * 9 * p.cin); p.c1b = take(p.cout);
    p.pw = take((int64_t)p.cout * (net->emb + 512)); p.pb = take(p.cout);
    p.n2w = take(p.cout); p.n2b = take(p.cout);
    p.c2w = take((int64_t)p.cout * 9 * p.cout); p.c2b = take(p.cout);
    p.ninw = p.ninb = -1;
    if (p.cin != p.cout) { p.ninw = take((int64_t)p.cout * p.cin); p.ninb = take(p.cout); }
    p.rb_col = rb_col;
    rb_col += p.cout;
  };
  auto name_at = [&](AtParams &a) {
    const int64_t C = a.C;
    a.nw = take(C); a.nb = take(C);
    a.qw = take(C * C); a.qb = take(C);
    a.kw = take(C * C); a.kb = take(C);
    a.vw = take(C * C); a.vb = take(C);
    a.pw = take(C * C); a.pb = take(C);
  };
  net->null_off = take(ch);
  net->t0w = take((int64_t)net->emb * ch); net->t0b = take(net->emb);
  net->t1w = take((int64_t)net->emb * net->emb); net->t1b = take(net->emb);
  net->cew = take((int64_t)c.n_classes * ch);
  net->c0w = take((int64_t)net->emb * ch); net->c0b = take(net->emb);
  net->c1w = take((int64_t)net->emb * net->emb); net->c1b = take(net->emb);
  const int64_t cin_w = take((int64_t)ch * 27), cin_b = take(ch);
  for (int l = 0; l < L; ++l) {
    for (auto &p : down[l].blocks) name_rb(p);
    for (auto &a : down[l].attns) name_at(a);
    if (l != L - 1) { down[l].rs_w = take((int64_t)down[l].rs_c * 9 * down[l].rs_c); down[l].rs_b = take(down[l].rs_c); }
  }
  name_rb(mid1);
  name_at(mida);
  name_rb(mid2);
  for (int l = 0; l < L; ++l) {
    for (auto &p : up[l].blocks) name_rb(p);
    for (auto &a : up[l].attns) name_at(a);
    if (l != 0) { up[l].rs_w = take((int64_t)up[l].rs_c * 9 * up[l].rs_c); up[l].rs_b = take(up[l].rs_c); }
  }
  const int64_t no_w = take(c_last), no_b = take(c_last);
  const int64_t co_w = take((int64_t)3 * 9 * c_last), co_b = take(3);
  net->n_params = off;
  net->ld_rb = rb_col;
  // ---- tape in EXECUTION order ----
  ArchBuilder B{net};
  std::vector<int> hs;
  const int t0 = B.add_tensor("conv_in", ch, S, false, false);
  B.add_conv(CK_IN, 3, ch, 3, S, cin_w, cin_b, -1, t0);
  hs.push_back(t0);
  char nm[96];
  for (int l = 0; l < L; ++l) {
    const int H = down[l].res;
    for (int i = 0; i < nrb; ++i) {
      snprintf(nm, sizeof nm, "down.%d.block.%d", l, i);
      int h = B.resblock(nm, hs.back(), down[l].blocks[i], H);
      net->rbs.push_back(down[l].blocks[i]);
      if (!down[l].attns.empty()) {
        snprintf(nm, sizeof nm, "down.%d.attn.%d", l, i);
        h = B.attn(nm, h, down[l].attns[i], H);
      }
      hs.push_back(h);
    }
    if (l != L - 1) {
      snprintf(nm, sizeof nm, "down.%d.downsample", l);
      const int C = down[l].rs_c;
      const int d = B.add_tensor(nm, C, H / 2, false, false);
      B.add_conv(CK_DOWN, C, C, 3, H / 2, down[l].rs_w, down[l].rs_b, hs.back(), d);
      hs.push_back(d);
    }
  }
  const int Hm = down[L - 1].res;
  int h = B.resblock("mid.block_1", hs.back(), mid1, Hm);
  net->rbs.push_back(mid1);
  h = B.attn("mid.attn_1", h, mida, Hm);
  h = B.resblock("mid.block_2", h, mid2, Hm);
  net->rbs.push_back(mid2);
  for (int l = L - 1; l >= 0; --l) {
    const int H = up[l].res;
    for (int i = 0; i < nrb + 1; ++i) {
      const int sk = hs.back();
      hs.pop_back();
      snprintf(nm, sizeof nm, "up.%d.block.%d", l, i);
      const int Ca = net->ts[h].C, Cb = net->ts[sk].C;
      const int cat = B.add_tensor(std::string(nm) + ".cat", Ca + Cb, H, false, false);
      net->ops.push_back(UOp{OP_CONCAT, 0, h, sk, cat});
      h = B.resblock(nm, cat, up[l].blocks[i], H);
      net->rbs.push_back(up[l].blocks[i]);
      if (!up[l].attns.empty()) {
        snprintf(nm, sizeof nm, "up.%d.attn.%d", l, i);
        h = B.attn(nm, h, up[l].attns[i], H);
      }
    }
    if (l != 0) {
      snprintf(nm, sizeof nm, "up.%d.upsample", l);
      const int C = up[l].rs_c;
      const int u = B.add_tensor(std::string(nm) + ".near", C, 2 * H, false, false);
      net->ops.push_back(UOp{OP_UP, 0, h, u, 0});
      const int o = B.add_tensor(nm, C, 2 * H, false, false);
      B.add_conv(CK_S1, C, C, 3, 2 * H, up[l].rs_w, up[l].rs_b, u, o);
      h = o;
    }
  }
  const int an = B.add_tensor("norm_out", c_last, S, false, true);
  B.add_gn(h, an, c_last, S, 1, 0, no_w, no_b);
  B.add_conv(CK_OUT, c_last, 3, 3, S, co_w, co_b, an, -1);
  // GroupNorm statistics out of the producers' epilogues: walk the tape backwards so that a concat that feeds a
  // GroupNorm marks its two sources (32-row groups must not straddle samples: H*H >= 32)
  {
    const char *e = getenv("SALUN_UNET_EPILOGUE_STATS");
    const bool on = e ? atoi(e) != 0 : true;
    for (int oi = (int)net->ops.size() - 1; on && oi >= 0; --oi) {
      const UOp &op = net->ops[oi];
      if (op.type == OP_GN) {
        UTensor &t = net->ts[net->gns[op.idx].in];
        if (t.H * t.H >= 32) t.want_stats = true;
      } else if (op.type == OP_CONCAT && net->ts[op.c].want_stats) {
        net->ts[op.a].want_stats = net->ts[op.b].want_stats = true;
      }
    }
  }
  return SALUN_OK;
}

template <typename T>
static int dmalloc(salun_unet *net, T **p, size_t count, bool zero = true) {
  void *q = nullptr;
  if (count == 0) count = 1;
  SALUN_CUDA_OK(cudaMalloc(&q, count * sizeof(T)));
  if (zero) SALUN_CUDA_OK(cudaMemset(q, 0, count * sizeof(T)));
  net->allocs.push_back(q);
  *p = (T *)q;
  return SALUN_OK;
}
static size_t tensor_elems(const salun_unet *net, const UTensor &t, bool flat) {
  return flat ? (size_t)net->nb * t.H * t.H * t.C : (size_t)net->nb * (t.H + 2) * (t.H + 2) * t.C;
}

// ------------------------------------------------------------------------------------------------------------------
// tensor maps per batch size
// ------------------------------------------------------------------------------------------------------------------
static int map_act(CUtensorMap *m, const bf16 *base, bool flat, int C, int H, int n, int pixels, int box_rows_flat) {
  if (flat) return make_tmap_2d_act(m, base, (uint64_t)n * H * H, C, box_rows_flat);
  TmapBox4 bx;
  TRY(conv_box(H, H, pixels, &bx));
  return make_tmap_4d_act(m, base, C, H + 2, H + 2, n, bx);
}
static int build_plan(salun_unet *net, int n, UPlan **out) {
  auto it = net->plans.find(n);
  if (it != net->plans.end()) {
    *out = &it->second;
    return SALUN_OK;
  }
  UPlan plan;
  plan.conv.resize(net->convs.size());
  plan.attn.resize(net->attns.size());
  for (size_t i = 0; i < net->convs.size(); ++i) {
    const UConv &L = net->convs[i];
    UConvMaps &m = plan.conv[i];
    const int64_t M = (int64_t)n * L.H * L.H;
    const int bn = u_pick_bn(L.cout_p, M);
    m.pair_fwd = use_pair(M, L.cout_p, bn) ? 1 : 0;
    m.pair_dg = 0;
    TRY(make_tmap_2d_wop(&m.fwdB, L.w_fwd, L.cout_p, L.kcp, m.pair_fwd ? bn / 2 : bn));
    if (L.kind == CK_S1 || L.kind == CK_OUT) {
      const UTensor &in = net->ts[L.in];
      TRY(map_act(&m.fwdA, in.v, in.vflat, L.cin, L.H, n, 128, 128));
      TRY(map_act(&m.wgB, in.v, in.vflat, L.cin, L.H, n, 64, 64));
      const bf16 *dy = L.kind == CK_OUT ? L.dy64 : net->ts[L.out].g;
      const bool dyflat = L.kind == CK_OUT ? false : net->ts[L.out].gflat;
      TRY(map_act(&m.dgA, dy, dyflat, L.cout_p, L.H, n, 128, 128));
      TRY(map_act(&m.wgA, dy, dyflat, L.cout_p, L.H, n, 64, 64));
      const int bnd = u_pick_bn(L.cin, M);
      m.pair_dg = use_pair(M, L.cin, bnd) ? 1 : 0;
      TRY(make_tmap_2d_wop(&m.dgB, L.w_dgrad, L.cin, (uint64_t)L.ks * L.ks * L.cout_p, m.pair_dg ? bnd / 2 : bnd));
    } else {
      // conv_in / downsample: explicit patch matrix col[M][kcp]
      TRY(make_tmap_2d_act(&m.fwdA, L.col, M, L.kcp, 128));
      TRY(make_tmap_2d_act(&m.wgB, L.col, M, L.kcp, 64));
      const UTensor &o = net->ts[L.out];
      TRY(map_act(&m.wgA, o.g, false, L.cout, L.H, n, 64, 64));
      if (L.kind == CK_DOWN) {
        TRY(map_act(&m.dgA, o.g, false, L.cout, L.H, n, 128, 128));
        const int bnd = u_pick_bn(L.kc, M);
        TRY(make_tmap_2d_wop(&m.dgB, L.w_dgrad, L.kc, L.cout, bnd));
      }
    }
  }
  for (size_t i = 0; i < net->attns.size(); ++i) {
    const UAttn &A = net->attns[i];
    UAttnMaps &m = plan.attn[i];
    const uint64_t M = (uint64_t)n * A.T;
    const uint64_t G = (M + A.Te - 1) / A.Te;
    m.bn_s = u_pick_bn(A.Te, M);
    m.bn_o = u_pick_bn(A.C, M);
    // A operands (box 128 rows) and B operands (box bn rows) over the same matrices
    // B operands that are themselves activations (K, V): the bf16 build reads them in place, the split build reads the
    // packed copy launch_pack_wop() leaves in net->wpk
    const wop_t *kb = kSplit ? net->wpk : reinterpret_cast<const wop_t *>(net->ts[A.k].v);
    const wop_t *vb = kSplit ? net->wpk : reinterpret_cast<const wop_t *>(net->ts[A.v].v);
    TRY(make_tmap_2d_act(&m.q, net->ts[A.q].v, M, A.C, 128));           // A of S = Q K^T
    TRY(make_tmap_2d_wop(&m.k, kb, G * A.Te, A.C, m.bn_s));             // B of S
    TRY(make_tmap_2d_wop(&m.v, vb, G * A.Te, A.C, m.bn_s));             // B of dP = dO V^T
    TRY(make_tmap_2d_act(&m.p, A.P, M, A.Te, 128));                     // A of O = P Vt^T
    TRY(make_tmap_2d_wop(&m.xt_b, net->xt1, G * A.C, A.Te, m.bn_o));    // B = a [C][Te] transpose (Vt, dOt, Kt, Qt)
    TRY(make_tmap_2d_act(&m.tt_a, net->tt, G * A.Te, A.Te, 128));       // A = a [Te][Te] transpose (Pt, dSt)
    TRY(make_tmap_2d_act(&m.og, net->ts[A.o].g, M, A.C, 128));          // A of dP
    TRY(make_tmap_2d_act(&m.ds, net->dS, M, A.Te, 128));                // A of dQ = dS Kt^T
  }
  {
    const int E8 = net->emb + 512, R = net->ld_rb;
    plan.bn_rb = u_pick_bn(R, n);
    plan.bn_de = u_pick_bn(E8, n);
    plan.bn_dw = u_pick_bn(E8, R);
    TRY(make_tmap_2d_act(&plan.e_a, net->E_bf, n, E8, 128));
    TRY(make_tmap_2d_wop(&plan.wcat_b, net->wcat, R, E8, plan.bn_rb));
    TRY(make_tmap_2d_act(&plan.drb_a, net->dRB_bf, n, R, 128));
    TRY(make_tmap_2d_wop(&plan.wcatT_b, net->wcatT, E8, R, plan.bn_de));
    // K = the batch: columns >= n of the A operand are zero-filled by TMA (the buffers keep stale rows of larger batches).
    // Split build: the B rows span the whole pitch ([dup(hi) | dup(lo)] of logical length nb32) and K runs over nb32.
    TRY(make_tmap_2d_bf16_ld(&plan.drbt_a, net->dRBt_bf, R, (uint64_t)n * kActK, (uint64_t)net->nb32 * kActK, 128, 64));
    if (kSplit)
      TRY(make_tmap_2d_wop(&plan.et_b, net->Et_bf, E8, net->nb32, plan.bn_dw));
    else
      TRY(make_tmap_2d_bf16_ld(&plan.et_b, net->Et_bf, E8, n, net->nb32, plan.bn_dw, 64));
  }
  {
    std::vector<SumEntry> tab;
    for (const UGn &g : net->gns) tab.push_back(SumEntry{g.persample, 2LL * g.C, n, 2, g.C, g.b_off, -1, g.g_off});
    for (const UConv &L : net->convs) {
      if (L.kind == CK_OUT) {
        tab.push_back(SumEntry{L.bias_partial, 3, n, 1, 3, L.b_off, -1, -1});
      } else if (L.rb_col >= 0) {  // per-sample sums live in dRB (they are also the projection's output gradient)
        tab.push_back(SumEntry{net->dRB + L.rb_col, net->ld_rb, n, 1, L.cout, L.b_off, L.pb_off, -1});
      } else {
        tab.push_back(SumEntry{L.bias_partial, L.cout, n * unet_slices_for(L.H, n), 1, L.cout, L.b_off, -1, -1});
      }
    }
    plan.n_sums = (int)tab.size();
    void *q = nullptr;
    SALUN_CUDA_OK(cudaMalloc(&q, tab.size() * sizeof(SumEntry)));
    net->allocs.push_back(q);
    plan.sum_table = (SumEntry *)q;
    SALUN_CUDA_OK(cudaMemcpy(q, tab.data(), tab.size() * sizeof(SumEntry), cudaMemcpyHostToDevice));
  }
  auto res = net->plans.emplace(n, std::move(plan));
  *out = &res.first->second;
  return SALUN_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
static void conv_geom(ConvGemmArgs &a, int mode_a, int C_k, int ks, int H) {
  a.mode_a = mode_a;
  if (mode_a == 1) {
    a.cin_blocks = C_k / 64;
    a.num_k_blocks = ks * ks * a.cin_blocks;
    a.kw = ks;
    a.tap_y0 = a.tap_x0 = ks == 3 ? 0 : 1;
    a.H = a.W = H;
  } else {
    a.num_k_blocks = C_k / 64;  // C_k = full reduction length
  }
}
static int conv_forward(salun_unet *net, const UConv &L, const UConvMaps &m, int n, const float *x_in, float *eps_out,
                        cudaStream_t st) {
  const int M = n * L.H * L.H;
  ConvGemmArgs a{};
  a.M = M;
  a.N = L.cout_p;
  a.fH = a.fW = L.H;
  if (L.kind == CK_IN) {
    const float zero3[3] = {0.f, 0.f, 0.f}, one3[3] = {1.f, 1.f, 1.f};
    launch_stem_im2col(x_in, L.col, n, L.H, L.H, zero3, one3, st);
    conv_geom(a, 0, L.kcp, 1, L.H);
  } else if (L.kind == CK_DOWN) {
    launch_down_im2col(net->ts[L.in].v, L.col, n, 2 * L.H, L.cin, st);
    conv_geom(a, 0, L.kcp, 1, L.H);
  } else {
    const UTensor &in = net->ts[L.in];
    if (in.vflat)
      conv_geom(a, 0, L.cin, 1, L.H);
    else
      conv_geom(a, 1, L.cin, L.ks, L.H);
  }
  if (L.kind == CK_OUT) {
    a.out_f32 = L.yf;
    a.ld_out = 64;
  } else {
    const UTensor &o = net->ts[L.out];
    a.out_bf16 = o.v;
    a.ld_out = L.cout;
    a.out_pad = o.vflat ? 0 : 1;
    a.bias = net->params + L.b_off;
    if (L.rb_col >= 0) {
      a.rowbias = net->RB + L.rb_col;
      a.rb_ld = net->ld_rb;
      a.rb_shift = ilog2(L.H * L.H);
    }
    if (L.addend >= 0) a.addend = net->ts[L.addend].v;  // same layout / width as the output
    if (o.want_stats) {
      a.stat_sum = o.st_sum;
      a.stat_sq = o.st_sq;
    }
  }
  a.pair = m.pair_fwd;
  TRY(launch_conv_gemm(m.fwdA, m.fwdB, a, u_pick_bn(L.cout_p, M), st));
  if (L.kind == CK_OUT) launch_eps_out(L.yf, net->params + L.b_off, eps_out, n, L.H, st);
  return SALUN_OK;
}

static int gemm_plain(const CUtensorMap &A, const CUtensorMap &B, int M, int N, int K, bf16 *out_bf16, float *out_f32,
                      int ld_out, int batch_a, int batch_b, int bn, cudaStream_t st, const float *bias = nullptr) {
  ConvGemmArgs a{};
  a.mode_a = 0;
  a.num_k_blocks = (K + 63) / 64;
  a.bias = bias;
  a.M = M;
  a.N = N;
  a.out_bf16 = out_bf16;
  a.out_f32 = out_f32;
  a.ld_out = ld_out;
  a.batch_rows_a = batch_a;
  a.batch_rows_b = batch_b;
  return launch_conv_gemm(A, B, a, bn, st);
}

static int attn_forward(salun_unet *net, const UAttn &A, const UAttnMaps &m, int n, cudaStream_t st) {
  const int M = n * A.T, G = (M + A.Te - 1) / A.Te;
  const float scale = 1.f / sqrtf((float)A.C);
  // S = Q K^T (per group), fp32
  launch_pack_wop(net->ts[A.k].v, net->wpk, (long long)G * A.Te, A.C, st);
  TRY(gemm_plain(m.q, m.k, M, A.Te, A.C, nullptr, net->S_f32, A.Te, A.Te, A.Te, m.bn_s, st));
  launch_softmax(net->S_f32, A.P, M, A.Te, A.T, scale, st);
  // O = P V : B operand = V^T per group
  launch_transpose_wop(net->ts[A.v].v, A.C, net->xt1, A.Te, A.C, G, st);
  TRY(gemm_plain(m.p, m.xt_b, M, A.C, A.Te, net->ts[A.o].v, nullptr, A.C, A.Te, A.C, m.bn_o, st));
  return SALUN_OK;
}

static int emb_forward(salun_unet *net, const UPlan &plan, int n, bool save, cudaStream_t st) {
  const int ch = net->cfg.ch, E4 = net->emb, E8 = net->emb + 512;
  const float *P = net->params;
  launch_emb_inputs(net->t_dev, net->c_dev, net->have_drop ? net->drop_dev : nullptr, P + net->cew, P + net->null_off,
                    net->sincos, net->ce, n, ch, net->cfg.n_classes, st);
  launch_sgemm(net->sincos, ch, 1, P + net->t0w, 1, ch, net->pre_t, E4, n, E4, ch, P + net->t0b, 0, st);
  launch_swish_f32(net->pre_t, net->h_t, (long long)n * E4, st);
  launch_sgemm(net->h_t, E4, 1, P + net->t1w, 1, E4, net->cat, E8, n, E4, E4, P + net->t1b, 0, st);
  launch_sgemm(net->ce, ch, 1, P + net->c0w, 1, ch, net->pre_c, E4, n, E4, ch, P + net->c0b, 0, st);
  launch_swish_f32(net->pre_c, net->h_c, (long long)n * E4, st);
  launch_sgemm(net->h_c, E4, 1, P + net->c1w, 1, E4, net->cat + E4, E8, n, E4, E4, P + net->c1b, 0, st);
  launch_swish_f32(net->cat, net->E, (long long)n * E8, st);
  // RB[n][ld_rb] = E . Wcat^T + bcat : the temb_cemb_proj Linear of every ResnetBlock (diffusion.py:131-132) in one GEMM
  launch_f32_to_bf16(net->E, E8, net->E_bf, E8, n, E8, st);
  launch_gather_proj(P, net->row_w, net->row_b, net->wcat_act, net->bcat, net->ld_rb, E8, st);
  launch_pack_wop(net->wcat_act, net->wcat, net->ld_rb, E8, st);   // bf16 build: wcat aliases wcat_act, nothing to do
  if (save) launch_transpose_wop(net->wcat_act, E8, net->wcatT, net->ld_rb, E8, 1, st);
  TRY(gemm_plain(plan.e_a, plan.wcat_b, n, net->ld_rb, E8, nullptr, net->RB, net->ld_rb, 0, 0, plan.bn_rb, st, net->bcat));
  return SALUN_OK;
}

static int prep_weights(salun_unet *net, bool need_dgrad, cudaStream_t st) {
  launch_prep_w_all(net->wprep_table, (int)net->convs.size(), net->params, need_dgrad ? 1 : 0, st);
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

static uint32_t gn_seed(const salun_unet *net, const UGn &g) { return net->seed * 0x9E3779B9u + (g.id + 1) * 0x85EBCA6Bu; }

static int forward_impl(salun_unet *net, const float *x, int n, int train, bool save, float *eps_out, cudaStream_t st) {
  UPlan *plan;
  TRY(build_plan(net, n, &plan));
  TRY(prep_weights(net, save, st));
  TRY(emb_forward(net, *plan, n, save, st));
  const float drop_p = train ? net->cfg.dropout : 0.f;
  net->drop_p = drop_p;
  for (const UOp &op : net->ops) {
    switch (op.type) {
      case OP_CONV:
        TRY(conv_forward(net, net->convs[op.idx], plan->conv[op.idx], n, x, eps_out, st));
        break;
      case OP_GN: {
        const UGn &g = net->gns[op.idx];
        const UTensor &in = net->ts[g.in], &out = net->ts[g.out];
        launch_gn_forward(in.v, in.want_stats ? in.st_sum : nullptr, in.st_sq, net->gn_partial, g.stats,
                          net->params + g.g_off, net->params + g.b_off, out.v,
                          out.vflat ? 1 : 0, g.swish, g.dropout ? drop_p : 0.f, gn_seed(net, g), n, g.H, g.C, 1e-6f, st);
        break;
      }
      case OP_CONCAT: {
        const UTensor &a = net->ts[op.a], &b = net->ts[op.b], &o = net->ts[op.c];
        launch_concat(a.v, a.C, b.v, b.C, o.v, n, o.H, st);
        if (o.want_stats)
          launch_concat_stats(a.st_sum, a.st_sq, a.C, b.st_sum, b.st_sq, b.C, o.st_sum, o.st_sq,
                              (long long)n * o.H * o.H / 32, st);
        break;
      }
      case OP_UP: {
        const UTensor &in = net->ts[op.a], &o = net->ts[op.b];
        launch_upsample2(in.v, o.v, n, in.H, in.C, st);
        break;
      }
      case OP_ATTN:
        TRY(attn_forward(net, net->attns[op.idx], plan->attn[op.idx], n, st));
        break;
    }
  }
  SALUN_CUDA_OK(cudaGetLastError());
  net->last_n = n;
  net->fwd_saved = save;
  return SALUN_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------
static int take_live(UTensor &t) {
  const int was = t.g_live ? 1 : 0;
  t.g_live = true;
  return was;
}

static int wgrad_conv(salun_unet *net, int ci, const UConvMaps &m, int n, cudaStream_t main_st) {
  const UConv &L = net->convs[ci];
  SALUN_CUDA_OK(cudaEventRecord(net->ev_fork, main_st));
  SALUN_CUDA_OK(cudaStreamWaitEvent(net->side, net->ev_fork, 0));
  const int64_t M = (int64_t)n * L.H * L.H;
  WgradArgs a{};
  bool dyflat, inflat;
  if (L.kind == CK_S1) {
    dyflat = net->ts[L.out].gflat;
    inflat = net->ts[L.in].vflat;
  } else if (L.kind == CK_OUT) {
    dyflat = false;
    inflat = false;
  } else {
    dyflat = false;
    inflat = true;  // col
  }
  a.mode_a = dyflat ? 0 : 1;
  a.mode_b = inflat ? 0 : 1;
  a.kb_total = (int)((M + 63) / 64);
  a.cin_blocks = L.cin >= 64 ? L.cin / 64 : 1;
  a.kw = (L.kind == CK_S1 || L.kind == CK_OUT) ? L.ks : 1;
  a.tap_y0 = a.tap_x0 = (a.kw == 3) ? 0 : 1;
  a.H = a.W = L.H;
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
  return launch_wgrad(m.wgA, m.wgB, a, co_tiles, groups, splits, net->side);
}

static int conv_backward(salun_unet *net, int ci, const UConvMaps &m, int n, const float *d_eps, cudaStream_t st) {
  const UConv &L = net->convs[ci];
  const int M = n * L.H * L.H;
  if (L.kind == CK_OUT) {
    launch_eps_in(d_eps, L.dy64, L.bias_partial, n, L.H, st);
  } else {
    UTensor &o = net->ts[L.out];
    launch_bias_partial(o.g, o.gflat ? 1 : 0, L.bias_partial, L.rb_col >= 0 ? net->dRB : nullptr, net->ld_rb, L.rb_col, n,
                        L.H, L.cout, st);
    if (L.addend >= 0) {
      UTensor &ad = net->ts[L.addend];
      launch_add_into(o.g, ad.g, take_live(ad), (long long)tensor_elems(net, ad, ad.gflat) / net->nb * n, st);
    }
  }
  // dgrad
  if (L.kind == CK_S1 || L.kind == CK_OUT) {
    UTensor &in = net->ts[L.in];
    const bool dyflat = L.kind == CK_OUT ? false : net->ts[L.out].gflat;
    ConvGemmArgs a{};
    a.M = M;
    a.N = L.cin;
    a.fH = a.fW = L.H;
    if (dyflat)
      conv_geom(a, 0, L.cout_p, 1, L.H);
    else
      conv_geom(a, 1, L.cout_p, L.ks, L.H);
    a.out_bf16 = in.g;
    a.ld_out = L.cin;
    a.out_pad = in.gflat ? 0 : 1;
    if (take_live(in)) a.addend = in.g;
    a.pair = m.pair_dg;
    TRY(launch_conv_gemm(m.dgA, m.dgB, a, u_pick_bn(L.cin, M), st));
  } else if (L.kind == CK_DOWN) {
    UTensor &in = net->ts[L.in];
    ConvGemmArgs a{};
    a.M = M;
    a.N = L.kc;
    conv_geom(a, 1, L.cout, 1, L.H);
    a.out_bf16 = L.dcol;
    a.ld_out = L.kc;
    TRY(launch_conv_gemm(m.dgA, m.dgB, a, u_pick_bn(L.kc, M), st));
    launch_down_col2im(L.dcol, in.g, take_live(in), n, 2 * L.H, L.cin, st);
  }
  return wgrad_conv(net, ci, m, n, st);
}

static int attn_backward(salun_unet *net, const UAttn &A, const UAttnMaps &m, int n, cudaStream_t st) {
  const int M = n * A.T, G = (M + A.Te - 1) / A.Te;
  const float scale = 1.f / sqrtf((float)A.C);
  UTensor &q = net->ts[A.q], &k = net->ts[A.k], &v = net->ts[A.v], &o = net->ts[A.o];
  // dV[j][c] = sum_i P[i][j] dO[i][c] : A = P^T, B = dO^T
  launch_transpose(A.P, A.Te, net->tt, A.Te, A.Te, G, st);
  launch_transpose_wop(o.g, A.C, net->xt1, A.Te, A.C, G, st);
  TRY(gemm_plain(m.tt_a, m.xt_b, M, A.C, A.Te, v.g, nullptr, A.C, A.Te, A.C, m.bn_o, st));
  // dP = dO V^T (fp32), dS = scale * P * (dP - rowsum(dP * P))
  launch_pack_wop(v.v, net->wpk, (long long)G * A.Te, A.C, st);
  TRY(gemm_plain(m.og, m.v, M, A.Te, A.C, nullptr, net->S_f32, A.Te, A.Te, A.Te, m.bn_s, st));
  launch_softmax_bwd(net->S_f32, A.P, net->dS, M, A.Te, scale, st);
  // dQ = dS K : B = K^T
  launch_transpose_wop(k.v, A.C, net->xt1, A.Te, A.C, G, st);
  TRY(gemm_plain(m.ds, m.xt_b, M, A.C, A.Te, q.g, nullptr, A.C, A.Te, A.C, m.bn_o, st));
  // dK[j][c] = sum_i dS[i][j] Q[i][c] : A = dS^T, B = Q^T
  launch_transpose(net->dS, A.Te, net->tt, A.Te, A.Te, G, st);
  launch_transpose_wop(q.v, A.C, net->xt1, A.Te, A.C, G, st);
  TRY(gemm_plain(m.tt_a, m.xt_b, M, A.C, A.Te, k.g, nullptr, A.C, A.Te, A.C, m.bn_o, st));
  q.g_live = k.g_live = v.g_live = true;
  return SALUN_OK;
}

static int emb_backward(salun_unet *net, const UPlan &plan, int n, float *gdst, cudaStream_t st) {
  const int ch = net->cfg.ch, E4 = net->emb, E8 = net->emb + 512;
  const float *P = net->params;
  // dE = dRB . Wcat ; dWcat = dRB^T . E  (two tensor-core GEMMs over all ResnetBlocks), rows scattered to the arena
  const int R = net->ld_rb;
  launch_f32_to_bf16(net->dRB, R, net->dRB_bf, R, n, R, st);
  launch_transpose(net->dRB_bf, R, net->dRBt_bf, net->nb32, R, 1, st);
  launch_transpose_wop(net->E_bf, E8, net->Et_bf, net->nb32, E8, 1, st);
  TRY(gemm_plain(plan.drb_a, plan.wcatT_b, n, E8, R, nullptr, net->dE, E8, 0, 0, plan.bn_de, st));
  TRY(gemm_plain(plan.drbt_a, plan.et_b, R, E8, kSplit ? net->nb32 : n, nullptr, net->dWcat, E8, 0, 0, plan.bn_dw, st));
  launch_scatter_rows(net->dWcat, net->row_w, gdst, R, E8, st);
  launch_dswish_f32(net->dE, net->cat, net->dcat, (long long)n * E8, st);
  for (int br = 0; br < 2; ++br) {  // 0: temb, 1: cemb
    const float *dy = net->dcat + br * E4;
    const int64_t w1 = br ? net->c1w : net->t1w, b1 = br ? net->c1b : net->t1b, w0 = br ? net->c0w : net->t0w,
                  b0 = br ? net->c0b : net->t0b;
    const float *hh = br ? net->h_c : net->h_t, *pre = br ? net->pre_c : net->pre_t, *xin = br ? net->ce : net->sincos;
    launch_sgemm(dy, 1, E8, hh, E4, 1, gdst + w1, E4, E4, E4, n, nullptr, 0, st);       // dW1 = dy^T h
    launch_colsum_f32(dy, E8, n, E4, gdst + b1, st);
    launch_sgemm(dy, E8, 1, P + w1, E4, 1, net->dh, E4, n, E4, E4, nullptr, 0, st);      // dh = dy W1
    launch_dswish_f32(net->dh, pre, net->dpre, (long long)n * E4, st);
    launch_sgemm(net->dpre, 1, E4, xin, ch, 1, gdst + w0, ch, E4, ch, n, nullptr, 0, st);  // dW0 = dpre^T x
    launch_colsum_f32(net->dpre, E4, n, E4, gdst + b0, st);
    if (br == 1) {
      launch_sgemm(net->dpre, E4, 1, P + w0, ch, 1, net->dce, ch, n, ch, E4, nullptr, 0, st);  // dce = dpre W0
      launch_emb_scatter(net->dce, net->c_dev, net->have_drop ? net->drop_dev : nullptr, gdst + net->cew,
                         gdst + net->null_off, n, ch, net->cfg.n_classes, st);
    }
  }
  return SALUN_OK;
}

__global__ void k_axpy_f32(const float *__restrict__ src, float *__restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] += src[i];
}

static int backward_impl(salun_unet *net, const float *d_eps, int accumulate, cudaStream_t st) {
  if (!net->fwd_saved) {
    set_error("salun_unet_backward called without a saved forward pass");
    return SALUN_ERR_STATE;
  }
  const int n = net->last_n;
  UPlan *plan;
  TRY(build_plan(net, n, &plan));
  float *gdst = accumulate ? net->gscratch : net->grads;
  for (UTensor &t : net->ts) t.g_live = false;
  for (int oi = (int)net->ops.size() - 1; oi >= 0; --oi) {
    const UOp &op = net->ops[oi];
    switch (op.type) {
      case OP_CONV:
        TRY(conv_backward(net, op.idx, plan->conv[op.idx], n, d_eps, st));
        break;
      case OP_GN: {
        const UGn &g = net->gns[op.idx];
        UTensor &in = net->ts[g.in];
        const UTensor &out = net->ts[g.out];
        launch_gn_backward(out.g, in.v, g.stats, net->params + g.g_off, net->params + g.b_off, g.swish,
                           g.dropout ? net->drop_p : 0.f, gn_seed(net, g), net->gn_partial, g.persample, in.g,
                           take_live(in), n, g.H, g.C, st);
        break;
      }
      case OP_CONCAT: {
        UTensor &a = net->ts[op.a], &b = net->ts[op.b];
        const UTensor &o = net->ts[op.c];
        const int la = take_live(a), lb = take_live(b);
        launch_split(o.g, a.g, a.C, la, b.g, b.C, lb, n, o.H, st);
        break;
      }
      case OP_UP: {
        UTensor &in = net->ts[op.a];
        const UTensor &o = net->ts[op.b];
        launch_upsample2_bwd(o.g, in.g, take_live(in), n, in.H, in.C, st);
        break;
      }
      case OP_ATTN:
        TRY(attn_backward(net, net->attns[op.idx], plan->attn[op.idx], n, st));
        break;
    }
  }
  launch_sum_rows_table(plan->sum_table, plan->n_sums, gdst, st);
  TRY(emb_backward(net, *plan, n, gdst, st));
  SALUN_CUDA_OK(cudaEventRecord(net->ev_join, net->side));
  SALUN_CUDA_OK(cudaStreamWaitEvent(st, net->ev_join, 0));
  for (size_t i = 0; i < net->convs.size(); ++i) {
    const UConv &L = net->convs[i];
    WgReduceEntry &e = net->wgred_host[i];
    e.ws = L.wg_ws;
    e.dst_off = L.w_off;
    e.count = (long long)L.cout * L.kc;
    e.splits = net->wg_splits[i];
    e.kc = L.kc;
  }
  SALUN_CUDA_OK(cudaMemcpyAsync(net->wgred_table, net->wgred_host, net->convs.size() * sizeof(WgReduceEntry),
                                cudaMemcpyHostToDevice, st));
  launch_wgrad_reduce(net->wgred_table, (int)net->convs.size(), gdst, st);
  if (accumulate) {
    k_axpy_f32<<<148 * 8, 256, 0, st>>>(net->gscratch, net->grads, net->n_params);
    ++g_launch_count;
  }
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

// NCHW fp32 export of a tape tensor (bring-up / parity tests)
__global__ void k_export_nchw(const bf16 *__restrict__ src, int flat, float *__restrict__ out, int n, int C, int H) {
  const long long total = (long long)n * C * H * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % H), y = (int)((i / H) % H), c = (int)((i / ((long long)H * H)) % C);
    const int b = (int)(i / ((long long)H * H * C));
    const size_t off = flat ? (((size_t)b * H + y) * H + x) * C + c
                            : (((size_t)b * (H + 2) + y + 1) * (H + 2) + x + 1) * C + c;
    out[i] = act_to_float(src[off]);
  }
}

}  // namespace salun

// =====================================================================================================================
// C ABI
// =====================================================================================================================
extern "C" {

int64_t salun_unet_param_count(const salun_unet_cfg *cfg) {
  if (!cfg) return -1;
  salun_unet tmp{};
  tmp.cfg = *cfg;
  if (build_arch(&tmp)) return -1;
  return tmp.n_params;
}

int salun_unet_destroy(salun_unet *net) {
  if (!net) return SALUN_OK;
  cudaSetDevice(net->ctx->device);
  for (void *p : net->allocs) cudaFree(p);
  if (net->wgred_host) cudaFreeHost(net->wgred_host);
  if (net->side) cudaStreamDestroy(net->side);
  if (net->ev_fork) cudaEventDestroy(net->ev_fork);
  if (net->ev_join) cudaEventDestroy(net->ev_join);
  delete net;
  return SALUN_OK;
}

int salun_unet_create(salun_ctx *ctx, const salun_unet_cfg *cfg, float *params, float *grads, salun_unet **out) {
  SALUN_REQUIRE(ctx && cfg && params && grads && out, "NULL argument");
  SALUN_REQUIRE(cfg->max_batch > 0 && cfg->n_classes > 0, "max_batch and n_classes must be positive");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  salun_unet *net = new salun_unet();
  net->ctx = ctx;
  net->cfg = *cfg;
  net->params = params;
  net->grads = grads;
  net->fwd_saved = false;
  net->last_n = 0;
  net->seed = 0;
  net->drop_p = 0.f;
  net->have_drop = false;
  int rc = build_arch(net);
  if (rc) {
    delete net;
    return rc;
  }
#define A(expr)                 \
  do {                          \
    int _rc = (expr);           \
    if (_rc) {                  \
      salun_unet_destroy(net);  \
      return _rc;               \
    }                           \
  } while (0)
  if (cudaStreamCreateWithFlags(&net->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&net->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&net->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    set_error("salun_unet: could not create the side stream / events");
    salun_unet_destroy(net);
    return SALUN_ERR_CUDA;
  }
  const int nb = (cfg->max_batch + 7) / 8 * 8;
  net->nb = nb;
  size_t max_part = 1, max_S = 1, max_xt = 1, max_tt = 1, max_pk = 1;
  for (UTensor &t : net->ts) {
    A(dmalloc(net, &t.v, tensor_elems(net, t, t.vflat)));
    if (t.need_g) A(dmalloc(net, &t.g, tensor_elems(net, t, t.gflat)));
    if (t.want_stats) {
      const size_t rows = ((size_t)nb * t.H * t.H + 127) / 128 * 4;
      A(dmalloc(net, &t.st_sum, rows * t.C));
      A(dmalloc(net, &t.st_sq, rows * t.C));
    }
  }
  for (UGn &g : net->gns) {
    A(dmalloc(net, &g.stats, (size_t)nb * kGnGroups * 2));
    A(dmalloc(net, &g.persample, (size_t)nb * 2 * g.C));
    const size_t part = (size_t)nb * unet_slices(g.H) * 2 * g.C;
    if (part > max_part) max_part = part;
  }
  for (UConv &L : net->convs) {
    const size_t M = (size_t)nb * L.H * L.H;
    A(dmalloc(net, &L.w_fwd, (size_t)L.cout_p * L.kcp * kWopK));
    if (L.kind != CK_IN) A(dmalloc(net, &L.w_dgrad, (size_t)L.cout_p * L.kc * kWopK));
    if (L.kind == CK_IN || L.kind == CK_DOWN) A(dmalloc(net, &L.col, M * L.kcp));
    if (L.kind == CK_DOWN) A(dmalloc(net, &L.dcol, M * L.kc));
    if (L.kind == CK_OUT) {
      A(dmalloc(net, &L.yf, M * 64));
      A(dmalloc(net, &L.dy64, (size_t)nb * (L.H + 2) * (L.H + 2) * 64));
    }
    const WgradGeom geo = wgrad_geometry(L.cout, L.kcp);
    const int tiles = geo.co_tiles * geo.groups;
    L.wg_splits_max = ctx->num_sms / tiles;
    if (L.wg_splits_max < 1) L.wg_splits_max = 1;
    A(dmalloc(net, &L.wg_ws, (size_t)L.wg_splits_max * wgrad_ws_elems(L.cout, L.kc), false));
    {
      size_t rows = 0;  // the largest n * slices_for(H, n) over n <= nb
      for (int b = 1; b <= nb; ++b) {
        const size_t r_ = (size_t)b * unet_slices_for(L.H, b);
        if (r_ > rows) rows = r_;
      }
      A(dmalloc(net, &L.bias_partial, rows * L.cout_p));
    }
  }
  for (UAttn &At : net->attns) {
    const size_t Mp = (size_t)nb * At.T;  // a multiple of Te (nb % 8 == 0)
    A(dmalloc(net, &At.P, Mp * At.Te));
    if (Mp * At.Te > max_S) max_S = Mp * At.Te;
    const size_t G = Mp / At.Te;
    if (G * At.C * At.Te > max_xt) max_xt = G * At.C * At.Te;
    if (G * At.Te * At.Te > max_tt) max_tt = G * At.Te * At.Te;
    if (Mp * At.C > max_pk) max_pk = Mp * At.C;
  }
  A(dmalloc(net, &net->gn_partial, max_part));
  A(dmalloc(net, &net->S_f32, max_S));
  A(dmalloc(net, &net->dS, max_S));
  A(dmalloc(net, &net->xt1, max_xt * kWopK));
  if (kSplit) A(dmalloc(net, &net->wpk, max_pk * kWopK));
  A(dmalloc(net, &net->tt, max_tt));
  A(dmalloc(net, &net->gscratch, (size_t)net->n_params));
  {
    const size_t ch = cfg->ch, E4 = net->emb, E8 = net->emb + 512;
    A(dmalloc(net, &net->sincos, nb * ch));
    A(dmalloc(net, &net->ce, nb * ch));
    A(dmalloc(net, &net->pre_t, nb * E4));
    A(dmalloc(net, &net->h_t, nb * E4));
    A(dmalloc(net, &net->pre_c, nb * E4));
    A(dmalloc(net, &net->h_c, nb * E4));
    A(dmalloc(net, &net->cat, nb * E8));
    A(dmalloc(net, &net->E, nb * E8));
    A(dmalloc(net, &net->dE, nb * E8));
    A(dmalloc(net, &net->dcat, nb * E8));
    A(dmalloc(net, &net->dh, nb * E4));
    A(dmalloc(net, &net->dpre, nb * E4));
    A(dmalloc(net, &net->dce, nb * ch));
    A(dmalloc(net, &net->RB, (size_t)nb * net->ld_rb));
    A(dmalloc(net, &net->dRB, (size_t)nb * net->ld_rb));
    {
      const size_t R = net->ld_rb;
      net->nb32 = (nb + 63) / 64 * 64;   // K pitch of the batch-contracting projection GEMM (whole k-blocks)
      A(dmalloc(net, &net->wcat_act, R * E8));
      if (kSplit)
        A(dmalloc(net, &net->wcat, R * E8 * kWopK));
      else
        net->wcat = reinterpret_cast<wop_t *>(net->wcat_act);
      A(dmalloc(net, &net->wcatT, R * E8 * kWopK));
      A(dmalloc(net, &net->bcat, R));
      A(dmalloc(net, &net->dWcat, R * E8));
      A(dmalloc(net, &net->E_bf, (size_t)net->nb32 * E8));
      A(dmalloc(net, &net->Et_bf, (size_t)net->nb32 * E8 * kWopK));
      A(dmalloc(net, &net->dRB_bf, (size_t)net->nb32 * R));
      A(dmalloc(net, &net->dRBt_bf, (size_t)net->nb32 * R));
      std::vector<long long> rw(R), rbv(R);
      for (const RbParams &p : net->rbs)
        for (int co = 0; co < p.cout; ++co) {
          rw[p.rb_col + co] = p.pw + (long long)co * E8;
          rbv[p.rb_col + co] = p.pb + co;
        }
      A(dmalloc(net, &net->row_w, R, false));
      A(dmalloc(net, &net->row_b, R, false));
      if (cudaMemcpy(net->row_w, rw.data(), R * sizeof(long long), cudaMemcpyHostToDevice) != cudaSuccess ||
          cudaMemcpy(net->row_b, rbv.data(), R * sizeof(long long), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("cudaMemcpy(projection row tables) failed");
        salun_unet_destroy(net);
        return SALUN_ERR_CUDA;
      }
    }
    A(dmalloc(net, &net->t_dev, (size_t)nb));
    A(dmalloc(net, &net->c_dev, (size_t)nb));
    A(dmalloc(net, &net->drop_dev, (size_t)nb));
  }
  net->wg_splits.assign(net->convs.size(), 1);
  A(dmalloc(net, &net->wgred_table, net->convs.size(), false));
  if (cudaMallocHost(&net->wgred_host, net->convs.size() * sizeof(WgReduceEntry)) != cudaSuccess) {
    set_error("cudaMallocHost(wgred_host) failed");
    salun_unet_destroy(net);
    return SALUN_ERR_CUDA;
  }
  {
    std::vector<WPrepEntry> tab;
    for (const UConv &L : net->convs) {
      WPrepEntry e{};
      e.w_off = L.w_off;
      e.w_fwd = L.w_fwd;
      e.w_dgrad = L.kind == CK_IN ? nullptr : L.w_dgrad;
      e.cout = L.cout;
      e.cin = L.cin;
      e.kc = L.kc;
      e.kcp = L.kcp;
      e.dgrad_mode = L.kind == CK_IN ? 0 : (L.kind == CK_DOWN ? 2 : 1);
      e.ldo = L.kind == CK_OUT ? L.cout_p : 0;
      tab.push_back(e);
    }
    A(dmalloc(net, &net->wprep_table, tab.size(), false));
    cudaError_t ce = cudaMemcpy(net->wprep_table, tab.data(), tab.size() * sizeof(WPrepEntry), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
      set_error("cudaMemcpy(wprep_table) failed: %s", cudaGetErrorString(ce));
      salun_unet_destroy(net);
      return SALUN_ERR_CUDA;
    }
  }
#undef A
  *out = net;
  return SALUN_OK;
}

int salun_unet_forward(salun_unet *net, const float *x, const float *t, const int64_t *c, const uint8_t *drop, int n,
                       int train, uint64_t seed, int save_for_backward, float *eps_out, void *stream) {
  SALUN_REQUIRE(net && x && t && c && eps_out, "NULL argument");
  SALUN_REQUIRE(n > 0 && n <= net->cfg.max_batch, "batch size out of range");
  SALUN_CUDA_OK(cudaSetDevice(net->ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  // the embedding backward re-reads t / c / drop: keep private copies (the caller may reuse its buffers)
  SALUN_CUDA_OK(cudaMemcpyAsync(net->t_dev, t, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  SALUN_CUDA_OK(cudaMemcpyAsync(net->c_dev, c, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  net->have_drop = drop != nullptr;
  if (drop) SALUN_CUDA_OK(cudaMemcpyAsync(net->drop_dev, drop, (size_t)n, cudaMemcpyDeviceToDevice, st));
  net->seed = (uint32_t)(seed ^ (seed >> 32));
  return forward_impl(net, x, n, train, save_for_backward != 0, eps_out, st);
}

int salun_unet_backward(salun_unet *net, const float *d_eps, int accumulate, void *stream) {
  SALUN_REQUIRE(net && d_eps, "NULL argument");
  SALUN_CUDA_OK(cudaSetDevice(net->ctx->device));
  return backward_impl(net, d_eps, accumulate, (cudaStream_t)stream);
}

int salun_ddpm_q_sample(salun_ctx *ctx, const float *x01, const float *e, const int64_t *t, const float *sqrt_abar,
                        const float *sqrt_1m_abar, int num_timesteps, int rescale, int n, int chw, float *x_t,
                        void *stream) {
  SALUN_REQUIRE(ctx && x01 && e && t && sqrt_abar && sqrt_1m_abar && x_t, "NULL argument");
  SALUN_REQUIRE(n >= 0 && chw > 0 && num_timesteps > 0, "bad sizes");
  if (n == 0) return SALUN_OK;
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  launch_q_sample(x01, e, t, sqrt_abar, sqrt_1m_abar, num_timesteps, rescale, n, chw, x_t, (cudaStream_t)stream);
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_ddpm_eps_loss_grad(salun_ctx *ctx, const float *eps, const float *target, const float *w, int n, int chw,
                             float *d_eps, float *sumsq_ps, float *loss_dev, void *stream) {
  SALUN_REQUIRE(ctx && eps && target && w && d_eps && sumsq_ps && loss_dev, "NULL argument");
  SALUN_REQUIRE(n > 0 && chw > 0, "bad sizes");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  launch_eps_loss_grad(eps, target, w, n, chw, d_eps, sumsq_ps, loss_dev, (cudaStream_t)stream);
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_unet_num_tensors(const salun_unet *net) { return net ? (int)net->ts.size() : -1; }

int salun_unet_tensor_info(const salun_unet *net, int idx, char *name_buf, int name_cap, int *C, int *H) {
  SALUN_REQUIRE(net && idx >= 0 && idx < (int)net->ts.size(), "tensor index out of range");
  const UTensor &t = net->ts[idx];
  if (name_buf && name_cap > 0) {
    strncpy(name_buf, t.name.c_str(), name_cap - 1);
    name_buf[name_cap - 1] = 0;
  }
  if (C) *C = t.C;
  if (H) *H = t.H;
  return SALUN_OK;
}

int salun_unet_export_tensor(salun_unet *net, int idx, int which, float *out_nchw, void *stream) {
  SALUN_REQUIRE(net && out_nchw && idx >= 0 && idx < (int)net->ts.size(), "bad argument");
  SALUN_REQUIRE(net->last_n > 0, "no forward pass yet");
  const UTensor &t = net->ts[idx];
  const bf16 *src = which ? t.g : t.v;
  SALUN_REQUIRE(src != nullptr, "tensor has no such buffer");
  k_export_nchw<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(src, which ? t.gflat : t.vflat, out_nchw, net->last_n, t.C,
                                                           t.H);
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

}  // extern "C"
