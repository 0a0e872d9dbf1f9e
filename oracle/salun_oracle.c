/*
 * oracle/salun_oracle.c -- CPU restatement of the HBM-bound tail of the SalUn hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under unlearn_saliency_b200/ may link, import or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs do, and only as the checker or as the timed CPU arm.
 *
 * Every function restates, in scalar C, what the reference computes with a chain of ATen
 * ops.  Citations are relative to /root/reference/.
 *
 * Parity pinning: the reference ships no golden vectors for this path, so
 * this restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the
 * build container by tests/golden/make_golden.py (which imports the unmodified reference) and
 * committed under tests/golden/.  tests/test_oracle_golden.py checks this file against them.
 *
 * Floating point: compiled with -ffp-contract=off so that every multiply and add below
 * rounds separately; the CUDA kernels use __fmul_rn/__fadd_rn in the same order, so the
 * CUDA path is compared BIT-EXACTLY against this file, and this file is compared against
 * torch.optim within 1e-6 relative (torch's own CPU/GPU kernels differ from each other in
 * FMA contraction, so "bit-exact against torch" is not a defined target).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * (i) saliency accumulation.
 * Classification/generate_mask.py:41-44   gradients[name] += param.grad.data   (signed!)
 * DDPM/runners/diffusion.py:992-996, SD/train-scripts/generate_mask.py:66-69 : same.
 * The |.| is taken once, AFTER the loop (generate_mask.py:46-48), see oracle_abs_inplace.
 * ------------------------------------------------------------------------------------------ */
void oracle_saliency_accumulate(float *accum, const float *grad, int64_t n) {
  for (int64_t i = 0; i < n; ++i) accum[i] = accum[i] + grad[i];
}

/* generate_mask.py:46-48  gradients[name] = torch.abs_(gradients[name]) */
void oracle_abs_inplace(float *a, int64_t n) {
  for (int64_t i = 0; i < n; ++i) a[i] = fabsf(a[i]);
}

/* ------------------------------------------------------------------------------------------
 * (i) global top-k mask.
 * generate_mask.py:57-80 :
 *     all_elements = -cat(flatten(|G|));  k = int(len * ratio)       (k is computed by the caller
 *     positions = argsort(all_elements);  ranks = argsort(positions)  in Python double arithmetic)
 *     mask = zeros_like(ranks) [int64];   mask[ranks < k] = 1
 * i.e. mask_i = 1  iff  the descending-|G| rank of element i is < k.
 *
 * Ties: torch.argsort is unstable, so the reference's choice among equal |G| values at the
 * threshold is implementation-defined.  Engine semantics, restated here:
 *     mask = { |G| > thr }  U  { the first (k - count(|G| > thr)) elements, in flat order,
 *                                with |G| == thr }
 * which is exactly what a STABLE descending argsort gives.  NaN sorts last in torch
 * (argsort of -|G| ascending puts NaN at the end), i.e. NaN has the lowest saliency; we key
 * it below +0.0.
 *
 * Key: for non-negative finite floats and +inf the IEEE bit pattern is monotone as uint32.
 * key = bits + 1, NaN -> 0.
 *
 * Implementation: restates the selection with a byte-wise radix select (same arithmetic the
 * CUDA kernel performs, but scalar and single pass per digit) -- result is independent of
 * the algorithm, tests additionally compare against numpy stable argsort on small cases.
 * Outputs: mask_i64 (may be NULL), mask_bits (may be NULL; bit i%32 of word i/32, zero padded),
 *          *thr_key_out = key of the k-th largest, *n_gt_out = count(key > thr),
 *          *n_eq_out = count(key == thr).   k == 0 -> all zero;  k >= n -> all one.
 * ------------------------------------------------------------------------------------------ */
static inline uint32_t sal_key(float a) {
  uint32_t b;
  memcpy(&b, &a, 4);
  b &= 0x7fffffffu; /* |a| */
  if (b > 0x7f800000u) return 0u; /* NaN */
  return b + 1u;
}

void oracle_topk_mask(const float *absg, int64_t n, int64_t k, int64_t *mask_i64,
                      uint32_t *mask_bits, uint32_t *thr_key_out, int64_t *n_gt_out,
                      int64_t *n_eq_out) {
  int64_t nwords = (n + 31) / 32;
  if (mask_bits) memset(mask_bits, 0, (size_t)nwords * 4);
  if (k <= 0 || n == 0) {
    if (mask_i64) memset(mask_i64, 0, (size_t)n * 8);
    if (thr_key_out) *thr_key_out = 0xffffffffu;
    if (n_gt_out) *n_gt_out = 0;
    if (n_eq_out) *n_eq_out = 0;
    return;
  }
  if (k >= n) {
    for (int64_t i = 0; i < n; ++i) {
      if (mask_i64) mask_i64[i] = 1;
      if (mask_bits) mask_bits[i >> 5] |= 1u << (i & 31);
    }
    if (thr_key_out) *thr_key_out = 0u;
    if (n_gt_out) *n_gt_out = n;
    if (n_eq_out) *n_eq_out = 0;
    return;
  }
  /* radix select, most significant byte first: find key T with count(key > T) < k <= count(key >= T) */
  uint32_t prefix = 0, prefix_mask = 0;
  int64_t remaining = k; /* rank (1-based, descending) still to locate inside the current bucket */
  for (int shift = 24; shift >= 0; shift -= 8) {
    int64_t hist[256];
    memset(hist, 0, sizeof hist);
    for (int64_t i = 0; i < n; ++i) {
      uint32_t key = sal_key(absg[i]);
      if ((key & prefix_mask) == prefix) hist[(key >> shift) & 255]++;
    }
    int d = 255;
    for (; d >= 0; --d) {
      if (remaining <= hist[d]) break;
      remaining -= hist[d];
    }
    prefix |= (uint32_t)d << shift;
    prefix_mask |= 255u << shift;
  }
  uint32_t thr = prefix;
  int64_t n_gt = 0, n_eq = 0;
  for (int64_t i = 0; i < n; ++i) {
    uint32_t key = sal_key(absg[i]);
    n_gt += key > thr;
    n_eq += key == thr;
  }
  int64_t need = k - n_gt; /* ties to take, in flat order */
  for (int64_t i = 0; i < n; ++i) {
    uint32_t key = sal_key(absg[i]);
    int64_t m = 0;
    if (key > thr) m = 1;
    else if (key == thr && need > 0) { m = 1; --need; }
    if (mask_i64) mask_i64[i] = m;
    if (mask_bits && m) mask_bits[i >> 5] |= 1u << (i & 31);
  }
  if (thr_key_out) *thr_key_out = thr;
  if (n_gt_out) *n_gt_out = n_gt;
  if (n_eq_out) *n_eq_out = n_eq;
}

/* pack an int64 {0,1} mask (the on-disk format, generate_mask.py:76-80) into bits */
void oracle_pack_mask(const int64_t *mask_i64, int64_t n, uint32_t *mask_bits) {
  int64_t nwords = (n + 31) / 32;
  memset(mask_bits, 0, (size_t)nwords * 4);
  for (int64_t i = 0; i < n; ++i)
    if (mask_i64[i] != 0) mask_bits[i >> 5] |= 1u << (i & 31);
}

/* ------------------------------------------------------------------------------------------
 * (ii) masked SGD step with restore  (Classification).
 *   unlearn/RL.py:11-14    _apply_mask_to_grads : g *= m
 *   unlearn/impl.py:68-73  torch.optim.SGD(lr, momentum, weight_decay), dampening 0, no nesterov:
 *        g' = g + wd*p ;  v = g' (first step) | v = mu*v + g' ;  p = p - lr*v
 *   unlearn/RL.py:17-34    _restore_masked_params : p = p*m + theta0*(1-m) ; v *= m
 * Net effect per coordinate:
 *   m == 1 : g' = g + wd*p ; v = mu*v + g' ; p = p - lr*v     (v starts at 0, so mu*0+g' == g')
 *   m == 0 : p unchanged (== theta0), v = 0
 * mask_bits == NULL means "no mask" (every coordinate updated), as main_forget.py:134-135 does.
 * ------------------------------------------------------------------------------------------ */
void oracle_masked_sgd_step(float *p, const float *g, float *v, const uint32_t *mask_bits,
                            int64_t n, float lr, float momentum, float wd) {
  for (int64_t i = 0; i < n; ++i) {
    int m = mask_bits ? (int)((mask_bits[i >> 5] >> (i & 31)) & 1u) : 1;
    if (m) {
      float gp = g[i] + wd * p[i];
      float vn = momentum * v[i] + gp;
      v[i] = vn;
      p[i] = p[i] - lr * vn;
    } else {
      v[i] = 0.0f;
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * (ii) clip + mask + Adam  (DDPM / SD).
 *   DDPM/runners/diffusion.py:582-587  clip_grad_norm_(params, optim.grad_clip=1.0)  BEFORE masking:
 *        total = sqrt(sum g^2) ; c = min(1, max_norm / (total + 1e-6)) ; g *= c
 *   DDPM/runners/diffusion.py:589-592  g *= mask
 *   DDPM/functions/__init__.py:9-18    optim.Adam(lr, betas=(beta1, 0.999), eps, weight_decay 0, amsgrad off)
 * torch.optim.Adam single-tensor arithmetic (torch/optim/adam.py, _single_tensor_adam):
 *        m1 = m1 + (g - m1)*(1-b1)                      (lerp_)
 *        m2 = m2*b2 + (1-b2)*g*g                        (mul_ ; addcmul_)
 *        bc1 = 1 - b1^t ; bc2 = 1 - b2^t  (python doubles) ; step_size = lr/bc1
 *        denom = sqrt(m2)/sqrt(bc2) + eps ; p = p + (-step_size) * (m1/denom)
 * max_norm < 0 disables clipping (SD scripts: train-esd.py:313-323 has no clip).
 * grad_sumsq is computed in double over all n (deterministic).  Returns the pre-clip norm.
 * ------------------------------------------------------------------------------------------ */
double oracle_grad_norm(const float *g, int64_t n) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += (double)g[i] * (double)g[i];
  return sqrt(s);
}

float oracle_clip_coef(double total_norm, float max_norm) {
  if (max_norm < 0.0f) return 1.0f;
  float c = max_norm / ((float)total_norm + 1e-6f);
  return c > 1.0f ? 1.0f : c;
}

void oracle_masked_adam_step(float *p, const float *g, float *m1, float *m2,
                             const uint32_t *mask_bits, int64_t n, float lr, float b1, float b2,
                             float eps, float wd, int64_t step, float clip_coef) {
  double bc1 = 1.0 - pow((double)b1, (double)step);
  double bc2 = 1.0 - pow((double)b2, (double)step);
  float step_size = (float)((double)lr / bc1);
  float bc2_sqrt = (float)sqrt(bc2);
  /* torch computes (1-beta) as a python double and narrows it to fp32 at the kernel boundary */
  float one_m_b1 = (float)(1.0 - (double)b1);
  float one_m_b2 = (float)(1.0 - (double)b2);
  for (int64_t i = 0; i < n; ++i) {
    int m = mask_bits ? (int)((mask_bits[i >> 5] >> (i & 31)) & 1u) : 1;
    float gi = g[i] * clip_coef;
    if (!m) gi = 0.0f * gi; /* g *= mask : keeps NaN/inf propagation of the reference */
    if (wd != 0.0f) gi = gi + wd * p[i];
    float d = gi - m1[i];
    float a = m1[i] + d * one_m_b1;
    float b = m2[i] * b2;
    float gg = gi * gi;
    b = b + one_m_b2 * gg;
    m1[i] = a;
    m2[i] = b;
    float denom = sqrtf(b) / bc2_sqrt + eps;
    float q = a / denom;
    p[i] = p[i] + (-step_size) * q;
  }
}

/* mask (.) grad only -- RL.py:11-14, runners/diffusion.py:589-592, train-esd.py:318-321 */
void oracle_apply_mask(float *g, const uint32_t *mask_bits, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    int m = (int)((mask_bits[i >> 5] >> (i & 31)) & 1u);
    if (!m) g[i] = g[i] * 0.0f;
  }
}

/* ------------------------------------------------------------------------------------------
 * FT_l1 penalty.  Classification/unlearn/FT.py:13-17  l1_regularization = ||cat(params)||_1 ;
 * :133-134  loss += current_alpha * l1_regularization(model)  =>  autograd adds
 * alpha * sign(theta) (sign(0) = 0) to every gradient BEFORE the mask multiply (:138-139).
 * Returns sum |theta| (double accumulation; the CUDA path's sum differs only in association).
 * ------------------------------------------------------------------------------------------ */
double oracle_l1_penalty_grad(const float *p, float *g, int64_t n, float alpha) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const float v = p[i];
    s += (double)fabsf(v);
    const float sg = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);
    g[i] = g[i] + alpha * sg;
  }
  return s;
}
