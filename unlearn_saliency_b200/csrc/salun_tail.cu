// salun_tail.cu -- HBM-bound tail of the SalUn hot path for sm_100a.
//
//   (i)  saliency accumulate, |.|, global top-k mask by 3-pass radix select   (include/salun.h)
//   (ii) mask (.) grad, fused masked SGD+restore, grad-norm clip, fused masked Adam
//
// All kernels are streaming kernels bounded by HBM bandwidth: 16-byte vector loads/stores,
// persistent grids of (#SMs x 8) CTAs x 256 threads, no shared-memory staging (no reuse).
// Floating point uses explicit __fmul_rn/__fadd_rn so the sequence of roundings equals
// oracle/salun_oracle.c (built with -ffp-contract=off) and results compare bit-exactly.
#include <stdlib.h>
#include <math.h>
#include <stdarg.h>

#include "salun_common.cuh"

namespace salun {

long long g_launch_count = 0;
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SALUN_PDL");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}
static thread_local char g_err[512] = "";
char *err_buf() { return g_err; }
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

constexpr int kThreads = 256;

static inline int grid_for(const salun_ctx *ctx, int64_t n_vec) {
  int64_t want = (n_vec + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)ctx->num_sms * 8;
  if (cap > kMaxPartials) cap = kMaxPartials;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

__device__ __forceinline__ uint32_t sal_key(float a) {
  uint32_t b = __float_as_uint(a) & 0x7fffffffu;
  return b > 0x7f800000u ? 0u : b + 1u;  // NaN ranks last (torch.argsort semantics), else monotone
}

// guarded 4-wide load: elements past n read as `pad`
__device__ __forceinline__ float4 load4(const float *__restrict__ a, int64_t i, int64_t n, float pad) {
  if (i + 3 < n) return *reinterpret_cast<const float4 *>(a + i);
  float4 r = make_float4(pad, pad, pad, pad);
  if (i < n) r.x = a[i];
  if (i + 1 < n) r.y = a[i + 1];
  if (i + 2 < n) r.z = a[i + 2];
  return r;
}
__device__ __forceinline__ void store4(float *__restrict__ a, int64_t i, int64_t n, float4 v) {
  if (i + 3 < n) {
    *reinterpret_cast<float4 *>(a + i) = v;
    return;
  }
  if (i < n) a[i] = v.x;
  if (i + 1 < n) a[i + 1] = v.y;
  if (i + 2 < n) a[i + 2] = v.z;
}
__device__ __forceinline__ uint32_t mask_nibble(const uint32_t *__restrict__ bits, int64_t i) {
  // i is a multiple of 4: the 4 mask bits of elements i..i+3 sit in one word
  return bits ? (__ldg(bits + (i >> 5)) >> (i & 31)) & 0xFu : 0xFu;
}

// ------------------------------------------------------------------------------------------
// accumulate / abs
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_accum_flat(const float *__restrict__ g, float *__restrict__ acc,
                                                         int64_t n, const float *__restrict__ scale) {
  const float s = scale ? *scale : 1.0f;
  const bool scaled = scale != nullptr;
  int64_t stride = (int64_t)gridDim.x * kThreads * 4;
  for (int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4; i < n; i += stride) {
    float4 a = load4(acc, i, n, 0.f), b = load4(g, i, n, 0.f);
    if (scaled) {
      b.x = __fmul_rn(b.x, s); b.y = __fmul_rn(b.y, s); b.z = __fmul_rn(b.z, s); b.w = __fmul_rn(b.w, s);
    }
    a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
    store4(acc, i, n, a);
  }
}

constexpr int kTabMax = 96;
struct TensorTable {
  const float *src[kTabMax];
  long long dst_off[kTabMax];
  long long numel[kTabMax];
};
// one grid row per tensor (blockIdx.y); tensors are separately allocated (param.grad), the
// destination offsets are arbitrary, so accesses are 4-byte coalesced rather than 16-byte.
__global__ void __launch_bounds__(kThreads) k_accum_multi(TensorTable tab, float *__restrict__ acc,
                                                          const float *__restrict__ scale) {
  const float s = scale ? *scale : 1.0f;
  const bool scaled = scale != nullptr;
  const float *__restrict__ src = tab.src[blockIdx.y];
  float *__restrict__ dst = acc + tab.dst_off[blockIdx.y];
  const long long n = tab.numel[blockIdx.y];
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    float b = src[i];
    if (scaled) b = __fmul_rn(b, s);
    dst[i] = __fadd_rn(dst[i], b);
  }
}

__global__ void __launch_bounds__(kThreads) k_abs(float *__restrict__ a, int64_t n) {
  int64_t stride = (int64_t)gridDim.x * kThreads * 4;
  for (int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4; i < n; i += stride) {
    float4 v = load4(a, i, n, 0.f);
    v.x = fabsf(v.x); v.y = fabsf(v.y); v.z = fabsf(v.z); v.w = fabsf(v.w);
    store4(a, i, n, v);
  }
}

// ------------------------------------------------------------------------------------------
// radix select: key = 32 bits split 11 | 11 | 10 (most significant first)
// sel[0]=prefix sel[1]=prefix_mask sel[2]=remaining sel[3]=thr sel[4]=n_gt sel[5]=n_eq
// sel[6]=need sel[7]=ordered_ties
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int pass_shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }
__device__ __forceinline__ int pass_bins(int pass) { return pass == 2 ? 1024 : 2048; }

__global__ void __launch_bounds__(kThreads) k_radix_hist(const float *__restrict__ a, int64_t n,
                                                         const unsigned long long *__restrict__ sel,
                                                         unsigned int *__restrict__ hist, int pass) {
  __shared__ unsigned int sh[kRadixBins];
  for (int i = threadIdx.x; i < kRadixBins; i += kThreads) sh[i] = 0;
  __syncthreads();
  const uint32_t prefix = (uint32_t)sel[0], pmask = (uint32_t)sel[1];
  const int shift = pass_shift(pass);
  const uint32_t dmask = pass_bins(pass) - 1;
  int64_t stride = (int64_t)gridDim.x * kThreads * 4;
  for (int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4; i < n; i += stride) {
    float4 v = load4(a, i, n, 0.f);
    uint32_t k0 = sal_key(v.x), k1 = sal_key(v.y), k2 = sal_key(v.z), k3 = sal_key(v.w);
    if ((k0 & pmask) == prefix) atomicAdd(&sh[(k0 >> shift) & dmask], 1u);
    if (i + 1 < n && (k1 & pmask) == prefix) atomicAdd(&sh[(k1 >> shift) & dmask], 1u);
    if (i + 2 < n && (k2 & pmask) == prefix) atomicAdd(&sh[(k2 >> shift) & dmask], 1u);
    if (i + 3 < n && (k3 & pmask) == prefix) atomicAdd(&sh[(k3 >> shift) & dmask], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRadixBins; i += kThreads)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// one warp: walk the histogram from the top digit, pick the digit holding rank `remaining`
__device__ __forceinline__ void radix_pick_warp(const unsigned int *__restrict__ hist, unsigned long long *__restrict__ sel,
                                                int pass, long long k, int lane) {
  const int bins = pass_bins(pass);
  const int per = bins / 32;
  const unsigned long long remaining = sel[2], prefix_in = sel[0], pmask_in = sel[1];  // read before any lane writes
  // lane 0 owns the TOP `per` digits
  const int top = bins - 1 - lane * per;
  unsigned long long mine = 0;
  for (int j = 0; j < per; ++j) mine += hist[top - j];
  unsigned long long incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  unsigned long long excl = incl - mine;
  const bool owner = remaining > excl && remaining <= incl;
  __syncwarp();
  if (owner) {
    unsigned long long r = remaining - excl;
    int d = top;
    for (int j = 0; j < per; ++j, --d) {
      unsigned long long h = hist[d];
      if (r <= h) break;
      r -= h;
    }
    const int shift = pass_shift(pass);
    unsigned long long prefix = prefix_in | ((unsigned long long)d << shift);
    sel[0] = prefix;
    sel[1] = pmask_in | ((unsigned long long)(bins - 1) << shift);
    sel[2] = r;
    if (pass == 2) {
      unsigned long long n_eq = hist[d];
      sel[3] = prefix;           // threshold key
      sel[4] = (unsigned long long)k - r;  // n_gt
      sel[5] = n_eq;
      sel[6] = r;                // ties to take
      sel[7] = r < n_eq ? 1ull : 0ull;  // flat-order tie resolution needed
    }
  }
  __syncwarp();
}
__global__ void k_radix_pick(unsigned int *__restrict__ hist, unsigned long long *__restrict__ sel, int pass,
                             long long k) {
  const int lane = threadIdx.x;
  radix_pick_warp(hist, sel, pass, k, lane);
  for (int i = lane; i < kRadixBins; i += 32) hist[i] = 0;  // ready for the next pass
}

__global__ void k_sel_init(unsigned long long *sel, unsigned int *hist, long long k) {
  int t = threadIdx.x + blockIdx.x * blockDim.x;
  if (t < 8) sel[t] = t == 2 ? (unsigned long long)k : 0ull;
  if (t < kRadixBins) hist[t] = 0;
}

// contiguous chunk of block b: [b*chunk, min(n, (b+1)*chunk)), chunk a multiple of 1024
__global__ void __launch_bounds__(kThreads) k_tie_count(const float *__restrict__ a, int64_t n, int64_t chunk,
                                                        const unsigned long long *__restrict__ sel,
                                                        unsigned int *__restrict__ block_ties) {
  __shared__ unsigned int warp_cnt[kThreads / 32];
  if (sel[7] == 0) return;  // all ties taken: no ordering needed (uniform across the grid)
  const uint32_t thr = (uint32_t)sel[3];
  const int64_t lo = (int64_t)blockIdx.x * chunk;
  const int64_t hi = min(n, lo + chunk);
  unsigned int c = 0;
  for (int64_t i = lo + (int64_t)threadIdx.x * 4; i < hi; i += kThreads * 4) {
    float4 v = load4(a, i, n, 0.f);
    c += (sal_key(v.x) == thr) + (i + 1 < n && sal_key(v.y) == thr) + (i + 2 < n && sal_key(v.z) == thr) +
         (i + 3 < n && sal_key(v.w) == thr);
  }
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += warp_cnt[w];
    block_ties[blockIdx.x] = t;
  }
}

__global__ void k_tie_scan(unsigned int *__restrict__ block_ties, int nblocks,
                           const unsigned long long *__restrict__ sel) {
  if (sel[7] == 0) return;
  // single thread: nblocks <= kMaxPartials (1184) -- negligible next to the N-element passes
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned int run = 0;
    for (int b = 0; b < nblocks; ++b) {
      unsigned int t = block_ties[b];
      block_ties[b] = run;
      run += t;
    }
  }
}

template <bool kVecI64>
__global__ void __launch_bounds__(kThreads) k_write_mask(const float *__restrict__ a, int64_t n, int64_t chunk,
                                                         const unsigned long long *__restrict__ sel,
                                                         const unsigned int *__restrict__ block_ties,
                                                         long long *__restrict__ mask_i64,
                                                         uint32_t *__restrict__ mask_bits, int all_ones) {
  __shared__ unsigned int warp_tot[kThreads / 32];
  __shared__ unsigned int iter_base;
  const uint32_t thr = all_ones ? 0u : (uint32_t)sel[3];
  const bool ordered = !all_ones && sel[7] != 0;
  const unsigned long long need = all_ones ? 0ull : sel[6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) iter_base = ordered ? block_ties[blockIdx.x] : 0u;
  __syncthreads();
  const int64_t lo = (int64_t)blockIdx.x * chunk;
  const int64_t hi = min(n, lo + chunk);
  for (int64_t base = lo; base < hi; base += kThreads * 4) {  // uniform trip count per block
    const int64_t i = base + (int64_t)threadIdx.x * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < hi) v = load4(a, i, n, 0.f);
    uint32_t key[4] = {sal_key(v.x), sal_key(v.y), sal_key(v.z), sal_key(v.w)};
    uint32_t gt = 0, eq = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool valid = i + j < n && i < hi;
      if (all_ones) {
        gt |= (valid ? 1u : 0u) << j;
      } else {
        gt |= ((valid && key[j] > thr) ? 1u : 0u) << j;
        eq |= ((valid && key[j] == thr) ? 1u : 0u) << j;
      }
    }
    uint32_t sel_bits = gt;
    if (!ordered) {
      sel_bits |= eq;  // need == n_eq: every tie is selected
    } else {
      // block-wide exclusive prefix of tie counts in flat order
      unsigned int c = __popc(eq), incl = c;
      for (int o = 1; o < 32; o <<= 1) {
        unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) warp_tot[warp] = incl;
      __syncthreads();
      unsigned int wbase = 0, total = 0;
      for (int w = 0; w < kThreads / 32; ++w) {
        unsigned int t = warp_tot[w];
        if (w < warp) wbase += t;
        total += t;
      }
      unsigned long long rank = (unsigned long long)iter_base + wbase + (incl - c);  // ties before my first element
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (eq & (1u << j)) {
          if (rank < need) sel_bits |= 1u << j;
          ++rank;
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) iter_base += total;
      // next iteration's readers are separated from this write by the __syncthreads above in that iteration
    }
    if (mask_i64 && i < hi) {
      long long m0 = sel_bits & 1u, m1 = (sel_bits >> 1) & 1u, m2 = (sel_bits >> 2) & 1u, m3 = (sel_bits >> 3) & 1u;
      if (kVecI64 && i + 3 < n) {
        reinterpret_cast<longlong2 *>(mask_i64 + i)[0] = make_longlong2(m0, m1);
        reinterpret_cast<longlong2 *>(mask_i64 + i)[1] = make_longlong2(m2, m3);
      } else {
        if (i < n) mask_i64[i] = m0;
        if (i + 1 < n) mask_i64[i + 1] = m1;
        if (i + 2 < n) mask_i64[i + 2] = m2;
        if (i + 3 < n) mask_i64[i + 3] = m3;
      }
    }
    if (mask_bits) {
      // 8 lanes x 4 bits -> one 32-bit word
      uint32_t w = sel_bits << (4 * (lane & 7));
      w |= __shfl_xor_sync(0xffffffffu, w, 1);
      w |= __shfl_xor_sync(0xffffffffu, w, 2);
      w |= __shfl_xor_sync(0xffffffffu, w, 4);
      if ((lane & 7) == 0 && i < hi && i < n) mask_bits[i >> 5] = w;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Several ratios in one sweep (the reference's threshold_list = 0.1 .. 1.0, generate_mask.py:50-82): the three histogram
// passes, the tie count and the mask write read the saliencies ONCE for all ratios instead of once per ratio, and the
// ~10 launches per ratio become ~12 per call.  Same selection rule as salun_topk_mask, ratio by ratio (same threshold
// key, same flat-order tie resolution): the outputs are bit-identical to R single calls.
// ------------------------------------------------------------------------------------------
struct MultiK {
  long long k[kMaxRatios];
  int mode[kMaxRatios];  // 0: select k of n, 1: all ones (k >= n), 2: all zeros (k == 0)
  int R;
};
struct MultiOut {
  long long *m64[kMaxRatios];
  uint32_t *bits[kMaxRatios];
  int vec64[kMaxRatios];
};

__global__ void k_msel_init(unsigned long long *msel, unsigned int *mhist, MultiK mk) {
  const int t = threadIdx.x + blockIdx.x * blockDim.x, nt = gridDim.x * blockDim.x;
  if (t < mk.R * 8) msel[t] = (t & 7) == 2 ? (unsigned long long)mk.k[t >> 3] : 0ull;
  for (int i = t; i < kMaxRatios * kRadixBins; i += nt) mhist[i] = 0;
}

// PASS 0: one histogram of the top digit for everybody.  PASS 1 / 2: one histogram per ratio of the next digit of the
// elements under that ratio's prefix; `which[top digit]` = bit set of the ratios whose prefix starts with that digit
template <int PASS>
__global__ void __launch_bounds__(kThreads) k_mhist(const float *__restrict__ a, int64_t n,
                                                    const unsigned long long *__restrict__ msel,
                                                    unsigned int *__restrict__ mhist, MultiK mk) {
  extern __shared__ unsigned int msm[];
  const int Rh = PASS == 0 ? 1 : mk.R;
  unsigned int *sh = msm, *which = msm + Rh * kRadixBins, *pref = which + kRadixBins;
  for (int i = threadIdx.x; i < Rh * kRadixBins; i += kThreads) sh[i] = 0;
  if (PASS > 0) {
    for (int i = threadIdx.x; i < kRadixBins; i += kThreads) which[i] = 0;
    __syncthreads();
    if (threadIdx.x < mk.R && mk.mode[threadIdx.x] == 0) {
      const uint32_t p = (uint32_t)msel[threadIdx.x * 8];
      pref[threadIdx.x] = p;
      atomicOr(&which[p >> 21], 1u << threadIdx.x);
    }
  }
  __syncthreads();
  constexpr int shift = PASS == 0 ? 21 : (PASS == 1 ? 10 : 0);
  constexpr uint32_t dmask = PASS == 2 ? 1023u : 2047u;
  constexpr uint32_t pmask = PASS == 2 ? 0xFFFFFC00u : 0xFFE00000u;
  const int64_t stride = (int64_t)gridDim.x * kThreads * 4;
  for (int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4; i < n; i += stride) {
    const float4 v = load4(a, i, n, 0.f);
    const uint32_t key[4] = {sal_key(v.x), sal_key(v.y), sal_key(v.z), sal_key(v.w)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i + j >= n) continue;
      const uint32_t kk = key[j];
      if (PASS == 0) {
        atomicAdd(&sh[kk >> 21], 1u);
      } else {
        unsigned int m = which[kk >> 21];
        while (m) {
          const int r = __ffs(m) - 1;
          m &= m - 1;
          if (PASS == 1 || (kk & pmask) == pref[r]) atomicAdd(&sh[r * kRadixBins + ((kk >> shift) & dmask)], 1u);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Rh * kRadixBins; i += kThreads)
    if (sh[i]) atomicAdd(&mhist[i], sh[i]);
}

// warp r picks for ratio r (pass 0 reads the shared histogram), then everybody clears the histograms for the next pass
__global__ void k_mpick(unsigned int *__restrict__ mhist, unsigned long long *__restrict__ msel, int pass, MultiK mk) {
  const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (r < mk.R && mk.mode[r] == 0)
    radix_pick_warp(pass == 0 ? mhist : mhist + r * kRadixBins, msel + r * 8, pass, mk.k[r], lane);
  __syncthreads();
  for (int i = threadIdx.x; i < kMaxRatios * kRadixBins; i += blockDim.x) mhist[i] = 0;
}

__global__ void __launch_bounds__(kThreads) k_mtie_count(const float *__restrict__ a, int64_t n, int64_t chunk,
                                                         const unsigned long long *__restrict__ msel,
                                                         unsigned int *__restrict__ mties, MultiK mk) {
  __shared__ unsigned int warp_cnt[kMaxRatios][kThreads / 32];
  __shared__ uint32_t thr[kMaxRatios];
  __shared__ int ordered[kMaxRatios];
  if (threadIdx.x < kMaxRatios) {
    const int r = threadIdx.x;
    const bool on = r < mk.R && mk.mode[r] == 0 && msel[r * 8 + 7] != 0;
    ordered[r] = on ? 1 : 0;
    thr[r] = on ? (uint32_t)msel[r * 8 + 3] : 0u;
  }
  __syncthreads();
  int any = 0;
  for (int r = 0; r < mk.R; ++r) any |= ordered[r];
  if (!any) return;  // uniform across the grid
  unsigned int c[kMaxRatios];
#pragma unroll
  for (int r = 0; r < kMaxRatios; ++r) c[r] = 0;
  const int64_t lo = (int64_t)blockIdx.x * chunk;
  const int64_t hi = min(n, lo + chunk);
  for (int64_t i = lo + (int64_t)threadIdx.x * 4; i < hi; i += kThreads * 4) {
    const float4 v = load4(a, i, n, 0.f);
    const uint32_t k0 = sal_key(v.x), k1 = sal_key(v.y), k2 = sal_key(v.z), k3 = sal_key(v.w);
#pragma unroll
    for (int r = 0; r < kMaxRatios; ++r)
      if (r < mk.R && ordered[r]) {
        const uint32_t t = thr[r];
        c[r] += (k0 == t) + (i + 1 < n && k1 == t) + (i + 2 < n && k2 == t) + (i + 3 < n && k3 == t);
      }
  }
#pragma unroll
  for (int r = 0; r < kMaxRatios; ++r) {
    unsigned int x = c[r];
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) warp_cnt[r][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x < mk.R) {
    unsigned int t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += warp_cnt[threadIdx.x][w];
    mties[threadIdx.x * (kMaxPartials + 1) + blockIdx.x] = t;
  }
}

__global__ void k_mtie_scan(unsigned int *__restrict__ mties, int nblocks, const unsigned long long *__restrict__ msel, MultiK mk) {
  const int r = threadIdx.x;
  if (r >= mk.R || mk.mode[r] != 0 || msel[r * 8 + 7] == 0) return;
  unsigned int *bt = mties + r * (kMaxPartials + 1);
  unsigned int run = 0;
  for (int b = 0; b < nblocks; ++b) {
    const unsigned int t = bt[b];
    bt[b] = run;
    run += t;
  }
}

__global__ void __launch_bounds__(kThreads) k_mwrite(const float *__restrict__ a, int64_t n, int64_t chunk,
                                                     const unsigned long long *__restrict__ msel,
                                                     const unsigned int *__restrict__ mties, MultiK mk, MultiOut out) {
  __shared__ unsigned int warp_tot[kThreads / 32];
  __shared__ unsigned int iter_base[kMaxRatios];
  __shared__ uint32_t thr[kMaxRatios];
  __shared__ unsigned long long need[kMaxRatios];
  __shared__ int ordered[kMaxRatios];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < kMaxRatios) {
    const int r = threadIdx.x;
    const bool sel = r < mk.R && mk.mode[r] == 0;
    thr[r] = sel ? (uint32_t)msel[r * 8 + 3] : 0u;
    need[r] = sel ? msel[r * 8 + 6] : 0ull;
    ordered[r] = (sel && msel[r * 8 + 7] != 0) ? 1 : 0;
    iter_base[r] = ordered[r] ? mties[r * (kMaxPartials + 1) + blockIdx.x] : 0u;
  }
  __syncthreads();
  const int64_t lo = (int64_t)blockIdx.x * chunk;
  const int64_t hi = min(n, lo + chunk);
  for (int64_t base = lo; base < hi; base += kThreads * 4) {  // uniform trip count per block
    const int64_t i = base + (int64_t)threadIdx.x * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < hi) v = load4(a, i, n, 0.f);
    const uint32_t key[4] = {sal_key(v.x), sal_key(v.y), sal_key(v.z), sal_key(v.w)};
    uint32_t validm = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) validm |= ((i + j < n && i < hi) ? 1u : 0u) << j;
    for (int r = 0; r < mk.R; ++r) {
      uint32_t sel_bits;
      if (mk.mode[r] == 1) {
        sel_bits = validm;
      } else if (mk.mode[r] == 2) {
        sel_bits = 0;
      } else {
        const uint32_t t = thr[r];
        uint32_t gt = 0, eq = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool valid = (validm >> j) & 1u;
          gt |= ((valid && key[j] > t) ? 1u : 0u) << j;
          eq |= ((valid && key[j] == t) ? 1u : 0u) << j;
        }
        sel_bits = gt;
        if (!ordered[r]) {
          sel_bits |= eq;
        } else {
          unsigned int c = __popc(eq), incl = c;
          for (int o = 1; o < 32; o <<= 1) {
            unsigned int tt = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += tt;
          }
          if (lane == 31) warp_tot[warp] = incl;
          __syncthreads();
          unsigned int wbase = 0, total = 0;
          for (int w = 0; w < kThreads / 32; ++w) {
            const unsigned int tt = warp_tot[w];
            if (w < warp) wbase += tt;
            total += tt;
          }
          unsigned long long rank = (unsigned long long)iter_base[r] + wbase + (incl - c);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (eq & (1u << j)) {
              if (rank < need[r]) sel_bits |= 1u << j;
              ++rank;
            }
          __syncthreads();
          if (threadIdx.x == 0) iter_base[r] += total;
          __syncthreads();
        }
      }
      long long *m64 = out.m64[r];
      if (m64 && i < hi) {
        const long long m0 = sel_bits & 1u, m1 = (sel_bits >> 1) & 1u, m2 = (sel_bits >> 2) & 1u, m3 = (sel_bits >> 3) & 1u;
        if (out.vec64[r] && i + 3 < n) {
          reinterpret_cast<longlong2 *>(m64 + i)[0] = make_longlong2(m0, m1);
          reinterpret_cast<longlong2 *>(m64 + i)[1] = make_longlong2(m2, m3);
        } else {
          if (i < n) m64[i] = m0;
          if (i + 1 < n) m64[i + 1] = m1;
          if (i + 2 < n) m64[i + 2] = m2;
          if (i + 3 < n) m64[i + 3] = m3;
        }
      }
      uint32_t *mb = out.bits[r];
      if (mb) {
        uint32_t w = sel_bits << (4 * (lane & 7));
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        w |= __shfl_xor_sync(0xffffffffu, w, 4);
        if ((lane & 7) == 0 && i < hi && i < n) mb[i >> 5] = w;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// mask pack / unpack / apply
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_pack_mask(const long long *__restrict__ m, int64_t n,
                                                        uint32_t *__restrict__ bits) {
  const int lane = threadIdx.x & 31;
  int64_t nw = (n + 31) / 32;
  int64_t warp_id = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * kThreads) >> 5;
  for (int64_t w = warp_id; w < nw; w += nwarps) {
    int64_t i = w * 32 + lane;
    bool on = i < n && m[i] != 0;
    uint32_t word = __ballot_sync(0xffffffffu, on);
    if (lane == 0) bits[w] = word;
  }
}
__global__ void __launch_bounds__(kThreads) k_unpack_mask(const uint32_t *__restrict__ bits, int64_t n,
                                                          long long *__restrict__ m) {
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    m[i] = (bits[i >> 5] >> (i & 31)) & 1u;
}
__global__ void __launch_bounds__(kThreads) k_apply_mask(float *__restrict__ g, const uint32_t *__restrict__ bits,
                                                         int64_t n) {
  int64_t stride = (int64_t)gridDim.x * kThreads * 4;
  for (int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4; i < n; i += stride) {
    uint32_t nib = mask_nibble(bits, i);
    if (nib == 0xFu) continue;
    float4 v = load4(g, i, n, 0.f);
    if (!(nib & 1u)) v.x = __fmul_rn(v.x, 0.f);
    if (!(nib & 2u)) v.y = __fmul_rn(v.y, 0.f);
    if (!(nib & 4u)) v.z = __fmul_rn(v.z, 0.f);
    if (!(nib & 8u)) v.w = __fmul_rn(v.w, 0.f);
    store4(g, i, n, v);
  }
}

// ------------------------------------------------------------------------------------------
// fused masked SGD + restore          (oracle_masked_sgd_step)
// bytes/param: read p,g,v (12) + 1/8 mask + write p,v (8) = 20.125
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void sgd1(float &p, float g, float &v, bool m, float lr, float mu, float wd) {
  if (m) {
    float gp = __fadd_rn(g, __fmul_rn(wd, p));
    float vn = __fadd_rn(__fmul_rn(mu, v), gp);
    v = vn;
    p = __fadd_rn(p, -__fmul_rn(lr, vn));
  } else {
    v = 0.f;
  }
}
__global__ void __launch_bounds__(kThreads) k_masked_sgd(float *__restrict__ p, const float *__restrict__ g,
                                                         float *__restrict__ v, const uint32_t *__restrict__ bits,
                                                         int64_t n, float lr, float mu, float wd) {
  int64_t stride = (int64_t)gridDim.x * kThreads * 4;
  for (int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4; i < n; i += stride) {
    uint32_t nib = mask_nibble(bits, i);
    float4 vv = load4(v, i, n, 0.f);
    if (nib == 0u) {  // whole vector masked out: p untouched, only v forced to 0
      store4(v, i, n, make_float4(0.f, 0.f, 0.f, 0.f));
      continue;
    }
    float4 pp = load4(p, i, n, 0.f), gg = load4(g, i, n, 0.f);
    sgd1(pp.x, gg.x, vv.x, nib & 1u, lr, mu, wd);
    sgd1(pp.y, gg.y, vv.y, nib & 2u, lr, mu, wd);
    sgd1(pp.z, gg.z, vv.z, nib & 4u, lr, mu, wd);
    sgd1(pp.w, gg.w, vv.w, nib & 8u, lr, mu, wd);
    store4(p, i, n, pp);
    store4(v, i, n, vv);
  }
}

// ------------------------------------------------------------------------------------------
// data-parallel fused step: reduce-scatter(grad) over NVLink peer memory -> masked SGD on the owned shard ->
// all-gather(param) by peer stores.  Replaces  all_reduce(grad); grad /= W; masked SGD  (3 passes over the arena + NCCL)
// with ONE kernel: every rank reads its 1/W shard of every peer's gradient arena (fixed rank order: bit-identical
// replicas), updates that shard of the weights with the same arithmetic as k_masked_sgd, and writes the new weights into
// every peer's parameter arena.  The momentum buffer exists only for the owned shard (ZeRO-1 style).
// Cross-rank ordering (gradients complete before, weights visible after) is the caller's job: a symmetric-memory
// barrier on the stream on both sides.
// ------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 8;
struct PeerPtrs {
  float *p[kMaxPeers];
  const float *g[kMaxPeers];
};
__global__ void __launch_bounds__(kThreads) k_dp_masked_sgd(PeerPtrs peers, int world, float *__restrict__ v_shard,
                                                            const uint32_t *__restrict__ bits, int64_t lo, int64_t hi,
                                                            float lr, float mu, float wd, float inv_world) {
  int64_t stride = (int64_t)gridDim.x * kThreads * 4;
  for (int64_t i = lo + ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4; i < hi; i += stride) {
    const uint32_t nib = mask_nibble(bits, i);
    float4 vv = load4(v_shard, i - lo, hi - lo, 0.f);
    if (nib == 0u) {
      store4(v_shard, i - lo, hi - lo, make_float4(0.f, 0.f, 0.f, 0.f));
      continue;  // masked-out: weights untouched on every replica
    }
    float4 gg = load4(peers.g[0], i, hi, 0.f);
    for (int r = 1; r < world; ++r) {
      const float4 t = load4(peers.g[r], i, hi, 0.f);
      gg.x = __fadd_rn(gg.x, t.x); gg.y = __fadd_rn(gg.y, t.y); gg.z = __fadd_rn(gg.z, t.z); gg.w = __fadd_rn(gg.w, t.w);
    }
    gg.x = __fmul_rn(gg.x, inv_world); gg.y = __fmul_rn(gg.y, inv_world);
    gg.z = __fmul_rn(gg.z, inv_world); gg.w = __fmul_rn(gg.w, inv_world);
    float4 pp = load4(peers.p[0], i, hi, 0.f);  // all replicas hold the same weights; read the first pointer (= local)
    sgd1(pp.x, gg.x, vv.x, nib & 1u, lr, mu, wd);
    sgd1(pp.y, gg.y, vv.y, nib & 2u, lr, mu, wd);
    sgd1(pp.z, gg.z, vv.z, nib & 4u, lr, mu, wd);
    sgd1(pp.w, gg.w, vv.w, nib & 8u, lr, mu, wd);
    store4(v_shard, i - lo, hi - lo, vv);
    for (int r = 0; r < world; ++r) store4(peers.p[r], i, hi, pp);
  }
}

// ------------------------------------------------------------------------------------------
// grad norm (double partials, fixed reduction tree -> deterministic), clip coefficient
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_sumsq_partial(const float *__restrict__ g, int64_t n,
                                                            double *__restrict__ partials) {
  __shared__ double sh[kThreads / 32];
  double s = 0.0;
  int64_t stride = (int64_t)gridDim.x * kThreads * 4;
  for (int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4; i < n; i += stride) {
    float4 v = load4(g, i, n, 0.f);
    s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += sh[w];
    partials[blockIdx.x] = t;
  }
}
// FT_l1 (Classification/unlearn/FT.py:13-17,133-134): loss += alpha * ||theta||_1  =>  g += alpha * sign(theta); the
// penalty value sum |theta| comes out of the same pass (double partials, fixed tree).  12 B/param.
__global__ void __launch_bounds__(kThreads) k_l1_penalty_grad(const float *__restrict__ p, float *__restrict__ g, int64_t n,
                                                              float alpha, double *__restrict__ partials) {
  __shared__ double sh[kThreads / 32];
  double s = 0.0;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const float v = p[i];
    s += (double)fabsf(v);
    const float sg = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);   // torch.sign: 0 at 0 (the subgradient autograd uses)
    g[i] = __fadd_rn(g[i], __fmul_rn(alpha, sg));
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += sh[w];
    partials[blockIdx.x] = t;
  }
}
__global__ void k_sumsq_final(const double *__restrict__ partials, int nparts, double *__restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partials[i];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    *out = t;
  }
}
__global__ void k_clip_coef(const double *__restrict__ sumsq, float max_norm, float *__restrict__ coef) {
  float total = (float)sqrt(*sumsq);
  float c = __fdiv_rn(max_norm, __fadd_rn(total, 1e-6f));
  *coef = c > 1.0f ? 1.0f : c;
}

// ------------------------------------------------------------------------------------------
// fused clip + mask + Adam            (oracle_masked_adam_step)
// bytes/param: read p,g,m1,m2 (16) + 1/8 + write p,m1,m2 (12) = 28.125 (+4 for the norm pass)
// ------------------------------------------------------------------------------------------
struct AdamK {
  float lr_step, bc2_sqrt, b2, one_m_b1, one_m_b2, eps, wd;
};
__device__ __forceinline__ void adam1(float &p, float g, float &m1, float &m2, bool m, float coef, const AdamK &k) {
  float gi = __fmul_rn(g, coef);
  if (!m) gi = __fmul_rn(0.f, gi);
  if (k.wd != 0.f) gi = __fadd_rn(gi, __fmul_rn(k.wd, p));
  float a = __fadd_rn(m1, __fmul_rn(__fadd_rn(gi, -m1), k.one_m_b1));
  float b = __fadd_rn(__fmul_rn(m2, k.b2), __fmul_rn(k.one_m_b2, __fmul_rn(gi, gi)));
  m1 = a;
  m2 = b;
  float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(b), k.bc2_sqrt), k.eps);
  p = __fadd_rn(p, __fmul_rn(-k.lr_step, __fdiv_rn(a, denom)));
}
__global__ void __launch_bounds__(kThreads) k_masked_adam(float *__restrict__ p, const float *__restrict__ g,
                                                          float *__restrict__ m1, float *__restrict__ m2,
                                                          const uint32_t *__restrict__ bits, int64_t n, AdamK k,
                                                          const float *__restrict__ coef_dev) {
  const float coef = coef_dev ? *coef_dev : 1.0f;
  int64_t stride = (int64_t)gridDim.x * kThreads * 4;
  for (int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4; i < n; i += stride) {
    uint32_t nib = mask_nibble(bits, i);
    float4 pp = load4(p, i, n, 0.f), gg = load4(g, i, n, 0.f), aa = load4(m1, i, n, 0.f), bb = load4(m2, i, n, 0.f);
    adam1(pp.x, gg.x, aa.x, bb.x, nib & 1u, coef, k);
    adam1(pp.y, gg.y, aa.y, bb.y, nib & 2u, coef, k);
    adam1(pp.z, gg.z, aa.z, bb.z, nib & 4u, coef, k);
    adam1(pp.w, gg.w, aa.w, bb.w, nib & 8u, coef, k);
    store4(p, i, n, pp);
    store4(m1, i, n, aa);
    store4(m2, i, n, bb);
  }
}

// ------------------------------------------------------------------------------------------
// data-parallel clip + mask + Adam over NVLink peer memory (ZeRO-1 style, SURVEY.md section 7.3 option (a)):
//   phase 1 (k_dp_reduce_sumsq): g_shard = sum_r grad_r[lo:hi] (peer loads, rank order) ; per-CTA partials of sum g_shard^2
//   phase 2 (k_dp_masked_adam) : total = sum_r norm_slot_r (every rank's shard norm, peer loads, rank order) ->
//                                clip coefficient -> mask (.) Adam on the shard -> new weights stored into every replica
// ------------------------------------------------------------------------------------------
struct PeerG {
  const float *g[kMaxPeers];
};
__global__ void __launch_bounds__(kThreads) k_dp_reduce_sumsq(PeerG peers, int world, float *__restrict__ g_shard,
                                                              int64_t lo, int64_t hi, float inv_world,
                                                              double *__restrict__ partials) {
  __shared__ double sh[kThreads / 32];
  double s = 0.0;
  int64_t stride = (int64_t)gridDim.x * kThreads * 4;
  for (int64_t i = lo + ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4; i < hi; i += stride) {
    float4 gg = load4(peers.g[0], i, hi, 0.f);
    for (int r = 1; r < world; ++r) {
      const float4 t = load4(peers.g[r], i, hi, 0.f);
      gg.x = __fadd_rn(gg.x, t.x); gg.y = __fadd_rn(gg.y, t.y); gg.z = __fadd_rn(gg.z, t.z); gg.w = __fadd_rn(gg.w, t.w);
    }
    gg.x = __fmul_rn(gg.x, inv_world); gg.y = __fmul_rn(gg.y, inv_world);
    gg.z = __fmul_rn(gg.z, inv_world); gg.w = __fmul_rn(gg.w, inv_world);
    store4(g_shard, i - lo, hi - lo, gg);
    s += (double)gg.x * gg.x + (double)gg.y * gg.y + (double)gg.z * gg.z + (double)gg.w * gg.w;
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += sh[w];
    partials[blockIdx.x] = t;
  }
}
struct PeerAdam {
  float *p[kMaxPeers];
  const double *norm[kMaxPeers];  // every rank's slot holding its shard's sum of squares
};
__global__ void __launch_bounds__(kThreads) k_dp_masked_adam(PeerAdam peers, int world, const float *__restrict__ g_shard,
                                                             float *__restrict__ m1, float *__restrict__ m2,
                                                             const uint32_t *__restrict__ bits, int64_t lo, int64_t hi,
                                                             AdamK k, float max_norm, float *__restrict__ coef_out) {
  float coef = 1.0f;
  if (max_norm > 0.f) {  // the same scalar arithmetic on every thread of every rank: identical replicas
    double tot = 0.0;
    for (int r = 0; r < world; ++r) tot += *peers.norm[r];
    const float c = __fdiv_rn(max_norm, __fadd_rn((float)sqrt(tot), 1e-6f));
    coef = c > 1.0f ? 1.0f : c;
    if (coef_out && blockIdx.x == 0 && threadIdx.x == 0) {
      coef_out[0] = coef;
      coef_out[1] = (float)sqrt(tot);
    }
  }
  int64_t stride = (int64_t)gridDim.x * kThreads * 4;
  for (int64_t i = lo + ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4; i < hi; i += stride) {
    const uint32_t nib = mask_nibble(bits, i);
    float4 pp = load4(peers.p[0], i, hi, 0.f), gg = load4(g_shard, i - lo, hi - lo, 0.f);
    float4 aa = load4(m1, i - lo, hi - lo, 0.f), bb = load4(m2, i - lo, hi - lo, 0.f);
    adam1(pp.x, gg.x, aa.x, bb.x, nib & 1u, coef, k);
    adam1(pp.y, gg.y, aa.y, bb.y, nib & 2u, coef, k);
    adam1(pp.z, gg.z, aa.z, bb.z, nib & 4u, coef, k);
    adam1(pp.w, gg.w, aa.w, bb.w, nib & 8u, coef, k);
    store4(m1, i - lo, hi - lo, aa);
    store4(m2, i - lo, hi - lo, bb);
    for (int r = 0; r < world; ++r) store4(peers.p[r], i, hi, pp);
  }
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace salun

using namespace salun;

// ============================================================================================
// C ABI
// ============================================================================================
extern "C" {

int salun_version(void) { return 1000; }
long long salun_launch_count(void) { return ::salun::g_launch_count; }
const char *salun_last_error(void) { return err_buf(); }

int salun_ctx_create(int device, salun_ctx **out) {
  SALUN_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  int ndev = 0;
  SALUN_CUDA_OK(cudaGetDeviceCount(&ndev));
  SALUN_REQUIRE(device >= 0 && device < ndev, "device index out of range");
  SALUN_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  SALUN_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("libsalun is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    return SALUN_ERR_UNSUPPORTED;
  }
  salun_ctx *c = new salun_ctx();
  memset(c, 0, sizeof *c);
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  SALUN_CUDA_OK(cudaMalloc(&c->hist, kRadixBins * sizeof(unsigned int)));
  SALUN_CUDA_OK(cudaMalloc(&c->sel, 8 * sizeof(unsigned long long)));
  SALUN_CUDA_OK(cudaMalloc(&c->block_ties, (kMaxPartials + 1) * sizeof(unsigned int)));
  SALUN_CUDA_OK(cudaMalloc(&c->partials, kMaxPartials * sizeof(double)));
  SALUN_CUDA_OK(cudaMallocHost(&c->mailbox_host, 8 * sizeof(unsigned long long)));
  *out = c;
  return SALUN_OK;
}

int salun_ctx_destroy(salun_ctx *c) {
  if (!c) return SALUN_OK;
  cudaSetDevice(c->device);
  cudaFree(c->hist);
  cudaFree(c->sel);
  cudaFree(c->block_ties);
  cudaFree(c->partials);
  cudaFreeHost(c->mailbox_host);
  if (c->msel) {
    cudaFree(c->msel);
    cudaFree(c->mhist);
    cudaFree(c->mties);
    cudaFreeHost(c->mmailbox_host);
  }
  delete c;
  return SALUN_OK;
}

#define SALUN_ENTER(ctx)                        \
  SALUN_REQUIRE((ctx) != nullptr, "ctx is NULL"); \
  SALUN_CUDA_OK(cudaSetDevice((ctx)->device));  \
  cudaStream_t st = (cudaStream_t)stream

int salun_saliency_accumulate_flat(salun_ctx *ctx, const float *grad, float *accum, int64_t n,
                                   const float *scale_dev, void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(n >= 0, "n < 0");
  if (n == 0) return SALUN_OK;
  SALUN_REQUIRE(grad && accum, "NULL buffer");
  SALUN_REQUIRE(aligned16(grad) && aligned16(accum), "buffers must be 16-byte aligned");
  { k_accum_flat<<<grid_for(ctx, (n + 3) / 4), kThreads, 0, st>>>(grad, accum, n, scale_dev); ++::salun::g_launch_count; }
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_saliency_accumulate(salun_ctx *ctx, const float *const *grads_host, const int64_t *numels_host,
                              int n_tensors, float *accum_flat, const float *scale_dev, void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(n_tensors >= 0, "n_tensors < 0");
  if (n_tensors == 0) return SALUN_OK;
  SALUN_REQUIRE(grads_host && numels_host && accum_flat, "NULL buffer");
  long long off = 0;
  int t = 0;
  while (t < n_tensors) {
    TensorTable tab;
    int cnt = 0;
    long long maxn = 0;
    while (t < n_tensors && cnt < kTabMax) {
      SALUN_REQUIRE(numels_host[t] >= 0, "negative numel");
      if (numels_host[t] > 0) {
        SALUN_REQUIRE(grads_host[t] != nullptr, "NULL grad pointer");
        tab.src[cnt] = grads_host[t];
        tab.dst_off[cnt] = off;
        tab.numel[cnt] = numels_host[t];
        if (numels_host[t] > maxn) maxn = numels_host[t];
        ++cnt;
      }
      off += numels_host[t];
      ++t;
    }
    if (cnt == 0) continue;
    long long gx = (maxn + kThreads * 8 - 1) / (kThreads * 8);
    if (gx > 128) gx = 128;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)cnt);
    { k_accum_multi<<<grid, kThreads, 0, st>>>(tab, accum_flat, scale_dev); ++::salun::g_launch_count; }
    SALUN_CUDA_OK(cudaGetLastError());
  }
  return SALUN_OK;
}

int salun_abs_inplace(salun_ctx *ctx, float *a, int64_t n, void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(n >= 0, "n < 0");
  if (n == 0) return SALUN_OK;
  SALUN_REQUIRE(a && aligned16(a), "buffer must be non-NULL and 16-byte aligned");
  { k_abs<<<grid_for(ctx, (n + 3) / 4), kThreads, 0, st>>>(a, n); ++::salun::g_launch_count; }
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_topk_mask(salun_ctx *ctx, const float *accum, int64_t n, int64_t k, int64_t *mask_i64,
                    uint32_t *mask_bits, salun_topk_info *info_host, void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(n >= 0 && k >= 0, "n or k negative");
  if (info_host) memset(info_host, 0, sizeof *info_host);
  if (n == 0) return SALUN_OK;
  SALUN_REQUIRE(accum && aligned16(accum), "accum must be non-NULL and 16-byte aligned");
  SALUN_REQUIRE(mask_i64 || mask_bits, "no output requested");
  const int grid = grid_for(ctx, (n + 3) / 4);
  // contiguous chunk per block, multiple of 1024 so that warps own whole mask words
  int64_t chunk = (n + grid - 1) / grid;
  chunk = (chunk + 1023) / 1024 * 1024;
  const int wgrid = (int)((n + chunk - 1) / chunk);
  const bool vec64 = mask_i64 && aligned16(mask_i64);
  if (k == 0) {
    if (mask_i64) SALUN_CUDA_OK(cudaMemsetAsync(mask_i64, 0, (size_t)n * 8, st));
    if (mask_bits) SALUN_CUDA_OK(cudaMemsetAsync(mask_bits, 0, (size_t)((n + 31) / 32) * 4, st));
    if (info_host) {
      SALUN_CUDA_OK(cudaStreamSynchronize(st));
      info_host->thr_key = 0xffffffffu;
      info_host->thr_value = INFINITY;
    }
    return SALUN_OK;
  }
  const int all_ones = k >= n;
  if (!all_ones) {
    { k_sel_init<<<(kRadixBins + 255) / 256, 256, 0, st>>>(ctx->sel, ctx->hist, (long long)k); ++::salun::g_launch_count; }
    for (int pass = 0; pass < 3; ++pass) {
      { k_radix_hist<<<grid, kThreads, 0, st>>>(accum, n, ctx->sel, ctx->hist, pass); ++::salun::g_launch_count; }
      { k_radix_pick<<<1, 32, 0, st>>>(ctx->hist, ctx->sel, pass, (long long)k); ++::salun::g_launch_count; }
    }
    { k_tie_count<<<wgrid, kThreads, 0, st>>>(accum, n, chunk, ctx->sel, ctx->block_ties); ++::salun::g_launch_count; }
    { k_tie_scan<<<1, 32, 0, st>>>(ctx->block_ties, wgrid, ctx->sel); ++::salun::g_launch_count; }
  }
  if (vec64)
    { k_write_mask<true><<<wgrid, kThreads, 0, st>>>(accum, n, chunk, ctx->sel, ctx->block_ties,
                                                   (long long *)mask_i64, mask_bits, all_ones); ++::salun::g_launch_count; }
  else
    { k_write_mask<false><<<wgrid, kThreads, 0, st>>>(accum, n, chunk, ctx->sel, ctx->block_ties,
                                                    (long long *)mask_i64, mask_bits, all_ones); ++::salun::g_launch_count; }
  SALUN_CUDA_OK(cudaGetLastError());
  if (info_host) {
    if (all_ones) {
      SALUN_CUDA_OK(cudaStreamSynchronize(st));
      info_host->thr_key = 0;
      info_host->thr_value = 0.f;
      info_host->n_greater = n;
      info_host->n_equal = 0;
    } else {
      SALUN_CUDA_OK(cudaMemcpyAsync(ctx->mailbox_host, ctx->sel, 8 * sizeof(unsigned long long),
                                    cudaMemcpyDeviceToHost, st));
      SALUN_CUDA_OK(cudaStreamSynchronize(st));
      info_host->thr_key = (uint32_t)ctx->mailbox_host[3];
      uint32_t vb = info_host->thr_key ? info_host->thr_key - 1u : 0x7fc00000u;
      memcpy(&info_host->thr_value, &vb, 4);
      info_host->n_greater = (int64_t)ctx->mailbox_host[4];
      info_host->n_equal = (int64_t)ctx->mailbox_host[5];
    }
  }
  return SALUN_OK;
}

int salun_topk_mask_multi(salun_ctx *ctx, const float *accum, int64_t n, const int64_t *ks_host, int n_ratios,
                          int64_t *const *mask_i64_host, uint32_t *const *mask_bits_host, salun_topk_info *infos_host,
                          void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(n >= 0 && n_ratios >= 0 && n_ratios <= kMaxRatios, "n < 0 or more than 16 ratios");
  if (infos_host) memset(infos_host, 0, sizeof(salun_topk_info) * (size_t)n_ratios);
  if (n == 0 || n_ratios == 0) return SALUN_OK;
  SALUN_REQUIRE(ks_host && (mask_i64_host || mask_bits_host), "NULL argument");
  SALUN_REQUIRE(accum && aligned16(accum), "accum must be non-NULL and 16-byte aligned");
  if (!ctx->msel) {
    SALUN_CUDA_OK(cudaMalloc(&ctx->msel, kMaxRatios * 8 * sizeof(unsigned long long)));
    SALUN_CUDA_OK(cudaMalloc(&ctx->mhist, (size_t)kMaxRatios * kRadixBins * sizeof(unsigned int)));
    SALUN_CUDA_OK(cudaMalloc(&ctx->mties, (size_t)kMaxRatios * (kMaxPartials + 1) * sizeof(unsigned int)));
    SALUN_CUDA_OK(cudaMallocHost(&ctx->mmailbox_host, kMaxRatios * 8 * sizeof(unsigned long long)));
  }
  MultiK mk{};
  MultiOut out{};
  mk.R = n_ratios;
  bool any_select = false;
  for (int r = 0; r < n_ratios; ++r) {
    SALUN_REQUIRE(ks_host[r] >= 0, "negative k");
    mk.k[r] = ks_host[r];
    mk.mode[r] = ks_host[r] == 0 ? 2 : (ks_host[r] >= n ? 1 : 0);
    any_select |= mk.mode[r] == 0;
    out.m64[r] = mask_i64_host ? (long long *)mask_i64_host[r] : nullptr;
    out.bits[r] = mask_bits_host ? mask_bits_host[r] : nullptr;
    out.vec64[r] = out.m64[r] && aligned16(out.m64[r]);
    SALUN_REQUIRE(out.m64[r] || out.bits[r], "a ratio without an output");
  }
  const int grid = grid_for(ctx, (n + 3) / 4);
  int64_t chunk = (n + grid - 1) / grid;
  chunk = (chunk + 1023) / 1024 * 1024;
  const int wgrid = (int)((n + chunk - 1) / chunk);
  if (any_select) {
    static bool attr_set = false;
    const size_t smem_full = ((size_t)kMaxRatios * kRadixBins + kRadixBins + kMaxRatios) * sizeof(unsigned int);
    if (!attr_set) {
      SALUN_CUDA_OK(cudaFuncSetAttribute(k_mhist<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_full));
      SALUN_CUDA_OK(cudaFuncSetAttribute(k_mhist<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_full));
      attr_set = true;
    }
    const size_t smem0 = ((size_t)kRadixBins * 2 + kMaxRatios) * sizeof(unsigned int);
    const size_t smemR = ((size_t)n_ratios * kRadixBins + kRadixBins + kMaxRatios) * sizeof(unsigned int);
    k_msel_init<<<32, 256, 0, st>>>(ctx->msel, ctx->mhist, mk);
    k_mhist<0><<<grid, kThreads, smem0, st>>>(accum, n, ctx->msel, ctx->mhist, mk);
    k_mpick<<<1, kMaxRatios * 32, 0, st>>>(ctx->mhist, ctx->msel, 0, mk);
    k_mhist<1><<<grid, kThreads, smemR, st>>>(accum, n, ctx->msel, ctx->mhist, mk);
    k_mpick<<<1, kMaxRatios * 32, 0, st>>>(ctx->mhist, ctx->msel, 1, mk);
    k_mhist<2><<<grid, kThreads, smemR, st>>>(accum, n, ctx->msel, ctx->mhist, mk);
    k_mpick<<<1, kMaxRatios * 32, 0, st>>>(ctx->mhist, ctx->msel, 2, mk);
    k_mtie_count<<<wgrid, kThreads, 0, st>>>(accum, n, chunk, ctx->msel, ctx->mties, mk);
    k_mtie_scan<<<1, 32, 0, st>>>(ctx->mties, wgrid, ctx->msel, mk);
    ::salun::g_launch_count += 9;
  }
  k_mwrite<<<wgrid, kThreads, 0, st>>>(accum, n, chunk, ctx->msel, ctx->mties, mk, out);
  ++::salun::g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  if (infos_host) {
    if (any_select)
      SALUN_CUDA_OK(cudaMemcpyAsync(ctx->mmailbox_host, ctx->msel, (size_t)n_ratios * 8 * sizeof(unsigned long long),
                                    cudaMemcpyDeviceToHost, st));
    SALUN_CUDA_OK(cudaStreamSynchronize(st));
    for (int r = 0; r < n_ratios; ++r) {
      salun_topk_info &inf = infos_host[r];
      if (mk.mode[r] == 2) {
        inf.thr_key = 0xffffffffu;
        inf.thr_value = INFINITY;
      } else if (mk.mode[r] == 1) {
        inf.thr_key = 0;
        inf.thr_value = 0.f;
        inf.n_greater = n;
      } else {
        const unsigned long long *m = ctx->mmailbox_host + r * 8;
        inf.thr_key = (uint32_t)m[3];
        uint32_t vb = inf.thr_key ? inf.thr_key - 1u : 0x7fc00000u;
        memcpy(&inf.thr_value, &vb, 4);
        inf.n_greater = (int64_t)m[4];
        inf.n_equal = (int64_t)m[5];
      }
    }
  }
  return SALUN_OK;
}

int salun_pack_mask(salun_ctx *ctx, const int64_t *mask_i64, int64_t n, uint32_t *mask_bits, void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(n >= 0, "n < 0");
  if (n == 0) return SALUN_OK;
  SALUN_REQUIRE(mask_i64 && mask_bits, "NULL buffer");
  { k_pack_mask<<<grid_for(ctx, n), kThreads, 0, st>>>((const long long *)mask_i64, n, mask_bits); ++::salun::g_launch_count; }
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_unpack_mask(salun_ctx *ctx, const uint32_t *mask_bits, int64_t n, int64_t *mask_i64, void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(n >= 0, "n < 0");
  if (n == 0) return SALUN_OK;
  SALUN_REQUIRE(mask_i64 && mask_bits, "NULL buffer");
  { k_unpack_mask<<<grid_for(ctx, n), kThreads, 0, st>>>(mask_bits, n, (long long *)mask_i64); ++::salun::g_launch_count; }
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_apply_mask(salun_ctx *ctx, float *g, const uint32_t *mask_bits, int64_t n, void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(n >= 0, "n < 0");
  if (n == 0) return SALUN_OK;
  SALUN_REQUIRE(g && mask_bits && aligned16(g), "g must be non-NULL, 16-byte aligned; mask_bits non-NULL");
  { k_apply_mask<<<grid_for(ctx, (n + 3) / 4), kThreads, 0, st>>>(g, mask_bits, n); ++::salun::g_launch_count; }
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_masked_sgd_step(salun_ctx *ctx, float *p, const float *g, float *v, const uint32_t *mask_bits,
                          int64_t n, float lr, float momentum, float wd, void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(n >= 0, "n < 0");
  if (n == 0) return SALUN_OK;
  SALUN_REQUIRE(p && g && v, "NULL buffer");
  SALUN_REQUIRE(aligned16(p) && aligned16(g) && aligned16(v), "p, g, v must be 16-byte aligned");
  { k_masked_sgd<<<grid_for(ctx, (n + 3) / 4), kThreads, 0, st>>>(p, g, v, mask_bits, n, lr, momentum, wd); ++::salun::g_launch_count; }
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_dp_shard(int64_t n, int rank, int world, int64_t *lo, int64_t *hi) {
  SALUN_REQUIRE(n >= 0 && world >= 1 && rank >= 0 && rank < world && lo && hi, "bad shard arguments");
  int64_t per = (n + world - 1) / world;
  per = (per + 127) / 128 * 128;  // whole mask words and 16-byte vectors per shard
  *lo = per * rank < n ? per * rank : n;
  *hi = per * (rank + 1) < n ? per * (rank + 1) : n;
  return SALUN_OK;
}

int salun_dp_masked_sgd_step(salun_ctx *ctx, float *const *param_peers_host, const float *const *grad_peers_host,
                             float *v_shard, const uint32_t *mask_bits, int64_t n, int rank, int world, float lr,
                             float momentum, float wd, void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(param_peers_host && grad_peers_host && v_shard, "NULL argument");
  SALUN_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "world must be 1..8");
  int64_t lo, hi;
  salun_dp_shard(n, rank, world, &lo, &hi);
  if (hi <= lo) return SALUN_OK;
  PeerPtrs pp;
  // local replica first: it is the one whose weights are read
  for (int r = 0; r < world; ++r) {
    const int src = (rank + r) % world;
    SALUN_REQUIRE(param_peers_host[src] && grad_peers_host[src], "NULL peer pointer");
    SALUN_REQUIRE(aligned16(param_peers_host[src]) && aligned16(grad_peers_host[src]), "peer arenas must be 16-byte aligned");
    pp.p[r] = param_peers_host[src];
  }
  for (int r = 0; r < world; ++r) pp.g[r] = grad_peers_host[r];  // gradients summed in rank order on every rank
  SALUN_REQUIRE(aligned16(v_shard), "v_shard must be 16-byte aligned");
  k_dp_masked_sgd<<<grid_for(ctx, (hi - lo + 3) / 4), kThreads, 0, st>>>(pp, world, v_shard, mask_bits, lo, hi, lr,
                                                                          momentum, wd, 1.0f / (float)world);
  ++::salun::g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_dp_grad_reduce_sumsq(salun_ctx *ctx, const float *const *grad_peers_host, float *g_shard, double *norm_slot,
                               int64_t n, int rank, int world, void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(grad_peers_host && g_shard && norm_slot, "NULL argument");
  SALUN_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "world must be 1..8");
  int64_t lo, hi;
  salun_dp_shard(n, rank, world, &lo, &hi);
  PeerG pg;
  for (int r = 0; r < world; ++r) {
    SALUN_REQUIRE(grad_peers_host[r] && aligned16(grad_peers_host[r]), "peer gradient arenas must be non-NULL and 16-byte aligned");
    pg.g[r] = grad_peers_host[r];  // summed in rank order on every rank
  }
  SALUN_REQUIRE(aligned16(g_shard), "g_shard must be 16-byte aligned");
  int grid = hi > lo ? grid_for(ctx, (hi - lo + 3) / 4) : 1;
  k_dp_reduce_sumsq<<<grid, kThreads, 0, st>>>(pg, world, g_shard, lo, hi > lo ? hi : lo, 1.0f / (float)world,
                                               ctx->partials);
  k_sumsq_final<<<1, 256, 0, st>>>(ctx->partials, grid, norm_slot);
  ::salun::g_launch_count += 2;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_dp_masked_adam_step(salun_ctx *ctx, float *const *param_peers_host, const double *const *norm_peers_host,
                              const float *g_shard, float *m1_shard, float *m2_shard, const uint32_t *mask_bits, int64_t n,
                              int rank, int world, float lr, float beta1, float beta2, float eps, float wd, int64_t step,
                              float max_norm, float *coef_norm_dev, void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(param_peers_host && norm_peers_host && g_shard && m1_shard && m2_shard, "NULL argument");
  SALUN_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "world must be 1..8");
  SALUN_REQUIRE(step >= 1, "step < 1");
  int64_t lo, hi;
  salun_dp_shard(n, rank, world, &lo, &hi);
  if (hi <= lo) return SALUN_OK;
  PeerAdam pa;
  for (int r = 0; r < world; ++r) {
    const int src = (rank + r) % world;  // local replica first: it is the one whose weights are read
    SALUN_REQUIRE(param_peers_host[src] && aligned16(param_peers_host[src]), "peer parameter arenas must be non-NULL and 16-byte aligned");
    pa.p[r] = param_peers_host[src];
    SALUN_REQUIRE(norm_peers_host[r], "NULL peer norm slot");
    pa.norm[r] = norm_peers_host[r];  // rank order
  }
  SALUN_REQUIRE(aligned16(g_shard) && aligned16(m1_shard) && aligned16(m2_shard), "shard buffers must be 16-byte aligned");
  AdamK k;
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  k.lr_step = (float)((double)lr / bc1);
  k.bc2_sqrt = (float)sqrt(bc2);
  k.b2 = beta2;
  k.one_m_b1 = (float)(1.0 - (double)beta1);
  k.one_m_b2 = (float)(1.0 - (double)beta2);
  k.eps = eps;
  k.wd = wd;
  k_dp_masked_adam<<<grid_for(ctx, (hi - lo + 3) / 4), kThreads, 0, st>>>(pa, world, g_shard, m1_shard, m2_shard, mask_bits,
                                                                           lo, hi, k, max_norm, coef_norm_dev);
  ++::salun::g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_grad_sumsq(salun_ctx *ctx, const float *g, int64_t n, double *sumsq_dev, void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(n >= 0 && sumsq_dev, "n < 0 or NULL output");
  if (n == 0) {
    SALUN_CUDA_OK(cudaMemsetAsync(sumsq_dev, 0, sizeof(double), st));
    return SALUN_OK;
  }
  SALUN_REQUIRE(g && aligned16(g), "g must be non-NULL and 16-byte aligned");
  const int grid = grid_for(ctx, (n + 3) / 4);
  { k_sumsq_partial<<<grid, kThreads, 0, st>>>(g, n, ctx->partials); ++::salun::g_launch_count; }
  { k_sumsq_final<<<1, 256, 0, st>>>(ctx->partials, grid, sumsq_dev); ++::salun::g_launch_count; }
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_l1_penalty_grad(salun_ctx *ctx, const float *p, float *g, int64_t n, float alpha, double *l1_dev, void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(n >= 0 && l1_dev, "n < 0 or NULL output");
  if (n == 0) {
    SALUN_CUDA_OK(cudaMemsetAsync(l1_dev, 0, sizeof(double), st));
    return SALUN_OK;
  }
  SALUN_REQUIRE(p && g, "NULL buffer");
  const int grid = grid_for(ctx, n);
  { k_l1_penalty_grad<<<grid, kThreads, 0, st>>>(p, g, n, alpha, ctx->partials); ++::salun::g_launch_count; }
  { k_sumsq_final<<<1, 256, 0, st>>>(ctx->partials, grid, l1_dev); ++::salun::g_launch_count; }
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_clip_coef(salun_ctx *ctx, const double *sumsq_dev, float max_norm, float *coef_dev, void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(sumsq_dev && coef_dev, "NULL buffer");
  { k_clip_coef<<<1, 1, 0, st>>>(sumsq_dev, max_norm, coef_dev); ++::salun::g_launch_count; }
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_masked_adam_step(salun_ctx *ctx, float *p, const float *g, float *m1, float *m2,
                           const uint32_t *mask_bits, int64_t n, float lr, float beta1, float beta2, float eps,
                           float wd, int64_t step, const float *coef_dev, void *stream) {
  SALUN_ENTER(ctx);
  SALUN_REQUIRE(n >= 0 && step >= 1, "n < 0 or step < 1");
  if (n == 0) return SALUN_OK;
  SALUN_REQUIRE(p && g && m1 && m2, "NULL buffer");
  SALUN_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m1) && aligned16(m2), "buffers must be 16-byte aligned");
  AdamK k;
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  k.lr_step = (float)((double)lr / bc1);
  k.bc2_sqrt = (float)sqrt(bc2);
  k.b2 = beta2;
  k.one_m_b1 = (float)(1.0 - (double)beta1);
  k.one_m_b2 = (float)(1.0 - (double)beta2);
  k.eps = eps;
  k.wd = wd;
  { k_masked_adam<<<grid_for(ctx, (n + 3) / 4), kThreads, 0, st>>>(p, g, m1, m2, mask_bits, n, k, coef_dev); ++::salun::g_launch_count; }
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

}  // extern "C"
