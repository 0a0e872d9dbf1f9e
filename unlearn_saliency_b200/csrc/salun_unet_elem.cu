// salun_unet_elem.cu -- HBM-bound kernels of the DDPM U-Net path (see salun_unet_elem.cuh).  bf16 storage, 8 channels
// (16 bytes) per thread, fp32 math, fp64 where a reduction feeds a variance.  Every kernel is deterministic (fixed
// reduction order, no atomics).
#include "salun_unet_elem.cuh"

#include <math.h>

#include "salun_common.cuh"

namespace salun {

typedef act_t bf16;  // activation element of this build (salun_act.cuh)

namespace {

constexpr int kT = 256;
constexpr int kU = 4;   // rows in flight per thread in the per-sample streaming kernels (memory-level parallelism)
constexpr int kUb = 2;  // ... in the register-heavy GroupNorm backward kernels

__device__ __forceinline__ void u_ld8(const bf16 *p, float (&f)[8]) { ld8(p, f); }
__device__ __forceinline__ avec u_ldraw(const bf16 *p) { return ldvec(p); }
__device__ __forceinline__ void u_cvt8(const avec &v, float (&f)[8]) { cvt8(v, f); }
__device__ __forceinline__ void u_st8(bf16 *p, const float (&f)[8]) { st8(p, f); }
__device__ __forceinline__ void st4(bf16 *p, const float4 &v) {   // 4 activation elements, p 4-element aligned
#ifdef SALUN_SPLIT
  *reinterpret_cast<uint4 *>(p) = make_uint4(act_pack_pair(v.x), act_pack_pair(v.y), act_pack_pair(v.z), act_pack_pair(v.w));
#else
  uint2 o;
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  o.x = *reinterpret_cast<uint32_t *>(&a);
  o.y = *reinterpret_cast<uint32_t *>(&b);
  *reinterpret_cast<uint2 *>(p) = o;
#endif
}
// element offset of pixel p (0 .. H*H-1) of image n, channel 0
__device__ __forceinline__ size_t pix_off(int n, int p, int H, int C, int flat) {
  if (flat) return ((size_t)n * H * H + p) * C;
  const int lh = 31 - __clz(H);  // H is a power of two
  const int y = p >> lh, x = p & (H - 1);
  return (((size_t)n * (H + 2) + y + 1) * (H + 2) + x + 1) * C;
}
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
// dropout multipliers of the 8 elements starting at flat index e0 = (n*HW + p)*C + c0 (a multiple of 8): four hashes,
// one 16-bit uniform per element, keep iff uniform >= thr16 = round(p * 65536).  The same function in forward and backward.
__device__ __forceinline__ void drop8(uint32_t seed, uint32_t e0, uint32_t thr16, float inv_keep, float (&m)[8]) {
  uint32_t h = hash32(e0 ^ seed);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m[2 * j] = (h & 0xffffu) >= thr16 ? inv_keep : 0.f;
    m[2 * j + 1] = (h >> 16) >= thr16 ? inv_keep : 0.f;
    h = hash32(h + 0x9E3779B9u);
  }
}
// sigmoid through one MUFU op (tanh.approx, rel. error ~2^-11: below the bf16 rounding of everything it feeds)
__device__ __forceinline__ float sigmoidf_(float x) {
#ifdef SALUN_SPLIT
  return 1.f / (1.f + expf(-x));  // the split build carries 16+ significand bits: full-precision transcendental
#endif
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}

// sums over the row lanes of a (vecs x rl) thread block: acc[k][8] per thread -> partial[k][C] of this (n, slice)
template <int K>
__device__ __forceinline__ void slice_reduce_store(float (&acc)[K][8], int v, int r, int rl, bool active, int C,
                                                   float *__restrict__ dst /* [K][C] */) {
  __shared__ float red[K * 2048];
  if (active) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) red[(k * rl + r) * C + v * 8 + i] = acc[k][i];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < K * C; j += kT) {
    const int k = j / C, c = j - k * C;
    float s = 0.f;
    for (int q = 0; q < rl; ++q) s += red[(k * rl + q) * C + c];
    dst[j] = s;
  }
}

}  // namespace

int unet_slices(int H) {
  int s = H * H / 64;
  if (s < 1) s = 1;
  if (s > 16) s = 16;
  return s;
}
// Slices actually launched for a batch of n: CTAs of 64 pixels are dominated by their fixed cost (launch, block
// reduction) -- measured 2 TB/s at n = 256 -- so use the fewest slices that still give ~4 CTAs per SM.
static int slices_for(int H, int n) {
  const int smax = unet_slices(H);
  int s = 1;
  while (s < smax && (long long)n * s < 592) s *= 2;
  return s;
}

// =================================================================================================================
// GroupNorm forward
// =================================================================================================================
__global__ void __launch_bounds__(kT) k_gn_stats(const bf16 *__restrict__ x, float *__restrict__ partial, int H, int C,
                                                 int S) {
  const int n = blockIdx.y, s = blockIdx.x;
  const int vecs = C >> 3, rl = kT / vecs;
  const int v = threadIdx.x % vecs, r = threadIdx.x / vecs;
  const bool active = r < rl;
  const int rps = H * H / S;
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (active) {
    const int end = (s + 1) * rps;
    for (int p = s * rps + r; p < end; p += rl * kU) {
      avec raw[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (p + u * rl < end) raw[u] = u_ldraw(x + pix_off(n, p + u * rl, H, C, 0) + v * 8);
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (p + u * rl >= end) continue;
        float f[8];
        u_cvt8(raw[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[0][i] += f[i];
          acc[1][i] += f[i] * f[i];
        }
      }
    }
  }
  slice_reduce_store<2>(acc, v, r, rl, active, C, partial + ((size_t)n * S + s) * 2 * C);
}
// group statistics of one sample from R rows of per-channel partial sums (sum x at psum[q*ld + c], sum x^2 at
// psq[q*ld + c]): the slice partials of k_gn_stats, or the per-(32-row group) column partials a producing GEMM's epilogue
// wrote.  Every CTA of the sample redoes this tiny reduction instead of a separate finalize launch; fp64 for the variance.
__device__ __forceinline__ void gn_group_stats(const float *__restrict__ psum, const float *__restrict__ psq, long long ld,
                                               int R, int C, float count, float eps, float (*sh_stat)[2], float *sh_a,
                                               float *sh_b) {
  const int cpg = C / kGnGroups;
  for (int c = threadIdx.x; c < C; c += kT) {
    float s1 = 0.f, s2 = 0.f;
    for (int q = 0; q < R; ++q) {
      s1 += psum[(size_t)q * ld + c];
      s2 += psq[(size_t)q * ld + c];
    }
    sh_a[c] = s1;
    sh_b[c] = s2;
  }
  __syncthreads();
  if (threadIdx.x < kGnGroups) {
    double a = 0.0, b = 0.0;
    for (int j = threadIdx.x * cpg; j < (threadIdx.x + 1) * cpg; ++j) {
      a += (double)sh_a[j];
      b += (double)sh_b[j];
    }
    const double mean = a / count;
    double var = b / count - mean * mean;  // biased variance, like torch.nn.GroupNorm
    if (var < 0.0) var = 0.0;
    sh_stat[threadIdx.x][0] = (float)mean;
    sh_stat[threadIdx.x][1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
}
__global__ void __launch_bounds__(kT) k_gn_apply(const bf16 *__restrict__ x, const float *__restrict__ psum,
                                                 const float *__restrict__ psq, long long pld, int R,
                                                 float *__restrict__ stats, const float *__restrict__ gamma,
                                                 const float *__restrict__ beta, bf16 *__restrict__ out, int out_flat,
                                                 int swish, uint32_t drop_thr, float inv_keep, uint32_t seed, int H,
                                                 int C, int S, float count, float eps) {
  __shared__ float sh_stat[kGnGroups][2], sh_a[512], sh_b[512];
  const int n = blockIdx.y, s = blockIdx.x;
  gn_group_stats(psum + (size_t)n * R * pld, psq + (size_t)n * R * pld, pld, R, C, count, eps, sh_stat, sh_a, sh_b);
  if (s == 0 && threadIdx.x < kGnGroups) {  // kept for the backward pass
    stats[((size_t)n * kGnGroups + threadIdx.x) * 2 + 0] = sh_stat[threadIdx.x][0];
    stats[((size_t)n * kGnGroups + threadIdx.x) * 2 + 1] = sh_stat[threadIdx.x][1];
  }
  const int vecs = C >> 3, rl = kT / vecs;
  const int v = threadIdx.x % vecs, r = threadIdx.x / vecs;
  if (r >= rl) return;
  const int cpg = C / kGnGroups, rps = H * H / S, c0 = v * 8;
  float a[8], b[8];  // y = a*x + b
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (c0 + i) / cpg;
    const float mean = sh_stat[g][0], rstd = sh_stat[g][1];
    a[i] = gamma[c0 + i] * rstd;
    b[i] = beta[c0 + i] - mean * a[i];
  }
  const int end = (s + 1) * rps;
  for (int p0 = s * rps + r; p0 < end; p0 += rl * kU) {
    avec raw[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (p0 + u * rl < end) raw[u] = u_ldraw(x + pix_off(n, p0 + u * rl, H, C, 0) + c0);
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int p = p0 + u * rl;
      if (p >= end) continue;
      float f[8];
      u_cvt8(raw[u], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float y = fmaf(a[i], f[i], b[i]);
        if (swish) y = y * sigmoidf_(y);
        f[i] = y;
      }
      if (drop_thr) {
        float m[8];
        drop8(seed, (uint32_t)(((size_t)n * H * H + p) * C + c0), drop_thr, inv_keep, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] *= m[i];
      }
      u_st8(out + pix_off(n, p, H, C, out_flat) + c0, f);
    }
  }
}
static inline uint32_t drop_threshold(float p) {
  if (p <= 0.f) return 0u;
  double t = (double)p * 65536.0 + 0.5;
  if (t > 65535.0) t = 65535.0;
  if (t < 1.0) t = 1.0;
  return (uint32_t)t;
}
void launch_gn_forward(const bf16 *x_pad, const float *ep_sum, const float *ep_sq, float *partial, float *stats,
                       const float *gamma, const float *beta, bf16 *out, int out_flat, int swish, float drop_p,
                       uint32_t drop_seed, int n, int H, int C, float eps, cudaStream_t st) {
  const int S = slices_for(H, n);
  const float *psum, *psq;
  long long pld;
  int R;
  if (ep_sum) {  // the producer's GEMM epilogue already reduced 32-row groups: H*H/32 partial rows per sample
    psum = ep_sum;
    psq = ep_sq;
    pld = C;
    R = H * H / 32;
  } else {
    k_gn_stats<<<dim3(S, n), kT, 0, st>>>(x_pad, partial, H, C, S);
    ++g_launch_count;
    psum = partial;
    psq = partial + C;
    pld = 2LL * C;
    R = S;
  }
  k_gn_apply<<<dim3(S, n), kT, 0, st>>>(x_pad, psum, psq, pld, R, stats, gamma, beta, out, out_flat, swish,
                                        drop_threshold(drop_p), drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f, drop_seed, H, C,
                                        S, (float)(C / kGnGroups) * H * H, eps);
  ++g_launch_count;
}
// cat_stat[r][0:Ca] = a_stat[r][:], cat_stat[r][Ca:] = b_stat[r][:]  (both planes) for r < rows
__global__ void k_concat_stats(const float *__restrict__ a1, const float *__restrict__ a2, int Ca,
                               const float *__restrict__ b1, const float *__restrict__ b2, int Cb, float *__restrict__ o1,
                               float *__restrict__ o2, long long total) {
  const int C = Ca + Cb;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    o1[i] = c < Ca ? a1[r * Ca + c] : b1[r * Cb + (c - Ca)];
    o2[i] = c < Ca ? a2[r * Ca + c] : b2[r * Cb + (c - Ca)];
  }
}
void launch_concat_stats(const float *a_sum, const float *a_sq, int Ca, const float *b_sum, const float *b_sq, int Cb,
                         float *o_sum, float *o_sq, long long rows, cudaStream_t st) {
  const long long total = rows * (Ca + Cb);
  long long g = (total + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  k_concat_stats<<<(int)g, 256, 0, st>>>(a_sum, a_sq, Ca, b_sum, b_sq, Cb, o_sum, o_sq, total);
  ++g_launch_count;
}

// =================================================================================================================
// GroupNorm backward
// =================================================================================================================
// dyh and xhat of 8 channels of one pixel
struct GnCh {
  float mean[8], rstd[8], gam[8], bet[8];
};
__device__ __forceinline__ void gn_load_ch(GnCh &ch, const float *stats, const float *gamma, const float *beta, int n,
                                           int c0, int cpg) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (c0 + i) / cpg;
    ch.mean[i] = stats[((size_t)n * kGnGroups + g) * 2];
    ch.rstd[i] = stats[((size_t)n * kGnGroups + g) * 2 + 1];
    ch.gam[i] = gamma[c0 + i];
    ch.bet[i] = beta[c0 + i];
  }
}
__device__ __forceinline__ void gn_dyh(const GnCh &ch, const float (&xf)[8], const float (&df)[8], int swish,
                                       uint32_t drop_thr, float inv_keep, uint32_t seed, uint32_t e0, float (&xhat)[8],
                                       float (&dyh)[8]) {
  float m[8];
  if (drop_thr) drop8(seed, e0, drop_thr, inv_keep, m);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    xhat[i] = (xf[i] - ch.mean[i]) * ch.rstd[i];
    float d = df[i];
    if (drop_thr) d *= m[i];
    if (swish) {
      const float y = fmaf(ch.gam[i], xhat[i], ch.bet[i]);
      const float sg = sigmoidf_(y);
      d *= sg * fmaf(y, 1.f - sg, 1.f);
    }
    dyh[i] = d;
  }
}
// reduces S1 = sum dyh, S2 = sum dyh * xhat and OVERWRITES dout with dyh (the apply pass then needs neither the
// activation derivative nor the dropout mask again)
__global__ void __launch_bounds__(kT) k_gn_bwd_reduce(bf16 *__restrict__ dout, const bf16 *__restrict__ x,
                                                      const float *__restrict__ stats, const float *__restrict__ gamma,
                                                      const float *__restrict__ beta, int swish, uint32_t drop_thr,
                                                      float inv_keep, uint32_t seed, float *__restrict__ partial, int H,
                                                      int C, int S) {
  const int n = blockIdx.y, s = blockIdx.x;
  const int vecs = C >> 3, rl = kT / vecs;
  const int v = threadIdx.x % vecs, r = threadIdx.x / vecs;
  const bool active = r < rl;
  const int cpg = C / kGnGroups, rps = H * H / S, c0 = v * 8;
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (active) {
    GnCh ch;
    gn_load_ch(ch, stats, gamma, beta, n, c0, cpg);
    const int end = (s + 1) * rps;
    for (int p0 = s * rps + r; p0 < end; p0 += rl * kUb) {
      avec rx[kUb], rd[kUb];
#pragma unroll
      for (int u = 0; u < kUb; ++u)
        if (p0 + u * rl < end) {
          rx[u] = u_ldraw(x + pix_off(n, p0 + u * rl, H, C, 0) + c0);
          rd[u] = u_ldraw(dout + pix_off(n, p0 + u * rl, H, C, 1) + c0);
        }
#pragma unroll
      for (int u = 0; u < kUb; ++u) {
        const int p = p0 + u * rl;
        if (p >= end) continue;
        float xf[8], df[8], xhat[8], dyh[8];
        u_cvt8(rx[u], xf);
        u_cvt8(rd[u], df);
        gn_dyh(ch, xf, df, swish, drop_thr, inv_keep, seed, (uint32_t)(((size_t)n * H * H + p) * C + c0), xhat, dyh);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[0][i] += dyh[i];
          acc[1][i] = fmaf(dyh[i], xhat[i], acc[1][i]);
        }
        if (swish || drop_thr) u_st8(dout + pix_off(n, p, H, C, 1) + c0, dyh);
      }
    }
  }
  slice_reduce_store<2>(acc, v, r, rl, active, C, partial + ((size_t)n * S + s) * 2 * C);
}
// per sample: slice sums -> persample[n][2][C] (S1, S2 per channel), and per group
// coef = (sum_c gamma_c S1[c], sum_c gamma_c S2[c]) / count
__global__ void __launch_bounds__(kT) k_gn_bwd_apply(const bf16 *__restrict__ dyh_flat, const bf16 *__restrict__ x,
                                                     const float *__restrict__ stats, const float *__restrict__ gamma,
                                                     const float *__restrict__ partial, float *__restrict__ persample,
                                                     bf16 *__restrict__ dx, int accumulate, int H, int C, int S,
                                                     float inv_count) {
  __shared__ float sh_coef[kGnGroups][2], sh_a[512], sh_b[512];
  const int n = blockIdx.y, s = blockIdx.x;
  const int cpg = C / kGnGroups;
  {
    // per-channel sums over the slices (S1 = sum dyh, S2 = sum dyh*xhat) -> persample (summed over samples later for
    // dbeta / dgamma) and the per-group coefficients (sum_c gamma S1, sum_c gamma S2) / count
    const float *pn = partial + (size_t)n * S * 2 * C;
    for (int c = threadIdx.x; c < C; c += kT) {
      float s1 = 0.f, s2 = 0.f;
      for (int q = 0; q < S; ++q) {
        s1 += pn[(size_t)q * 2 * C + c];
        s2 += pn[(size_t)q * 2 * C + C + c];
      }
      if (s == 0) {
        persample[((size_t)n * 2 + 0) * C + c] = s1;
        persample[((size_t)n * 2 + 1) * C + c] = s2;
      }
      sh_a[c] = gamma[c] * s1;
      sh_b[c] = gamma[c] * s2;
    }
    __syncthreads();
    if (threadIdx.x < kGnGroups) {
      float a = 0.f, b = 0.f;
      for (int j = threadIdx.x * cpg; j < (threadIdx.x + 1) * cpg; ++j) {
        a += sh_a[j];
        b += sh_b[j];
      }
      sh_coef[threadIdx.x][0] = a * inv_count;
      sh_coef[threadIdx.x][1] = b * inv_count;
    }
    __syncthreads();
  }
  const int vecs = C >> 3, rl = kT / vecs;
  const int v = threadIdx.x % vecs, r = threadIdx.x / vecs;
  if (r >= rl) return;
  const int rps = H * H / S, c0 = v * 8;
  // dx = rstd*gamma*dyh - rstd*cb*xhat - rstd*ca  with xhat = (x - mean)*rstd:  dx = k1*dyh + k2*x + k3
  float k1[8], k2[8], k3[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (c0 + i) / cpg;
    const float mean = stats[((size_t)n * kGnGroups + g) * 2], rstd = stats[((size_t)n * kGnGroups + g) * 2 + 1];
    const float ca = sh_coef[g][0], cb = sh_coef[g][1];
    k1[i] = rstd * gamma[c0 + i];
    k2[i] = -rstd * rstd * cb;
    k3[i] = rstd * (rstd * cb * mean - ca);
  }
  const int end = (s + 1) * rps;
  for (int p0 = s * rps + r; p0 < end; p0 += rl * kU) {
    avec rx[kU], rd[kU], ro[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (p0 + u * rl < end) {
        const size_t po = pix_off(n, p0 + u * rl, H, C, 0) + c0;
        rx[u] = u_ldraw(x + po);
        rd[u] = u_ldraw(dyh_flat + pix_off(n, p0 + u * rl, H, C, 1) + c0);
        ro[u] = accumulate ? u_ldraw(dx + po) : avec_zero();
      }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int p = p0 + u * rl;
      if (p >= end) continue;
      float xf[8], df[8], o[8];
      u_cvt8(rx[u], xf);
      u_cvt8(rd[u], df);
      u_cvt8(ro[u], o);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += fmaf(k1[i], df[i], fmaf(k2[i], xf[i], k3[i]));
      u_st8(dx + pix_off(n, p, H, C, 0) + c0, o);
    }
  }
}
void launch_gn_backward(bf16 *dout_flat, const bf16 *x_pad, const float *stats, const float *gamma, const float *beta,
                        int swish, float drop_p, uint32_t drop_seed, float *partial, float *persample, bf16 *dx_pad,
                        int accumulate, int n, int H, int C, cudaStream_t st) {
  const int S = slices_for(H, n);
  const uint32_t thr = drop_threshold(drop_p);
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  k_gn_bwd_reduce<<<dim3(S, n), kT, 0, st>>>(dout_flat, x_pad, stats, gamma, beta, swish, thr, inv_keep, drop_seed,
                                             partial, H, C, S);
  k_gn_bwd_apply<<<dim3(S, n), kT, 0, st>>>(dout_flat, x_pad, stats, gamma, partial, persample, dx_pad, accumulate, H, C,
                                            S, 1.f / ((float)(C / kGnGroups) * H * H));
  g_launch_count += 2;
}

// cross-sample sums of the whole backward pass in ONE launch: entry e sums rows of src into the gradient arena
__global__ void __launch_bounds__(1024) k_sum_rows_table(const SumEntry *__restrict__ tab, float *__restrict__ gbase) {
  __shared__ float red[2][32][33];
  const SumEntry e = tab[blockIdx.y];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if ((int)blockIdx.x * 32 >= e.C) return;
  const int c = blockIdx.x * 32 + tx;
  float s0 = 0.f, s1 = 0.f;
  if (c < e.C) {
    for (int i = ty; i < e.rows; i += 32) {
      s0 += e.src[(size_t)i * e.ld + c];
      if (e.K > 1) s1 += e.src[(size_t)i * e.ld + e.C + c];
    }
  }
  red[0][ty][tx] = s0;
  red[1][ty][tx] = s1;
  __syncthreads();
  if (ty == 0 && c < e.C) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      a += red[0][q][tx];
      b += red[1][q][tx];
    }
    if (e.d0 >= 0) gbase[e.d0 + c] = a;
    if (e.d0b >= 0) gbase[e.d0b + c] = a;
    if (e.d1 >= 0 && e.K > 1) gbase[e.d1 + c] = b;
  }
}
void launch_sum_rows_table(const SumEntry *table_dev, int n_entries, float *grad_base, cudaStream_t st) {
  if (n_entries <= 0) return;
  k_sum_rows_table<<<dim3(16, n_entries), 1024, 0, st>>>(table_dev, grad_base);
  ++g_launch_count;
}

// =================================================================================================================
// bias gradients
// =================================================================================================================
__global__ void __launch_bounds__(kT) k_colsum(const bf16 *__restrict__ dy, int flat, float *__restrict__ partial, int H,
                                               int C, int S) {
  const int n = blockIdx.y, s = blockIdx.x;
  const int vecs = C >> 3, rl = kT / vecs;
  const int v = threadIdx.x % vecs, r = threadIdx.x / vecs;
  const bool active = r < rl;
  const int rps = H * H / S;
  float acc[1][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = 0.f;
  if (active) {
    const int end = (s + 1) * rps;
    for (int p = s * rps + r; p < end; p += rl * kU) {
      avec raw[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (p + u * rl < end) raw[u] = u_ldraw(dy + pix_off(n, p + u * rl, H, C, flat) + v * 8);
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (p + u * rl >= end) continue;
        float f[8];
        u_cvt8(raw[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[0][i] += f[i];
      }
    }
  }
  slice_reduce_store<1>(acc, v, r, rl, active, C, partial + ((size_t)n * S + s) * C);
}
__global__ void k_rowsum_slices(const float *__restrict__ partial, float *__restrict__ rowsum, int ld, int col0, int C,
                                int S) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) a += partial[((size_t)n * S + s) * C + c];
    rowsum[(size_t)n * ld + col0 + c] = a;
  }
}
int launch_bias_partial(const bf16 *dy, int dy_flat, float *partial, float *rowsum, int rowsum_ld, int rowsum_col0, int n,
                        int H, int C, cudaStream_t st) {
  const int S = slices_for(H, n);
  k_colsum<<<dim3(S, n), kT, 0, st>>>(dy, dy_flat, partial, H, C, S);
  ++g_launch_count;
  if (rowsum) {
    k_rowsum_slices<<<n, 256, 0, st>>>(partial, rowsum, rowsum_ld, rowsum_col0, C, S);
    ++g_launch_count;
  }
  return n * S;
}
int unet_slices_for(int H, int n) { return slices_for(H, n); }

// =================================================================================================================
// copies
// =================================================================================================================
static inline int grid_for(long long total) {
  long long g = (total + kT - 1) / kT;
  if (g > 148 * 8) g = 148 * 8;
  return (int)(g < 1 ? 1 : g);
}
// i -> (pixel m in n*H*H, vector v of `vecs`)
__global__ void __launch_bounds__(kT) k_concat(const bf16 *__restrict__ a, int Ca, const bf16 *__restrict__ b, int Cb,
                                               bf16 *__restrict__ out, long long total, int H) {
  const int C = Ca + Cb, vecs = C >> 3, hw = H * H;
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
    const int v = (int)(i % vecs);
    const int m = (int)(i / vecs);
    const int n = m / hw, p = m - n * hw;
    const int c = v * 8;
    const avec val = c < Ca ? ldvec(a + pix_off(n, p, H, Ca, 0) + c) : ldvec(b + pix_off(n, p, H, Cb, 0) + (c - Ca));
    stvec(out + pix_off(n, p, H, C, 0) + c, val);
  }
}
void launch_concat(const bf16 *a_pad, int Ca, const bf16 *b_pad, int Cb, bf16 *out_pad, int n, int H, cudaStream_t st) {
  const long long total = (long long)n * H * H * ((Ca + Cb) >> 3);
  k_concat<<<grid_for(total), kT, 0, st>>>(a_pad, Ca, b_pad, Cb, out_pad, total, H);
  ++g_launch_count;
}
__global__ void __launch_bounds__(kT) k_split(const bf16 *__restrict__ dcat, bf16 *__restrict__ da, int Ca, int acc_a,
                                              bf16 *__restrict__ db, int Cb, int acc_b, long long total, int H) {
  const int C = Ca + Cb, vecs = C >> 3, hw = H * H;
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
    const int v = (int)(i % vecs);
    const int m = (int)(i / vecs);
    const int n = m / hw, p = m - n * hw;
    const int c = v * 8;
    float f[8];
    u_ld8(dcat + pix_off(n, p, H, C, 0) + c, f);
    bf16 *dst = c < Ca ? da + pix_off(n, p, H, Ca, 0) + c : db + pix_off(n, p, H, Cb, 0) + (c - Ca);
    if (c < Ca ? acc_a : acc_b) {
      float o[8];
      u_ld8(dst, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += o[j];
    }
    u_st8(dst, f);
  }
}
void launch_split(const bf16 *dcat_pad, bf16 *da_pad, int Ca, int acc_a, bf16 *db_pad, int Cb, int acc_b, int n, int H,
                  cudaStream_t st) {
  const long long total = (long long)n * H * H * ((Ca + Cb) >> 3);
  k_split<<<grid_for(total), kT, 0, st>>>(dcat_pad, da_pad, Ca, acc_a, db_pad, Cb, acc_b, total, H);
  ++g_launch_count;
}
__global__ void __launch_bounds__(kT) k_add_into(const bf16 *__restrict__ src, bf16 *__restrict__ dst, int accumulate,
                                                 long long nvec) {
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < nvec; i += (long long)gridDim.x * kT) {
    if (!accumulate) {
      stvec(dst + i * 8, ldvec(src + i * 8));
    } else {
      float a[8], b[8];
      u_ld8(src + i * 8, a);
      u_ld8(dst + i * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
      u_st8(dst + i * 8, a);
    }
  }
}
void launch_add_into(const bf16 *src, bf16 *dst, int accumulate, long long count, cudaStream_t st) {
  k_add_into<<<grid_for(count >> 3), kT, 0, st>>>(src, dst, accumulate, count >> 3);
  ++g_launch_count;
}
// out side 2H
__global__ void __launch_bounds__(kT) k_upsample2(const bf16 *__restrict__ in, bf16 *__restrict__ out, long long total,
                                                  int H, int C) {
  const int vecs = C >> 3, Ho = 2 * H, hwo = Ho * Ho;
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
    const int v = (int)(i % vecs);
    const int m = (int)(i / vecs);
    const int n = m / hwo, p = m - n * hwo;
    const int y = p / Ho, x = p - y * Ho;
    const avec val = ldvec(in + pix_off(n, (y >> 1) * H + (x >> 1), H, C, 0) + v * 8);
    stvec(out + pix_off(n, p, Ho, C, 0) + v * 8, val);
  }
}
void launch_upsample2(const bf16 *in_pad, bf16 *out_pad, int n, int H, int C, cudaStream_t st) {
  const long long total = (long long)n * 4 * H * H * (C >> 3);
  k_upsample2<<<grid_for(total), kT, 0, st>>>(in_pad, out_pad, total, H, C);
  ++g_launch_count;
}
__global__ void __launch_bounds__(kT) k_upsample2_bwd(const bf16 *__restrict__ dout, bf16 *__restrict__ din,
                                                      int accumulate, long long total, int H, int C) {
  const int vecs = C >> 3, Ho = 2 * H, hw = H * H;
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
    const int v = (int)(i % vecs);
    const int m = (int)(i / vecs);
    const int n = m / hw, p = m - n * hw;
    const int y = p / H, x = p - y * H;
    float s[8];
    bf16 *dst = din + pix_off(n, p, H, C, 0) + v * 8;
    if (accumulate) {
      u_ld8(dst, s);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = 0.f;
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float f[8];
        u_ld8(dout + pix_off(n, (2 * y + dy) * Ho + 2 * x + dx, Ho, C, 0) + v * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += f[j];
      }
    u_st8(dst, s);
  }
}
void launch_upsample2_bwd(const bf16 *dout_pad, bf16 *din_pad, int accumulate, int n, int H, int C, cudaStream_t st) {
  const long long total = (long long)n * H * H * (C >> 3);
  k_upsample2_bwd<<<grid_for(total), kT, 0, st>>>(dout_pad, din_pad, accumulate, total, H, C);
  ++g_launch_count;
}
// Downsample: out(oy, ox) = sum_{ky,kx} w[ky][kx] . in(2oy + ky, 2ox + kx), in(H, .) = in(., H) = 0: padded coordinate
// (2oy + ky + 1, 2ox + kx + 1), the last row / column falls into the zero halo.
__global__ void __launch_bounds__(kT) k_down_im2col(const bf16 *__restrict__ in, bf16 *__restrict__ col, long long total,
                                                    int H, int C) {
  const int vecs = C >> 3, Ho = H / 2, hwo = Ho * Ho, Hp = H + 2;
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
    const int v = (int)(i % vecs);
    long long t = i / vecs;
    const int tap = (int)(t % 9);
    const int mo = (int)(t / 9);
    const int n = mo / hwo, p = mo - n * hwo;
    const int oy = p / Ho, ox = p - oy * Ho;
    const int ky = tap / 3, kx = tap - ky * 3;
    const int py = 2 * oy + ky + 1, px = 2 * ox + kx + 1;
    const avec val = ldvec(in + (((size_t)n * Hp + py) * Hp + px) * C + v * 8);
    stvec(col + ((size_t)mo * 9 + tap) * C + v * 8, val);
  }
}
void launch_down_im2col(const bf16 *in_pad, bf16 *col, int n, int H, int C, cudaStream_t st) {
  const long long total = (long long)n * (H / 2) * (H / 2) * 9 * (C >> 3);
  k_down_im2col<<<grid_for(total), kT, 0, st>>>(in_pad, col, total, H, C);
  ++g_launch_count;
}
// gather form: input pixel (y, x) receives dcol[(oy, ox)][tap (ky, kx)] for every 2oy + ky == y, 2ox + kx == x
__global__ void __launch_bounds__(kT) k_down_col2im(const bf16 *__restrict__ dcol, bf16 *__restrict__ din,
                                                    int accumulate, long long total, int H, int C) {
  const int vecs = C >> 3, Ho = H / 2, hw = H * H;
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
    const int v = (int)(i % vecs);
    const int m = (int)(i / vecs);
    const int n = m / hw, p = m - n * hw;
    const int y = p / H, x = p - y * H;
    float s[8];
    bf16 *dst = din + pix_off(n, p, H, C, 0) + v * 8;
    if (accumulate) {
      u_ld8(dst, s);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = 0.f;
    }
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = y - ky;
      if (ty < 0 || (ty & 1)) continue;
      const int oy = ty >> 1;
      if (oy >= Ho) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = x - kx;
        if (tx < 0 || (tx & 1)) continue;
        const int ox = tx >> 1;
        if (ox >= Ho) continue;
        float f[8];
        u_ld8(dcol + (((size_t)n * Ho * Ho + oy * Ho + ox) * 9 + ky * 3 + kx) * C + v * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += f[j];
      }
    }
    u_st8(dst, s);
  }
}
void launch_down_col2im(const bf16 *dcol, bf16 *din_pad, int accumulate, int n, int H, int C, cudaStream_t st) {
  const long long total = (long long)n * H * H * (C >> 3);
  k_down_col2im<<<grid_for(total), kT, 0, st>>>(dcol, din_pad, accumulate, total, H, C);
  ++g_launch_count;
}

__global__ void __launch_bounds__(kT) k_eps_out(const float *__restrict__ y, const float *__restrict__ bias3,
                                                float *__restrict__ eps, int total, int hw) {
  const int i = blockIdx.x * kT + threadIdx.x;  // over n*3*hw, output order
  if (i >= total) return;
  const int p = i % hw, c = (i / hw) % 3, n = i / (3 * hw);
  eps[i] = y[((size_t)n * hw + p) * 64 + c] + bias3[c];
}
void launch_eps_out(const float *y, const float *bias3, float *eps_nchw, int n, int H, cudaStream_t st) {
  const int total = n * 3 * H * H;
  k_eps_out<<<(total + kT - 1) / kT, kT, 0, st>>>(y, bias3, eps_nchw, total, H * H);
  ++g_launch_count;
}
__global__ void __launch_bounds__(kT) k_eps_in(const float *__restrict__ deps, bf16 *__restrict__ dy, int total, int H) {
  const int m = blockIdx.x * kT + threadIdx.x;  // over n*hw pixels
  if (m >= total) return;
  const int hw = H * H, n = m / hw, p = m - n * hw;
  bf16 *dst = dy + pix_off(n, p, H, 64, 0);
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = 0.f;
  for (int c = 0; c < 3; ++c) f[c] = deps[((size_t)n * 3 + c) * hw + p];
  u_st8(dst, f);
}
__global__ void __launch_bounds__(256) k_eps_bias_partial(const float *__restrict__ deps, float *__restrict__ partial,
                                                          int hw) {
  __shared__ float red[256];
  const int c = blockIdx.x, i = blockIdx.y;
  float s = 0.f;
  for (int p = threadIdx.x; p < hw; p += 256) s += deps[((size_t)i * 3 + c) * hw + p];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(size_t)i * 3 + c] = red[0];
}
void launch_eps_in(const float *deps_nchw, bf16 *dy_pad, float *dbias_partial, int n, int H, cudaStream_t st) {
  const int total = n * H * H;
  k_eps_in<<<(total + kT - 1) / kT, kT, 0, st>>>(deps_nchw, dy_pad, total, H);
  k_eps_bias_partial<<<dim3(3, n), 256, 0, st>>>(deps_nchw, dbias_partial, H * H);
  g_launch_count += 2;
}

// =================================================================================================================
// attention: softmax rows (one warp per row), batched transposes
// =================================================================================================================
__global__ void __launch_bounds__(256) k_softmax(const float *__restrict__ S, bf16 *__restrict__ P, int M, int Te, int T,
                                                 float scale) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const int blk0 = ((row % Te) / T) * T;  // first key column of this row's sample inside its group
  const float *s = S + (size_t)row * Te;
  float v[8];
  const int per = Te / 32;  // 4 or 8
  float mx = -INFINITY;
  for (int i = 0; i < per; ++i) {
    const int j = lane + 32 * i;
    const bool ok = j >= blk0 && j < blk0 + T;
    v[i] = ok ? s[j] * scale : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int i = 0; i < per; ++i) {
    v[i] = v[i] == -INFINITY ? 0.f : (kSplit ? expf(v[i] - mx) : __expf(v[i] - mx));
    sum += v[i];
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  for (int i = 0; i < per; ++i) P[(size_t)row * Te + lane + 32 * i] = act_from_float(v[i] * inv);
}
void launch_softmax(const float *S, bf16 *P, int M, int Te, int T, float scale, cudaStream_t st) {
  k_softmax<<<(M + 7) / 8, 256, 0, st>>>(S, P, M, Te, T, scale);
  ++g_launch_count;
}
__global__ void __launch_bounds__(256) k_softmax_bwd(const float *__restrict__ dP, const bf16 *__restrict__ P,
                                                     bf16 *__restrict__ dS, int M, int Te, float scale) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const int per = Te / 32;
  float p[8], d[8], dot = 0.f;
  for (int i = 0; i < per; ++i) {
    const size_t j = (size_t)row * Te + lane + 32 * i;
    p[i] = act_to_float(P[j]);
    d[i] = dP[j];
    dot += p[i] * d[i];
  }
  for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  for (int i = 0; i < per; ++i)
    dS[(size_t)row * Te + lane + 32 * i] = act_from_float(scale * p[i] * (d[i] - dot));
}
void launch_softmax_bwd(const float *dP, const bf16 *P, bf16 *dS, int M, int Te, float scale, cudaStream_t st) {
  k_softmax_bwd<<<(M + 7) / 8, 256, 0, st>>>(dP, P, dS, M, Te, scale);
  ++g_launch_count;
}
// kWop: the result is a prepared weight-side operand (rows of logical length R, salun_act.cuh) instead of an activation
template <bool kWop>
__global__ void __launch_bounds__(256) k_transpose(const bf16 *__restrict__ in, int ld_in, void *__restrict__ out, int R,
                                                   int Cc) {
  __shared__ float tile[32][33];
  const int g = blockIdx.z, c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const bf16 *src = in + (size_t)g * R * ld_in;
#pragma unroll
  for (int j = 0; j < 4; ++j) tile[ty + 8 * j][tx] = act_to_float(src[(size_t)(r0 + ty + 8 * j) * ld_in + c0 + tx]);
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const size_t orow = (size_t)g * Cc + c0 + ty + 8 * j;
    if (kWop)
      wop_store(reinterpret_cast<wop_t *>(out) + orow * R * kWopK, r0 + tx, R, tile[tx][ty + 8 * j]);
    else
      reinterpret_cast<bf16 *>(out)[orow * R + r0 + tx] = act_from_float(tile[tx][ty + 8 * j]);
  }
}
void launch_transpose(const bf16 *in, int ld_in, bf16 *out, int R, int Cc, int G, cudaStream_t st) {
  k_transpose<false><<<dim3(Cc / 32, R / 32, G), 256, 0, st>>>(in, ld_in, out, R, Cc);
  ++g_launch_count;
}
void launch_transpose_wop(const bf16 *in, int ld_in, wop_t *out, int R, int Cc, int G, cudaStream_t st) {
  k_transpose<true><<<dim3(Cc / 32, R / 32, G), 256, 0, st>>>(in, ld_in, out, R, Cc);
  ++g_launch_count;
}
#ifdef SALUN_SPLIT
__global__ void __launch_bounds__(256) k_pack_wop(const bf16 *__restrict__ in, wop_t *__restrict__ out, long long total, int K) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long r = i / K;
    const int k = (int)(i - r * K);
    const bf16 v = in[i];   // the (hi, lo) pair is already split: no re-rounding
    wop_t *row = out + (size_t)r * K * kWopK;
    *reinterpret_cast<__nv_bfloat162 *>(row + 2 * k) = __nv_bfloat162(v.hi, v.hi);
    *reinterpret_cast<__nv_bfloat162 *>(row + 2 * (size_t)K + 2 * k) = __nv_bfloat162(v.lo, v.lo);
  }
}
#endif
const wop_t *launch_pack_wop(const bf16 *in, wop_t *out, long long rows, int K, cudaStream_t st) {
#ifdef SALUN_SPLIT
  const long long total = rows * K;
  long long g = (total + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  k_pack_wop<<<(int)g, 256, 0, st>>>(in, out, total, K);
  ++g_launch_count;
  return out;
#else
  (void)out; (void)rows; (void)K; (void)st;
  return in;
#endif
}

// =================================================================================================================
// embedding MLPs
// =================================================================================================================
__global__ void k_emb_inputs(const float *__restrict__ t, const int64_t *__restrict__ c, const uint8_t *__restrict__ drop,
                             const float *__restrict__ class_emb, const float *__restrict__ null_emb,
                             float *__restrict__ sincos, float *__restrict__ ce, int n, int ch, int n_classes) {
  const int i = blockIdx.x, d = threadIdx.x;
  if (d >= ch) return;
  const int half = ch / 2;
  const float k = -(float)(log(10000.0) / (double)(half - 1));
  const int j = d < half ? d : d - half;
  const float ang = t[i] * expf((float)j * k);
  sincos[(size_t)i * ch + d] = d < half ? sinf(ang) : cosf(ang);
  const bool dr = drop && drop[i];
  long long ci = c[i];  // an out-of-range label must not read outside the embedding table
  ci = ci < 0 ? 0 : (ci >= n_classes ? n_classes - 1 : ci);
  ce[(size_t)i * ch + d] = dr ? null_emb[d] : class_emb[(size_t)ci * ch + d];
}
void launch_emb_inputs(const float *t, const int64_t *c, const uint8_t *drop, const float *class_emb,
                       const float *null_emb, float *sincos, float *ce, int n, int ch, int n_classes, cudaStream_t st) {
  k_emb_inputs<<<n, ch, 0, st>>>(t, c, drop, class_emb, null_emb, sincos, ce, n, ch, n_classes);
  ++g_launch_count;
}
// 32 x 32 tile, 32-deep k steps, 2 x 2 outputs per thread: the embedding MLPs are 256 x 512 x 512 and smaller, so small
// tiles are what fills the SMs (128 CTAs instead of 32 for the 64 x 64 version, which took 41 us per launch)
__global__ void __launch_bounds__(256) k_sgemm(const float *__restrict__ A, long long sai, long long sak,
                                               const float *__restrict__ B, long long sbk, long long sbj,
                                               float *__restrict__ C, int ldc, int M, int N, int K,
                                               const float *__restrict__ bias, int accumulate) {
  __shared__ float As[32][33], Bs[32][33];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = threadIdx.x + 256 * q;
      // consecutive threads walk the dimension that is contiguous in memory
      int ai, ak;
      if (sak == 1) { ak = e & 31; ai = e >> 5; } else { ai = e & 31; ak = e >> 5; }
      As[ak][ai] = (i0 + ai < M && k0 + ak < K) ? A[(long long)(i0 + ai) * sai + (long long)(k0 + ak) * sak] : 0.f;
      int bj, bk;
      if (sbk == 1) { bk = e & 31; bj = e >> 5; } else { bj = e & 31; bk = e >> 5; }
      Bs[bk][bj] = (j0 + bj < N && k0 + bk < K) ? B[(long long)(k0 + bk) * sbk + (long long)(j0 + bj) * sbj] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float a0 = As[k][ty * 2], a1 = As[k][ty * 2 + 1], b0 = Bs[k][tx * 2], b1 = Bs[k][tx * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]);
      acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]);
      acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int i = i0 + ty * 2 + p;
    if (i >= M) continue;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int j = j0 + tx * 2 + q;
      if (j >= N) continue;
      float v = acc[p][q] + (bias ? bias[j] : 0.f);
      if (accumulate) v += C[(size_t)i * ldc + j];
      C[(size_t)i * ldc + j] = v;
    }
  }
}
void launch_sgemm(const float *A, long long sai, long long sak, const float *B, long long sbk, long long sbj, float *C,
                  int ldc, int M, int N, int K, const float *bias, int accumulate, cudaStream_t st) {
  k_sgemm<<<dim3((N + 31) / 32, (M + 31) / 32), 256, 0, st>>>(A, sai, sak, B, sbk, sbj, C, ldc, M, N, K, bias,
                                                             accumulate);
  ++g_launch_count;
}
__global__ void k_swish_f32(const float *__restrict__ in, float *__restrict__ out, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    const float x = in[i];
    out[i] = x / (1.f + expf(-x));
  }
}
void launch_swish_f32(const float *in, float *out, long long count, cudaStream_t st) {
  k_swish_f32<<<(int)((count + 255) / 256), 256, 0, st>>>(in, out, count);
  ++g_launch_count;
}
__global__ void k_dswish_f32(const float *__restrict__ dy, const float *__restrict__ pre, float *__restrict__ out,
                             long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    const float x = pre[i];
    const float sg = 1.f / (1.f + expf(-x));
    out[i] = dy[i] * sg * (1.f + x * (1.f - sg));
  }
}
void launch_dswish_f32(const float *dy, const float *pre, float *out, long long count, cudaStream_t st) {
  k_dswish_f32<<<(int)((count + 255) / 256), 256, 0, st>>>(dy, pre, out, count);
  ++g_launch_count;
}
__global__ void k_colsum_f32(const float *__restrict__ a, int ld, int rows, int cols, float *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  float s = 0.f;
  for (int i = 0; i < rows; ++i) s += a[(size_t)i * ld + j];
  out[j] = s;
}
void launch_colsum_f32(const float *a, int ld, int rows, int cols, float *out, cudaStream_t st) {
  k_colsum_f32<<<(cols + 127) / 128, 128, 0, st>>>(a, ld, rows, cols, out);
  ++g_launch_count;
}
__global__ void k_emb_scatter(const float *__restrict__ dce, const int64_t *__restrict__ c,
                              const uint8_t *__restrict__ drop, float *__restrict__ d_class_emb,
                              float *__restrict__ d_null, int n, int ch, int n_classes) {
  const int k = blockIdx.x, d = threadIdx.x;  // k == n_classes: the null embedding
  if (d >= ch) return;
  float s = 0.f;
  for (int i = 0; i < n; ++i) {
    const bool dr = drop && drop[i];
    const bool mine = k == n_classes ? dr : (!dr && (int)c[i] == k);
    if (mine) s += dce[(size_t)i * ch + d];
  }
  if (k == n_classes)
    d_null[d] = s;
  else
    d_class_emb[(size_t)k * ch + d] = s;
}
void launch_emb_scatter(const float *dce, const int64_t *c, const uint8_t *drop, float *d_class_emb, float *d_null, int n,
                        int ch, int n_classes, cudaStream_t st) {
  k_emb_scatter<<<n_classes + 1, ch, 0, st>>>(dce, c, drop, d_class_emb, d_null, n, ch, n_classes);
  ++g_launch_count;
}


// ---- temb/cemb projection of all ResnetBlocks as ONE tensor-core GEMM: operand staging ----
// Wcat[r][k] = bf16(params[row_w[r] + k]) ; bcat[r] = params[row_b[r]]   (r walks the blocks' output channels)
__global__ void __launch_bounds__(256) k_gather_proj(const float *__restrict__ params, const long long *__restrict__ row_w,
                                                     const long long *__restrict__ row_b, bf16 *__restrict__ wcat,
                                                     float *__restrict__ bcat, int K) {
  const int r = blockIdx.x;
  const float *src = params + row_w[r];
  for (int k = threadIdx.x * 4; k < K; k += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4 *>(src + k);
    st4(wcat + (size_t)r * K + k, v);
  }
  if (threadIdx.x == 0) bcat[r] = params[row_b[r]];
}
void launch_gather_proj(const float *params, const long long *row_w, const long long *row_b, bf16 *wcat, float *bcat,
                        int rows, int K, cudaStream_t st) {
  k_gather_proj<<<rows, 256, 0, st>>>(params, row_w, row_b, wcat, bcat, K);
  ++g_launch_count;
}
// dst[row_w[r] + k] = src[r][k]
__global__ void __launch_bounds__(256) k_scatter_rows(const float *__restrict__ src, const long long *__restrict__ row_w,
                                                      float *__restrict__ dst, int K) {
  const int r = blockIdx.x;
  float *d = dst + row_w[r];
  for (int k = threadIdx.x * 4; k < K; k += 256 * 4)
    *reinterpret_cast<float4 *>(d + k) = *reinterpret_cast<const float4 *>(src + (size_t)r * K + k);
}
void launch_scatter_rows(const float *src, const long long *row_w, float *dst, int rows, int K, cudaStream_t st) {
  k_scatter_rows<<<rows, 256, 0, st>>>(src, row_w, dst, K);
  ++g_launch_count;
}
// out[i*ld_out + j] = bf16(in[i*ld_in + j]), cols % 4 == 0
__global__ void __launch_bounds__(256) k_f32_to_bf16(const float *__restrict__ in, int ld_in, bf16 *__restrict__ out,
                                                     int ld_out, int rows, int cols) {
  const int c4 = cols >> 2;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < (long long)rows * c4; i += (long long)gridDim.x * 256) {
    const int r = (int)(i / c4), j = (int)(i % c4) * 4;
    const float4 v = *reinterpret_cast<const float4 *>(in + (size_t)r * ld_in + j);
    st4(out + (size_t)r * ld_out + j, v);
  }
}
void launch_f32_to_bf16(const float *in, int ld_in, bf16 *out, int ld_out, int rows, int cols, cudaStream_t st) {
  k_f32_to_bf16<<<grid_for((long long)rows * (cols >> 2)), 256, 0, st>>>(in, ld_in, out, ld_out, rows, cols);
  ++g_launch_count;
}


// =================================================================================================================
// q-sample and the eps-prediction losses (DDPM/functions/losses.py:21-37, runners/diffusion.py:533-572)
// =================================================================================================================
// x_t = x0 * sqrt(abar_t) + e * sqrt(1 - abar_t), x0 = 2 * x01 - 1 when rescale (data_transform): the same fp32 operation
// sequence as torch (separately rounded multiplies and adds, no contraction) -> bit-identical to the reference statements
__global__ void __launch_bounds__(256) k_q_sample(const float *__restrict__ x01, const float *__restrict__ e,
                                                  const int64_t *__restrict__ t, const float *__restrict__ sa,
                                                  const float *__restrict__ sb, int num_t, int rescale, long long total,
                                                  int chw, float *__restrict__ xt) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int n = (int)(i / chw);
    long long tl = t[n];  // an out-of-range timestep must not read outside the schedule tables
    const int tt = (int)(tl < 0 ? 0 : (tl >= num_t ? num_t - 1 : tl));
    float x0 = x01[i];
    if (rescale) x0 = __fadd_rn(__fmul_rn(2.0f, x0), -1.0f);
    xt[i] = __fadd_rn(__fmul_rn(x0, sa[tt]), __fmul_rn(e[i], sb[tt]));
  }
}
void launch_q_sample(const float *x01, const float *e, const int64_t *t, const float *sqrt_abar, const float *sqrt_1m_abar,
                     int num_t, int rescale, int n, int chw, float *xt, cudaStream_t st) {
  const long long total = (long long)n * chw;
  k_q_sample<<<grid_for(total), 256, 0, st>>>(x01, e, t, sqrt_abar, sqrt_1m_abar, num_t, rescale, total, chw, xt);
  ++g_launch_count;
}
// per sample: ss[n] = sum_chw (eps - target)^2 ; d_eps = 2 * w[n] * (eps - target)
__global__ void __launch_bounds__(256) k_eps_loss_grad(const float *__restrict__ eps, const float *__restrict__ target,
                                                       const float *__restrict__ w, int chw, float *__restrict__ d_eps,
                                                       float *__restrict__ ss) {
  __shared__ float red[256];
  const int n = blockIdx.x;
  const float w2 = 2.0f * w[n];
  float s = 0.f;
  for (int i = threadIdx.x; i < chw; i += 256) {
    const float d = eps[(size_t)n * chw + i] - target[(size_t)n * chw + i];
    s = fmaf(d, d, s);
    d_eps[(size_t)n * chw + i] = w2 * d;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) ss[n] = red[0];
}
// loss = sum_n w[n] * ss[n]   (one warp, fixed order)
__global__ void k_weighted_sum(const float *__restrict__ w, const float *__restrict__ ss, int n, float *__restrict__ out) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s = fmaf(w[i], ss[i], s);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) *out = s;
}
void launch_eps_loss_grad(const float *eps, const float *target, const float *w, int n, int chw, float *d_eps, float *ss,
                          float *loss, cudaStream_t st) {
  k_eps_loss_grad<<<n, 256, 0, st>>>(eps, target, w, chw, d_eps, ss);
  k_weighted_sum<<<1, 32, 0, st>>>(w, ss, n, loss);
  g_launch_count += 2;
}

}  // namespace salun
