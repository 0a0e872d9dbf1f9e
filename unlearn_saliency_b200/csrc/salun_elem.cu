// salun_elem.cu -- HBM-bound kernels of the ResNet path (see salun_elem.cuh).  All activations are NHWC bf16,
// 8 channels (16 bytes) per thread, fp32 math.  Replaces the ATen/cuDNN elementwise chain of
// Classification/models/ResNet.py:108-124 (BatchNorm2d train/eval, ReLU, residual add), :303-322
// (normalize, avgpool, fc) and their autograd backward.
#include <cstdlib>
#include "salun_elem.cuh"

#include <math.h>

#include "salun_common.cuh"

namespace salun {

constexpr int kET = 256;
constexpr int kRedY = 32;  // row lanes of the (32 x kRedY)-thread column-reduction blocks

__device__ __forceinline__ size_t pad_off(int m, int H, int W, int C) {
  const int hw = H * W;
  const int n = m / hw, r = m - n * hw;
  const int y = r / W, x = r - y * W;
  return ((size_t)(n * (H + 2) + y + 1) * (W + 2) + x + 1) * C;
}

// ------------------------------------------------------------------------------------------------
// forward BN statistics: [rows][C] fp32 partials (GEMM epilogue) -> [kStatSlices][2][C] doubles
// ------------------------------------------------------------------------------------------------
__global__ void k_bn_stats_reduce(const float *__restrict__ ssum, const float *__restrict__ ssq, int rows, int C,
                                  double *__restrict__ slices) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sh[2][kRedY][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int s = blockIdx.y;
  const int lo = (int)((long long)rows * s / kStatSlices), hi = (int)((long long)rows * (s + 1) / kStatSlices);
  double a = 0.0, b = 0.0;
  if (c < C) {
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;  // <= 4 fp32 adds per accumulator before widening
    int r = lo + threadIdx.y;
    for (; r + kRedY < hi; r += 2 * kRedY) {
      a0 += ssum[(size_t)r * C + c];
      b0 += ssq[(size_t)r * C + c];
      a1 += ssum[(size_t)(r + kRedY) * C + c];
      b1 += ssq[(size_t)(r + kRedY) * C + c];
      a += (double)a0 + (double)a1;
      b += (double)b0 + (double)b1;
      a0 = a1 = b0 = b1 = 0.f;
    }
    if (r < hi) {
      a += (double)ssum[(size_t)r * C + c];
      b += (double)ssq[(size_t)r * C + c];
    }
  }
  sh[0][threadIdx.y][threadIdx.x] = a;
  sh[1][threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int j = 1; j < kRedY; ++j) {
      a += sh[0][j][threadIdx.x];
      b += sh[1][j][threadIdx.x];
    }
    slices[((size_t)s * 2 + 0) * C + c] = a;
    slices[((size_t)s * 2 + 1) * C + c] = b;
  }
}
void launch_bn_stats_reduce(const float *stat_sum, const float *stat_sq, int rows, int C, double *slices,
                            cudaStream_t st) {
  dim3 grid((C + 31) / 32, kStatSlices), block(32, kRedY);
  { ::salun::launch_pdl(k_bn_stats_reduce, dim3(grid), dim3(block), 0, st, stat_sum, stat_sq, rows, C, slices); ++::salun::g_launch_count; }
}

// ------------------------------------------------------------------------------------------------
// BN apply (+ second BN branch, + residual, + ReLU) -> padded NHWC
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bn_prologue(const BnFwd &p, float *sc, float *sh, int C, int train, double count,
                                            float eps, float momentum) {
  if (train && p.count_dev) count = *p.count_dev;   // sync-BN: the global pixel count
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, invstd;
    if (train) {
      double s = 0.0, q = 0.0;
      for (int i = 0; i < kStatSlices; ++i) {
        s += p.slices[((size_t)i * 2 + 0) * C + c];
        q += p.slices[((size_t)i * 2 + 1) * C + c];
      }
      const double m = s / count;
      double var = q / count - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      invstd = (float)(1.0 / sqrt(var + (double)eps));
      if (blockIdx.x == 0) {  // nn.BatchNorm2d buffers: momentum update with the unbiased variance
        const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
        p.running_mean[c] = (1.f - momentum) * p.running_mean[c] + momentum * mean;
        p.running_var[c] = (1.f - momentum) * p.running_var[c] + momentum * (float)unb;
      }
    } else {
      mean = p.running_mean[c];
      invstd = (float)(1.0 / sqrt((double)p.running_var[c] + (double)eps));
    }
    if (blockIdx.x == 0) {
      p.saved_mean[c] = mean;
      p.saved_invstd[c] = invstd;
    }
    const float scale = p.gamma[c] * invstd;
    sc[c] = scale;
    sh[c] = p.beta[c] - mean * scale;
  }
}

template <int kU, int kMinB>
__global__ void __launch_bounds__(kET, kMinB) k_bn_apply(BnFwd a, BnFwd b, int has_b, const act_t *__restrict__ resid,
                                                  act_t *__restrict__ out, uint8_t *__restrict__ rmask, int M,
                                                  int H, int W, int C, int relu, int train, float eps, float momentum) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float smf[];
  float *sc_a = smf, *sh_a = smf + C, *sc_b = smf + 2 * C, *sh_b = smf + 3 * C;
  bn_prologue(a, sc_a, sh_a, C, train, (double)M, eps, momentum);
  if (has_b) bn_prologue(b, sc_b, sh_b, C, train, (double)M, eps, momentum);
  __syncthreads();
  const int tpr = C >> 3, rpb = kET / tpr;
  const int rl = threadIdx.x / tpr, c0 = (threadIdx.x - rl * tpr) * 8;
  // kU rows per thread per trip, all loads issued before the first use: one 16-byte load per thread and trip leaves the
  // SM with ~32 KB in flight, below what HBM3e latency x bandwidth needs (profiles/README.md section 3)
  const int stride = gridDim.x * rpb;
  for (int m0 = blockIdx.x * rpb + rl; m0 < M; m0 += kU * stride) {
    avec ra[kU], rb[kU], rr[kU];
    size_t po[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int m = m0 + u * stride;
      if (m < M) {
        ra[u] = ldraw(a.y + (size_t)m * C + c0);
        if (has_b) rb[u] = ldraw(b.y + (size_t)m * C + c0);
        po[u] = pad_off(m, H, W, C) + c0;
        if (resid) rr[u] = ldraw(resid + po[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int m = m0 + u * stride;
      if (m >= M) break;
      float v[8], t[8];
      cvt8(ra[u], v);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], sc_a[c0 + i], sh_a[c0 + i]);
      if (has_b) {
        cvt8(rb[u], t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += fmaf(t[i], sc_b[c0 + i], sh_b[c0 + i]);
      }
      if (resid) {
        cvt8(rr[u], t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += t[i];
      }
      if (relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      st8(out + po[u], v);
      if (rmask) {  // the stored bf16 value decides (a positive fp32 that rounds to +0 cannot occur: bf16 keeps the exponent)
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) bits |= (v[i] > 0.f ? 1u : 0u) << i;
        rmask[(size_t)m * (C >> 3) + (c0 >> 3)] = (uint8_t)bits;
      }
    }
  }
}
// (rows per trip, resident blocks per SM) of the BN elementwise kernels; SALUN_ELEM_VARIANT picks another one for tuning
static int elem_variant() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SALUN_ELEM_VARIANT");
    v = e ? atoi(e) : 3;
    if (v < 0 || v > 6) v = 3;
  }
  return v;
}
static inline int elem_grid(int M, int C, int blocks_per_sm = 8) {
  const int rpb = kET / (C >> 3);
  long long g = ((long long)M + rpb - 1) / rpb;
  if (g > 148 * blocks_per_sm) g = 148 * blocks_per_sm;  // = resident blocks: one full wave, no ragged second wave
  return (int)(g < 1 ? 1 : g);
}
void launch_bn_apply(const BnFwd &a, const BnFwd *b, const act_t *resid_padded, act_t *out_padded,
                     uint8_t *relu_mask_out, int n_img, int H, int W, int C, int relu, int train, float eps,
                     float momentum, cudaStream_t st) {
  const int M = n_img * H * W;
  BnFwd bb = b ? *b : a;
#define SALUN_BN_APPLY(U, B)                                                                                          \
  ::salun::launch_pdl(k_bn_apply<U, B>, dim3(elem_grid(M, C, B)), dim3(kET), 4 * C * sizeof(float), st, a, bb, b != nullptr, resid_padded, out_padded, \
                                                                            relu_mask_out, M, H, W, C, relu, train, eps, \
                                                                            momentum)
  switch (elem_variant()) {
    case 0: SALUN_BN_APPLY(1, 8); break;
    case 1: SALUN_BN_APPLY(2, 4); break;
    case 2: SALUN_BN_APPLY(4, 3); break;
    case 4: SALUN_BN_APPLY(8, 1); break;
    case 5: SALUN_BN_APPLY(2, 2); break;
    case 6: SALUN_BN_APPLY(4, 1); break;
    default: SALUN_BN_APPLY(4, 2); break;
  }
#undef SALUN_BN_APPLY
  ++::salun::g_launch_count;
}

// ------------------------------------------------------------------------------------------------
// BN backward
// ------------------------------------------------------------------------------------------------
template <int kU>
__global__ void __launch_bounds__(kET, 2) k_bn_bwd_reduce(const act_t *__restrict__ dout,
                                                       const uint8_t *__restrict__ rmask,
                                                       const act_t *__restrict__ y,
                                                       const float *__restrict__ mean, const float *__restrict__ invstd,
                                                       float *__restrict__ partials, int M, int H, int W, int C) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float smf[];  // [2][rpb][C]
  const int tpr = C >> 3, rpb = kET / tpr;
  const int rl = threadIdx.x / tpr, c0 = (threadIdx.x - rl * tpr) * 8;
  float s1[8], s2[8], mu[8], is[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s1[i] = s2[i] = 0.f;
    mu[i] = mean[c0 + i];
    is[i] = invstd[c0 + i];
  }
  // kU independent loads per trip; rows are still accumulated in ascending order (same sums as kU = 1)
  const int stride = gridDim.x * rpb;
  for (int m0 = blockIdx.x * rpb + rl; m0 < M; m0 += kU * stride) {
    avec rd[kU], ry[kU];
    uint32_t rbits[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int m = m0 + u * stride;
      if (m < M) {
        rd[u] = ldraw(dout + (size_t)m * C + c0);
        ry[u] = ldraw(y + (size_t)m * C + c0);
        rbits[u] = rmask ? (uint32_t)__ldg(rmask + (size_t)m * (C >> 3) + (c0 >> 3)) : 0xffu;
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (m0 + u * stride >= M) break;
      float d[8], yy[8];
      cvt8(rd[u], d);
      cvt8(ry[u], yy);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = (rbits[u] >> i) & 1u ? d[i] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s1[i] += d[i];
        s2[i] += d[i] * ((yy[i] - mu[i]) * is[i]);
      }
    }
  }
  float *b1 = smf, *b2 = smf + rpb * C;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    b1[rl * C + c0 + i] = s1[i];
    b2[rl * C + c0 + i] = s2[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kET) {
    float a = 0.f, b = 0.f;
    for (int r = 0; r < rpb; ++r) {
      a += b1[r * C + c];
      b += b2[r * C + c];
    }
    partials[((size_t)blockIdx.x * 2 + 0) * C + c] = a;
    partials[((size_t)blockIdx.x * 2 + 1) * C + c] = b;
  }
}
static inline int bwd_rows(int M, int C) {
  const int rpb = kET / (C >> 3);
  int g = (M + rpb - 1) / rpb;
  return g < kBwdPartialRows ? (g < 1 ? 1 : g) : kBwdPartialRows;
}
void launch_bn_bwd_reduce(const act_t *dout, const uint8_t *relu_mask, const act_t *y,
                          const float *saved_mean, const float *saved_invstd, float *partials, int n_img, int H, int W,
                          int C, cudaStream_t st) {
  const int M = n_img * H * W;
  const int rpb = kET / (C >> 3);
#define SALUN_BN_BWD_REDUCE(U)                                                                                  \
  ::salun::launch_pdl(k_bn_bwd_reduce<U>, dim3(bwd_rows(M, C)), dim3(kET), 2 * rpb * C * sizeof(float), st, dout, relu_mask, y, saved_mean, \
                                                                               saved_invstd, partials, M, H, W, C)
  switch (elem_variant()) {
    case 0: SALUN_BN_BWD_REDUCE(1); break;
    case 1: SALUN_BN_BWD_REDUCE(2); break;
    case 2: SALUN_BN_BWD_REDUCE(4); break;
    case 5: SALUN_BN_BWD_REDUCE(2); break;
    case 6: SALUN_BN_BWD_REDUCE(4); break;
    default: SALUN_BN_BWD_REDUCE(8); break;
  }
#undef SALUN_BN_BWD_REDUCE
  ++::salun::g_launch_count;
}

__global__ void k_bn_bwd_finalize(const float *__restrict__ partials, int rows, int C, const float *__restrict__ gamma,
                                  const float *__restrict__ invstd, float count, int train, float *__restrict__ dgamma,
                                  float *__restrict__ dbeta, float *__restrict__ coef, double *__restrict__ xchg) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sh[2][kRedY][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  double a = 0.0, b = 0.0;
  if (c < C) {
    int r = threadIdx.y;
    for (; r + kRedY < rows; r += 2 * kRedY) {
      const float a0 = partials[((size_t)r * 2 + 0) * C + c], b0 = partials[((size_t)r * 2 + 1) * C + c];
      const float a1 = partials[((size_t)(r + kRedY) * 2 + 0) * C + c], b1 = partials[((size_t)(r + kRedY) * 2 + 1) * C + c];
      a += (double)a0 + (double)a1;
      b += (double)b0 + (double)b1;
    }
    if (r < rows) {
      a += (double)partials[((size_t)r * 2 + 0) * C + c];
      b += (double)partials[((size_t)r * 2 + 1) * C + c];
    }
  }
  sh[0][threadIdx.y][threadIdx.x] = a;
  sh[1][threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int j = 1; j < kRedY; ++j) {
      a += sh[0][j][threadIdx.x];
      b += sh[1][j][threadIdx.x];
    }
    dbeta[c] = (float)a;
    dgamma[c] = (float)b;
    coef[c] = gamma[c] * invstd[c];
    coef[C + c] = train ? (float)(a / (double)count) : 0.f;
    coef[2 * C + c] = train ? (float)(b / (double)count) : 0.f;
    if (xchg) {   // sync-BN: the exchange kernel recomputes coef[C..3C) from the global sums
      xchg[c] = a;
      xchg[C + c] = b;
    }
  }
}
void launch_bn_bwd_finalize(const float *partials, int rows, int C, const float *gamma, const float *saved_invstd,
                            float count, int train, float *dgamma, float *dbeta, float *coef, cudaStream_t st, double *xchg_out) {
  const int M = (int)count;
  dim3 grid((C + 31) / 32), block(32, kRedY);
  { ::salun::launch_pdl(k_bn_bwd_finalize, dim3(grid), dim3(block), 0, st, partials, rows > 0 ? rows : bwd_rows(M, C), C, gamma, saved_invstd, count, train, dgamma,
                                            dbeta, coef, xchg_out); ++::salun::g_launch_count; }
}

// ------------------------------------------------------------------------------------------------
// sync-BN exchange over NVLink peer memory (see salun_elem.cuh)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(512) k_syncbn_exchange(SyncBnPeers p, unsigned long long epoch, int mode, int C,
                                                         double *__restrict__ slices, double local_count,
                                                         double *__restrict__ count_out, float *__restrict__ coef) {
  double *mine = p.slot[p.rank];
  if (mode == 0) {   // publish this rank's sums: the kStatSlices row slices collapse to one
    for (int j = threadIdx.x; j < 2 * C; j += blockDim.x) {
      double s = 0.0;
      for (int i = 0; i < kStatSlices; ++i) s += slices[(size_t)i * 2 * C + j];
      mine[j] = s;
    }
  }
  if (threadIdx.x == 0) mine[2 * C] = local_count;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < p.world) {
    st_release_sys(p.flags[threadIdx.x] + p.rank, epoch);          // "rank has published epoch" into every peer's memory
    const unsigned long long *f = p.flags[p.rank] + threadIdx.x;   // wait until peer threadIdx.x has published it too
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < epoch) {
      if (clock64() - t0 > 20000000000LL) __trap();                // ~10 s: a missing rank must fail loudly, not hang
    }
  }
  __syncthreads();
  double cnt = 0.0;
  for (int r = 0; r < p.world; ++r) cnt += p.slot[r][2 * C];
  for (int j = threadIdx.x; j < 2 * C; j += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < p.world; ++r) s += p.slot[r][j];           // rank order: the same sum on every rank
    if (mode == 0) {
      slices[j] = s;
      for (int i = 1; i < kStatSlices; ++i) slices[(size_t)i * 2 * C + j] = 0.0;
    } else {
      coef[C + j] = (float)(s / cnt);                                // coef[C + c] = mean dZ, coef[2C + c] = mean dZ * xhat
    }
  }
  if (mode == 0 && threadIdx.x == 0) *count_out = cnt;
}
void launch_syncbn_exchange(const SyncBnPeers &p, unsigned long long epoch, int mode, int C, double *slices,
                            double local_count, double *count_out, float *coef, cudaStream_t st) {
  k_syncbn_exchange<<<1, 512, 0, st>>>(p, epoch, mode, C, slices, local_count, count_out, coef);
  ++::salun::g_launch_count;
}

template <int kU, int kMinB>
__global__ void __launch_bounds__(kET, kMinB) k_bn_bwd_apply(const act_t *__restrict__ dout,
                                                      const uint8_t *__restrict__ rmask,
                                                      const act_t *__restrict__ y,
                                                      const float *__restrict__ mean, const float *__restrict__ invstd,
                                                      const float *__restrict__ coef, act_t *__restrict__ dy,
                                                      int dy_padded, act_t *__restrict__ dz_flat, int M, int H,
                                                      int W, int C) {
  pdl_trigger();
  pdl_wait();
  const int tpr = C >> 3, rpb = kET / tpr;
  const int rl = threadIdx.x / tpr, c0 = (threadIdx.x - rl * tpr) * 8;
  float mu[8], is[8], k1[8], m1[8], m2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mu[i] = mean[c0 + i];
    is[i] = invstd[c0 + i];
    k1[i] = coef[c0 + i];
    m1[i] = coef[C + c0 + i];
    m2[i] = coef[2 * C + c0 + i];
  }
  const int stride = gridDim.x * rpb;
  for (int m0 = blockIdx.x * rpb + rl; m0 < M; m0 += kU * stride) {
    avec rd[kU], ry[kU];
    uint32_t rbits[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int m = m0 + u * stride;
      if (m < M) {
        rd[u] = ldraw(dout + (size_t)m * C + c0);
        ry[u] = ldraw(y + (size_t)m * C + c0);
        rbits[u] = rmask ? (uint32_t)__ldg(rmask + (size_t)m * (C >> 3) + (c0 >> 3)) : 0xffu;
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int m = m0 + u * stride;
      if (m >= M) break;
      float d[8], yy[8];
      cvt8(rd[u], d);
      cvt8(ry[u], yy);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = (rbits[u] >> i) & 1u ? d[i] : 0.f;
      if (dz_flat) st8(dz_flat + (size_t)m * C + c0, d);
#pragma unroll
      for (int i = 0; i < 8; ++i) yy[i] = k1[i] * (d[i] - m1[i] - (yy[i] - mu[i]) * is[i] * m2[i]);
      st8(dy_padded ? dy + pad_off(m, H, W, C) + c0 : dy + (size_t)m * C + c0, yy);
    }
  }
}
void launch_bn_bwd_apply(const act_t *dout, const uint8_t *relu_mask, const act_t *y,
                         const float *saved_mean, const float *saved_invstd, const float *coef, act_t *dy,
                         int dy_padded, act_t *dz_flat, int n_img, int H, int W, int C, cudaStream_t st) {
  const int M = n_img * H * W;
#define SALUN_BN_BWD_APPLY(U, B)                                                                              \
  ::salun::launch_pdl(k_bn_bwd_apply<U, B>, dim3(elem_grid(M, C, B)), dim3(kET), 0, st, dout, relu_mask, y, saved_mean, saved_invstd, coef, dy, \
                                                           dy_padded, dz_flat, M, H, W, C)
  switch (elem_variant()) {
    case 0: SALUN_BN_BWD_APPLY(1, 6); break;
    case 1: SALUN_BN_BWD_APPLY(2, 4); break;
    case 2: SALUN_BN_BWD_APPLY(4, 3); break;
    case 4: SALUN_BN_BWD_APPLY(8, 1); break;
    case 5: SALUN_BN_BWD_APPLY(2, 2); break;
    case 6: SALUN_BN_BWD_APPLY(4, 1); break;
    default: SALUN_BN_BWD_APPLY(4, 2); break;
  }
#undef SALUN_BN_BWD_APPLY
  ++::salun::g_launch_count;
}

// ------------------------------------------------------------------------------------------------
// stem / stride-2 patch kernels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kET) k_stem_im2col(const float *__restrict__ x, act_t *__restrict__ col, int M,
                                                     int H, int W, float m0, float m1, float m2, float i0, float i1,
                                                     float i2) {
  pdl_trigger();
  pdl_wait();
  const int m = blockIdx.x * kET + threadIdx.x;
  if (m >= M) return;
  const int hw = H * W;
  const int n = m / hw, r = m - n * hw, y = r / W, xx = r - y * W;
  const float mean[3] = {m0, m1, m2}, inv[3] = {i0, i1, i2};
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = y + ky - 1, xs = xx + kx - 1;
      const bool ok = yy >= 0 && yy < H && xs >= 0 && xs < W;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float t = 0.f;  // zero padding is applied AFTER normalisation (conv pads the normalised tensor)
        if (ok) t = (x[((size_t)(n * 3 + c) * H + yy) * W + xs] - mean[c]) * inv[c];
        v[(ky * 3 + kx) * 3 + c] = t;
      }
    }
  act_t *dst = col + (size_t)m * 64;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = v[8 * j + i];
    st8(dst + 8 * j, f);
  }
}
void launch_stem_im2col(const float *x, act_t *col, int n_img, int H, int W, const float *mean3,
                        const float *inv_std3, cudaStream_t st) {
  const int M = n_img * H * W;
  { ::salun::launch_pdl(k_stem_im2col, dim3((M + kET - 1) / kET), dim3(kET), 0, st, x, col, M, H, W, mean3[0], mean3[1], mean3[2], inv_std3[0],
                                                     inv_std3[1], inv_std3[2]); ++::salun::g_launch_count; }
}

__global__ void __launch_bounds__(kET) k_im2col_s2(const act_t *__restrict__ in, act_t *__restrict__ col,
                                                   long long total, int Hin, int Win, int C, int ks) {
  pdl_trigger();
  pdl_wait();
  const int Ho = Hin / 2, Wo = Win / 2, cgs = C >> 3, taps = ks * ks;
  for (long long i = (long long)blockIdx.x * kET + threadIdx.x; i < total; i += (long long)gridDim.x * kET) {
    const int cg = (int)(i % cgs);
    long long t = i / cgs;
    const int tap = (int)(t % taps);
    const int mo = (int)(t / taps);
    const int n = mo / (Ho * Wo), r = mo - n * Ho * Wo, oy = r / Wo, ox = r - oy * Wo;
    const int ky = tap / ks, kx = tap - ky * ks;
    const int off = ks == 3 ? 0 : 1;  // padded coordinates of input pixel (2oy+ky-pad, 2ox+kx-pad)
    const int py = 2 * oy + ky + off, px = 2 * ox + kx + off;
    const avec v = ldvec(in + ((size_t)(n * (Hin + 2) + py) * (Win + 2) + px) * C + cg * 8);
    stvec(col + ((size_t)mo * taps + tap) * C + cg * 8, v);
  }
}
void launch_im2col_s2(const act_t *in_padded, act_t *col, int n_img, int Hin, int Win, int C, int ks,
                      cudaStream_t st) {
  const long long total = (long long)n_img * (Hin / 2) * (Win / 2) * ks * ks * (C >> 3);
  long long g = (total + kET - 1) / kET;
  if (g > 148 * 16) g = 148 * 16;
  { ::salun::launch_pdl(k_im2col_s2, dim3((int)g), dim3(kET), 0, st, in_padded, col, total, Hin, Win, C, ks); ++::salun::g_launch_count; }
}

__global__ void __launch_bounds__(kET) k_col2im_s2(const act_t *__restrict__ dcol3,
                                                   const act_t *__restrict__ dcol1,
                                                   act_t *__restrict__ dx, long long total, int Hin, int Win,
                                                   int C) {
  pdl_trigger();
  pdl_wait();
  const int Ho = Hin / 2, Wo = Win / 2, cgs = C >> 3;
  for (long long i = (long long)blockIdx.x * kET + threadIdx.x; i < total; i += (long long)gridDim.x * kET) {
    const int cg = (int)(i % cgs);
    const int m = (int)(i / cgs);
    const int n = m / (Hin * Win), r = m - n * Hin * Win, y = r / Win, x = r - y * Win;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = y + 1 - ky;
      if (ty < 0 || (ty & 1) || (ty >> 1) >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = x + 1 - kx;
        if (tx < 0 || (tx & 1) || (tx >> 1) >= Wo) continue;
        const int mo = (n * Ho + (ty >> 1)) * Wo + (tx >> 1);
        float t[8];
        ld8(dcol3 + ((size_t)mo * 9 + ky * 3 + kx) * C + cg * 8, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += t[j];
      }
    }
    if (dcol1 && !(y & 1) && !(x & 1)) {
      const int mo = (n * Ho + (y >> 1)) * Wo + (x >> 1);
      float t[8];
      ld8(dcol1 + (size_t)mo * C + cg * 8, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += t[j];
    }
    st8(dx + (size_t)m * C + cg * 8, acc);
  }
}
void launch_col2im_s2(const act_t *dcol3, const act_t *dcol1, act_t *dx, int n_img, int Hin,
                      int Win, int C, cudaStream_t st) {
  const long long total = (long long)n_img * Hin * Win * (C >> 3);
  long long g = (total + kET - 1) / kET;
  if (g > 148 * 16) g = 148 * 16;
  { ::salun::launch_pdl(k_col2im_s2, dim3((int)g), dim3(kET), 0, st, dcol3, dcol1, dx, total, Hin, Win, C); ++::salun::g_launch_count; }
}

// ------------------------------------------------------------------------------------------------
// generic flat-activation kernels (Bottleneck / ImageNet-stem nets: any H, W, stride, padding)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kET) k_im2col_flat(const act_t *__restrict__ in, act_t *__restrict__ col,
                                                     long long total, int Hin, int Win, int C, int ks, int stride, int pad,
                                                     int Hout, int Wout) {
  const int cgs = C >> 3, taps = ks * ks;
  for (long long i = (long long)blockIdx.x * kET + threadIdx.x; i < total; i += (long long)gridDim.x * kET) {
    const int cg = (int)(i % cgs);
    long long t = i / cgs;
    const int tap = (int)(t % taps);
    const int mo = (int)(t / taps);
    const int n = mo / (Hout * Wout), r = mo - n * Hout * Wout, oy = r / Wout, ox = r - oy * Wout;
    const int ky = tap / ks, kx = tap - ky * ks;
    const int y = oy * stride + ky - pad, x = ox * stride + kx - pad;
    avec v = avec_zero();
    if (y >= 0 && y < Hin && x >= 0 && x < Win) v = ldvec(in + ((size_t)(n * Hin + y) * Win + x) * C + cg * 8);
    stvec(col + ((size_t)mo * taps + tap) * C + cg * 8, v);
  }
}
// 3x3 taps of one (output pixel, 8-channel group) per thread: nine independent 16-byte loads in flight, then nine stores
// (the one-tap-per-thread form above spends its time in index arithmetic: 2.7 TB/s on the 56x56 layers of ResNet-50)
__global__ void __launch_bounds__(kET) k_im2col_flat3(const act_t *__restrict__ in, act_t *__restrict__ col, long long total,
                                                      int Hin, int Win, int C, int stride, int Hout, int Wout) {
  const int cgs = C >> 3;
  for (long long i = (long long)blockIdx.x * kET + threadIdx.x; i < total; i += (long long)gridDim.x * kET) {
    const int cg = (int)(i % cgs);
    const int mo = (int)(i / cgs);
    const int n = mo / (Hout * Wout), r = mo - n * Hout * Wout, oy = r / Wout, ox = r - oy * Wout;
    const int y0 = oy * stride - 1, x0 = ox * stride - 1;
    const long long off0 = ((long long)(n * Hin + y0) * Win + x0) * C + cg * 8;  // may be negative: only used in bounds
    avec v[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int y = y0 + ky, x = x0 + kx;
        v[ky * 3 + kx] = (y >= 0 && y < Hin && x >= 0 && x < Win) ? ldvec(in + (off0 + ((long long)ky * Win + kx) * C)) : avec_zero();
      }
    act_t *dst = col + (size_t)mo * 9 * C + cg * 8;
#pragma unroll
    for (int t = 0; t < 9; ++t) stvec(dst + (size_t)t * C, v[t]);
  }
}
void launch_im2col_flat(const act_t *in_flat, act_t *col, int n_img, int Hin, int Win, int C, int ks,
                        int stride, int pad, int Hout, int Wout, cudaStream_t st) {
  if (ks == 3 && pad == 1) {
    const long long total = (long long)n_img * Hout * Wout * (C >> 3);
    long long g = (total + kET - 1) / kET;
    if (g > 148 * 8) g = 148 * 8;
    k_im2col_flat3<<<(int)g, kET, 0, st>>>(in_flat, col, total, Hin, Win, C, stride, Hout, Wout);
    ++::salun::g_launch_count;
    return;
  }
  const long long total = (long long)n_img * Hout * Wout * ks * ks * (C >> 3);
  long long g = (total + kET - 1) / kET;
  if (g > 148 * 16) g = 148 * 16;
  { k_im2col_flat<<<(int)g, kET, 0, st>>>(in_flat, col, total, Hin, Win, C, ks, stride, pad, Hout, Wout); ++::salun::g_launch_count; }
}

__global__ void __launch_bounds__(kET) k_stem_im2col_generic(const float *__restrict__ x, act_t *__restrict__ col,
                                                             long long total, int Hin, int Win, int ks, int stride,
                                                             int pad, int Hout, int Wout, int kcp, float m0, float m1,
                                                             float m2, float i0, float i1, float i2) {
  // one thread per (output pixel, 8-column group of the patch row): columns j = tap*3 + c.  The column -> (ky, kx, c)
  // decode is a shared-memory table (the divisions per element made this kernel ALU bound: 0.64 TB/s)
  extern __shared__ int lut[];  // [kcp]: ky | kx << 8 | c << 16, or -1 for the zero padding columns
  const int groups = kcp >> 3, kc = ks * ks * 3;
  for (int j = threadIdx.x; j < kcp; j += kET) {
    const int tap = j / 3, c = j - tap * 3, ky = tap / ks, kx = tap - ky * ks;
    lut[j] = j < kc ? (ky | (kx << 8) | (c << 16)) : -1;
  }
  __syncthreads();
  const size_t plane = (size_t)Hin * Win;
  for (long long i = (long long)blockIdx.x * kET + threadIdx.x; i < total; i += (long long)gridDim.x * kET) {
    const int gidx = (int)(i % groups);
    const int mo = (int)(i / groups);
    const int n = mo / (Hout * Wout), r = mo - n * Hout * Wout, oy = r / Wout, ox = r - oy * Wout;
    const int yb = oy * stride - pad, xb = ox * stride - pad;
    const float *xn = x + (size_t)n * 3 * plane;
    float f[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = lut[gidx * 8 + q];
      float v = 0.f;
      if (e >= 0) {
        const int c = e >> 16, y = yb + (e & 0xff), xx = xb + ((e >> 8) & 0xff);
        if (y >= 0 && y < Hin && xx >= 0 && xx < Win) {
          const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), inv = c == 0 ? i0 : (c == 1 ? i1 : i2);
          v = (__ldg(xn + c * plane + (size_t)y * Win + xx) - mean) * inv;
        }
      }
      f[q] = v;
    }
    st8(col + (size_t)mo * kcp + gidx * 8, f);
  }
}
void launch_stem_im2col_generic(const float *x, act_t *col, int n_img, int Hin, int Win, int ks, int stride,
                                int pad, int Hout, int Wout, int kcp, const float *mean3, const float *inv_std3,
                                cudaStream_t st) {
  const long long total = (long long)n_img * Hout * Wout * (kcp >> 3);
  long long g = (total + kET - 1) / kET;
  if (g > 148 * 16) g = 148 * 16;
  { k_stem_im2col_generic<<<(int)g, kET, kcp * sizeof(int), st>>>(x, col, total, Hin, Win, ks, stride, pad, Hout, Wout, kcp, mean3[0],
                                                  mean3[1], mean3[2], inv_std3[0], inv_std3[1], inv_std3[2]); ++::salun::g_launch_count; }
}

__global__ void __launch_bounds__(kET) k_col2im_flat(const act_t *__restrict__ dcol,
                                                     const act_t *__restrict__ addend,
                                                     act_t *__restrict__ dx, long long total, int Hin, int Win,
                                                     int C, int ks, int stride, int pad, int Hout, int Wout) {
  const int cgs = C >> 3, taps = ks * ks;
  for (long long i = (long long)blockIdx.x * kET + threadIdx.x; i < total; i += (long long)gridDim.x * kET) {
    const int cg = (int)(i % cgs);
    const int m = (int)(i / cgs);
    const int n = m / (Hin * Win), r = m - n * Hin * Win, y = r / Win, x = r - y * Win;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (addend) ld8(addend + (size_t)m * C + cg * 8, acc);
    for (int ky = 0; ky < ks; ++ky) {
      const int ty = y + pad - ky;
      if (ty < 0 || ty % stride != 0) continue;
      const int oy = ty / stride;
      if (oy >= Hout) continue;
      for (int kx = 0; kx < ks; ++kx) {
        const int tx = x + pad - kx;
        if (tx < 0 || tx % stride != 0) continue;
        const int ox = tx / stride;
        if (ox >= Wout) continue;
        const int mo = (n * Hout + oy) * Wout + ox;
        float t[8];
        ld8(dcol + ((size_t)mo * taps + ky * ks + kx) * C + cg * 8, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += t[j];
      }
    }
    st8(dx + (size_t)m * C + cg * 8, acc);
  }
}
// 3x3 / padding 1: the (up to) nine patch-gradient vectors of one input pixel are independent loads -- issue them all, then add
// in the tap order of the generic kernel (same summation order, same result)
template <int kStride>
__global__ void __launch_bounds__(kET) k_col2im_flat3(const act_t *__restrict__ dcol, const act_t *__restrict__ addend,
                                                      act_t *__restrict__ dx, long long total, int Hin, int Win, int C,
                                                      int Hout, int Wout) {
  const int cgs = C >> 3;
  for (long long i = (long long)blockIdx.x * kET + threadIdx.x; i < total; i += (long long)gridDim.x * kET) {
    const int cg = (int)(i % cgs);
    const int m = (int)(i / cgs);
    const int n = m / (Hin * Win), r = m - n * Hin * Win, y = r / Win, x = r - y * Win;
    avec v[9];
    bool ok[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ty = y + 1 - ky, tx = x + 1 - kx;
        const int oy = ty / kStride, ox = tx / kStride;
        const bool in = ty >= 0 && tx >= 0 && (kStride == 1 || (ty % kStride == 0 && tx % kStride == 0)) && oy < Hout && ox < Wout;
        ok[ky * 3 + kx] = in;
        if (in) v[ky * 3 + kx] = ldvec(dcol + ((size_t)((n * Hout + oy) * Wout + ox) * 9 + ky * 3 + kx) * C + cg * 8);
      }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (addend) ld8(addend + (size_t)m * C + cg * 8, acc);
#pragma unroll
    for (int t = 0; t < 9; ++t)
      if (ok[t]) {
        float f[8];
        cvt8(v[t], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    st8(dx + (size_t)m * C + cg * 8, acc);
  }
}
void launch_col2im_flat(const act_t *dcol, const act_t *addend, act_t *dx, int n_img, int Hin,
                        int Win, int C, int ks, int stride, int pad, int Hout, int Wout, cudaStream_t st) {
  const long long total = (long long)n_img * Hin * Win * (C >> 3);
  long long g = (total + kET - 1) / kET;
  if (ks == 3 && pad == 1 && (stride == 1 || stride == 2)) {
    if (g > 148 * 8) g = 148 * 8;
    if (stride == 1)
      k_col2im_flat3<1><<<(int)g, kET, 0, st>>>(dcol, addend, dx, total, Hin, Win, C, Hout, Wout);
    else
      k_col2im_flat3<2><<<(int)g, kET, 0, st>>>(dcol, addend, dx, total, Hin, Win, C, Hout, Wout);
    ++::salun::g_launch_count;
    return;
  }
  if (g > 148 * 16) g = 148 * 16;
  { k_col2im_flat<<<(int)g, kET, 0, st>>>(dcol, addend, dx, total, Hin, Win, C, ks, stride, pad, Hout, Wout); ++::salun::g_launch_count; }
}

__global__ void __launch_bounds__(kET) k_bn_apply_flat(BnFwd a, BnFwd b, int has_b, const act_t *__restrict__ resid,
                                                       act_t *__restrict__ out, uint8_t *__restrict__ rmask, int M,
                                                       int C, int relu, int train, float eps, float momentum) {
  extern __shared__ float smf[];
  float *sc_a = smf, *sh_a = smf + C, *sc_b = smf + 2 * C, *sh_b = smf + 3 * C;
  bn_prologue(a, sc_a, sh_a, C, train, (double)M, eps, momentum);
  if (has_b) bn_prologue(b, sc_b, sh_b, C, train, (double)M, eps, momentum);
  __syncthreads();
  const int tpr = C >> 3;
  // C up to 2048 (tpr 256): one row per pass when tpr == kET, several rows otherwise; four passes' loads are issued before the
  // first is consumed (one 16-byte load per thread in flight left this kernel at 3.0 TB/s on the 1.2 GB tensors of ResNet-50)
  const int rpb = kET / tpr;
  const int rl = threadIdx.x / tpr, c0 = (threadIdx.x - rl * tpr) * 8;
  const long long step = (long long)gridDim.x * rpb;
  constexpr int U = 4;
  for (long long m0 = (long long)blockIdx.x * rpb + rl; m0 < M; m0 += step * U) {
    avec ya[U], yb[U], rs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long m = m0 + u * step;
      if (m < M) {
        const size_t o = (size_t)m * C + c0;
        ya[u] = ldvec(a.y + o);
        if (has_b) yb[u] = ldvec(b.y + o);
        if (resid) rs[u] = ldvec(resid + o);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long m = m0 + u * step;
      if (m >= M) break;
      const size_t o = (size_t)m * C + c0;
      float v[8], t[8];
      cvt8(ya[u], v);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], sc_a[c0 + i], sh_a[c0 + i]);
      if (has_b) {
        cvt8(yb[u], t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += fmaf(t[i], sc_b[c0 + i], sh_b[c0 + i]);
      }
      if (resid) {
        cvt8(rs[u], t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += t[i];
      }
      if (relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      st8(out + o, v);
      if (rmask) {
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) bits |= (v[i] > 0.f ? 1u : 0u) << i;
        rmask[(size_t)m * (C >> 3) + (c0 >> 3)] = (uint8_t)bits;
      }
    }
  }
}
void launch_bn_apply_flat(const BnFwd &a, const BnFwd *b, const act_t *resid_flat, act_t *out_flat,
                          uint8_t *relu_mask_out, int M, int C, int relu, int train, float eps, float momentum,
                          cudaStream_t st) {
  BnFwd bb = b ? *b : a;
  { k_bn_apply_flat<<<elem_grid(M, C), kET, 4 * C * sizeof(float), st>>>(a, bb, b != nullptr, resid_flat, out_flat,
                                                                       relu_mask_out, M, C, relu, train, eps, momentum); ++::salun::g_launch_count; }
}

__global__ void __launch_bounds__(kET) k_maxpool_fwd(const act_t *__restrict__ in, act_t *__restrict__ out,
                                                     uint8_t *__restrict__ argmax, long long total, int Hin, int Win,
                                                     int C) {
  const int Ho = (Hin + 2 - 3) / 2 + 1, Wo = (Win + 2 - 3) / 2 + 1, cgs = C >> 3;
  for (long long i = (long long)blockIdx.x * kET + threadIdx.x; i < total; i += (long long)gridDim.x * kET) {
    const int cg = (int)(i % cgs);
    const int mo = (int)(i / cgs);
    const int n = mo / (Ho * Wo), r = mo - n * Ho * Wo, oy = r / Wo, ox = r - oy * Wo;
    float best[8];
    int arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -INFINITY;
      arg[j] = 0;
    }
    for (int ky = 0; ky < 3; ++ky) {
      const int y = oy * 2 + ky - 1;
      if (y < 0 || y >= Hin) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int x = ox * 2 + kx - 1;
        if (x < 0 || x >= Win) continue;
        float t[8];
        ld8(in + ((size_t)(n * Hin + y) * Win + x) * C + cg * 8, t);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (t[j] > best[j]) {  // first maximum in scan order wins, like torch's max_pool2d
            best[j] = t[j];
            arg[j] = ky * 3 + kx;
          }
      }
    }
    st8(out + (size_t)mo * C + cg * 8, best);
    uint2 packed;
    packed.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
    packed.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
    *reinterpret_cast<uint2 *>(argmax + (size_t)mo * C + cg * 8) = packed;
  }
}
void launch_maxpool_fwd(const act_t *in_flat, act_t *out_flat, uint8_t *argmax, int n_img, int Hin,
                        int Win, int C, cudaStream_t st) {
  const int Ho = (Hin + 2 - 3) / 2 + 1, Wo = (Win + 2 - 3) / 2 + 1;
  const long long total = (long long)n_img * Ho * Wo * (C >> 3);
  long long g = (total + kET - 1) / kET;
  if (g > 148 * 16) g = 148 * 16;
  { k_maxpool_fwd<<<(int)g, kET, 0, st>>>(in_flat, out_flat, argmax, total, Hin, Win, C); ++::salun::g_launch_count; }
}
__global__ void __launch_bounds__(kET) k_maxpool_bwd(const act_t *__restrict__ dout,
                                                     const uint8_t *__restrict__ argmax, act_t *__restrict__ dx,
                                                     long long total, int Hin, int Win, int C) {
  const int Ho = (Hin + 2 - 3) / 2 + 1, Wo = (Win + 2 - 3) / 2 + 1, cgs = C >> 3;
  for (long long i = (long long)blockIdx.x * kET + threadIdx.x; i < total; i += (long long)gridDim.x * kET) {
    const int cg = (int)(i % cgs);
    const int m = (int)(i / cgs);
    const int n = m / (Hin * Win), r = m - n * Hin * Win, y = r / Win, x = r - y * Win;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = y + 1 - ky;
      if (ty < 0 || (ty & 1) || (ty >> 1) >= Ho) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = x + 1 - kx;
        if (tx < 0 || (tx & 1) || (tx >> 1) >= Wo) continue;
        const int mo = (n * Ho + (ty >> 1)) * Wo + (tx >> 1);
        const uint2 a = *reinterpret_cast<const uint2 *>(argmax + (size_t)mo * C + cg * 8);
        float t[8];
        ld8(dout + (size_t)mo * C + cg * 8, t);
        const int want = ky * 3 + kx;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t w = j < 4 ? a.x : a.y;
          if ((int)((w >> (8 * (j & 3))) & 0xffu) == want) acc[j] += t[j];
        }
      }
    }
    st8(dx + (size_t)m * C + cg * 8, acc);
  }
}
void launch_maxpool_bwd(const act_t *dout_flat, const uint8_t *argmax, act_t *dx_flat, int n_img, int Hin,
                        int Win, int C, cudaStream_t st) {
  const long long total = (long long)n_img * Hin * Win * (C >> 3);
  long long g = (total + kET - 1) / kET;
  if (g > 148 * 16) g = 148 * 16;
  { k_maxpool_bwd<<<(int)g, kET, 0, st>>>(dout_flat, argmax, dx_flat, total, Hin, Win, C); ++::salun::g_launch_count; }
}
__global__ void k_avgpool_flat(const act_t *__restrict__ act, float *__restrict__ pooled, int n_img, int pix, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * C) return;
  const int n = i / C, c = i - n * C;
  float s = 0.f;
  for (int p = 0; p < pix; ++p) s += act_to_float(act[((size_t)n * pix + p) * C + c]);
  pooled[i] = s / (float)pix;
}
void launch_avgpool_flat(const act_t *act_flat, float *pooled, int n_img, int pix, int C, cudaStream_t st) {
  { k_avgpool_flat<<<(n_img * C + 255) / 256, 256, 0, st>>>(act_flat, pooled, n_img, pix, C); ++::salun::g_launch_count; }
}

// ------------------------------------------------------------------------------------------------
// weight re-layout (fp32 master, native [Cout][tap][Cin]) -> bf16 GEMM operands
// ------------------------------------------------------------------------------------------------
// all convolutions of the network in ONE launch: blockIdx.y walks the table.
//   forward operand : plain cast (+ zero padding of the stem's 27 -> 64 columns), coalesced both ways
//   dgrad operand   : out[ci][taps-1-t][co] = w[co][t][ci]  (stride-1: flipped taps)   or   out[j][co] = w[co][j]
//                     (stride-2: taps == 1, "ci" walks all kc columns) -- a batched matrix transpose, done through
//                     32x33 shared-memory tiles so that both the fp32 reads and the bf16 writes are coalesced.
__global__ void __launch_bounds__(256) k_prep_w_all(const WPrepEntry *__restrict__ tab, const float *__restrict__ params,
                                                    int need_dgrad) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][33];
  const WPrepEntry e = tab[blockIdx.y];
  const float *__restrict__ w = params + e.w_off;
  // forward operand: one output row per block iteration, threads along the row (no integer division per element)
  for (int co = blockIdx.x; co < e.cout; co += gridDim.x) {
    const float *__restrict__ src = w + (size_t)co * e.kc;
    wop_t *__restrict__ dst = e.w_fwd + (size_t)co * e.kcp * kWopK;
    for (int j = threadIdx.x; j < e.kcp; j += blockDim.x) wop_store(dst, j, e.kcp, j < e.kc ? src[j] : 0.f);
  }
  if (!need_dgrad || e.dgrad_mode == 0) return;
  const int taps = e.dgrad_mode == 1 ? e.kc / e.cin : 1;
  const int cin = e.dgrad_mode == 1 ? e.cin : e.kc;
  const int tiles_ci = (cin + 31) / 32, tiles_co = (e.cout + 31) / 32;
  const int ntiles = taps * tiles_ci * tiles_co;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int ldo = e.ldo ? e.ldo : e.cout;
  for (int tIdx = blockIdx.x; tIdx < ntiles; tIdx += gridDim.x) {
    const int t = tIdx / (tiles_ci * tiles_co);
    const int r = tIdx - t * tiles_ci * tiles_co;
    const int ci0 = (r % tiles_ci) * 32, co0 = (r / tiles_ci) * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // read w[co][t][ci]: ci contiguous
      const int co = co0 + ty + 8 * j, ci = ci0 + tx;
      tile[ty + 8 * j][tx] = (co < e.cout && ci < cin) ? w[((size_t)co * taps + t) * cin + ci] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // write out[ci][taps-1-t][co]: co contiguous
      const int ci = ci0 + ty + 8 * j, co = co0 + tx;
      if (ci < cin && co < e.cout)
        wop_store(e.w_dgrad + (size_t)ci * taps * ldo * kWopK, (taps - 1 - t) * ldo + co, taps * ldo, tile[tx][ty + 8 * j]);
    }
    __syncthreads();
  }
}
void launch_prep_w_all(const WPrepEntry *table_dev, int n_convs, const float *params, int need_dgrad, cudaStream_t st) {
  { ::salun::launch_pdl(k_prep_w_all, dim3(dim3(592, n_convs)), dim3(256), 0, st, table_dev, params, need_dgrad); ++::salun::g_launch_count; }
}

// ------------------------------------------------------------------------------------------------
// head: global average pool, FC, cross-entropy (mean over the batch), and their backward
// ------------------------------------------------------------------------------------------------
__global__ void k_avgpool(const act_t *__restrict__ act, float *__restrict__ pooled, int n_img, int H, int W,
                          int C) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * C) return;
  const int n = i / C, c = i - n * C;
  float s = 0.f;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) s += act_to_float(act[((size_t)(n * (H + 2) + y + 1) * (W + 2) + x + 1) * C + c]);
  pooled[i] = s / (float)(H * W);
}
void launch_avgpool(const act_t *act_padded, float *pooled, int n_img, int H, int W, int C, cudaStream_t st) {
  { ::salun::launch_pdl(k_avgpool, dim3((n_img * C + 255) / 256), dim3(256), 0, st, act_padded, pooled, n_img, H, W, C); ++::salun::g_launch_count; }
}

__global__ void __launch_bounds__(128) k_fc_ce(const float *__restrict__ pooled, const float *__restrict__ w,
                                               const float *__restrict__ bias, const int64_t *__restrict__ labels,
                                               float *__restrict__ logits, float *__restrict__ dlogits,
                                               float *__restrict__ loss_ps, int n_img, int C, int K, float sign) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float lg[];  // [K]
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *x = pooled + (size_t)b * C;
  for (int k = warp; k < K; k += 4) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += x[c] * w[(size_t)k * C + c];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) lg[k] = s + bias[k];
  }
  __syncthreads();
  if (logits)
    for (int k = threadIdx.x; k < K; k += 128) logits[(size_t)b * K + k] = lg[k];
  if (!labels) return;
  __shared__ float red[2];
  if (threadIdx.x == 0) {
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, lg[k]);
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(lg[k] - mx);
    red[0] = mx;
    red[1] = se;
    const int y = (int)labels[b];
    loss_ps[b] = (logf(se) + mx) - lg[y];
  }
  __syncthreads();
  const float mx = red[0], inv = 1.f / red[1];
  const int y = (int)labels[b];
  for (int k = threadIdx.x; k < K; k += 128)
    dlogits[(size_t)b * K + k] = sign * (expf(lg[k] - mx) * inv - (k == y ? 1.f : 0.f)) / (float)n_img;
}
void launch_fc_ce(const float *pooled, const float *w, const float *b, const int64_t *labels, float *logits,
                  float *dlogits, float *loss_per_sample, int n_img, int C, int K, float sign, cudaStream_t st) {
  { ::salun::launch_pdl(k_fc_ce, dim3(n_img), dim3(128), K * sizeof(float), st, pooled, w, b, labels, logits, dlogits, loss_per_sample, n_img, C, K,
                                                  sign); ++::salun::g_launch_count; }
}
// Wide classifier heads (ImageNet: 1000 x 2048): the logits come from an fp32 GEMM (launch_sgemm), this kernel does the
// per-sample cross entropy and dL/dlogits of k_fc_ce on them.  One warp per sample, fixed summation order.
__global__ void __launch_bounds__(256) k_ce_rows(const float *__restrict__ logits, const int64_t *__restrict__ labels,
                                                 float *__restrict__ dlogits, float *__restrict__ loss_ps, int n_img, int K,
                                                 float sign) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= n_img) return;
  const float *lg = logits + (size_t)b * K;
  float mx = -INFINITY;
  for (int k = lane; k < K; k += 32) mx = fmaxf(mx, lg[k]);
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.f;
  for (int k = lane; k < K; k += 32) se += expf(lg[k] - mx);
  for (int o = 16; o; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  int y = (int)labels[b];
  y = y < 0 ? 0 : (y >= K ? K - 1 : y);
  if (lane == 0) loss_ps[b] = (logf(se) + mx) - lg[y];
  const float inv = 1.f / se;
  for (int k = lane; k < K; k += 32)
    dlogits[(size_t)b * K + k] = sign * (expf(lg[k] - mx) * inv - (k == y ? 1.f : 0.f)) / (float)n_img;
}
void launch_ce_rows(const float *logits, const int64_t *labels, float *dlogits, float *loss_per_sample, int n_img, int K,
                    float sign, cudaStream_t st) {
  k_ce_rows<<<(n_img + 7) / 8, 256, 0, st>>>(logits, labels, dlogits, loss_per_sample, n_img, K, sign);
  ++g_launch_count;
}
// gradient of the global average pool: dact[(b*pix + p)][c] = dpooled[b][c] / pix for every pixel p
__global__ void k_pool_grad_bcast(const float *__restrict__ dpooled, act_t *__restrict__ dact, long long total, int C, int pix) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % C);
    const long long b = i / ((long long)pix * C);
    dact[i] = act_from_float(dpooled[b * C + c] / (float)pix);
  }
}
void launch_pool_grad_bcast(const float *dpooled, act_t *dact_flat, int n_img, int C, int pix, cudaStream_t st) {
  const long long total = (long long)n_img * pix * C;
  long long g = (total + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  k_pool_grad_bcast<<<(int)g, 256, 0, st>>>(dpooled, dact_flat, total, C, pix);
  ++g_launch_count;
}

__global__ void k_loss_sum(const float *__restrict__ l, int n, float sign, float *__restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += l[i];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    *out = sign * t / (float)n;
  }
}
void launch_loss_sum(const float *loss_per_sample, int n_img, float sign, float *loss_out, cudaStream_t st) {
  { ::salun::launch_pdl(k_loss_sum, dim3(1), dim3(256), 0, st, loss_per_sample, n_img, sign, loss_out); ++::salun::g_launch_count; }
}

__global__ void k_fc_bwd_w(const float *__restrict__ pooled, const float *__restrict__ dl, float *__restrict__ dw,
                           float *__restrict__ db, int n_img, int C, int K) {
  pdl_trigger();
  pdl_wait();
  // grid (K, C/64), block (64 channels, 8 batch lanes): fixed summation order -> deterministic
  __shared__ float sh[8][64];
  const int k = blockIdx.x, c = blockIdx.y * 64 + threadIdx.x;
  float s = 0.f, sb = 0.f;
  for (int b = threadIdx.y; b < n_img; b += 8) {
    const float d = dl[(size_t)b * K + k];
    if (c < C) s += d * pooled[(size_t)b * C + c];
    sb += d;
  }
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int j = 1; j < 8; ++j) s += sh[j][threadIdx.x];
    dw[(size_t)k * C + c] = s;
  }
  __syncthreads();
  if (blockIdx.y == 0 && threadIdx.x == 0) {
    sh[threadIdx.y][0] = sb;
  }
  __syncthreads();
  if (blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) {
    float t = 0.f;
    for (int j = 0; j < 8; ++j) t += sh[j][0];
    db[k] = t;
  }
}
__global__ void k_fc_bwd_x(const float *__restrict__ dl, const float *__restrict__ w, act_t *__restrict__ dact,
                           int n_img, int C, int K, int pix) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * C) return;
  const int b = i / C, c = i - b * C;
  float s = 0.f;
  for (int k = 0; k < K; ++k) s += dl[(size_t)b * K + k] * w[(size_t)k * C + c];
  const act_t v = act_from_float(s / (float)pix);
  for (int p = 0; p < pix; ++p) dact[((size_t)b * pix + p) * C + c] = v;
}
void launch_fc_bwd(const float *pooled, const float *dlogits, const float *w, float *dw, float *db,
                   act_t *dact_flat, int n_img, int C, int K, int pix, cudaStream_t st) {
  { ::salun::launch_pdl(k_fc_bwd_w, dim3(dim3(K, (C + 63) / 64)), dim3(dim3(64, 8)), 0, st, pooled, dlogits, dw, db, n_img, C, K); ++::salun::g_launch_count; }
  { ::salun::launch_pdl(k_fc_bwd_x, dim3((n_img * C + 255) / 256), dim3(256), 0, st, dlogits, w, dact_flat, n_img, C, K, pix); ++::salun::g_launch_count; }
}

}  // namespace salun
