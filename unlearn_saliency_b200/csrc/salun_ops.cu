// salun_ops.cu -- op-level C ABI over the tcgen05 GEMM / convolution kernels and the elementwise kernels, plus the layers the
// Stable-Diffusion (LDM) U-Net adds to the DDPM one: LayerNorm, GEGLU, multi-head self / cross attention, the cos|sin timestep
// embedding, GroupNorm at any width.  The SD U-Net forward (SD/ldm/modules/diffusionmodules/openaimodel.py:814-846 with
// ResBlock :268-288, SpatialTransformer / BasicTransformerBlock / CrossAttention SD/ldm/modules/attention.py:168-303, GEGLU
// :37-44, timestep_embedding util.py:173-197) is composed from these ops by the host mirror (unlearn_saliency_b200/sd/engine.py)
// and replayed from a CUDA graph; the ops are also the unit the parity tests check one by one.
//
// Activations are act_t (bf16, or bf16 hi/lo pairs in the split build; salun_act.cuh): padded NHWC [n][H+2][W+2][C] with a zero
// halo for 3x3 convolutions, flat [rows][C] for token matrices.  Weights are prepared once per load into tensor-core operand
// layout (salun_op_prep_weight).  Every call enqueues on the caller's stream; nothing synchronises.
#include <math.h>
#include <stdlib.h>

#include "salun_elem.cuh"
#include "salun_gemm.cuh"
#include "salun_unet_elem.cuh"

namespace salun {

static inline int grid1d(long long total, int threads = 256, int cap = 148 * 8) {
  long long g = (total + threads - 1) / threads;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}
static int ilog2i(long long v) {
  int s = 0;
  while ((1LL << s) < v) ++s;
  return s;
}
__device__ __forceinline__ size_t pad_off4(int n, int y, int x, int H, int W, int C) {
  return (((size_t)n * (H + 2) + y + 1) * (W + 2) + x + 1) * C;
}

// ------------------------------------------------------------------------------------------------ conversions
__global__ void k_f32_to_act(const float *__restrict__ src, long long ld_src, act_t *__restrict__ dst, long long ld_dst,
                             long long rows, long long cols) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    dst[r * ld_dst + c] = act_from_float(src[r * ld_src + c]);
  }
}
__global__ void k_act_to_f32(const act_t *__restrict__ src, long long ld_src, float *__restrict__ dst, long long ld_dst,
                             long long rows, long long cols) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    dst[r * ld_dst + c] = act_to_float(src[r * ld_src + c]);
  }
}
// x fp32 NCHW -> padded NHWC act with Cp >= C channels (extra channels zero); the halo is not touched (zeroed at allocation)
__global__ void k_nchw_to_padded(const float *__restrict__ x, act_t *__restrict__ out, long long total, int C, int Cp, int H,
                                 int W) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    long long r = i / Cp;
    const int xx = (int)(r % W);
    r /= W;
    const int y = (int)(r % H), n = (int)(r / H);
    const float v = c < C ? x[(((size_t)n * C + c) * H + y) * W + xx] : 0.f;
    out[pad_off4(n, y, xx, H, W, Cp) + c] = act_from_float(v);
  }
}
__global__ void k_padded_to_nchw(const act_t *__restrict__ in, float *__restrict__ out, long long total, int C, int H, int W) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    long long r = i / W;
    const int y = (int)(r % H);
    r /= H;
    const int c = (int)(r % C), n = (int)(r / C);
    out[i] = act_to_float(in[pad_off4(n, y, xx, H, W, C) + c]);
  }
}
// out NCHW [n][C][H][W] = y[(n*H + yy)*W + xx][c] + bias[c]   (fp32 rows of a GEMM whose N was padded to ld)
__global__ void k_rows_to_nchw(const float *__restrict__ y, int ld, const float *__restrict__ bias, float *__restrict__ out,
                               long long total, int C, int H, int W) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    long long r = i / W;
    const int yy = (int)(r % H);
    r /= H;
    const int c = (int)(r % C), n = (int)(r / C);
    out[i] = y[(((size_t)n * H + yy) * W + xx) * ld + c] + (bias ? bias[c] : 0.f);
  }
}
// PyTorch conv weight OIHW fp32 (or Linear [out][in], ks = 1) -> operand rows [cout_pad][ks*ks*cin_pad] (tap-major, then
// channel), zero padded, in the weight-operand layout of this build
__global__ void k_prep_weight(const float *__restrict__ w, wop_t *__restrict__ out, int cout, int cin, int ks, int cout_pad,
                              int cin_pad) {
  const int taps = ks * ks, Kp = taps * cin_pad;
  const long long total = (long long)cout_pad * Kp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / Kp), j = (int)(i - (long long)co * Kp);
    const int tap = j / cin_pad, ci = j - tap * cin_pad;
    float v = 0.f;
    if (co < cout && ci < cin) v = w[((size_t)co * cin + ci) * taps + tap];
    wop_store(out + (size_t)co * Kp * kWopK, j, Kp, v);
  }
}

// ------------------------------------------------------------------------------------------------ GroupNorm (any C % 8 == 0)
// Both kernels give every thread ONE fixed 8-channel vector (two for C > 2048) and walk pixels: 16-byte coalesced loads, one
// 32-bit division per load (pixel -> row / column of the padded image) and nothing else in the loop.  (The first version
// decoded (pixel, channel) per ELEMENT with 64-bit divisions and ran at 0.27 TB/s: profiles/r2_ncu_summary_sd.txt.)
//   k_gn2_partial  grid (S, n): CTA z sums its 1/S slice of the pixels of sample n for all channels (fp32 per thread over a few
//                  pixels, fp64 across threads and channels), writes (sum, sum of squares) per group as doubles
//   k_gn2_apply    grid (X, n): prologue folds the S partials of the sample's 32 groups into mean / rstd and each thread
//                  turns them into scale / shift of its own 8 channels (registers); then y = x * scale + shift (+ SiLU).
// Deterministic: fixed partition, fixed summation order, no atomics.
constexpr int kGnMaxSplits = 64;
static inline int gn_splits(int HW) { return HW < kGnMaxSplits ? HW : kGnMaxSplits; }

template <int VPT>
__global__ void __launch_bounds__(256) k_gn2_partial(const act_t *__restrict__ x, double *__restrict__ part, int H, int W, int C,
                                                     int S) {
  extern __shared__ float gsm[];  // [R][C][2] per-row-lane channel sums
  const int vecs = C >> 3, n = blockIdx.y, z = blockIdx.x, HW = H * W, cpg = C / 32;
  const int R = VPT == 1 ? 256 / vecs : 1;
  const int r = VPT == 1 ? threadIdx.x / vecs : 0, v = VPT == 1 ? threadIdx.x - r * vecs : threadIdx.x;
  const int p0 = (int)((long long)HW * z / S), p1 = (int)((long long)HW * (z + 1) / S);
  float a[VPT][8], b[VPT][8];
#pragma unroll
  for (int k = 0; k < VPT; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) a[k][j] = b[k][j] = 0.f;
  if (r < R) {
    constexpr int UN = VPT == 1 ? 4 : 2;  // pixels in flight per thread: the loads are issued before the first is consumed
    for (int pb = p0 + r; pb < p1; pb += R * UN) {
      avec q[UN][VPT];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int p = pb + u * R;
        if (p < p1) {
          const int y = p / W, xx = p - y * W;
          const act_t *row = x + pad_off4(n, y, xx, H, W, C);
#pragma unroll
          for (int k = 0; k < VPT; ++k)
            if (v + k * 256 < vecs) q[u][k] = ldvec(row + (v + k * 256) * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        if (pb + u * R >= p1) break;
#pragma unroll
        for (int k = 0; k < VPT; ++k)
          if (v + k * 256 < vecs) {
            float f[8];
            cvt8(q[u][k], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              a[k][j] += f[j];
              b[k][j] = fmaf(f[j], f[j], b[k][j]);
            }
          }
      }
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int vv = v + k * 256;
      if (vv < vecs) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          gsm[((size_t)r * C + vv * 8 + j) * 2] = a[k][j];
          gsm[((size_t)r * C + vv * 8 + j) * 2 + 1] = b[k][j];
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {   // group g: its cpg channels x R row lanes, in a fixed order, in fp64
    const int g = threadIdx.x;
    double s0 = 0.0, s1 = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c)
      for (int q = 0; q < R; ++q) {
        s0 += gsm[((size_t)q * C + c) * 2];
        s1 += gsm[((size_t)q * C + c) * 2 + 1];
      }
    part[(((size_t)n * 32 + g) * S + z) * 2] = s0;
    part[(((size_t)n * 32 + g) * S + z) * 2 + 1] = s1;
  }
}

template <int VPT>
__global__ void __launch_bounds__(256) k_gn2_apply(const act_t *__restrict__ x, const double *__restrict__ part,
                                                   const float *__restrict__ gamma, const float *__restrict__ beta,
                                                   act_t *__restrict__ out, int out_flat, int swish, int H, int W, int C, int S,
                                                   float eps) {
  __shared__ float2 st[32];  // (mean, rstd) of the sample's groups
  const int vecs = C >> 3, n = blockIdx.y, HW = H * W, cpg = C / 32;
  {  // 8 lanes per group: each sums every 8th partial, then a fixed xor-tree over the 8 lanes (deterministic); a single
     // thread per group walking 2 x 64 dependent L2 loads cost more than the normalisation itself
    const int g = threadIdx.x >> 3, l = threadIdx.x & 7;
    double s0 = 0.0, s1 = 0.0;
    const double *pg = part + ((size_t)n * 32 + g) * S * 2;
    for (int z = l; z < S; z += 8) {
      s0 += pg[z * 2];
      s1 += pg[z * 2 + 1];
    }
    for (int o = 4; o; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (l == 0) {
      const double cnt = (double)HW * cpg, mean = s0 / cnt;
      double var = s1 / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      st[g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
  }
  __syncthreads();
  const int R = VPT == 1 ? 256 / vecs : 1;
  const int r = VPT == 1 ? threadIdx.x / vecs : 0, v = VPT == 1 ? threadIdx.x - r * vecs : threadIdx.x;
  if (r >= R) return;
  float sc[VPT][8], sh[VPT][8];
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int vv = v + k * 256;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = vv * 8 + j;
      if (vv < vecs) {
        const float2 mr = st[c / cpg];
        sc[k][j] = mr.y * gamma[c];
        sh[k][j] = beta[c] - mr.x * sc[k][j];
      } else {
        sc[k][j] = sh[k][j] = 0.f;
      }
    }
  }
  const int p0 = (int)((long long)HW * blockIdx.x / gridDim.x), p1 = (int)((long long)HW * (blockIdx.x + 1) / gridDim.x);
  for (int p = p0 + r; p < p1; p += R) {
    const int y = p / W, xx = p - y * W;
    const size_t ip = pad_off4(n, y, xx, H, W, C);
    const size_t op = out_flat ? ((size_t)n * HW + p) * C : ip;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int vv = v + k * 256;
      if (vv < vecs) {
        float f[8];
        ld8(x + ip + vv * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float yv = fmaf(f[j], sc[k][j], sh[k][j]);
          if (swish) yv = yv / (1.f + expf(-yv));
          f[j] = yv;
        }
        st8(out + op + vv * 8, f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm / GEGLU / embedding
// one warp per token row, C <= 2048, C % 8 == 0 (nn.LayerNorm(dim), attention.py:221-223: biased variance, eps 1e-5)
__global__ void __launch_bounds__(256) k_layernorm(const act_t *__restrict__ x, const float *__restrict__ gamma,
                                                   const float *__restrict__ beta, act_t *__restrict__ out, long long rows, int C,
                                                   float eps) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int vecs = C >> 3;
  float f[8][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int v = lane + 32 * k;
    if (v < vecs) {
      ld8(x + row * C + v * 8, f[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[k][j];
    }
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int v = lane + 32 * k;
    if (v < vecs) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[k][j] - mean;
        q += d * d;
      }
    }
  }
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)C + eps);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int v = lane + 32 * k;
    if (v < vecs) {
      float o8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o8[j] = (f[k][j] - mean) * rstd * gamma[v * 8 + j] + beta[v * 8 + j];
      st8(out + row * C + v * 8, o8);
    }
  }
}
// out[r][c] = proj[r][c] * gelu(proj[r][Ci + c])   (GEGLU, attention.py:37-44; exact erf GELU like F.gelu)
__global__ void __launch_bounds__(256) k_geglu(const act_t *__restrict__ proj, act_t *__restrict__ out, long long total, int Ci) {
  const int vecs = Ci >> 3;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long r = i / vecs;
    const int c0 = (int)(i - r * vecs) * 8;
    float a[8], g[8];
    ld8(proj + r * 2 * Ci + c0, a);
    ld8(proj + r * 2 * Ci + Ci + c0, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= 0.5f * g[j] * (1.f + erff(g[j] * 0.70710678118654752f));
    st8(out + r * Ci + c0, a);
  }
}
// timestep_embedding (util.py:173-197): freqs = exp(-ln(max_period) * i / half), emb = [cos(t f) | sin(t f)] (+ 0 if dim is odd)
__global__ void k_timestep_embedding(const float *__restrict__ t, float *__restrict__ out, int n, int dim, float max_period) {
  const int half = dim / 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * dim; i += gridDim.x * blockDim.x) {
    const int b = i / dim, j = i - b * dim;
    float v = 0.f;
    if (j < 2 * half) {
      const int k = j < half ? j : j - half;
      const float freq = expf(-logf(max_period) * (float)k / (float)half);
      const float arg = t[b] * freq;
      v = j < half ? cosf(arg) : sinf(arg);
    }
    out[i] = v;
  }
}

// out[r][j] = sum_k f(x[r][k]) w[j][k] + b[j], f = SiLU or identity: the embedding MLPs (a handful of rows against a wide fp32
// weight).  One warp per output feature streams its weight row once (float4, coalesced) for up to 8 rows of x at a time.
__global__ void __launch_bounds__(256) k_linear_rows_f32(const float *__restrict__ x, const float *__restrict__ w,
                                                         const float *__restrict__ b, float *__restrict__ out, int n, int K, int N,
                                                         int silu_in) {
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= N) return;
  const float *wr = w + (size_t)j * K;
  for (int r0 = 0; r0 < n; r0 += 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int nr = min(8, n - r0);
    if ((K & 3) == 0) {
      for (int k = lane * 4; k < K; k += 128) {
        const float4 wv = __ldg(reinterpret_cast<const float4 *>(wr + k));
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (r < nr) {
            float4 xv = __ldg(reinterpret_cast<const float4 *>(x + (size_t)(r0 + r) * K + k));
            if (silu_in) {
              xv.x = xv.x / (1.f + expf(-xv.x));
              xv.y = xv.y / (1.f + expf(-xv.y));
              xv.z = xv.z / (1.f + expf(-xv.z));
              xv.w = xv.w / (1.f + expf(-xv.w));
            }
            acc[r] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[r]))));
          }
      }
    } else {
      for (int k = lane; k < K; k += 32) {
        const float wv = __ldg(wr + k);
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (r < nr) {
            float xv = __ldg(x + (size_t)(r0 + r) * K + k);
            if (silu_in) xv = xv / (1.f + expf(-xv));
            acc[r] = fmaf(wv, xv, acc[r]);
          }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float v = acc[r];
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && r < nr) out[(size_t)(r0 + r) * N + j] = v + (b ? b[j] : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------ multi-head attention
// (sample, head) unit g = b * heads + h.  Q rows are padded to Tqp (multiple of 128), keys to Tkp (multiple of 128), the head
// width to dp (multiple of 64): zero padding, masked in the softmax.
// Head packs as ONE launch (blocks [0, gq) pack Q as the A operand, [gq, gq + gk) pack K as the B operand of S = Q K^T -- operand
// row (g * Tkp + j), logical length dp --, the rest pack V^T as the B operand of O = P V -- operand row (g * dp + c), logical
// length Tkp), reading q / k / v with their own row strides: they may be column slices of one fused projection output.
__global__ void __launch_bounds__(256) k_heads_pack_all(const act_t *__restrict__ q, int ldq, const act_t *__restrict__ k, int ldk,
                                                        const act_t *__restrict__ v, int ldv, act_t *__restrict__ Qh,
                                                        wop_t *__restrict__ Kh, wop_t *__restrict__ Vt, int gq, int gk, long long totq,
                                                        long long totk, long long totv, int Tq, int Tqp, int Tk, int Tkp, int heads,
                                                        int d, int dp) {
  if ((int)blockIdx.x < gq) {
    const int vecs = dp >> 3;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < totq; i += (long long)gq * 256) {
      const int vv = (int)(i % vecs);
      long long r = i / vecs;
      const int t = (int)(r % Tqp);
      const long long g = r / Tqp;
      const int h = (int)(g % heads);
      const long long b = g / heads;
      avec val = avec_zero();
      if (t < Tq && vv * 8 < d) val = ldvec(q + (b * Tq + t) * ldq + h * d + vv * 8);
      stvec(Qh + (g * Tqp + t) * dp + vv * 8, val);
    }
  } else if ((int)blockIdx.x < gq + gk) {
    // K: one 8-channel vector of one key per thread (totk counts vectors)
    const int dv = dp >> 3;
    for (long long i = (long long)(blockIdx.x - gq) * 256 + threadIdx.x; i < totk; i += (long long)gk * 256) {
      const int vv = (int)(i % dv);
      long long r = i / dv;
      const int j = (int)(r % Tkp);
      const long long g = r / Tkp;
      const int h = (int)(g % heads);
      const long long b = g / heads;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (j < Tk && vv * 8 < d) ld8(k + (b * Tk + j) * ldk + h * d + vv * 8, f);
      wop_store8(Kh + (size_t)(g * Tkp + j) * dp * kWopK, vv * 8, dp, f);
    }
  } else {
    // V^T: 64 keys x 64 channels tiles through shared memory (coalesced reads along channels, 16-byte stores along keys);
    // totv counts tiles
    __shared__ float tile[64][65];
    const int gv = gridDim.x - gq - gk, jt_n = Tkp >> 6, ct_n = dp >> 6;
    for (long long t = blockIdx.x - gq - gk; t < totv; t += gv) {
      const int ct = (int)(t % ct_n);
      long long r = t / ct_n;
      const int jt = (int)(r % jt_n);
      const long long g = r / jt_n;
      const int h = (int)(g % heads);
      const long long b = g / heads;
      __syncthreads();
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int idx = threadIdx.x + 256 * w, jj = idx >> 3, vv = idx & 7;
        const int j = jt * 64 + jj, c0 = ct * 64 + vv * 8;
        float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (j < Tk && c0 < d) ld8(v + (b * Tk + j) * ldv + h * d + c0, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) tile[jj][vv * 8 + e] = f[e];
      }
      __syncthreads();
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int idx = threadIdx.x + 256 * w, cc = idx >> 3, jv = idx & 7;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = tile[jv * 8 + e][cc];
        wop_store8(Vt + (size_t)(g * dp + ct * 64 + cc) * Tkp * kWopK, jt * 64 + jv * 8, Tkp, f);
      }
    }
  }
}
// P = softmax(scale * S) over the Tk valid keys; padded query rows and padded keys get 0.  One warp per row.
__global__ void __launch_bounds__(256) k_softmax_rows(const float *__restrict__ S, act_t *__restrict__ P, long long rows, int Tq,
                                                      int Tqp, int Tk, int Tkp, float scale) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float *s = S + row * Tkp;
  act_t *p = P + row * Tkp;
  if ((int)(row % Tqp) >= Tq) {
    for (int j = lane; j < Tkp; j += 32) p[j] = act_from_float(0.f);
    return;
  }
  float mx = -INFINITY;
  for (int j = lane; j < Tk; j += 32) mx = fmaxf(mx, s[j] * scale);
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < Tk; j += 32) sum += expf(s[j] * scale - mx);
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  for (int j = lane; j < Tkp; j += 32) p[j] = act_from_float(j < Tk ? expf(s[j] * scale - mx) * inv : 0.f);
}
__global__ void __launch_bounds__(256) k_heads_merge(const act_t *__restrict__ oh, act_t *__restrict__ out, long long total, int Tq,
                                                     int Tqp, int C, int heads, int d, int dp) {
  const int vecs = d >> 3;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int v = (int)(i % vecs);
    long long r = i / vecs;
    const int h = (int)(r % heads);
    r /= heads;
    const int t = (int)(r % Tq);
    const long long b = r / Tq;
    const long long g = b * heads + h;
    stvec(out + (b * Tq + t) * C + h * d + v * 8, ldvec(oh + (g * Tqp + t) * dp + v * 8));
  }
}

static int pick_bn_ops(int N, long long M) {
  static int forced = -1;  // SALUN_OPS_BN=64|128|160|256: A/B timing of the column-tile width
  if (forced < 0) {
    const char *e = getenv("SALUN_OPS_BN");
    forced = e ? atoi(e) : 0;
  }
  if (forced) return forced;
  if (N % 256 == 0 && ((M + 127) / 128) * (N / 256) >= 96) return 256;
  if (N % 128 == 0) return 128;
  // N = 320, 960, ...: 160-wide tiles re-read the A tile from L2 N/160 times instead of N/64 times
  return N % 160 == 0 ? 160 : 64;
}
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace salun

using namespace salun;

extern "C" {

int salun_act_bytes(void) { return (int)sizeof(act_t); }
int salun_wop_k(void) { return kWopK; }

int salun_op_f32_to_act(salun_ctx *ctx, const float *src, int64_t ld_src, void *dst, int64_t ld_dst, int64_t rows, int64_t cols,
                        void *stream) {
  SALUN_REQUIRE(ctx && src && dst && rows >= 0 && cols >= 0, "bad argument");
  if (rows * cols == 0) return SALUN_OK;
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  k_f32_to_act<<<grid1d(rows * cols), 256, 0, (cudaStream_t)stream>>>(src, ld_src, (act_t *)dst, ld_dst, rows, cols);
  ++g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}
int salun_op_act_to_f32(salun_ctx *ctx, const void *src, int64_t ld_src, float *dst, int64_t ld_dst, int64_t rows, int64_t cols,
                        void *stream) {
  SALUN_REQUIRE(ctx && src && dst && rows >= 0 && cols >= 0, "bad argument");
  if (rows * cols == 0) return SALUN_OK;
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  k_act_to_f32<<<grid1d(rows * cols), 256, 0, (cudaStream_t)stream>>>((const act_t *)src, ld_src, dst, ld_dst, rows, cols);
  ++g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}
int salun_op_nchw_to_padded(salun_ctx *ctx, const float *x, void *out_pad, int n, int C, int Cp, int H, int W, void *stream) {
  SALUN_REQUIRE(ctx && x && out_pad && n > 0 && C > 0 && Cp >= C && Cp % 8 == 0, "bad argument");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  const long long total = (long long)n * H * W * Cp;
  k_nchw_to_padded<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(x, (act_t *)out_pad, total, C, Cp, H, W);
  ++g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}
int salun_op_padded_to_nchw(salun_ctx *ctx, const void *in_pad, float *out, int n, int C, int H, int W, void *stream) {
  SALUN_REQUIRE(ctx && in_pad && out && n > 0, "bad argument");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  const long long total = (long long)n * C * H * W;
  k_padded_to_nchw<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>((const act_t *)in_pad, out, total, C, H, W);
  ++g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}
int salun_op_rows_to_nchw(salun_ctx *ctx, const float *y, int ld, const float *bias, float *out, int n, int C, int H, int W,
                          void *stream) {
  SALUN_REQUIRE(ctx && y && out && n > 0 && C > 0 && ld >= C, "bad argument");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  const long long total = (long long)n * C * H * W;
  k_rows_to_nchw<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(y, ld, bias, out, total, C, H, W);
  ++g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}
int salun_op_prep_weight(salun_ctx *ctx, const float *w, void *wop, int cout, int cin, int ks, int cout_pad, int cin_pad,
                         void *stream) {
  SALUN_REQUIRE(ctx && w && wop, "NULL argument");
  SALUN_REQUIRE((ks == 1 || ks == 3) && cout_pad >= cout && cin_pad >= cin && cout_pad % 64 == 0 && (ks * ks * cin_pad) % 64 == 0,
                "ks in {1,3}; cout_pad % 64 == 0; ks*ks*cin_pad % 64 == 0");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  k_prep_weight<<<grid1d((long long)cout_pad * ks * ks * cin_pad), 256, 0, (cudaStream_t)stream>>>(w, (wop_t *)wop, cout, cin, ks,
                                                                                                cout_pad, cin_pad);
  ++g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_op_conv(salun_ctx *ctx, const void *in, int in_flat, const void *wop, const float *bias, const float *rowbias, int rb_ld,
                  const void *addend, void *out, int out_pad, float *out_f32, int n, int H, int W, int cin, int cout, int ks,
                  void *stream) {
  SALUN_REQUIRE(ctx && in && wop && (out || out_f32), "NULL argument");
  SALUN_REQUIRE(ks == 1 || ks == 3, "ks must be 1 or 3");
  SALUN_REQUIRE(cin % 64 == 0 && cout % 64 == 0, "cin and cout must be multiples of 64 (pad the weight operand)");
  SALUN_REQUIRE(!(in_flat && ks != 1), "a flat input is a token matrix: ks must be 1");
  SALUN_REQUIRE(!(out_f32 && (out_pad || addend)), "the fp32 output is flat and takes no addend");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  const long long M = (long long)n * H * W;
  const int bn = pick_bn_ops(cout, M);
  CUtensorMap tmA, tmB;
  int rc;
  ConvGemmArgs a{};
  if (in_flat) {
    if ((rc = make_tmap_2d_act(&tmA, (const act_t *)in, (uint64_t)M, cin, 128))) return rc;
    a.mode_a = 0;
    a.num_k_blocks = cin / 64;
  } else {
    TmapBox4 bx;
    if ((rc = conv_box(H, W, 128, &bx))) return rc;
    if ((rc = make_tmap_4d_act(&tmA, (const act_t *)in, cin, W + 2, H + 2, n, bx))) return rc;
    a.mode_a = 1;
    a.cin_blocks = cin / 64;
    a.num_k_blocks = ks * ks * a.cin_blocks;
    a.kw = ks;
    a.tap_y0 = a.tap_x0 = ks == 3 ? 0 : 1;
    a.H = H;
    a.W = W;
  }
  if ((rc = make_tmap_2d_wop(&tmB, (const wop_t *)wop, cout, (uint64_t)ks * ks * cin, bn))) return rc;
  a.M = (int)M;
  a.N = cout;
  a.fH = H;
  a.fW = W;
  a.out_bf16 = (act_t *)out;
  a.out_f32 = out_f32;
  a.ld_out = cout;
  a.out_pad = out_pad;
  a.bias = bias;
  a.addend = (const act_t *)addend;
  if (rowbias) {
    SALUN_REQUIRE(((long long)H * W & ((long long)H * W - 1)) == 0, "rowbias needs a power-of-two pixel count per sample");
    a.rowbias = rowbias;
    a.rb_ld = rb_ld;
    a.rb_shift = ilog2i((long long)H * W);
  }
  a.splitk_ws = ctx->op_scratch;
  a.splitk_ws_floats = ctx->op_scratch_floats;
  return launch_conv_gemm(tmA, tmB, a, bn, (cudaStream_t)stream);
}

// Downsample (openaimodel.py:131-160: conv 3x3, stride 2, padding 1): patches of the padded input -> GEMM
int salun_op_conv_s2(salun_ctx *ctx, const void *in_pad, void *col_scratch, const void *wop, const float *bias, void *out_padded,
                     int n, int Hin, int Win, int cin, int cout, void *stream) {
  SALUN_REQUIRE(ctx && in_pad && col_scratch && wop && out_padded, "NULL argument");
  SALUN_REQUIRE(cin % 64 == 0 && cout % 64 == 0 && Hin % 2 == 0 && Win % 2 == 0, "cin, cout % 64 == 0; even image size");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  launch_im2col_s2((const act_t *)in_pad, (act_t *)col_scratch, n, Hin, Win, cin, 3, st);
  const int Ho = Hin / 2, Wo = Win / 2;
  const long long M = (long long)n * Ho * Wo;
  const int bn = pick_bn_ops(cout, M);
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tmap_2d_act(&tmA, (const act_t *)col_scratch, (uint64_t)M, 9 * (uint64_t)cin, 128))) return rc;
  if ((rc = make_tmap_2d_wop(&tmB, (const wop_t *)wop, cout, 9 * (uint64_t)cin, bn))) return rc;
  ConvGemmArgs a{};
  a.mode_a = 0;
  a.num_k_blocks = 9 * cin / 64;
  a.M = (int)M;
  a.N = cout;
  a.fH = Ho;
  a.fW = Wo;
  a.out_bf16 = (act_t *)out_padded;
  a.ld_out = cout;
  a.out_pad = 1;
  a.bias = bias;
  a.splitk_ws = ctx->op_scratch;
  a.splitk_ws_floats = ctx->op_scratch_floats;
  return launch_conv_gemm(tmA, tmB, a, bn, st);
}

// Caller-owned scratch for the split-K path of salun_op_conv / salun_op_conv_s2 (small-M, deep-K GEMMs: the 8x8 and 16x16
// levels of the U-Net at batch 1-2 have 10-40 output tiles for 148 SMs).  NULL / 0 turns the path off.  The buffer must stay
// valid, and ops of one context must not run concurrently on several streams, while it is set.
int salun_op_set_scratch(salun_ctx *ctx, void *scratch, int64_t bytes) {
  SALUN_REQUIRE(ctx && bytes >= 0 && (scratch || bytes == 0), "bad argument");
  SALUN_REQUIRE(((uintptr_t)scratch & 15) == 0, "scratch must be 16-byte aligned");
  ctx->op_scratch = (float *)scratch;
  ctx->op_scratch_floats = bytes / 4;
  return SALUN_OK;
}
int64_t salun_op_groupnorm_ws_floats(int n) { return (int64_t)n * 32 * kGnMaxSplits * 4; }
int salun_op_groupnorm(salun_ctx *ctx, const void *in_pad, const float *gamma, const float *beta, float *stats_ws, void *out,
                       int out_flat, int n, int H, int W, int C, float eps, int swish, void *stream) {
  SALUN_REQUIRE(ctx && in_pad && gamma && beta && stats_ws && out, "NULL argument");
  SALUN_REQUIRE(C % 32 == 0 && C % 8 == 0 && n > 0 && n <= 65535 && C <= 4096, "C must be a multiple of 32 (<= 4096)");
  SALUN_REQUIRE(((uintptr_t)stats_ws & 7) == 0, "stats_ws must be 8-byte aligned");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  double *part = reinterpret_cast<double *>(stats_ws);
  const int HW = H * W, S = gn_splits(HW), vecs = C >> 3;
  const int R = vecs <= 256 ? 256 / vecs : 1;
  const size_t smem = (size_t)R * C * 2 * sizeof(float);
  int gx = (HW + 2 * R - 1) / (2 * R);           // >= 2 pixels per row lane and CTA
  const int cap = (148 * 6 + n - 1) / n;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  if (vecs <= 256) {
    k_gn2_partial<1><<<dim3(S, n), 256, smem, st>>>((const act_t *)in_pad, part, H, W, C, S);
    k_gn2_apply<1><<<dim3(gx, n), 256, 0, st>>>((const act_t *)in_pad, part, gamma, beta, (act_t *)out, out_flat, swish, H, W, C, S, eps);
  } else {
    k_gn2_partial<2><<<dim3(S, n), 256, smem, st>>>((const act_t *)in_pad, part, H, W, C, S);
    k_gn2_apply<2><<<dim3(gx, n), 256, 0, st>>>((const act_t *)in_pad, part, gamma, beta, (act_t *)out, out_flat, swish, H, W, C, S, eps);
  }
  g_launch_count += 2;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_op_upsample2(salun_ctx *ctx, const void *in_pad, void *out_pad, int n, int H, int C, void *stream) {
  SALUN_REQUIRE(ctx && in_pad && out_pad && C % 8 == 0 && (H & (H - 1)) == 0, "square power-of-two images, C % 8 == 0");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  launch_upsample2((const act_t *)in_pad, (act_t *)out_pad, n, H, C, (cudaStream_t)stream);
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}
int salun_op_concat(salun_ctx *ctx, const void *a_pad, int Ca, const void *b_pad, int Cb, void *out_pad, int n, int H, void *stream) {
  SALUN_REQUIRE(ctx && a_pad && b_pad && out_pad && Ca % 8 == 0 && Cb % 8 == 0 && (H & (H - 1)) == 0, "bad argument");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  launch_concat((const act_t *)a_pad, Ca, (const act_t *)b_pad, Cb, (act_t *)out_pad, n, H, (cudaStream_t)stream);
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}
// out[n][N] = act_in(x)[n][K] . w[N][K]^T + b   (fp32, CUDA cores: the embedding MLPs); silu_in: x <- x * sigmoid(x) first (tmp)
int salun_op_linear_f32(salun_ctx *ctx, const float *x, const float *w, const float *b, float *out, float *tmp, int n, int K, int N,
                        int silu_in, void *stream) {
  SALUN_REQUIRE(ctx && x && w && out, "NULL argument");
  SALUN_REQUIRE(n > 0 && K > 0 && N > 0, "bad sizes");
  (void)tmp;  // kept in the signature: SiLU is applied on the fly
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  k_linear_rows_f32<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, w, b, out, n, K, N, silu_in);
  ++g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}
int salun_sd_timestep_embedding(salun_ctx *ctx, const float *t, float *out, int n, int dim, float max_period, void *stream) {
  SALUN_REQUIRE(ctx && t && out && n > 0 && dim > 1, "bad argument");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  k_timestep_embedding<<<grid1d((long long)n * dim), 256, 0, (cudaStream_t)stream>>>(t, out, n, dim, max_period);
  ++g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}
int salun_sd_layernorm(salun_ctx *ctx, const void *x, const float *gamma, const float *beta, void *out, int64_t rows, int C,
                       float eps, void *stream) {
  SALUN_REQUIRE(ctx && x && gamma && beta && out, "NULL argument");
  SALUN_REQUIRE(C % 8 == 0 && C <= 2048, "C % 8 == 0 and C <= 2048");
  if (rows == 0) return SALUN_OK;
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  k_layernorm<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const act_t *)x, gamma, beta, (act_t *)out, rows, C, eps);
  ++g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}
int salun_sd_geglu(salun_ctx *ctx, const void *proj, void *out, int64_t rows, int Ci, void *stream) {
  SALUN_REQUIRE(ctx && proj && out && Ci % 8 == 0, "bad argument");
  if (rows == 0) return SALUN_OK;
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  const long long total = rows * (Ci >> 3);
  k_geglu<<<grid1d(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>((const act_t *)proj, (act_t *)out, total, Ci);
  ++g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

// head widths <= 64 take the fused kernel (salun_attn.cu: S and P stay in tensor / shared memory); SALUN_FLASH_ATTN=0 forces the
// unfused product - softmax - product chain (bring-up and A/B timing)
static bool use_flash_attn(int d) {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SALUN_FLASH_ATTN");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1 && flash_attn_supported(d);
}
// workspace carve-up of salun_sd_attention
static void attn_sizes(int n, int Tq, int Tk, int heads, int d, int *Tqp, int *Tkp, int *dp, size_t off[6], size_t *total) {
  *Tqp = (Tq + 127) / 128 * 128;
  *Tkp = (Tk + 127) / 128 * 128;
  *dp = (d + 63) / 64 * 64;
  const size_t G = (size_t)n * heads;
  size_t o = 0;
  off[0] = o; o = align_up(o + G * *Tqp * *dp * sizeof(act_t), 1024);                 // Qh  (A operand)
  off[1] = o; o = align_up(o + G * *Tkp * *dp * kWopK * sizeof(wop_t), 1024);         // Kh  (B operand of S)
  off[2] = o; o = align_up(o + G * *dp * *Tkp * kWopK * sizeof(wop_t), 1024);         // Vt  (B operand of O)
  off[3] = o; o = align_up(o + G * *Tqp * (size_t)*Tkp * sizeof(float), 1024);        // S   fp32
  off[4] = o; o = align_up(o + G * *Tqp * (size_t)*Tkp * sizeof(act_t), 1024);        // P
  off[5] = o; o = align_up(o + G * *Tqp * *dp * sizeof(act_t), 1024);                 // Oh
  *total = o;
}
int64_t salun_sd_attention_ws_bytes(int n, int Tq, int Tk, int heads, int d) {
  int a, b, c;
  size_t off[6], total;
  attn_sizes(n, Tq, Tk, heads, d, &a, &b, &c, off, &total);
  return (int64_t)total;
}
// out[n*Tq][heads*d] = softmax(q_h k_h^T / sqrt(d)) v_h per (sample, head)   (CrossAttention.forward, attention.py:168-192;
// q [n*Tq][C], k / v [n*Tk][C], C = heads * d; self-attention: k, v from the same tokens, Tk = Tq)
int salun_sd_attention(salun_ctx *ctx, void *ws, int64_t ws_bytes, const void *q, const void *k, const void *v, void *out, int n,
                       int Tq, int Tk, int heads, int d, void *stream) {
  return salun_sd_attention_ld(ctx, ws, ws_bytes, q, heads * d, k, heads * d, v, heads * d, out, n, Tq, Tk, heads, d, stream);
}
// the same with row strides (in elements) for q, k, v: column slices of a fused q | k | v (or k | v) projection output
int salun_sd_attention_ld(salun_ctx *ctx, void *ws, int64_t ws_bytes, const void *q, int ldq, const void *k, int ldk, const void *v,
                          int ldv, void *out, int n, int Tq, int Tk, int heads, int d, void *stream) {
  SALUN_REQUIRE(ctx && ws && q && k && v && out, "NULL argument");
  SALUN_REQUIRE(n > 0 && Tq > 0 && Tk > 0 && heads > 0 && d > 0 && d % 8 == 0, "bad sizes (d % 8 == 0)");
  SALUN_REQUIRE(ldq >= heads * d && ldk >= heads * d && ldv >= heads * d && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0,
                "row strides must cover heads * d and be multiples of 8");
  int Tqp, Tkp, dp;
  size_t off[6], total;
  attn_sizes(n, Tq, Tk, heads, d, &Tqp, &Tkp, &dp, off, &total);
  SALUN_REQUIRE((size_t)ws_bytes >= total, "workspace too small (salun_sd_attention_ws_bytes)");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int C = heads * d;
  const long long G = (long long)n * heads;
  char *w = (char *)ws;
  act_t *Qh = (act_t *)(w + off[0]);
  wop_t *Kh = (wop_t *)(w + off[1]), *Vt = (wop_t *)(w + off[2]);
  float *S = (float *)(w + off[3]);
  act_t *P = (act_t *)(w + off[4]), *Oh = (act_t *)(w + off[5]);
  {
    const long long totq = G * Tqp * (dp >> 3), totk = G * Tkp * (dp >> 3), totv = G * (Tkp >> 6) * (dp >> 6);  // vectors | vectors | tiles
    const int gq = grid1d(totq, 256, 148 * 4), gk = grid1d(totk, 256, 148 * 4);
    const int gv = (int)(totv < 148 * 4 ? totv : 148 * 4);
    k_heads_pack_all<<<gq + gk + gv, 256, 0, st>>>((const act_t *)q, ldq, (const act_t *)k, ldk, (const act_t *)v, ldv, Qh, Kh, Vt, gq,
                                                  gk, totq, totk, totv, Tq, Tqp, Tk, Tkp, heads, d, dp);
    ++g_launch_count;
  }
  if (use_flash_attn(d)) return launch_flash_attn(Qh, Kh, Vt, (act_t *)out, n, Tq, Tqp, Tk, Tkp, heads, d, st);
  const long long M = G * Tqp;
  int rc;
  {  // S = Qh Kh^T per unit
    const int bn = pick_bn_ops(Tkp, M);
    CUtensorMap tmA, tmB;
    if ((rc = make_tmap_2d_act(&tmA, Qh, (uint64_t)M, dp, 128))) return rc;
    if ((rc = make_tmap_2d_wop(&tmB, Kh, (uint64_t)G * Tkp, dp, bn))) return rc;
    ConvGemmArgs a{};
    a.mode_a = 0;
    a.num_k_blocks = dp / 64;
    a.M = (int)M;
    a.N = Tkp;
    a.out_f32 = S;
    a.ld_out = Tkp;
    a.batch_rows_a = Tqp;
    a.batch_rows_b = Tkp;
    if ((rc = launch_conv_gemm(tmA, tmB, a, bn, st))) return rc;
  }
  k_softmax_rows<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(S, P, M, Tq, Tqp, Tk, Tkp, 1.f / sqrtf((float)d));
  ++g_launch_count;
  {  // Oh = P Vt^T per unit
    const int bn = pick_bn_ops(dp, M);
    CUtensorMap tmA, tmB;
    if ((rc = make_tmap_2d_act(&tmA, P, (uint64_t)M, Tkp, 128))) return rc;
    if ((rc = make_tmap_2d_wop(&tmB, Vt, (uint64_t)G * dp, Tkp, bn))) return rc;
    ConvGemmArgs a{};
    a.mode_a = 0;
    a.num_k_blocks = Tkp / 64;
    a.M = (int)M;
    a.N = dp;
    a.out_bf16 = Oh;
    a.ld_out = dp;
    a.batch_rows_a = Tqp;
    a.batch_rows_b = dp;
    if ((rc = launch_conv_gemm(tmA, tmB, a, bn, st))) return rc;
  }
  const long long tot = (long long)n * Tq * heads * (d >> 3);
  k_heads_merge<<<grid1d(tot, 256, 148 * 16), 256, 0, st>>>(Oh, (act_t *)out, tot, Tq, Tqp, C, heads, d, dp);
  ++g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

}  // extern "C"
