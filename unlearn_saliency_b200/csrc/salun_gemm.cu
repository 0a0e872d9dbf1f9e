// salun_gemm.cu -- tcgen05 / TMEM / TMA GEMM kernels for the convolution forward, dgrad and wgrad
// of the SalUn ResNet path (replaces the cuDNN calls behind Classification/models/ResNet.py:58-74,
// 108-124 forward and loss.backward() at Classification/unlearn/RL.py:128-132).
//
// k_conv_gemm : D[M][N] = A[M][K] . B[N][K]^T, both operands K-major bf16 in 128B-swizzled smem tiles
//               written by TMA; accumulators in TMEM; warp-specialised:
//                 warp 0      TMA producer (one elected lane)
//                 warp 1      TMEM allocator + tcgen05.mma issuer (one elected lane)
//                 warps 2..5  epilogue: tcgen05.ld -> registers -> bf16/fp32 global stores (+ per-column
//                             sum / sum-of-squares partials for the BatchNorm batch statistics)
//               For stride-1 convolutions A is never materialised: k-block kb = (tap, 64 input channels) is
//               fetched straight from the halo-padded NHWC activation by a 4-D TMA box whose (x, y) origin is
//               shifted by the tap -- TMA does the im2col.
// k_wgrad     : dW[co][kc] += sum_pixels dY[p][co] * X[p][kc]; the pixel dimension is the MMA K dimension, so both
//               operands are MN-major tiles ([pixels][64 channels], the same TMA boxes as above); split-K over
//               pixel ranges with fp32 red.global.add.
#include <stdlib.h>

#include <vector>

#include "salun_gemm.cuh"
#include "salun_sm100.cuh"

namespace salun {
using namespace sm100;

constexpr int kGemmThreads = 192;
// k_conv_gemm_p: 8 epilogue warps (two per TMEM lane quarter, alternating 32-column chunks).  The epilogue is a latency
// chain per warp (tcgen05.ld -> wait -> adds -> 4 row-scattered 16-byte stores); with short K (the 1x1 convolutions and
// attention products of the U-Net, K = 256) it is the slow side of the kernel: role timing showed the MMA issuer waiting on
// the accumulator hand-back for 43 % of its life with 4 epilogue warps (profiles/README.md section 5).
constexpr int kGemmPThreads = 320;
constexpr int kBM = 128;
constexpr int kBK = 64;  // bf16 elements per k-block = one 128-byte swizzle row

// Transposing reduction: every lane holds v[0..31] (32 columns of its own row); afterwards lane j holds the
// sum over the 32 lanes (rows) of column j.  31 shuffles instead of 160.
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const bool up = lane & 16;
    float send = up ? v[i] : v[i + 16];
    float keep = up ? v[i + 16] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 8;
    float send = up ? v[i] : v[i + 8];
    float keep = up ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 4;
    float send = up ? v[i] : v[i + 4];
    float keep = up ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 2;
    float send = up ? v[i] : v[i + 2];
    float keep = up ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const bool up = lane & 1;
    float send = up ? v[0] : v[1];
    float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return v[0];
}

// r[0..31] (fp32 accumulators of 32 consecutive columns) += 32 consecutive activation elements at src
__device__ __forceinline__ void epi_add_act32(uint32_t (&r)[32], const act_t *src) {
#ifdef SALUN_SPLIT
  const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 v = s4[j];
    r[4 * j + 0] = __float_as_uint(__uint_as_float(r[4 * j + 0]) + act_unpack_pair(v.x));
    r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + act_unpack_pair(v.y));
    r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + act_unpack_pair(v.z));
    r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + act_unpack_pair(v.w));
  }
#else
  const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 v = s4[j];
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(h[i]);
      r[8 * j + 2 * i] = __float_as_uint(__uint_as_float(r[8 * j + 2 * i]) + t.x);
      r[8 * j + 2 * i + 1] = __float_as_uint(__uint_as_float(r[8 * j + 2 * i + 1]) + t.y);
    }
  }
#endif
}
// 32 consecutive activation elements at dst <- r[0..31] (row-per-lane store straight out of the TMEM layout)
__device__ __forceinline__ void epi_store_act32(act_t *dst, const uint32_t (&r)[32]) {
  uint4 *d4 = reinterpret_cast<uint4 *>(dst);
#ifdef SALUN_SPLIT
#pragma unroll
  for (int j = 0; j < 8; ++j)
    d4[j] = make_uint4(act_pack_pair(__uint_as_float(r[4 * j])), act_pack_pair(__uint_as_float(r[4 * j + 1])),
                       act_pack_pair(__uint_as_float(r[4 * j + 2])), act_pack_pair(__uint_as_float(r[4 * j + 3])));
#else
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 v;
    v.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1]));
    v.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
    v.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
    v.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
    d4[j] = v;
  }
#endif
}

#ifndef SALUN_SPLIT
// BN-backward partial sums of one 32-column chunk held in r[] (fp32 dX of row `row`): see BnBwdFuse.
__device__ __forceinline__ void bn_bwd_fuse_chunk(const BnBwdFuse &f, const uint32_t (&r)[32], bool row_ok,
                                                  size_t pad_off_elems, size_t flat_off_elems, int col0, int N,
                                                  int stat_row, int lane) {
  uint32_t maskbits = 0;
  uint4 yv[4];
  if (row_ok) {
    const uint4 *ap = reinterpret_cast<const uint4 *>(f.act + pad_off_elems + col0);
    const uint4 *yp = reinterpret_cast<const uint4 *>(f.y + flat_off_elems + col0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 av = ap[j];
      yv[j] = yp[j];
      const uint32_t w[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // bf16 > 0  <=>  sign bit clear and magnitude non-zero (activations are post-ReLU: never negative, no NaN)
        maskbits |= ((w[i] & 0x7fffu) != 0 ? 1u : 0u) << (8 * j + 2 * i);
        maskbits |= ((w[i] & 0x7fff0000u) != 0 ? 1u : 0u) << (8 * j + 2 * i + 1);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) yv[j] = make_uint4(0, 0, 0, 0);
  }
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = (maskbits >> j) & 1u ? __uint_as_float(r[j]) : 0.f;
  const float s1 = warp_transpose_reduce(v, lane);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const uint32_t *yw = reinterpret_cast<const uint32_t *>(&yv[j >> 3]);
    const uint32_t pair = yw[(j & 7) >> 1];
    const float yy = __uint_as_float((j & 1) ? (pair & 0xffff0000u) : (pair << 16));
    const float xh = (yy - __ldg(f.mean + col0 + j)) * __ldg(f.invstd + col0 + j);
    v[j] = (maskbits >> j) & 1u ? __uint_as_float(r[j]) * xh : 0.f;
  }
  const float s2 = warp_transpose_reduce(v, lane);
  f.partials[((size_t)stat_row * 2 + 0) * N + col0 + lane] = s1;
  f.partials[((size_t)stat_row * 2 + 1) * N + col0 + lane] = s2;
}
#else
__device__ __forceinline__ void bn_bwd_fuse_chunk(const BnBwdFuse &, const uint32_t (&)[32], bool, size_t, size_t, int, int,
                                                  int, int) {}  // the fused BatchNorm-backward partials are a bf16-build option
#endif
__device__ __forceinline__ size_t pad_row_off(int m, int H, int W, int C) {
  const int hw = H * W;
  const int n = m / hw, rr = m - n * hw;
  const int y = rr / W, x = rr - y * W;
  return ((size_t)(n * (H + 2) + y + 1) * (W + 2) + x + 1) * C;
}

#ifndef SALUN_SPLIT
// =================================================================================================
// conv / plain GEMM, K-major operands
// =================================================================================================
template <int BN, int kStages>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_conv_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvGemmArgs a) {
  constexpr uint32_t kABytes = kBM * kBK * 2;  // 16 KB
  constexpr uint32_t kBBytes = BN * kBK * 2;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kStages * kStageBytes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kStages, tfull = empty0 + 8 * kStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, n_tile = blockIdx.y;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      int n0 = 0, y0 = 0;
      if (a.mode_a == 1) {
        const int pix0 = m_tile * kBM;
        n0 = pix0 / (a.H * a.W);
        y0 = (pix0 % (a.H * a.W)) / a.W;
      }
      for (int kb = 0; kb < a.num_k_blocks; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t sa = smem_base + s * kStageBytes, sb = sa + kABytes;
        mbar_arrive_expect_tx(full0 + 8 * s, kStageBytes);
        if (a.mode_a == 1) {
          const int tap = kb / a.cin_blocks, cb = kb - tap * a.cin_blocks;
          const int ky = tap / a.kw, kx = tap - ky * a.kw;
          tma_load_4d(sa, &tmA, full0 + 8 * s, cb * kBK, a.tap_x0 + kx, a.tap_y0 + y0 + ky, n0);
        } else {
          tma_load_2d(sa, &tmA, full0 + 8 * s, kb * kBK, m_tile * kBM);
        }
        tma_load_2d(sb, &tmB, full0 + 8 * s, kb * kBK, n_tile * BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, 0, 0);
      for (int kb = 0; kb < a.num_k_blocks; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * kStageBytes, sb = sa + kABytes;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          const uint64_t ad = umma_desc_sw128(sa + k * 32, 16, 1024);
          const uint64_t bd = umma_desc_sw128(sb + k * 32, 16, 1024);
          tc_mma_f16(tmem_base, ad, bd, idesc, (kb | k) != 0);
        }
        tc_commit(empty0 + 8 * s);  // frees the smem stage when the MMAs above retire
      }
      tc_commit(tfull);
    }
  } else {
    // -------------------------------------------------------------------- epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    mbar_wait(tfull, 0);
    tc_fence_after();
    const int row = m_tile * kBM + q * 32 + lane;
    const bool row_ok = row < a.M;
    const int stat_row = (m_tile * 4 + q);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tc_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, r);
      tc_wait_ld();
      const int col0 = n_tile * BN + c * 32;
      if (col0 >= a.N) break;
      if (row_ok && a.addend) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.addend + (size_t)row * a.ld_out + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 v = src[j];
          const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 t = __bfloat1622float2(h[i]);
            r[8 * j + 2 * i] = __float_as_uint(__uint_as_float(r[8 * j + 2 * i]) + t.x);
            r[8 * j + 2 * i + 1] = __float_as_uint(__uint_as_float(r[8 * j + 2 * i + 1]) + t.y);
          }
        }
      }
      if (row_ok) {
        if (a.out_bf16) {
          uint4 *dst = reinterpret_cast<uint4 *>(a.out_bf16 + (size_t)row * a.ld_out + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1]));
            v.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
            v.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
            v.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
            dst[j] = v;
          }
        }
        if (a.out_f32) {
          uint4 *dst = reinterpret_cast<uint4 *>(a.out_f32 + (size_t)row * a.ld_out + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        }
      }
      if (a.stat_sum) {
        float v[32], w[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = row_ok ? __uint_as_float(r[j]) : 0.f;
          v[j] = x;
          w[j] = x * x;
        }
        float s1 = warp_transpose_reduce(v, lane);
        float s2 = warp_transpose_reduce(w, lane);
        if (col0 + lane < a.N) {
          a.stat_sum[(size_t)stat_row * a.N + col0 + lane] = s1;
          a.stat_sq[(size_t)stat_row * a.N + col0 + lane] = s2;
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

#endif  // !SALUN_SPLIT

// =================================================================================================
// k_conv_gemm_p: PERSISTENT variant of k_conv_gemm.  One CTA per SM walks output tiles (n fastest, so that CTAs
// running concurrently share the A tile in L2); two TMEM accumulators let the epilogue of tile i run under the MMAs of
// tile i+1; barriers / TMEM / descriptors are set up once per CTA instead of once per tile (the short-K GEMMs of the
// stem and of the stride-2 dgrad spent most of their time in that setup).
// =================================================================================================
template <int BN, int kStages>
__global__ void __launch_bounds__(kGemmPThreads, 1)
k_conv_gemm_p(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvGemmArgs a,
              int m_tiles, int n_tiles) {
  constexpr uint32_t kABytes = kBM * kBK * 2;
  constexpr uint32_t kBBytes = BN * kBK * 2;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kStages * kStageBytes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kStages;
  const uint32_t tfull0 = empty0 + 8 * kStages, tempty0 = tfull0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ksplit = a.ksplit > 1 ? a.ksplit : 1;
  const int total_units = m_tiles * n_tiles * ksplit;  // (tile, k-slice) work units, k-slice fastest
  constexpr uint32_t kTmemCols = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);  // power of two >= two accumulators

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above is global-memory free (barriers, tensor memory, descriptor prefetch)

  const bool dbg_on = a.dbg != nullptr;
  const long long t_begin = dbg_on ? clock64() : 0;
  if (warp == 0) {
    if (lane < 2) {
      // lane 0 fetches the A tile, lane 1 the B tile (the issuing threads are instruction-latency bound: two lanes
      // halve the per-stage issue time; no integer division inside the k loop)
      long long w_empty = 0;
      uint32_t it = 0;
      const int hw = a.H * a.W;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const int tile = unit / ksplit, kz = unit - tile * ksplit;
        const int m_tile = tile / n_tiles, n_tile = tile - m_tile * n_tiles;
        int n0 = 0, y0 = 0;
        if (a.mode_a == 1) {
          const int pix0 = m_tile * kBM;
          n0 = pix0 / hw;
          y0 = (pix0 - n0 * hw) / a.W;
        }
        // k-blocks [kb0, kb1) of this unit (split-K: slice kz of ksplit; otherwise the whole loop)
        const int kb0 = (int)((long long)a.num_k_blocks * kz / ksplit), kb1 = (int)((long long)a.num_k_blocks * (kz + 1) / ksplit);
        int ka = (a.k_wrap > 0 && kb0 >= a.k_wrap) ? kb0 - a.k_wrap : kb0;
        int cb = 0, kx = 0, ky = 0;
        if (a.mode_a == 1 && ka) {
          cb = ka % a.cin_blocks;
          const int tap = ka / a.cin_blocks;
          ky = tap / a.kw;
          kx = tap - ky * a.kw;
        }
        const int b_row0 = n_tile * BN + (a.batch_rows_a ? (m_tile * kBM / a.batch_rows_a) * a.batch_rows_b : 0);
        for (int kb = kb0; kb < kb1; ++kb, ++ka, ++it) {
          if (kb == a.k_wrap && kb != kb0) ka = cb = kx = ky = 0;  // split build: second pass over the same activation tile
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait_t(empty0 + 8 * s, ph ^ 1, w_empty, dbg_on);
          const uint32_t sa = smem_base + s * kStageBytes, sb = sa + kABytes;
          if (lane == 0) {
            mbar_arrive_expect_tx(full0 + 8 * s, kStageBytes);
            if (a.mode_a == 1)
              tma_load_4d(sa, &tmA, full0 + 8 * s, cb * kBK, a.tap_x0 + kx, a.tap_y0 + y0 + ky, n0);
            else
              tma_load_2d(sa, &tmA, full0 + 8 * s, ka * kBK, m_tile * kBM);
          } else {
            tma_load_2d(sb, &tmB, full0 + 8 * s, kb * kBK, b_row0);
          }
          if (++cb == a.cin_blocks) {
            cb = 0;
            if (++kx == a.kw) {
              kx = 0;
              ++ky;
            }
          }
        }
      }
      if (dbg_on && lane == 0) {
        a.dbg[blockIdx.x * 8 + 0] = w_empty;
        a.dbg[blockIdx.x * 8 + 1] = clock64() - t_begin;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, 0, 0);
      const uint32_t dhi = umma_desc_hi_sw128(1024), a_lo0 = umma_desc_lo(smem_base, 16);
      long long w_full = 0, w_tempty = 0;
      uint32_t it = 0, ti = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++ti) {
        const uint32_t acc = ti & 1, aph = (ti >> 1) & 1;
        mbar_wait_t(tempty0 + 8 * acc, aph ^ 1, w_tempty, dbg_on);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int kz = unit % ksplit;
        const int kb0 = (int)((long long)a.num_k_blocks * kz / ksplit), kb1 = (int)((long long)a.num_k_blocks * (kz + 1) / ksplit);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait_t(full0 + 8 * s, ph, w_full, dbg_on);
          tc_fence_after();
          const uint32_t alo = a_lo0 + s * (kStageBytes >> 4), blo = alo + (kABytes >> 4);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            tc_mma_f16(d_tmem, umma_desc_pack(alo + 2 * k, dhi), umma_desc_pack(blo + 2 * k, dhi), idesc, ((kb - kb0) | k) != 0);
          tc_commit(empty0 + 8 * s);
        }
        tc_commit(tfull0 + 8 * acc);
      }
      if (dbg_on) {
        a.dbg[blockIdx.x * 8 + 2] = w_full;
        a.dbg[blockIdx.x * 8 + 3] = w_tempty;
        a.dbg[blockIdx.x * 8 + 4] = clock64() - t_begin;
      }
    }
  } else {
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2; // warps 2..5 take the even 32-column chunks, warps 6..9 the odd ones
    uint8_t *stg = smem + kStages * kStageBytes + 256 + (warp - 2) * 4096;  // this warp's 4 KB store-transpose buffer
    long long w_tfull = 0;
    uint32_t ti = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++ti) {
      const int tile = unit / ksplit, kz = unit - tile * ksplit;
      const int m_tile = tile / n_tiles, n_tile = tile - m_tile * n_tiles;
      float *const out_f32 = a.out_f32 ? a.out_f32 + (size_t)kz * a.split_stride : nullptr;  // split-K: slab kz of the scratch
      const uint32_t acc = ti & 1, aph = (ti >> 1) & 1;
      mbar_wait_t(tfull0 + 8 * acc, aph, w_tfull, dbg_on);
      tc_fence_after();
      const int row = m_tile * kBM + q * 32 + lane;
      const bool row_ok = row < a.M;
      const int stat_row = (m_tile * 4 + q);
      const size_t fuse_pad = (a.f1.act && row_ok) ? pad_row_off(row, a.fH, a.fW, a.N) : 0;
      const size_t out_row = !row_ok ? 0 : (a.out_pad ? pad_row_off(row, a.fH, a.fW, a.ld_out) : (size_t)row * a.ld_out);
      const float *rb_row = (a.rowbias && row_ok) ? a.rowbias + (size_t)(row >> a.rb_shift) * a.rb_ld : nullptr;
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        uint32_t r[32];
        tc_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c * 32, r);
        tc_wait_ld();
        const int col0 = n_tile * BN + c * 32;
        if (col0 < a.N) {
          if (a.bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = __ldg(reinterpret_cast<const float4 *>(a.bias + col0) + j);
              r[4 * j + 0] = __float_as_uint(__uint_as_float(r[4 * j + 0]) + b4.x);
              r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + b4.y);
              r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + b4.z);
              r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + b4.w);
            }
          }
          if (rb_row) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = __ldg(reinterpret_cast<const float4 *>(rb_row + col0) + j);
              r[4 * j + 0] = __float_as_uint(__uint_as_float(r[4 * j + 0]) + b4.x);
              r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + b4.y);
              r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + b4.z);
              r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + b4.w);
            }
          }
          if (row_ok && a.addend) epi_add_act32(r, a.addend + out_row + col0);
          if (a.stat_sum) {  // column partials of the FINAL value (after bias / projection / residual)
            float v[32], w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = row_ok ? __uint_as_float(r[j]) : 0.f;
              v[j] = x;
              w[j] = x * x;
            }
            float s1 = warp_transpose_reduce(v, lane);
            float s2 = warp_transpose_reduce(w, lane);
            if (col0 + lane < a.N) {
              a.stat_sum[(size_t)stat_row * a.N + col0 + lane] = s1;
              a.stat_sq[(size_t)stat_row * a.N + col0 + lane] = s2;
            }
          }
          if (a.f1.act) {
            bn_bwd_fuse_chunk(a.f1, r, row_ok, fuse_pad, (size_t)row * a.N, col0, a.N, stat_row, lane);
            if (a.f2.act) bn_bwd_fuse_chunk(a.f2, r, row_ok, fuse_pad, (size_t)row * a.N, col0, a.N, stat_row, lane);
          }
          // Stores go through a per-warp shared-memory transpose: straight out of the TMEM layout (lane = row) a warp
          // store touches 32 different rows with 16 bytes each (32 LSU wavefronts per instruction -- the measured bound
          // of the short-K GEMMs); transposed, one instruction writes 8 rows x 64 B (bf16) / 4 rows x 128 B (fp32).
#ifdef SALUN_SPLIT
          if (a.out_bf16) {  // (hi, lo) pairs are 4 bytes per element: the fp32 staging geometry, 4 rows x 128 B per store
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4 *>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_uint4(act_pack_pair(__uint_as_float(r[4 * j])), act_pack_pair(__uint_as_float(r[4 * j + 1])),
                             act_pack_pair(__uint_as_float(r[4 * j + 2])), act_pack_pair(__uint_as_float(r[4 * j + 3])));
            __syncwarp();
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const int rl = 4 * jj + (lane >> 3), u4 = lane & 7;
              const uint4 v = *reinterpret_cast<const uint4 *>(stg + rl * 128 + ((u4 ^ (rl & 7)) << 4));
              const unsigned long long orow = __shfl_sync(0xffffffffu, (unsigned long long)out_row, rl);
              const int ok = __shfl_sync(0xffffffffu, row_ok ? 1 : 0, rl);
              if (ok) *reinterpret_cast<uint4 *>(a.out_bf16 + orow + col0 + u4 * 4) = v;
            }
            __syncwarp();
          }
#else
          if (a.out_bf16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 v;
              v.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1]));
              v.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
              v.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
              v.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
              *reinterpret_cast<uint4 *>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = v;
            }
            __syncwarp();
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int rl = 8 * jj + (lane >> 2), u4 = lane & 3;
              const uint4 v = *reinterpret_cast<const uint4 *>(stg + rl * 64 + ((u4 ^ ((rl >> 1) & 3)) << 4));
              const unsigned long long orow = __shfl_sync(0xffffffffu, (unsigned long long)out_row, rl);
              const int ok = __shfl_sync(0xffffffffu, row_ok ? 1 : 0, rl);
              if (ok) *reinterpret_cast<uint4 *>(a.out_bf16 + orow + col0 + u4 * 8) = v;
            }
            __syncwarp();
          }
#endif
          if (out_f32) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4 *>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            __syncwarp();
            const int row_base = m_tile * kBM + q * 32;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const int rl = 4 * jj + (lane >> 3), u4 = lane & 7;
              const uint4 v = *reinterpret_cast<const uint4 *>(stg + rl * 128 + ((u4 ^ (rl & 7)) << 4));
              if (row_base + rl < a.M)
                *reinterpret_cast<uint4 *>(out_f32 + (size_t)(row_base + rl) * a.ld_out + col0 + u4 * 4) = v;
            }
            __syncwarp();
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
    }
    if (dbg_on && warp == 4 && lane == 0) {
      a.dbg[blockIdx.x * 8 + 5] = w_tfull;
      a.dbg[blockIdx.x * 8 + 6] = clock64() - t_begin;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// =================================================================================================
// k_gemm2: CTA-PAIR variant of k_conv_gemm_p (cta_group::2).  A cluster of two CTAs (two SMs of one TPC) computes a
// 256 x BN tile: CTA r holds rows [r*128, +128) of A and rows [r*BN/2, +BN/2) of B in its own shared memory, the leader
// (rank 0) issues tcgen05.mma.cta_group::2 (M = 256) into both CTAs' tensor memory.  Per CTA a k-block moves
// 16 KB (A) + BN/2 * 128 B (half of B) through the L2->SM port for 128 x BN x 64 MACs: half the B bytes of the 1-CTA
// kernel, which that port bounds (profiles/README.md section 3).
//   full[s]   : leader's barrier; both CTAs' TMA loads signal it (peer-bit-masked address), leader arrives with expect_tx
//   empty[s]  : one per CTA; the leader's tcgen05.commit multicasts the arrive to both
//   tfull[a]  : one per CTA (multicast commit); tempty[a]: leader's, 8 arrivals (4 epilogue warps x 2 CTAs)
// =================================================================================================
template <int BN, int kStages>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvGemmArgs a,
        int m_pairs, int n_tiles) {
  constexpr uint32_t kABytes = kBM * kBK * 2;
  constexpr uint32_t kBBytes = (BN / 2) * kBK * 2;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kStages * kStageBytes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kStages;
  const uint32_t tfull0 = empty0 + 8 * kStages, tempty0 = tfull0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int total_tiles = m_pairs * n_tiles;
  constexpr uint32_t kTmemCols = 2 * BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem2_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem2_relinquish();
  }
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers initialised and tensor memory allocated before anyone signals a peer
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane < 2) {
      uint32_t it = 0;
      const int hw = a.H * a.W;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        const int m_pair = tile / n_tiles, n_tile = tile - m_pair * n_tiles;
        const int m0 = m_pair * 256 + (int)rank * 128;
        int n0 = 0, y0 = 0;
        if (a.mode_a == 1) {
          n0 = m0 / hw;
          y0 = (m0 - n0 * hw) / a.W;
        }
        int cb = 0, kx = 0, ky = 0;
        const int b_row0 = n_tile * BN + (int)rank * (BN / 2) +
                           (a.batch_rows_a ? (m_pair * 256 / a.batch_rows_a) * a.batch_rows_b : 0);
        for (int kb = 0, ka = 0; kb < a.num_k_blocks; ++kb, ++ka, ++it) {
          if (kb == a.k_wrap) ka = cb = kx = ky = 0;  // split build: second pass over the same activation tile
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t sa = smem_base + s * kStageBytes, sb = sa + kABytes;
          if (lane == 0) {
            if (leader) mbar_arrive_expect_tx(full0 + 8 * s, 2 * kStageBytes);  // both CTAs' bytes land on this barrier
            if (a.mode_a == 1)
              tma2_load_4d(sa, &tmA, full0 + 8 * s, cb * kBK, a.tap_x0 + kx, a.tap_y0 + y0 + ky, n0);
            else
              tma2_load_2d(sa, &tmA, full0 + 8 * s, ka * kBK, m0);
          } else {
            tma2_load_2d(sb, &tmB, full0 + 8 * s, kb * kBK, b_row0);
          }
          if (++cb == a.cin_blocks) {
            cb = 0;
            if (++kx == a.kw) {
              kx = 0;
              ++ky;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
      const uint32_t dhi = umma_desc_hi_sw128(1024), a_lo0 = umma_desc_lo(smem_base, 16);
      uint32_t it = 0, ti = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs, ++ti) {
        const uint32_t acc = ti & 1, aph = (ti >> 1) & 1;
        mbar_wait(tempty0 + 8 * acc, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < a.num_k_blocks; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t alo = a_lo0 + s * (kStageBytes >> 4), blo = alo + (kABytes >> 4);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            tc2_mma_f16(d_tmem, umma_desc_pack(alo + 2 * k, dhi), umma_desc_pack(blo + 2 * k, dhi), idesc, (kb | k) != 0);
          tc2_commit_mc(empty0 + 8 * s, 3);  // frees this stage in BOTH CTAs
        }
        tc2_commit_mc(tfull0 + 8 * acc, 3);
      }
    }
  } else {
    const int q = warp & 3;
    uint32_t ti = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs, ++ti) {
      const int m_pair = tile / n_tiles, n_tile = tile - m_pair * n_tiles;
      const uint32_t acc = ti & 1, aph = (ti >> 1) & 1;
      mbar_wait(tfull0 + 8 * acc, aph);
      tc_fence_after();
      const int row = m_pair * 256 + (int)rank * 128 + q * 32 + lane;
      const bool row_ok = row < a.M;
      const int stat_row = (m_pair * 2 + (int)rank) * 4 + q;
      const size_t out_row = !row_ok ? 0 : (a.out_pad ? pad_row_off(row, a.fH, a.fW, a.ld_out) : (size_t)row * a.ld_out);
      const float *rb_row = (a.rowbias && row_ok) ? a.rowbias + (size_t)(row >> a.rb_shift) * a.rb_ld : nullptr;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tc_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c * 32, r);
        tc_wait_ld();
        const int col0 = n_tile * BN + c * 32;
        if (col0 < a.N) {
          if (a.bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = __ldg(reinterpret_cast<const float4 *>(a.bias + col0) + j);
              r[4 * j + 0] = __float_as_uint(__uint_as_float(r[4 * j + 0]) + b4.x);
              r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + b4.y);
              r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + b4.z);
              r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + b4.w);
            }
          }
          if (rb_row) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = __ldg(reinterpret_cast<const float4 *>(rb_row + col0) + j);
              r[4 * j + 0] = __float_as_uint(__uint_as_float(r[4 * j + 0]) + b4.x);
              r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + b4.y);
              r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + b4.z);
              r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + b4.w);
            }
          }
          if (row_ok && a.addend) epi_add_act32(r, a.addend + out_row + col0);
          if (a.stat_sum) {
            float v[32], w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = row_ok ? __uint_as_float(r[j]) : 0.f;
              v[j] = x;
              w[j] = x * x;
            }
            float s1 = warp_transpose_reduce(v, lane);
            float s2 = warp_transpose_reduce(w, lane);
            if (col0 + lane < a.N) {
              a.stat_sum[(size_t)stat_row * a.N + col0 + lane] = s1;
              a.stat_sq[(size_t)stat_row * a.N + col0 + lane] = s2;
            }
          }
          if (row_ok) {
            if (a.out_bf16) epi_store_act32(a.out_bf16 + out_row + col0, r);
            if (a.out_f32) {
              uint4 *dst = reinterpret_cast<uint4 *>(a.out_f32 + (size_t)row * a.ld_out + col0);
#pragma unroll
              for (int j = 0; j < 8; ++j) dst[j] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      if (lane == 0) mbar_arrive_remote(tempty0 + 8 * acc, 0);  // the leader's MMA thread owns the accumulator hand-back
    }
  }
  tc_fence_before();
  cluster_sync_all();  // no CTA may exit (or free tensor memory) while its peer can still signal it
  if (warp == 1) tmem2_dealloc(tmem_base, kTmemCols);
}

#ifndef SALUN_SPLIT
// =================================================================================================
// k_conv_rw: persistent stride-1 3x3 convolution with the weight slice RESIDENT in shared memory.
//
// The large-image layers (layer1: 64 ch @32x32, layer2: 128 ch @16x16) are bound by L2->smem operand traffic in
// k_conv_gemm (every one of the 9 taps re-fetches a 16 KB activation tile and an 8-16 KB weight tile).  Here
//   * one CTA per SM keeps its [64 out-ch][9*Cin] weight slice in smem for its whole life (72 / 144 KB),
//   * an activation box of (rows+2) image rows is fetched once per (kx, 64-channel slab) and serves the three
//     ky taps: tap ky starts ky*W*128 bytes into the box, a multiple of the 1024-byte swizzle atom, so the UMMA
//     descriptor simply moves its start address,
//   * two TMEM accumulators let the epilogue of tile i overlap the MMAs of tile i+1.
// Operand traffic per 128x64 output tile drops from 9*(16+8) = 216 KB to 3*24 = 72 KB (Cin = 64).
// =================================================================================================
template <int kW, int kCinBlocks>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_conv_rw(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvRwArgs a) {
  constexpr int kRows = 128 / kW;                         // image rows per 128-pixel tile
  constexpr uint32_t kABytes = (kRows + 2) * kW * 128;    // activation box incl. the two halo rows
  constexpr int kStages = kCinBlocks == 1 ? 6 : 3;
  constexpr uint32_t kSlab = 64 * 128;                    // one [64 out-ch][64 in-ch] weight slab
  constexpr uint32_t kWBytes = 9 * kCinBlocks * kSlab;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t w_base = smem_u32(smem);
  const uint32_t a_base = w_base + kWBytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kWBytes + kStages * kABytes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 5);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kStages, wfull = empty0 + 8 * kStages;
  const uint32_t tfull0 = wfull + 8, tempty0 = tfull0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.y;
  constexpr int kSteps = 3 * kCinBlocks;  // (kx, channel slab) load steps per tile

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(wfull, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const bool dbg_on = a.dbg != nullptr;
  const long long t_begin = dbg_on ? clock64() : 0;
  const int cta = blockIdx.y * gridDim.x + blockIdx.x;
  if (warp == 0) {
    if (lane == 0) {
      long long w_empty = 0;
      // resident weights: 9*cin_blocks slabs, slab (tap, cb) = columns [(tap*cin_blocks + cb)*64, +64) of rows n_tile*64..
      mbar_arrive_expect_tx(wfull, kWBytes);
      for (int j = 0; j < 9 * kCinBlocks; ++j) tma_load_2d(w_base + j * kSlab, &tmB, wfull, j * 64, n_tile * 64);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int pix0 = tile * 128;
        const int n0 = pix0 / (a.H * a.W), y0 = (pix0 % (a.H * a.W)) / a.W;
        for (int st = 0; st < kSteps; ++st, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          const int kx = st / kCinBlocks, cb = st - kx * kCinBlocks;
          mbar_wait_t(empty0 + 8 * s, ph ^ 1, w_empty, dbg_on);
          mbar_arrive_expect_tx(full0 + 8 * s, kABytes);
          tma_load_4d(a_base + s * kABytes, &tmA, full0 + 8 * s, cb * 64, kx, y0, n0);
        }
      }
      if (dbg_on) {
        a.dbg[cta * 8 + 0] = w_empty;
        a.dbg[cta * 8 + 1] = clock64() - t_begin;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      long long w_full = 0, w_tempty = 0;
      const uint32_t dhi = umma_desc_hi_sw128(1024), a_lo0 = umma_desc_lo(a_base, 16), w_lo0 = umma_desc_lo(w_base, 16);
      mbar_wait_t(wfull, 0, w_full, dbg_on);
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++ti) {
        const uint32_t acc = ti & 1, aph = (ti >> 1) & 1;
        mbar_wait_t(tempty0 + 8 * acc, aph ^ 1, w_tempty, dbg_on);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 64;
        for (int st = 0; st < kSteps; ++st, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          const int kx = st / kCinBlocks, cb = st - kx * kCinBlocks;
          mbar_wait_t(full0 + 8 * s, ph, w_full, dbg_on);
          tc_fence_after();
          const uint32_t alo = a_lo0 + s * (kABytes >> 4);
          const uint32_t blo = w_lo0 + (kx * kCinBlocks + cb) * (kSlab >> 4);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_f16(d_tmem, umma_desc_pack(alo + ky * ((kW * 128) >> 4) + 2 * k, dhi),
                         umma_desc_pack(blo + ky * (3 * kCinBlocks * (kSlab >> 4)) + 2 * k, dhi), idesc,
                         (st | ky | k) != 0);
          }
          tc_commit(empty0 + 8 * s);
        }
        tc_commit(tfull0 + 8 * acc);
      }
      if (dbg_on) {
        a.dbg[cta * 8 + 2] = w_full;
        a.dbg[cta * 8 + 3] = w_tempty;
        a.dbg[cta * 8 + 4] = clock64() - t_begin;
      }
    }
  } else {
    const int q = warp & 3;
    long long w_tfull = 0;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++ti) {
      const uint32_t acc = ti & 1, aph = (ti >> 1) & 1;
      mbar_wait_t(tfull0 + 8 * acc, aph, w_tfull, dbg_on);
      tc_fence_after();
      const int row = tile * 128 + q * 32 + lane;
      const bool row_ok = row < a.M;
      const int stat_row = tile * 4 + q;
      const size_t fuse_pad = (a.f1.act && row_ok) ? pad_row_off(row, a.H, a.W, a.N) : 0;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tc_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 64 + c * 32, r);
        tc_wait_ld();
        const int col0 = n_tile * 64 + c * 32;
        if (a.stat_sum) {  // BatchNorm batch statistics from the unrounded fp32 accumulators
          float v[32], w[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = row_ok ? __uint_as_float(r[j]) : 0.f;
            v[j] = x;
            w[j] = x * x;
          }
          float s1 = warp_transpose_reduce(v, lane);
          float s2 = warp_transpose_reduce(w, lane);
          a.stat_sum[(size_t)stat_row * a.N + col0 + lane] = s1;
          a.stat_sq[(size_t)stat_row * a.N + col0 + lane] = s2;
        }
        if (row_ok) {
          if (a.addend) {
            const uint4 *src = reinterpret_cast<const uint4 *>(a.addend + (size_t)row * a.ld_out + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 v = src[j];
              const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 t = __bfloat1622float2(h[i]);
                r[8 * j + 2 * i] = __float_as_uint(__uint_as_float(r[8 * j + 2 * i]) + t.x);
                r[8 * j + 2 * i + 1] = __float_as_uint(__uint_as_float(r[8 * j + 2 * i + 1]) + t.y);
              }
            }
          }
          uint4 *dst = reinterpret_cast<uint4 *>(a.out_bf16 + (size_t)row * a.ld_out + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1]));
            v.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
            v.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
            v.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
            dst[j] = v;
          }
        }
        if (a.f1.act) {
          bn_bwd_fuse_chunk(a.f1, r, row_ok, fuse_pad, (size_t)row * a.N, col0, a.N, stat_row, lane);
          if (a.f2.act) bn_bwd_fuse_chunk(a.f2, r, row_ok, fuse_pad, (size_t)row * a.N, col0, a.N, stat_row, lane);
        }
        __syncwarp();
      }
      // this warp has drained its quarter of the accumulator: hand it back to the MMA issuer
      tc_fence_before();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
    }
    if (dbg_on && q == 0 && lane == 0) {
      a.dbg[cta * 8 + 5] = w_tfull;
      a.dbg[cta * 8 + 6] = clock64() - t_begin;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

#endif  // !SALUN_SPLIT

// =================================================================================================
// wgrad: MN-major operands (pixel dimension is K)
// =================================================================================================
constexpr int kWgStages = 4;
constexpr int kWgPix = 64;                         // pixels per k-block
constexpr uint32_t kWgBlockBytes = kWgPix * 128;   // one [64 px][64 ch] bf16 slab = 8 KB

__global__ void __launch_bounds__(kGemmThreads, 1)
k_wgrad(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = (2 + a.n_blocks) * kWgBlockBytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kWgStages * 6 * kWgBlockBytes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kWgStages + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kWgStages, tfull = empty0 + 8 * kWgStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co_tile = blockIdx.x, grp = blockIdx.y, split = blockIdx.z;
  const int kb_begin = split * a.kb_per_split;
  const int kb_end = min(a.kb_total, kb_begin + a.kb_per_split);
  const int nkb = kb_end - kb_begin;
  const int a_blocks = (a.Cout - co_tile * 128) >= 128 ? 2 : 1;  // valid 64-row slabs of dY in this tile
  const uint32_t tmem_cols = a.n_blocks <= 1 ? 64 : (a.n_blocks == 2 ? 128 : 256);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool dbg_on = a.dbg != nullptr;
  const long long t_begin = dbg_on ? clock64() : 0;
  const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  if (nkb > 0) {
    if (warp == 0) {
      const int nbox = a_blocks + a.n_blocks;
      if (lane < nbox) {
        // one lane per TMA box (2 dY slabs + up to 4 X slabs): address math and issue run SIMT-parallel
        long long w_empty = 0;
        const bool is_a = lane < a_blocks;
        const int j = is_a ? lane : lane - a_blocks;
        const uint32_t dst_off = is_a ? j * kWgBlockBytes : (2 + j) * kWgBlockBytes;
        const int hw = a.H * a.W;
        int c0, dx = 1, dy = 1;   // channel offset and padded-coordinate shift of this lane's box
        if (is_a) {
          c0 = co_tile * 128 + j * 64;
        } else {
          const int b = grp * a.n_blocks + j;
          if (a.mode_b == 1) {
            const int tap = b / a.cin_blocks, cb = b - tap * a.cin_blocks;
            const int ky = tap / a.kw, kx = tap - ky * a.kw;
            c0 = cb * 64;
            dx = a.tap_x0 + kx;
            dy = a.tap_y0 + ky;
          } else {
            c0 = b * 64;
          }
        }
        const bool four_d = is_a ? (a.mode_a == 1) : (a.mode_b == 1);
        const CUtensorMap *tm = is_a ? &tmA : &tmB;
        for (int i = 0; i < nkb; ++i) {
          const int s = i % kWgStages;
          const uint32_t ph = (i / kWgStages) & 1;
          mbar_wait_t(empty0 + 8 * s, ph ^ 1, w_empty, dbg_on);
          const uint32_t dst = smem_base + s * stage_bytes + dst_off;
          if (lane == 0) mbar_arrive_expect_tx(full0 + 8 * s, nbox * kWgBlockBytes);
          const int pix0 = (kb_begin + i) * kWgPix;
          if (four_d) {
            const int n0 = pix0 / hw;   // hw is a power of two on every supported shape: compiles to a shift-free
            const int y0 = (pix0 - n0 * hw) / a.W;  // udiv, but only two per k-block per lane, in parallel lanes
            tma_load_4d(dst, tm, full0 + 8 * s, c0, dx, dy + y0, n0);
          } else {
            tma_load_2d(dst, tm, full0 + 8 * s, c0, pix0);
          }
        }
        if (dbg_on && lane == 0) {
          a.dbg[cta * 8 + 0] = w_empty;
          a.dbg[cta * 8 + 1] = clock64() - t_begin;
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, 64 * a.n_blocks, 1, 1);
        // MN-major, 128B swizzle: 64-element (128 B) rows, 8 pixel rows per 1024-byte atom (SBO),
        // next 64-channel slab kWgBlockBytes further (LBO)
        const uint32_t lbo = a.swap_lbo_sbo ? 1024u : kWgBlockBytes;
        const uint32_t sbo = a.swap_lbo_sbo ? kWgBlockBytes : 1024u;
        const uint32_t dhi = umma_desc_hi_sw128(sbo), a_lo0 = umma_desc_lo(smem_base, lbo);
        long long w_full = 0;
        for (int i = 0; i < nkb; ++i) {
          const int s = i % kWgStages;
          const uint32_t ph = (i / kWgStages) & 1;
          mbar_wait_t(full0 + 8 * s, ph, w_full, dbg_on);
          tc_fence_after();
          const uint32_t alo = a_lo0 + s * (stage_bytes >> 4), blo = alo + ((2 * kWgBlockBytes) >> 4);
#pragma unroll
          for (int k = 0; k < kWgPix / 16; ++k)
            tc_mma_f16(tmem_base, umma_desc_pack(alo + k * (2048 >> 4), dhi), umma_desc_pack(blo + k * (2048 >> 4), dhi),
                       idesc, (i | k) != 0);
          tc_commit(empty0 + 8 * s);
        }
        tc_commit(tfull);
        if (dbg_on) {
          a.dbg[cta * 8 + 2] = w_full;
          a.dbg[cta * 8 + 4] = clock64() - t_begin;
        }
      }
    } else {
      const int q = warp & 3;
      long long w_tfull = 0;
      mbar_wait_t(tfull, 0, w_tfull, dbg_on);
      tc_fence_after();
      const int co = co_tile * 128 + q * 32 + lane;
      const bool row_ok = co < a.Cout;
      for (int c = 0; c < a.n_blocks * 2; ++c) {
        uint32_t r[32];
        tc_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, r);
        tc_wait_ld();
        const int col0 = (grp * a.n_blocks) * 64 + c * 32;
        float *dst = a.dw + (size_t)split * a.split_stride + (size_t)co * a.ldw + col0;
        if (!row_ok) {
          // junk rows of a half-empty co tile: nothing to store
        } else if (a.split_stride > 0) {
          // workspace mode: plain stores of this split's partial tile
          if (col0 + 32 <= a.kvalid && (a.ldw & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<uint4 *>(dst)[j] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < a.kvalid) dst[j] = __uint_as_float(r[j]);
          }
        } else if (col0 + 32 <= a.kvalid && (a.ldw & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j),
                         "f"(__uint_as_float(r[4 * j])), "f"(__uint_as_float(r[4 * j + 1])),
                         "f"(__uint_as_float(r[4 * j + 2])), "f"(__uint_as_float(r[4 * j + 3]))
                         : "memory");
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < a.kvalid) atomicAdd(dst + j, __uint_as_float(r[j]));
        }
        __syncwarp();
      }
      if (dbg_on && q == 0 && lane == 0) {
        a.dbg[cta * 8 + 5] = w_tfull;
        a.dbg[cta * 8 + 6] = clock64() - t_begin;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// =================================================================================================
// host side
// =================================================================================================
// cuTensorMapEncodeTiled is resolved through the runtime (no link-time dependency on libcuda.so.1, so the
// library still loads -- and reports a clean error -- on a machine without a driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
#define SALUN_ENCODE_OR_FAIL()                                                   \
  EncodeTiledFn enc = encode_tiled();                                            \
  if (!enc) {                                                                    \
    set_error("cuTensorMapEncodeTiled not available (no CUDA driver?)");         \
    return SALUN_ERR_CUDA;                                                       \
  }
int make_tmap_2d_bf16(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                      uint32_t box_cols) {
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  SALUN_ENCODE_OR_FAIL();
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstr, box,
                                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d rows=%llu cols=%llu box=%ux%u) failed: %d", (unsigned long long)rows,
              (unsigned long long)cols, box_rows, box_cols, (int)r);
    return SALUN_ERR_CUDA;
  }
  return SALUN_OK;
}

int make_tmap_2d_bf16_ld(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                         uint32_t box_cols) {
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  SALUN_ENCODE_OR_FAIL();
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstr, box,
                                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d rows=%llu cols=%llu ld=%llu box=%ux%u) failed: %d", (unsigned long long)rows,
              (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols, (int)r);
    return SALUN_ERR_CUDA;
  }
  return SALUN_OK;
}

int make_tmap_4d_bf16(CUtensorMap *m, const void *base, uint64_t C, uint64_t Wp, uint64_t Hp, uint64_t N,
                      TmapBox4 bx) {
  cuuint64_t gdim[4] = {C, Wp, Hp, N};
  cuuint64_t gstr[3] = {C * 2, Wp * C * 2, Hp * Wp * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)bx.c, (cuuint32_t)bx.w, (cuuint32_t)bx.h, (cuuint32_t)bx.n};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  SALUN_ENCODE_OR_FAIL();
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), gdim, gstr, box,
                                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d C=%llu Wp=%llu Hp=%llu N=%llu box=%d,%d,%d,%d) failed: %d",
              (unsigned long long)C, (unsigned long long)Wp, (unsigned long long)Hp, (unsigned long long)N, bx.c, bx.w,
              bx.h, bx.n, (int)r);
    return SALUN_ERR_CUDA;
  }
  return SALUN_OK;
}

int conv_box(int H, int W, int pixels, TmapBox4 *box) {
  // `pixels` consecutive output pixels in (n, y, x) order must form a (w=W, h, n) box
  if (W <= 0 || H <= 0 || pixels % W != 0) {
    set_error("conv_box: W=%d does not divide the %d-pixel tile", W, pixels);
    return SALUN_ERR_UNSUPPORTED;
  }
  int rows = pixels / W;
  int h = rows < H ? rows : H;
  if (rows % h != 0 || H % h != 0) {
    set_error("conv_box: H=%d W=%d cannot tile %d pixels", H, W, pixels);
    return SALUN_ERR_UNSUPPORTED;
  }
  box->c = 64;
  box->w = W;
  box->h = h;
  box->n = rows / h;
  return SALUN_OK;
}

// optional per-launch timing of the two tensor-core kernels (bench.py roofline): CUDA events on the launch stream
struct ProfRec {
  cudaEvent_t a, b;
  int cat;
  double flops;
};
static long long *g_dbg = nullptr;  // role-timing buffer attached to every GEMM launch while set (salun_debug_role_timing)
static bool g_prof = false;
static std::vector<ProfRec> g_recs;
static void prof_open(int cat, double flops, cudaStream_t st) {
  if (!g_prof) return;
  ProfRec r;
  r.cat = cat;
  r.flops = flops;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
}
static void prof_close(cudaStream_t st) {
  if (!g_prof) return;
  cudaEventRecord(g_recs.back().b, st);
}

#ifndef SALUN_SPLIT
template <int BN, int S>
static int launch_conv_gemm_t(const CUtensorMap &tmA, const CUtensorMap &tmB, const ConvGemmArgs &a, cudaStream_t st) {
  constexpr size_t smem = (size_t)S * (kBM * kBK * 2 + BN * kBK * 2) + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    SALUN_CUDA_OK(cudaFuncSetAttribute(k_conv_gemm<BN, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((a.M + kBM - 1) / kBM, (a.N + BN - 1) / BN);
  { k_conv_gemm<BN, S><<<grid, kGemmThreads, smem, st>>>(tmA, tmB, a); ++::salun::g_launch_count; }
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

#endif

template <int BN, int S>
static int launch_conv_gemm_p_t(const CUtensorMap &tmA, const CUtensorMap &tmB, const ConvGemmArgs &a, cudaStream_t st) {
  constexpr size_t smem = (size_t)S * (kBM * kBK * 2 + BN * kBK * 2) + 1024 + 256 + 8 * 4096;  // + store-transpose buffers
  static bool attr_set = false;
  static int num_sms = 0;
  if (!attr_set) {
    SALUN_CUDA_OK(cudaFuncSetAttribute(k_conv_gemm_p<BN, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0;
    SALUN_CUDA_OK(cudaGetDevice(&dev));
    SALUN_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  const int m_tiles = (a.M + kBM - 1) / kBM, n_tiles = (a.N + BN - 1) / BN;
  int grid = m_tiles * n_tiles * (a.ksplit > 1 ? a.ksplit : 1);
  if (grid > num_sms) grid = num_sms;
  ConvGemmArgs aa = a;
  aa.dbg = g_dbg;
  { SALUN_CUDA_OK(::salun::launch_pdl(k_conv_gemm_p<BN, S>, dim3(grid), dim3(kGemmPThreads), smem, st, tmA, tmB, aa, m_tiles, n_tiles)); ++::salun::g_launch_count; }
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

template <int BN, int S>
static int launch_gemm2_t(const CUtensorMap &tmA, const CUtensorMap &tmB, const ConvGemmArgs &a, cudaStream_t st) {
  constexpr size_t smem = (size_t)S * (kBM * kBK * 2 + (BN / 2) * kBK * 2) + 1024 + 256;
  static bool attr_set = false;
  static int num_sms = 0;
  if (!attr_set) {
    SALUN_CUDA_OK(cudaFuncSetAttribute(k_gemm2<BN, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0;
    SALUN_CUDA_OK(cudaGetDevice(&dev));
    SALUN_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  const int m_pairs = (a.M + 255) / 256, n_tiles = (a.N + BN - 1) / BN;
  int pairs = m_pairs * n_tiles;
  if (pairs > num_sms / 2) pairs = num_sms / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ConvGemmArgs aa = a;
  aa.dbg = nullptr;
  SALUN_CUDA_OK(cudaLaunchKernelEx(&cfg, k_gemm2<BN, S>, tmA, tmB, aa, m_pairs, n_tiles));
  ++::salun::g_launch_count;
  return SALUN_OK;
}

// The runtimes describe a GEMM in activation / weight ELEMENTS (K / 64 k-blocks, Cin / 64 blocks per tap).  In the split
// build every activation element is two bf16 operand elements and the weight operand holds [dup(hi) | dup(lo)]: twice
// the k-blocks per pass, two passes, the A operand rewinding at the start of the second (salun_act.cuh).
static ConvGemmArgs operand_units(const ConvGemmArgs &a) {
  ConvGemmArgs aa = a;
  aa.k_wrap = 0;
  if (kSplit) {
    aa.k_wrap = a.num_k_blocks * 2;
    aa.num_k_blocks = a.num_k_blocks * 4;
    aa.cin_blocks = a.cin_blocks * 2;
  }
  return aa;
}

// CTA-pair GEMM: B tensor map must have box rows BN/2.  bn in {128, 256}.
static int launch_gemm2_units(const CUtensorMap &tmA, const CUtensorMap &tmB, const ConvGemmArgs &a, double flops, int bn,
                              cudaStream_t st) {
  prof_open(0, flops, st);
  int rc;
  switch (bn) {
    case 128: rc = launch_gemm2_t<128, 8>(tmA, tmB, a, st); break;
    case 256: rc = launch_gemm2_t<256, 6>(tmA, tmB, a, st); break;
    default: set_error("launch_gemm2: unsupported BN=%d", bn); rc = SALUN_ERR_INVALID;
  }
  prof_close(st);
  return rc;
}
int launch_gemm2(const CUtensorMap &tmA, const CUtensorMap &tmB, const ConvGemmArgs &a, int bn, cudaStream_t st) {
  return launch_gemm2_units(tmA, tmB, operand_units(a), 2.0 * a.M * a.N * (double)a.num_k_blocks * 64.0, bn, st);
}

static bool gemm_persistent() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SALUN_GEMM_PERSIST");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1 || kSplit;
}

static bool gemm_log() {  // SALUN_GEMM_LOG=1: one stderr line per tensor-core launch (shape), to pair with an ncu launch list
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SALUN_GEMM_LOG");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// Split-K epilogue: D[row][col] = sum_z part[z][row][col] (+ bias + rowbias + addend) -> activation-typed and / or fp32 output,
// the same terms in the same order as the in-kernel epilogue of k_conv_gemm_p.  One thread per 8 columns of one row.
__global__ void __launch_bounds__(256) k_splitk_epilogue(const float *__restrict__ part, int ksplit, long long split_stride, int ldp,
                                                         const ConvGemmArgs a) {
  const int vecs = (a.N + 7) >> 3;
  const long long total = (long long)a.M * vecs;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int row = (int)(i / vecs), col0 = (int)(i - (long long)row * vecs) * 8;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float *p = part + (size_t)row * ldp + col0;
    for (int z = 0; z < ksplit; ++z, p += split_stride) {
      const float4 u = *reinterpret_cast<const float4 *>(p), v = *reinterpret_cast<const float4 *>(p + 4);
      f[0] += u.x; f[1] += u.y; f[2] += u.z; f[3] += u.w;
      f[4] += v.x; f[5] += v.y; f[6] += v.z; f[7] += v.w;
    }
    if (a.bias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += __ldg(a.bias + col0 + j);
    }
    if (a.rowbias) {
      const float *rb = a.rowbias + (size_t)(row >> a.rb_shift) * a.rb_ld + col0;
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += __ldg(rb + j);
    }
    const size_t out_row = a.out_pad ? pad_row_off(row, a.fH, a.fW, a.ld_out) : (size_t)row * a.ld_out;
    if (a.addend) {
      float g[8];
      ld8(a.addend + out_row + col0, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += g[j];
    }
    if (a.out_bf16) st8(a.out_bf16 + out_row + col0, f);
    if (a.out_f32) {
      float *o = a.out_f32 + (size_t)row * a.ld_out + col0;
      *reinterpret_cast<float4 *>(o) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4 *>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
  }
}

// k-slices per output tile for a launch of `tiles` tiles and `nkb` k-blocks (operand units): fill ~one wave of 148 CTAs,
// keep >= 4 k-blocks per slice, stay inside the scratch.  1 = do not split.
static int pick_ksplit(const ConvGemmArgs &a, int bn, int nkb) {
  if (!a.splitk_ws || a.stat_sum || a.f1.act || a.f2.act || a.pair || (a.N & 7) || (a.ld_out & 7)) return 1;
  const int m_tiles = (a.M + kBM - 1) / kBM, n_tiles = (a.N + bn - 1) / bn;
  const long long tiles = (long long)m_tiles * n_tiles;
  if (tiles > 74 || nkb < 8) return 1;
  long long ks = 148 / tiles;
  if (ks > nkb / 4) ks = nkb / 4;
  if (ks > 32) ks = 32;
  const long long slab = (long long)m_tiles * kBM * n_tiles * bn;
  if (ks * slab > a.splitk_ws_floats) ks = a.splitk_ws_floats / slab;
  return ks < 2 ? 1 : (int)ks;
}

int launch_conv_gemm(const CUtensorMap &tmA, const CUtensorMap &tmB, const ConvGemmArgs &a0, int bn, cudaStream_t st) {
  if (gemm_log())
    fprintf(stderr, "GEMMLOG M=%d N=%d K=%d mode=%d bn=%d H=%d batched=%d\n", a0.M, a0.N, a0.num_k_blocks * 64, a0.mode_a, bn, a0.H,
            a0.batch_rows_a);
  const double flops = 2.0 * a0.M * a0.N * (double)a0.num_k_blocks * 64.0;  // algorithmic (one product per element pair)
  const ConvGemmArgs a = operand_units(a0);
  if (a.pair) return launch_gemm2_units(tmA, tmB, a, flops, bn, st);  // the caller encoded tmB with bn/2 box rows
  prof_open(0, flops, st);
  int rc;
  if (gemm_persistent()) {
    const int ks = pick_ksplit(a, bn, a.num_k_blocks);
    ConvGemmArgs g = a;
    if (ks > 1) {  // partial tiles to the scratch, everything else of the epilogue in k_splitk_epilogue
      const int m_tiles = (a.M + kBM - 1) / kBM, n_tiles = (a.N + bn - 1) / bn;
      g.ksplit = ks;
      g.split_stride = (long long)m_tiles * kBM * n_tiles * bn;
      g.out_f32 = a.splitk_ws;
      g.ld_out = n_tiles * bn;
      g.out_bf16 = nullptr;
      g.addend = nullptr;
      g.bias = g.rowbias = nullptr;
      g.out_pad = 0;
    }
    switch (bn) {
      case 64: rc = launch_conv_gemm_p_t<64, 8>(tmA, tmB, g, st); break;
      case 128: rc = launch_conv_gemm_p_t<128, 6>(tmA, tmB, g, st); break;
      case 160: rc = launch_conv_gemm_p_t<160, 5>(tmA, tmB, g, st); break;  // N = 320 (SD level 0): 2 column tiles, not 5 x 64
      case 256: rc = launch_conv_gemm_p_t<256, 4>(tmA, tmB, g, st); break;
      default: set_error("launch_conv_gemm: unsupported BN=%d", bn); rc = SALUN_ERR_INVALID;
    }
    if (ks > 1 && rc == SALUN_OK) {
      const long long total = (long long)a.M * ((a.N + 7) >> 3);
      long long blocks = (total + 255) / 256;
      if (blocks > 148 * 8) blocks = 148 * 8;
      k_splitk_epilogue<<<(unsigned)blocks, 256, 0, st>>>(a.splitk_ws, ks, g.split_stride, g.ld_out, a);
      ++::salun::g_launch_count;
      if (cudaGetLastError() != cudaSuccess) rc = SALUN_ERR_CUDA;
    }
    prof_close(st);
    return rc;
  }
#ifndef SALUN_SPLIT
  if (a.bias || a.rowbias || a.out_pad || a.batch_rows_a) {
    set_error("launch_conv_gemm: bias / rowbias / out_pad / batched operands need the persistent kernel (SALUN_GEMM_PERSIST=1)");
    prof_close(st);
    return SALUN_ERR_UNSUPPORTED;
  }
  switch (bn) {
    case 64: rc = launch_conv_gemm_t<64, 4>(tmA, tmB, a, st); break;
    case 128: rc = launch_conv_gemm_t<128, 3>(tmA, tmB, a, st); break;
    case 256: rc = launch_conv_gemm_t<256, 4>(tmA, tmB, a, st); break;
    default: set_error("launch_conv_gemm: unsupported BN=%d", bn); rc = SALUN_ERR_INVALID;
  }
#else
  rc = SALUN_ERR_UNSUPPORTED;
#endif
  prof_close(st);
  return rc;
}

#ifndef SALUN_SPLIT
template <int kW, int kCB>
static int launch_conv_rw_t(const CUtensorMap &tmA, const CUtensorMap &tmB, const ConvRwArgs &a, int num_sms,
                            cudaStream_t st) {
  constexpr int kRows = 128 / kW;
  constexpr int kStages = kCB == 1 ? 6 : 3;
  constexpr size_t smem = (size_t)9 * kCB * 8192 + (size_t)kStages * (kRows + 2) * kW * 128 + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    SALUN_CUDA_OK(cudaFuncSetAttribute(k_conv_rw<kW, kCB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int n_tiles = a.N / 64;
  int gx = num_sms / n_tiles;
  if (gx > a.num_tiles) gx = a.num_tiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, n_tiles);
  prof_open(0, 2.0 * a.M * a.N * 9.0 * kCB * 64.0, st);
  ConvRwArgs aa = a;
  aa.dbg = g_dbg;
  { SALUN_CUDA_OK(::salun::launch_pdl(k_conv_rw<kW, kCB>, dim3(grid), dim3(kGemmThreads), smem, st, tmA, tmB, aa)); ++::salun::g_launch_count; }
  prof_close(st);
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

bool conv_rw_supported(int W, int cin, int cout) {
  return (W == 32 || W == 16) && (cin == 64 || cin == 128) && cout % 64 == 0;
}

int launch_conv_rw(const CUtensorMap &tmA, const CUtensorMap &tmB, const ConvRwArgs &a, int num_sms, cudaStream_t st) {
  if (a.W == 32 && a.cin_blocks == 1) return launch_conv_rw_t<32, 1>(tmA, tmB, a, num_sms, st);
  if (a.W == 32 && a.cin_blocks == 2) return launch_conv_rw_t<32, 2>(tmA, tmB, a, num_sms, st);
  if (a.W == 16 && a.cin_blocks == 1) return launch_conv_rw_t<16, 1>(tmA, tmB, a, num_sms, st);
  if (a.W == 16 && a.cin_blocks == 2) return launch_conv_rw_t<16, 2>(tmA, tmB, a, num_sms, st);
  set_error("launch_conv_rw: unsupported W=%d cin_blocks=%d", a.W, a.cin_blocks);
  return SALUN_ERR_UNSUPPORTED;
}
#else
bool conv_rw_supported(int, int, int) { return false; }  // resident-weight kernel: bf16 build only
int launch_conv_rw(const CUtensorMap &, const CUtensorMap &, const ConvRwArgs &, int, cudaStream_t) {
  set_error("launch_conv_rw: not available in the split-precision build");
  return SALUN_ERR_UNSUPPORTED;
}
#endif

#ifndef SALUN_SPLIT
__global__ void __launch_bounds__(256) k_wgrad_reduce(const WgReduceEntry *__restrict__ tab, float *__restrict__ grads) {
  const WgReduceEntry e = tab[blockIdx.y];
  float *__restrict__ dst = grads + e.dst_off;
  const long long n4 = e.count >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = reinterpret_cast<const float4 *>(e.ws)[i];
    for (int s = 1; s < e.splits; ++s) {
      const float4 v = reinterpret_cast<const float4 *>(e.ws + (size_t)s * e.count)[i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4 *>(dst)[i] = acc;
  }
  if (blockIdx.x == 0 && threadIdx.x < (e.count & 3)) {  // tail (the 64x27 stem weight is not a multiple of 4... it is; kept general)
    const long long i = (n4 << 2) + threadIdx.x;
    float acc = 0.f;
    for (int s = 0; s < e.splits; ++s) acc += e.ws[(size_t)s * e.count + i];
    dst[i] = acc;
  }
}
#else
// split build: slab s holds D'[2 cout][2 kc], the four partial products (dY_hi | dY_lo) x (X_hi | X_lo) of every weight:
// dW[co][j] = sum_s  D'[2co][2j] + D'[2co][2j+1] + D'[2co+1][2j] + D'[2co+1][2j+1]   (fixed order: deterministic)
__global__ void __launch_bounds__(256) k_wgrad_reduce(const WgReduceEntry *__restrict__ tab, float *__restrict__ grads) {
  const WgReduceEntry e = tab[blockIdx.y];
  float *__restrict__ dst = grads + e.dst_off;
  const long long slab = 4 * e.count;
  const int kc = e.kc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e.count; i += (long long)gridDim.x * blockDim.x) {
    const long long co = i / kc;
    const int j = (int)(i - co * kc);
    const float *__restrict__ p0 = e.ws + (size_t)(2 * co) * (2 * kc) + 2 * j;
    float acc = 0.f;
    for (int s = 0; s < e.splits; ++s) {
      const float2 r0 = *reinterpret_cast<const float2 *>(p0 + (size_t)s * slab);
      const float2 r1 = *reinterpret_cast<const float2 *>(p0 + (size_t)s * slab + 2 * kc);
      acc += (r0.x + r0.y) + (r1.x + r1.y);
    }
    dst[i] = acc;
  }
}
#endif
void launch_wgrad_reduce(const WgReduceEntry *table_dev, int n_entries, float *grads, cudaStream_t st) {
  { k_wgrad_reduce<<<dim3(kSplit ? 256 : 64, n_entries), 256, 0, st>>>(table_dev, grads); ++::salun::g_launch_count; }
}

int wgrad_pick_blocks(int total_blocks) {
  if (total_blocks % 4 == 0) return 4;
  if (total_blocks % 3 == 0) return 3;
  if (total_blocks % 2 == 0) return 2;
  return 1;
}

WgradGeom wgrad_geometry(int cout, int kcp) {
  WgradGeom g;
  g.total_blocks = kcp * kActK / 64;
  g.n_blocks = wgrad_pick_blocks(g.total_blocks);
  g.co_tiles = (cout * kActK + 127) / 128;
  g.groups = g.total_blocks / g.n_blocks;
  return g;
}

// `a0` in weight / activation ELEMENTS (Cout, ldw, kvalid, cin_blocks = Cin / 64, split_stride = cout * kc); total_blocks,
// n_blocks and the grid (co_tiles, col_groups) from wgrad_geometry().  The split build widens rows and columns by two.
int launch_wgrad(const CUtensorMap &tmA, const CUtensorMap &tmB, const WgradArgs &a0, int co_tiles, int col_groups,
                 int splits, cudaStream_t st) {
  constexpr size_t smem = (size_t)kWgStages * 6 * kWgBlockBytes + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    SALUN_CUDA_OK(cudaFuncSetAttribute(k_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  WgradArgs aa = a0;
  if (kSplit) {
    aa.Cout = a0.Cout * 2;
    aa.ldw = a0.ldw * 2;
    aa.kvalid = a0.kvalid * 2;
    aa.cin_blocks = a0.cin_blocks * 2;
    aa.split_stride = a0.split_stride * 4;
  }
  dim3 grid(co_tiles, col_groups, splits);
  if (gemm_log())
    fprintf(stderr, "WGLOG pixels=%d Cout=%d Kc=%d grid=%d,%d,%d\n", a0.kb_total * 64, a0.Cout, a0.kvalid, co_tiles, col_groups, splits);
  prof_open(1, 2.0 * (double)a0.kb_total * 64.0 * a0.Cout * (double)a0.kvalid, st);
  aa.dbg = g_dbg;
  { k_wgrad<<<grid, kGemmThreads, smem, st>>>(tmA, tmB, aa); ++::salun::g_launch_count; }
  prof_close(st);
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

}  // namespace salun

// =================================================================================================
// C ABI: stand-alone entry points (used by the parity tests and by integrators who want one conv)
// =================================================================================================
using namespace salun;

extern "C" {

// per-launch CUDA-event timing of the tensor-core kernels: category 0 = k_conv_gemm (forward + dgrad),
// 1 = k_wgrad.  begin() arms it, end() synchronises the device and returns summed milliseconds, launch counts and
// algorithmic FLOPs per category.
// bring-up aid: while a buffer is attached, every GEMM launch writes per-CTA role timings into it
// ([cta][8] cycles: 0 producer wait-empty, 1 producer total, 2 MMA wait-full, 3 MMA wait-tmem-empty, 4 MMA total,
//  5 epilogue wait-tmem-full, 6 epilogue total).  Pass NULL to detach.
int salun_debug_role_timing(long long *buf_dev) {
  g_dbg = buf_dev;
  return SALUN_OK;
}
int salun_profile_begin(void) {
  g_recs.clear();
  g_prof = true;
  return SALUN_OK;
}
int salun_profile_end(double *ms_by_cat, int64_t *launches_by_cat, double *flops_by_cat) {
  g_prof = false;
  SALUN_CUDA_OK(cudaDeviceSynchronize());
  for (int c = 0; c < 2; ++c) {
    if (ms_by_cat) ms_by_cat[c] = 0.0;
    if (launches_by_cat) launches_by_cat[c] = 0;
    if (flops_by_cat) flops_by_cat[c] = 0.0;
  }
  for (ProfRec &r : g_recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    if (ms_by_cat) ms_by_cat[r.cat] += ms;
    if (launches_by_cat) launches_by_cat[r.cat] += 1;
    if (flops_by_cat) flops_by_cat[r.cat] += r.flops;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_recs.clear();
  return SALUN_OK;
}

// D[M][N] (fp32 and/or bf16) = A[M][K] . B[N][K]^T, bf16 row-major operands, K % 64 == 0, N % 64 == 0
int salun_gemm_bf16_tn(salun_ctx *ctx, const void *A, const void *B, float *out_f32, void *out_bf16, int64_t M,
                       int64_t N, int64_t K, void *stream) {
#ifdef SALUN_SPLIT
  set_error("salun_gemm_bf16_tn: raw-bf16 entry point, served by libsalun.so (this is the split-precision build)");
  return SALUN_ERR_UNSUPPORTED;
#else
  SALUN_REQUIRE(ctx && A && B && (out_f32 || out_bf16), "NULL argument");
  SALUN_REQUIRE(M > 0 && N > 0 && K > 0 && K % 64 == 0 && N % 64 == 0, "need K % 64 == 0 and N % 64 == 0");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  const int bn = (N % 256 == 0 && M * N >= 148ll * 128 * 256) ? 256 : (N % 128 == 0 ? 128 : 64);
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmA, A, M, K, kBM, kBK))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmB, B, N, K, bn, kBK))) return rc;
  ConvGemmArgs a{};
  a.mode_a = 0;
  a.num_k_blocks = (int)(K / 64);
  a.M = (int)M;
  a.N = (int)N;
  a.out_bf16 = (act_t *)out_bf16;
  a.out_f32 = out_f32;
  a.ld_out = (int)N;
  return launch_conv_gemm(tmA, tmB, a, bn, (cudaStream_t)stream);
#endif
}

// CTA-pair (cta_group::2) variant of salun_gemm_bf16_tn: N % 128 == 0.
int salun_gemm2_bf16_tn(salun_ctx *ctx, const void *A, const void *B, float *out_f32, void *out_bf16, int64_t M,
                        int64_t N, int64_t K, void *stream) {
#ifdef SALUN_SPLIT
  set_error("salun_gemm2_bf16_tn: raw-bf16 entry point, served by libsalun.so (this is the split-precision build)");
  return SALUN_ERR_UNSUPPORTED;
#else
  SALUN_REQUIRE(ctx && A && B && (out_f32 || out_bf16), "NULL argument");
  SALUN_REQUIRE(M > 0 && N > 0 && K > 0 && K % 64 == 0 && N % 128 == 0, "need K % 64 == 0 and N % 128 == 0");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  const int bn = N % 256 == 0 ? 256 : 128;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmA, A, M, K, kBM, kBK))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmB, B, N, K, bn / 2, kBK))) return rc;
  ConvGemmArgs a{};
  a.mode_a = 0;
  a.num_k_blocks = (int)(K / 64);
  a.cin_blocks = 1 << 30;
  a.M = (int)M;
  a.N = (int)N;
  a.out_bf16 = (act_t *)out_bf16;
  a.out_f32 = out_f32;
  a.ld_out = (int)N;
  return launch_gemm2(tmA, tmB, a, bn, (cudaStream_t)stream);
#endif
}

// Y[batch*H*W][Cout] = conv(X, Wk) for a stride-1 kh x kw convolution (3x3/pad 1 or 1x1/pad 0).
//   xpad : bf16 [batch][H+2][W+2][Cin]  halo-padded NHWC, halo = 0        (Cin % 64 == 0)
//   wk   : bf16 [Cout][kh*kw*Cin]       (tap-major, then input channel)   (Cout % 64 == 0)
//   y    : bf16 [batch*H*W][Cout]  ;  stat_sum / stat_sq: optional fp32 [(batch*H*W/128)*4][Cout] partials
// replaces F.conv2d forward (cuDNN) behind Classification/models/ResNet.py:111-119.
int salun_conv_fwd_bf16(salun_ctx *ctx, const void *xpad, const void *wk, void *y_bf16, float *y_f32, float *stat_sum,
                        float *stat_sq, int batch, int H, int W, int Cin, int Cout, int ksize, void *stream) {
#ifdef SALUN_SPLIT
  set_error("salun_conv_fwd_bf16: raw-bf16 entry point, served by libsalun.so (this is the split-precision build)");
  return SALUN_ERR_UNSUPPORTED;
#else
  SALUN_REQUIRE(ctx && xpad && wk && (y_bf16 || y_f32), "NULL argument");
  SALUN_REQUIRE(ksize == 3 || ksize == 1, "ksize must be 1 or 3");
  SALUN_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "Cin and Cout must be multiples of 64");
  SALUN_REQUIRE(((int64_t)batch * H * W) % 128 == 0, "batch*H*W must be a multiple of 128");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  const int bn = Cout % 128 == 0 ? 128 : 64;
  TmapBox4 bx;
  int rc;
  if ((rc = conv_box(H, W, kBM, &bx))) return rc;
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap_4d_bf16(&tmA, xpad, Cin, W + 2, H + 2, batch, bx))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmB, wk, Cout, (uint64_t)ksize * ksize * Cin, bn, kBK))) return rc;
  ConvGemmArgs a{};
  a.mode_a = 1;
  a.cin_blocks = Cin / 64;
  a.num_k_blocks = ksize * ksize * a.cin_blocks;
  a.kw = ksize;
  a.tap_y0 = a.tap_x0 = ksize == 3 ? 0 : 1;
  a.H = H;
  a.W = W;
  a.M = batch * H * W;
  a.N = Cout;
  a.out_bf16 = (act_t *)y_bf16;
  a.out_f32 = y_f32;
  a.ld_out = Cout;
  a.stat_sum = stat_sum;
  a.stat_sq = stat_sq;
  return launch_conv_gemm(tmA, tmB, a, bn, (cudaStream_t)stream);
#endif
}

// Same contract as salun_conv_fwd_bf16 (3x3 only) through the persistent resident-weight kernel k_conv_rw;
// W (= H) in {16, 32}, Cin in {64, 128}.
int salun_conv_rw_fwd_bf16(salun_ctx *ctx, const void *xpad, const void *wk, void *y_bf16, float *stat_sum,
                           float *stat_sq, int batch, int H, int W, int Cin, int Cout, void *stream) {
#ifdef SALUN_SPLIT
  set_error("salun_conv_rw_fwd_bf16: raw-bf16 entry point, served by libsalun.so (this is the split-precision build)");
  return SALUN_ERR_UNSUPPORTED;
#else
  SALUN_REQUIRE(ctx && xpad && wk && y_bf16, "NULL argument");
  SALUN_REQUIRE(H == W && conv_rw_supported(W, Cin, Cout), "shape not served by k_conv_rw");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  CUtensorMap tmA, tmB;
  int rc;
  TmapBox4 bx{64, W, 128 / W + 2, 1};
  if ((rc = make_tmap_4d_bf16(&tmA, xpad, Cin, W + 2, H + 2, batch, bx))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmB, wk, Cout, (uint64_t)9 * Cin, 64, 64))) return rc;
  ConvRwArgs r{};
  r.H = H;
  r.W = W;
  r.cin_blocks = Cin / 64;
  r.M = batch * H * W;
  r.num_tiles = (r.M + 127) / 128;
  r.N = Cout;
  r.out_bf16 = (act_t *)y_bf16;
  r.ld_out = Cout;
  r.stat_sum = stat_sum;
  r.stat_sq = stat_sq;
  return launch_conv_rw(tmA, tmB, r, ctx->num_sms, (cudaStream_t)stream);
#endif
}

// dW[Cout][kh*kw*Cin] (fp32, tap-major) += sum over pixels dY[p][Cout]^T . X_tap[p][Cin]
//   dy : bf16 [batch*H*W][Cout] ; xpad : bf16 halo-padded NHWC ; dw must be zeroed by the caller.
// replaces the cuDNN wgrad inside loss.backward() (Classification/unlearn/RL.py:132).
int salun_conv_wgrad_bf16(salun_ctx *ctx, const void *dy, const void *xpad, float *dw, int batch, int H, int W, int Cin,
                          int Cout, int ksize, int splits, int swap_lbo_sbo, void *stream) {
#ifdef SALUN_SPLIT
  set_error("salun_conv_wgrad_bf16: raw-bf16 entry point, served by libsalun.so (this is the split-precision build)");
  return SALUN_ERR_UNSUPPORTED;
#else
  SALUN_REQUIRE(ctx && dy && xpad && dw, "NULL argument");
  SALUN_REQUIRE(ksize == 3 || ksize == 1, "ksize must be 1 or 3");
  SALUN_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "Cin and Cout must be multiples of 64");
  const int64_t M = (int64_t)batch * H * W;
  SALUN_REQUIRE(M % 64 == 0, "batch*H*W must be a multiple of 64");
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  TmapBox4 bx;
  int rc;
  if ((rc = conv_box(H, W, kWgPix, &bx))) return rc;
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap_2d_bf16(&tmA, dy, M, Cout, kWgPix, 64))) return rc;
  if ((rc = make_tmap_4d_bf16(&tmB, xpad, Cin, W + 2, H + 2, batch, bx))) return rc;
  WgradArgs a{};
  a.mode_b = 1;
  a.kb_total = (int)(M / kWgPix);
  a.cin_blocks = Cin / 64;
  a.kw = ksize;
  a.tap_y0 = a.tap_x0 = ksize == 3 ? 0 : 1;
  a.H = H;
  a.W = W;
  a.total_blocks = ksize * ksize * a.cin_blocks;
  a.n_blocks = wgrad_pick_blocks(a.total_blocks);
  a.Cout = Cout;
  a.ldw = ksize * ksize * Cin;
  a.kvalid = a.ldw;
  a.dw = dw;
  a.swap_lbo_sbo = swap_lbo_sbo;
  const int co_tiles = (Cout + 127) / 128, groups = a.total_blocks / a.n_blocks;
  if (splits <= 0) {
    splits = (2 * ctx->num_sms + co_tiles * groups - 1) / (co_tiles * groups);
    if (splits < 1) splits = 1;
  }
  if (splits > a.kb_total) splits = a.kb_total;
  a.kb_per_split = (a.kb_total + splits - 1) / splits;
  splits = (a.kb_total + a.kb_per_split - 1) / a.kb_per_split;
  return launch_wgrad(tmA, tmB, a, co_tiles, groups, splits, (cudaStream_t)stream);
#endif
}

}  // extern "C"
