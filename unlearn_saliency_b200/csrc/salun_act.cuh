// salun_act.cuh -- the activation storage type of the engines and its 8-element vector accessors.
//
// Two builds of the library (csrc/Makefile):
//   default       act_t = bf16.  Activations, raw conv outputs and activation gradients carry 8 significand bits.
//   -DSALUN_SPLIT act_t = (hi, lo) bf16 PAIR, value = float(hi) + float(lo) with hi = bf16_rn(v), lo = bf16_rn(v - hi):
//                 16+ significand bits in 4 bytes.  In memory an activation row [C] is the bf16 row
//                 [h0 l0 h1 l1 ... ] of length 2C, so the tensor-core kernels read it unchanged as a K-major (or MN-major)
//                 bf16 operand of twice the channel count.  With the weight operand prepared as
//                 [w0h w0h w1h w1h ... | w0l w0l w1l w1l ...] (salun_elem.cu: k_prep_w_all) two passes over the SAME
//                 activation tile accumulate (h + l) * wh + (h + l) * wl = all four partial products of the split
//                 operands in the fp32 TMEM accumulator: fp32-class convolution results out of kind::f16 MMAs (what
//                 cuBLAS calls bf16x emulation), used for the precision-critical saliency pass whose top-k index set
//                 has to reproduce the fp32 reference (Classification/generate_mask.py:30-82).
// Every elementwise kernel is written against ld8 / st8 / ldraw / cvt8 below and compiles for either type.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace salun {

typedef __nv_bfloat16 wop_t;  // tensor-core WEIGHT operand element (always bf16; see kWopK)

#ifdef SALUN_SPLIT
struct __align__(4) act_t {
  __nv_bfloat16 hi, lo;
};
constexpr int kActK = 2;   // bf16 GEMM-K (or MN) elements per activation element
constexpr int kWopK = 4;   // bf16 elements a prepared weight operand spends per weight ([dup(hi) | dup(lo)])
struct avec {              // 8 activation elements
  uint4 a, b;
};
#else
typedef __nv_bfloat16 act_t;
constexpr int kActK = 1;
constexpr int kWopK = 1;
struct avec {
  uint4 a;
};
#endif
constexpr bool kSplit = kActK == 2;

#ifdef __CUDACC__
// one fp32 -> (hi, lo) pair packed as hi | lo << 16 (hi at the lower address)
__device__ __forceinline__ uint32_t act_pack_pair(float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  return (uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(l) << 16);
}
__device__ __forceinline__ float act_unpack_pair(uint32_t w) {
  return __uint_as_float(w << 16) + __uint_as_float(w & 0xffff0000u);
}
__device__ __forceinline__ float act_to_float(const act_t &v) {
#ifdef SALUN_SPLIT
  return __bfloat162float(v.hi) + __bfloat162float(v.lo);
#else
  return __bfloat162float(v);
#endif
}
__device__ __forceinline__ act_t act_from_float(float v) {
#ifdef SALUN_SPLIT
  act_t r;
  r.hi = __float2bfloat16_rn(v);
  r.lo = __float2bfloat16_rn(v - __bfloat162float(r.hi));
  return r;
#else
  return __float2bfloat16(v);
#endif
}

__device__ __forceinline__ void cvt8(const avec &v, float (&f)[8]) {
#ifdef SALUN_SPLIT
  const uint32_t w[8] = {v.a.x, v.a.y, v.a.z, v.a.w, v.b.x, v.b.y, v.b.z, v.b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = act_unpack_pair(w[i]);
#else
  const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v.a);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
#endif
}
// raw 8-element load through the read-only path / plain load / store (p must be 8-element aligned)
__device__ __forceinline__ avec ldraw(const act_t *p) {
  avec v;
  v.a = __ldg(reinterpret_cast<const uint4 *>(p));
#ifdef SALUN_SPLIT
  v.b = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
#endif
  return v;
}
__device__ __forceinline__ avec ldvec(const act_t *p) {
  avec v;
  v.a = *reinterpret_cast<const uint4 *>(p);
#ifdef SALUN_SPLIT
  v.b = *(reinterpret_cast<const uint4 *>(p) + 1);
#endif
  return v;
}
__device__ __forceinline__ void stvec(act_t *p, const avec &v) {
  *reinterpret_cast<uint4 *>(p) = v.a;
#ifdef SALUN_SPLIT
  *(reinterpret_cast<uint4 *>(p) + 1) = v.b;
#endif
}
__device__ __forceinline__ avec avec_zero() {
  avec v;
  v.a = make_uint4(0, 0, 0, 0);
#ifdef SALUN_SPLIT
  v.b = make_uint4(0, 0, 0, 0);
#endif
  return v;
}
__device__ __forceinline__ void ld8(const act_t *p, float (&f)[8]) { cvt8(ldvec(p), f); }
__device__ __forceinline__ avec pack8(const float (&f)[8]) {
  avec v;
#ifdef SALUN_SPLIT
  v.a = make_uint4(act_pack_pair(f[0]), act_pack_pair(f[1]), act_pack_pair(f[2]), act_pack_pair(f[3]));
  v.b = make_uint4(act_pack_pair(f[4]), act_pack_pair(f[5]), act_pack_pair(f[6]), act_pack_pair(f[7]));
#else
  __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&v.a);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
#endif
  return v;
}
__device__ __forceinline__ void st8(act_t *p, const float (&f)[8]) { stvec(p, pack8(f)); }

// one weight value at logical column j of a prepared weight-operand row of logical length Kp:
//   bf16 build : row[j] = bf16(w)
//   split build: row[2j] = row[2j+1] = hi(w) ; row[2Kp + 2j] = row[2Kp + 2j + 1] = lo(w)      ([dup(hi) | dup(lo)])
__device__ __forceinline__ void wop_store(wop_t *__restrict__ row, int j, int Kp, float w) {
#ifdef SALUN_SPLIT
  const __nv_bfloat16 h = __float2bfloat16_rn(w);
  const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
  *reinterpret_cast<__nv_bfloat162 *>(row + 2 * (size_t)j) = __nv_bfloat162(h, h);
  *reinterpret_cast<__nv_bfloat162 *>(row + 2 * (size_t)Kp + 2 * (size_t)j) = __nv_bfloat162(l, l);
#else
  (void)Kp;
  row[j] = __float2bfloat16(w);
#endif
}
// eight consecutive operand elements j0 .. j0+7 (j0 % 8 == 0, row 16-byte aligned): 16-byte stores
__device__ __forceinline__ void wop_store8(wop_t *__restrict__ row, int j0, int Kp, const float (&f)[8]) {
#ifdef SALUN_SPLIT
  uint32_t hh[8], ll[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __nv_bfloat16 h = __float2bfloat16_rn(f[i]);
    const __nv_bfloat16 l = __float2bfloat16_rn(f[i] - __bfloat162float(h));
    const uint32_t hb = __bfloat16_as_ushort(h), lb = __bfloat16_as_ushort(l);
    hh[i] = hb | (hb << 16);
    ll[i] = lb | (lb << 16);
  }
  uint4 *ph = reinterpret_cast<uint4 *>(row + 2 * (size_t)j0), *pl = reinterpret_cast<uint4 *>(row + 2 * (size_t)Kp + 2 * (size_t)j0);
  ph[0] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
  ph[1] = make_uint4(hh[4], hh[5], hh[6], hh[7]);
  pl[0] = make_uint4(ll[0], ll[1], ll[2], ll[3]);
  pl[1] = make_uint4(ll[4], ll[5], ll[6], ll[7]);
#else
  (void)Kp;
  __nv_bfloat162 h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4 *>(row + j0) = *reinterpret_cast<const uint4 *>(h);
#endif
}
#endif  // __CUDACC__

}  // namespace salun
