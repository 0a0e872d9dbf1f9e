// salun_attn.cu -- fused (flash-style) multi-head attention forward on tcgen05 for head widths <= 64:
//   out[b, t, h*d : (h+1)*d] = softmax(q_h k_h^T / sqrt(d)) v_h        (CrossAttention.forward, SD/ldm/modules/attention.py:168-192)
// One CTA owns 128 query rows of one (sample, head) unit and walks the keys in blocks of 128 operand columns:
//   S = Q K^T          tcgen05.mma into tensor memory (128 lanes = query rows, one fp32 column per key)
//   softmax            one thread per query row reads its S row from tensor memory, keeps the running max / sum, writes
//                      P = exp2(S*c - max) as the A operand of the second product straight into swizzled shared memory
//   O_blk = P V        tcgen05.mma into a second tensor-memory region; the row's thread folds it into its fp32 registers
//                      (O = O * exp2(old max - new max) + O_blk) -- neither S nor P ever exists in HBM.
// The unfused path (salun_ops.cu) writes S (fp32) and P for 16 x 4096 x 4096 scores per layer: ~3 GB of HBM traffic and
// 0.93 ms per full-resolution self-attention of SD v1.4 at batch 2 (profiles/r2_sd_launches_bf16.csv).
//
// Operands are the packed per-head buffers of salun_sd_attention (k_heads_pack_*): Qh as activations, Kh / Vt in weight-
// operand layout, so the same kernel serves both builds.  In the split build an activation element is a (hi, lo) bf16 pair
// (2 operand columns) and a weight-operand row is [dup(hi) | dup(lo)] (4 operand columns per element): a key block is 64
// keys, each product runs over both halves of the B operand with the A operand rewound -- the four partial products of
// (hi + lo) x (hi + lo), as in k_conv_gemm_p.
//
// Roles: warps 0-3 = softmax (thread t <-> query row t <-> TMEM lane t), warp 4 lane 0 = TMA producer and MMA issuer.
// K / V tiles are double buffered; S(j+1) is issued right behind PV(j) so it runs under the O update of block j.
#include <math.h>

#include "salun_gemm.cuh"
#include "salun_sm100.cuh"

namespace salun {
using namespace sm100;

namespace {
constexpr int kFaThreads = 160;
constexpr int kFaDp = 64;  // padded head width (operand columns of Q per bf16 unit)

template <int U, int W>
struct FaCfg {
  static constexpr int BKV = 128 / U;                    // keys per block (128 operand columns of P)
  static constexpr int QCB = kFaDp * U / 64;             // 64-column blocks of the Q tile          1 | 2
  static constexpr int KCB = kFaDp * W / 64;             // ... of a K tile row                     1 | 4
  static constexpr int PCB = 2;                          // ... of the P tile (128 operand columns)
  static constexpr int VCB = BKV * W / 64;               // ... of a V^T tile row                   2 | 4
  static constexpr uint32_t kQBlock = 128 * 128;         // [128 rows][128 B]
  static constexpr uint32_t kKBlock = BKV * 128;
  static constexpr uint32_t kVBlock = kFaDp * 128;
  static constexpr uint32_t kQBytes = QCB * kQBlock;     // 16 | 32 KB
  static constexpr uint32_t kKBytes = KCB * kKBlock;     // 16 | 32 KB
  static constexpr uint32_t kVBytes = VCB * kVBlock;     // 16 | 32 KB
  static constexpr uint32_t kPBytes = PCB * kQBlock;     // 32 KB
  static constexpr uint32_t kStageBytes = kKBytes + kVBytes;
  static constexpr uint32_t kDataBytes = kQBytes + 2 * kStageBytes + kPBytes;  // 112 | 192 KB
  static constexpr uint32_t kSmemBytes = kDataBytes + 128;                     // + barriers, tensor-memory slot
  static constexpr uint32_t kTmemCols = (BKV + kFaDp) <= 128 ? 128 : 256;      // S | O_blk
};

struct FaArgs {
  act_t *out;   // [n*Tq][C] merged heads
  int Tq, Tqp, Tk, Tkp, heads, d, C;
  float c2;     // log2(e) / sqrt(d)
};

// tcgen05.wait::ld that also names the destination registers of the load it completes: the compiler must treat them as
// written HERE, so it cannot read or move them between the (asynchronous) load and the wait when other work is interleaved
__device__ __forceinline__ void tc_wait_ld_regs(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int U, int W>
__global__ void __launch_bounds__(kFaThreads, U == 1 ? 2 : 1)
k_flash_attn(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
             const __grid_constant__ CUtensorMap tmV, const FaArgs a) {
  using Cfg = FaCfg<U, W>;
  constexpr int BKV = Cfg::BKV;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  if (smem_base & 1023u) __trap();  // the swizzled tiles need 1 KB alignment; no slack is reserved (2 CTAs per SM)
  const uint32_t sQ = smem_base, sKV = sQ + Cfg::kQBytes, sP = sKV + 2 * Cfg::kStageBytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kDataBytes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 12);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_full = bar0, kv_full0 = bar0 + 8, kv_empty0 = bar0 + 24, s_full = bar0 + 40, p_full = bar0 + 48,
                 o_full = bar0 + 56;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, g = blockIdx.y;
  const int nblk = a.Tkp / BKV;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_full0 + 8 * s, 1);
      mbar_init(kv_empty0 + 8 * s, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + BKV;

  if (warp == 4) {
    if (lane == 0) {
      const uint32_t dhi = umma_desc_hi_sw128(1024);
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, BKV, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kFaDp, 0, 0);
      auto load_kv = [&](int j, int s) {
        const uint32_t sK = sKV + s * Cfg::kStageBytes, sV = sK + Cfg::kKBytes, bar = kv_full0 + 8 * s;
        mbar_arrive_expect_tx(bar, Cfg::kStageBytes);
#pragma unroll
        for (int cb = 0; cb < Cfg::KCB; ++cb) tma_load_2d(sK + cb * Cfg::kKBlock, &tmK, bar, cb * 64, g * a.Tkp + j * BKV);
#pragma unroll
        for (int cb = 0; cb < Cfg::VCB; ++cb)
          tma_load_2d(sV + cb * Cfg::kVBlock, &tmV, bar, (cb / Cfg::PCB) * (a.Tkp * U) + j * 128 + 64 * (cb % Cfg::PCB), g * kFaDp);
      };
      auto mma_s = [&](int s) {  // S = Q K_j^T
        const uint32_t sK = sKV + s * Cfg::kStageBytes;
#pragma unroll
        for (int cb = 0; cb < Cfg::KCB; ++cb) {
          const uint32_t alo = umma_desc_lo(sQ + (cb % Cfg::QCB) * Cfg::kQBlock, 16), blo = umma_desc_lo(sK + cb * Cfg::kKBlock, 16);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tmem_S, umma_desc_pack(alo + 2 * k, dhi), umma_desc_pack(blo + 2 * k, dhi), idesc_s, (cb | k) != 0);
        }
        tc_commit(s_full);
      };
      mbar_arrive_expect_tx(q_full, Cfg::kQBytes);
#pragma unroll
      for (int cb = 0; cb < Cfg::QCB; ++cb) tma_load_2d(sQ + cb * Cfg::kQBlock, &tmQ, q_full, cb * 64, g * a.Tqp + m_tile * 128);
      load_kv(0, 0);
      if (nblk > 1) load_kv(1, 1);
      mbar_wait(q_full, 0);
      mbar_wait(kv_full0, 0);
      tc_fence_after();
      mma_s(0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j & 1;
        const uint32_t sV = sKV + s * Cfg::kStageBytes + Cfg::kKBytes;
        mbar_wait(p_full, j & 1);  // P(j) is in shared memory; S and O_blk of block j-1 have been read
        tc_fence_after();
#pragma unroll
        for (int cb = 0; cb < Cfg::VCB; ++cb) {  // O_blk = P V_j
          const uint32_t alo = umma_desc_lo(sP + (cb % Cfg::PCB) * Cfg::kQBlock, 16), blo = umma_desc_lo(sV + cb * Cfg::kVBlock, 16);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tmem_O, umma_desc_pack(alo + 2 * k, dhi), umma_desc_pack(blo + 2 * k, dhi), idesc_o, (cb | k) != 0);
        }
        tc_commit(o_full);
        tc_commit(kv_empty0 + 8 * s);
        if (j + 1 < nblk) {
          mbar_wait(kv_full0 + 8 * (s ^ 1), ((j + 1) >> 1) & 1);
          tc_fence_after();
          mma_s(s ^ 1);
        }
        if (j + 2 < nblk) {
          mbar_wait(kv_empty0 + 8 * s, (j >> 1) & 1);  // PV(j) retired: stage s is free
          load_kv(j + 2, s);
        }
      }
    }
  } else {
    const int row = threadIdx.x;  // query row of the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    float O[kFaDp];
#pragma unroll
    for (int i = 0; i < kFaDp; ++i) O[i] = 0.f;
    uint8_t *p_row = smem + (sP - smem_base) + row * 128;
    constexpr int NC = BKV / 32;  // 32-column chunks of an S row: 4 | 2 (even: the two register buffers alternate statically)
    uint32_t r[2][32];
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int key0 = j * BKV;
      // Tensor-memory loads are software pipelined (chunk c+1 is in flight while chunk c is consumed) and the max / sum
      // run as four independent chains: with two softmax warps per scheduler nothing else hides those latencies.
      // ---- pass 1: row maximum
      float mx4[4] = {m_run, -INFINITY, -INFINITY, -INFINITY};
      tc_ld_32x32b_x32(tmem_S + lane_off, r[0]);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        tc_wait_ld_regs(r[c & 1]);
        tc_ld_32x32b_x32(tmem_S + lane_off + ((c + 1) % NC) * 32, r[(c + 1) & 1]);  // next chunk, or chunk 0 again for pass 2
        const uint32_t(&x)[32] = r[c & 1];
        if (key0 + c * 32 + 32 <= a.Tk) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(x[i]) * a.c2);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (key0 + c * 32 + i < a.Tk) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(x[i]) * a.c2);
        }
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float alpha = ex2(m_run - mx);  // first block: exp2(-inf) = 0
      // ---- pass 2: P = exp2(S c - max), row sum, A operand of P V
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        tc_wait_ld_regs(r[c & 1]);
        if (c + 1 < NC) tc_ld_32x32b_x32(tmem_S + lane_off + (c + 1) * 32, r[(c + 1) & 1]);
        const uint32_t(&x)[32] = r[c & 1];
        const bool full = key0 + c * 32 + 32 <= a.Tk;
        // eight scores at a time: exp2, row sum, pack, one 16-byte store of the A operand of P V (row `row`, operand columns
        // [c*32*U, +32*U) -> 16-byte chunks of the 128B-swizzled [128][128 B] blocks); keeps the live registers low
#pragma unroll
        for (int q8 = 0; q8 < 4; ++q8) {
          float p[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v = ex2(fmaf(__uint_as_float(x[q8 * 8 + i]), a.c2, -mx));
            if (!full && key0 + c * 32 + q8 * 8 + i >= a.Tk) v = 0.f;
            p[i] = v;
            rs4[i & 3] += v;
          }
          if (U == 1) {
            uint8_t *blk = p_row + (c >> 1) * Cfg::kQBlock;
            const int chunk = (c & 1) * 4 + q8;
            *reinterpret_cast<uint4 *>(blk + ((chunk ^ (row & 7)) << 4)) =
                make_uint4(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]), pack_bf16x2(p[4], p[5]), pack_bf16x2(p[6], p[7]));
          } else {
            uint8_t *blk = p_row + c * Cfg::kQBlock;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int chunk = q8 * 2 + hh;
              *reinterpret_cast<uint4 *>(blk + ((chunk ^ (row & 7)) << 4)) =
                  make_uint4(act_pack_pair(p[4 * hh]), act_pack_pair(p[4 * hh + 1]), act_pack_pair(p[4 * hh + 2]),
                             act_pack_pair(p[4 * hh + 3]));
            }
          }
        }
      }
      l_run = fmaf(l_run, alpha, (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      m_run = mx;
      tc_fence_before();
      fence_proxy_async();  // generic-proxy stores of P -> visible to the tensor core's async-proxy reads
      mbar_arrive(p_full);
      mbar_wait(o_full, j & 1);
      tc_fence_after();
      tc_ld_32x32b_x32(tmem_O + lane_off, r[0]);
      tc_wait_ld_regs(r[0]);
      tc_ld_32x32b_x32(tmem_O + lane_off + 32, r[1]);
#pragma unroll
      for (int i = 0; i < 32; ++i) O[i] = fmaf(O[i], alpha, __uint_as_float(r[0][i]));
      tc_wait_ld_regs(r[1]);
#pragma unroll
      for (int i = 0; i < 32; ++i) O[32 + i] = fmaf(O[32 + i], alpha, __uint_as_float(r[1][i]));
    }
    const int tq = m_tile * 128 + row;
    if (tq < a.Tq) {
      const float inv = 1.f / l_run;
      const int b = g / a.heads, h = g - b * a.heads;
      act_t *dst = a.out + ((size_t)b * a.Tq + tq) * a.C + h * a.d;
#pragma unroll
      for (int v = 0; v < kFaDp / 8; ++v)
        if (v * 8 < a.d) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = O[v * 8 + i] * inv;
          st8(dst + v * 8, f);
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}
}  // namespace

bool flash_attn_supported(int d) { return d > 0 && d <= kFaDp && d % 8 == 0; }

// Qh: act [G*Tqp][64]; Kh: wop rows [G*Tkp], logical length 64; Vt: wop rows [G*64], logical length Tkp (k_heads_pack_*)
int launch_flash_attn(const act_t *Qh, const wop_t *Kh, const wop_t *Vt, act_t *out, int n, int Tq, int Tqp, int Tk, int Tkp, int heads,
                      int d, cudaStream_t st) {
  using Cfg = FaCfg<kActK, kWopK>;
  const long long G = (long long)n * heads;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap_2d_act(&tmQ, Qh, (uint64_t)G * Tqp, kFaDp, 128))) return rc;
  if ((rc = make_tmap_2d_wop(&tmK, Kh, (uint64_t)G * Tkp, kFaDp, Cfg::BKV))) return rc;
  if ((rc = make_tmap_2d_wop(&tmV, Vt, (uint64_t)G * kFaDp, Tkp, kFaDp))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    SALUN_CUDA_OK(cudaFuncSetAttribute(k_flash_attn<kActK, kWopK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes));
    attr_set = true;
  }
  FaArgs a;
  a.out = out;
  a.Tq = Tq;
  a.Tqp = Tqp;
  a.Tk = Tk;
  a.Tkp = Tkp;
  a.heads = heads;
  a.d = d;
  a.C = heads * d;
  a.c2 = 1.4426950408889634f / sqrtf((float)d);
  k_flash_attn<kActK, kWopK><<<dim3(Tqp / 128, (unsigned)G), kFaThreads, Cfg::kSmemBytes, st>>>(tmQ, tmK, tmV, a);
  ++g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

}  // namespace salun
