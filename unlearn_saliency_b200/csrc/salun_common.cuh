// salun_common.cuh -- shared host-side plumbing for libsalun.so (error reporting, context).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/salun.h"

namespace salun {

// thread-local error string behind salun_last_error()
char *err_buf();
void set_error(const char *fmt, ...);

#define SALUN_CUDA_OK(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::salun::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                  \
                         cudaGetErrorString(_e));                                            \
      return SALUN_ERR_CUDA;                                                                 \
    }                                                                                        \
  } while (0)

#define SALUN_REQUIRE(cond, msg)                                                             \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      ::salun::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg);            \
      return SALUN_ERR_INVALID;                                                              \
    }                                                                                        \
  } while (0)

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------
// A step is ~170 dependent launches of 3-35 us kernels; the drain -> launch -> prologue bubble between two of them is
// 1.5-2 us.  Kernels of the step path start with pdl_trigger() (the NEXT kernel of the stream may begin launching: its
// CTAs become resident as this kernel's CTAs retire and run their global-memory-free prologue) and call pdl_wait()
// before their first global-memory access (returns once the PREVIOUS kernel has completed and its writes are visible).
// launch_pdl() must only be used for kernels that call pdl_wait(); both instructions are no-ops under a plain launch.
// SALUN_PDL=0 turns the launch attribute off.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();
template <typename... Exp, typename... Act>
inline cudaError_t launch_pdl(void (*kernel)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Act &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Exp>(args)...);
}

extern long long g_launch_count;  // kernels launched by this library (bench.py reports it as gpu_launches)

constexpr int kRadixBins = 2048;      // 11-bit digits: 3 passes over a 32-bit key
constexpr int kMaxRatios = 16;        // ratios one salun_topk_mask_multi call can serve (the reference sweeps ten)
constexpr int kMaxPartials = 148 * 8; // one partial per CTA of the persistent reduction grids

}  // namespace salun

struct salun_ctx {
  int device;
  int num_sms;
  // radix-select workspace (device)
  unsigned int *hist;          // [kRadixBins]
  unsigned long long *sel;     // [8]: prefix, prefix_mask, remaining, thr, n_gt, n_eq, need, ordered_ties
  unsigned int *block_ties;    // [kMaxPartials + 1] ties per contiguous chunk, then exclusive scan
  double *partials;            // [kMaxPartials] reduction partials
  // pinned host mailbox for info read-back
  unsigned long long *mailbox_host;  // [8]
  // multi-ratio select workspace (salun_topk_mask_multi), allocated on first use
  unsigned long long *msel;          // [kMaxRatios][8]
  unsigned int *mhist;               // [kMaxRatios][kRadixBins]
  unsigned int *mties;               // [kMaxRatios][kMaxPartials + 1]
  unsigned long long *mmailbox_host; // [kMaxRatios][8] pinned
  // caller-owned scratch of the op-level entry points (salun_op_set_scratch): split-K partial tiles of small-M GEMMs
  float *op_scratch;
  long long op_scratch_floats;
};
