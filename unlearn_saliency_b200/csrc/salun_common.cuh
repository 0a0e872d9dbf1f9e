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

extern long long g_launch_count;  // kernels launched by this library (bench.py reports it as gpu_launches)

constexpr int kRadixBins = 2048;      // 11-bit digits: 3 passes over a 32-bit key
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
};
