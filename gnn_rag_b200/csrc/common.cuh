// Shared helpers for libgnnrag_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gnnrag_b200.h"

namespace gr {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_error(const char* fmt, ...);

#define GR_CHECK_ARG(cond, msg)                                                     \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      gr::set_error("%s: invalid argument: %s", __func__, msg);                     \
      return GR_ERR_INVALID_ARG;                                                    \
    }                                                                               \
  } while (0)

#define GR_CHECK_CUDA(expr)                                                         \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      gr::set_error("%s: CUDA error %s at %s:%d", __func__, cudaGetErrorString(_e), \
                    __FILE__, __LINE__);                                            \
      return GR_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

#define GR_CHECK_LAUNCH() GR_CHECK_CUDA(cudaGetLastError())

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// number of SMs of the current device (148 on B200); cached
int sm_count();

// true the first time it is called for (flag array, current device): kernel attributes such as the dynamic
// shared-memory opt-in are per device, a process may drive several
inline bool first_use_on_device(bool (&done)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

}  // namespace gr
