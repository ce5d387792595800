// Error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

int sdrm_fail(int code, const char* msg);  // records the thread-local message, returns code

#define SDRM_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      char _m[256];                                                                             \
      snprintf(_m, sizeof _m, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
               __LINE__);                                                                       \
      return sdrm_fail(-3, _m);                                                                 \
    }                                                                                           \
  } while (0)
