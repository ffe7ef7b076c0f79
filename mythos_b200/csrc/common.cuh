// common.cuh -- error plumbing and small launch helpers shared by the kernels' translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

#include "../../include/mythos_b200.h"

namespace mb {

void set_error(const char* fmt, ...);

#define MB_CUDA_CHECK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      mb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MB_ECUDA;                                                                   \
    }                                                                                    \
  } while (0)

#define MB_REQUIRE(cond, code, msg)   \
  do {                                \
    if (!(cond)) {                    \
      mb::set_error("%s", msg);       \
      return code;                    \
    }                                 \
  } while (0)

inline int ceil_div(long long a, long long b) { return int((a + b - 1) / b); }

}  // namespace mb
