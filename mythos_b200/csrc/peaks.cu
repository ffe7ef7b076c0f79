// peaks.cu -- issue-rate micro-benchmarks used as roofline denominators (measured on the box, not datasheet):
// dependent-free FMA chains in fp64 / fp32, 8 independent accumulators per thread, every SM saturated.
#include "common.cuh"

namespace mb {

template <class T>
__global__ void k_fma_peak(T* out, int iters, T a, T b) {
  T x0 = T(threadIdx.x), x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
    x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace mb

// Enqueues `blocks` x 256 threads each doing iters*8 FMAs; the caller times it with CUDA events.
// flops = blocks * 256 * iters * 8 * 2.  `scratch` must hold blocks*256 reals.
extern "C" int mythos_b200_fma_peak_f64(void* stream, void* scratch, int blocks, int iters) {
  mb::k_fma_peak<double><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<double*>(scratch), iters, 0.999999, 1e-7);
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}
extern "C" int mythos_b200_fma_peak_f32(void* stream, void* scratch, int blocks, int iters) {
  mb::k_fma_peak<float><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float*>(scratch), iters, 0.999999f, 1e-7f);
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Issue cost of the special functions the energy kernels call, measured the same way as the FMA peak: every thread runs 8
// independent dependent-chains x <- f(x) * a + b (the FMA keeps the argument in the function's ordinary domain and is
// subtracted by the caller: cost(f) = t(f chain) / t(FMA chain) - 1 FMA slots).  These are the weights SURVEY 8(d) asks
// for ("measured from a micro-kernel and committed next to the peaks"); bench.py measures them in every run.
namespace mb {

enum SpecialKind { kDiv = 0, kSqrt = 1, kExp = 2, kLog = 3, kAcos = 4, kRsqrt = 5, kRcp = 6, kCos = 7, kFmod = 8, kNKinds = 9 };

template <class T, int KIND>
__device__ __forceinline__ T special_op(T x, T c) {
  if (KIND == kDiv) return c / x;
  if (KIND == kSqrt) return sqrt(x);
  if (KIND == kExp) return exp(x);
  if (KIND == kLog) return log(x);
  if (KIND == kAcos) return acos(x);
  if (KIND == kRsqrt) return T(1) / sqrt(x);
  if (KIND == kRcp) return T(1) / x;
  if (KIND == kCos) return cos(x);
  return fmod(x * T(37.0), c);
}

template <class T, int KIND>
__global__ void k_special_rate(T* out, int iters, T a, T b, T c) {
  T x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = b + T(0.01) * T((threadIdx.x + k) & 7);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = special_op<T, KIND>(x[k], c) * a + b;
  }
  T s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class T>
static int launch_special(cudaStream_t s, T* out, int blocks, int iters, int kind) {
  // (a, b) chosen so that the iterates stay inside the domain the energy kernels use: acos on (-1,1), exp on O(1)
  // arguments, log / sqrt / div on O(1) positive numbers
  switch (kind) {
    case kDiv: k_special_rate<T, kDiv><<<blocks, 256, 0, s>>>(out, iters, T(0.5), T(0.7), T(1.3)); break;
    case kSqrt: k_special_rate<T, kSqrt><<<blocks, 256, 0, s>>>(out, iters, T(0.5), T(0.7), T(0)); break;
    case kExp: k_special_rate<T, kExp><<<blocks, 256, 0, s>>>(out, iters, T(-1.1), T(0.3), T(0)); break;
    case kLog: k_special_rate<T, kLog><<<blocks, 256, 0, s>>>(out, iters, T(0.1), T(1.5), T(0)); break;
    case kAcos: k_special_rate<T, kAcos><<<blocks, 256, 0, s>>>(out, iters, T(0.3), T(-0.4), T(0)); break;
    case kRsqrt: k_special_rate<T, kRsqrt><<<blocks, 256, 0, s>>>(out, iters, T(0.5), T(0.7), T(0)); break;
    case kRcp: k_special_rate<T, kRcp><<<blocks, 256, 0, s>>>(out, iters, T(0.5), T(0.7), T(0)); break;
    case kCos: k_special_rate<T, kCos><<<blocks, 256, 0, s>>>(out, iters, T(1.7), T(0.4), T(0)); break;
    case kFmod: k_special_rate<T, kFmod><<<blocks, 256, 0, s>>>(out, iters, T(0.01), T(0.7), T(20.0)); break;
    default: set_error("special_rate: unknown kind %d", kind); return MB_EINVAL_SHAPE;
  }
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}

}  // namespace mb

// blocks x 256 threads, each iters * 8 evaluations of (special(x) * a + b); ops = blocks*256*iters*8.  kind: 0 div,
// 1 sqrt, 2 exp, 3 log, 4 acos, 5 1/sqrt, 6 1/x, 7 cos, 8 fmod.  scratch: blocks*256 reals.
extern "C" int mythos_b200_special_rate_f64(void* stream, void* scratch, int blocks, int iters, int kind) {
  return mb::launch_special<double>(static_cast<cudaStream_t>(stream), static_cast<double*>(scratch), blocks, iters, kind);
}
extern "C" int mythos_b200_special_rate_f32(void* stream, void* scratch, int blocks, int iters, int kind) {
  return mb::launch_special<float>(static_cast<cudaStream_t>(stream), static_cast<float*>(scratch), blocks, iters, kind);
}
