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
