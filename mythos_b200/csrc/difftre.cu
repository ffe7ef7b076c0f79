// difftre.cu -- Boltzmann reweighting of stored frames (DiffTRe): weights, effective sample size.
//
// compute_weights_and_neff, mythos/optimization/objective.py:139-163:
//   w_k = exp(-beta_k (E_k - Eref_k)) / sum_k exp(...),   n_eff = exp(-sum_k w_k ln w_k) / F
// evaluated with the usual max-shift (mathematically identical, finite for large |beta dE|).
// The per-frame energies and dE/dparams rows come from mythos_b200_energy_* with n_frames = F; this kernel is the
// (F,) -> (F,) tail of the pass.  F is at most a few 10^4, so one block walks the vector three times.
#include "common.cuh"

namespace mb {

constexpr int kWBlock = 1024;

template <class T>
__device__ __forceinline__ T block_reduce(T v, T* sh, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    const T y = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? (y > v ? y : v) : v + y;
  }
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  T r = sh[0];
  for (int w = 1; w < kWBlock / 32; ++w) r = is_max ? (sh[w] > r ? sh[w] : r) : r + sh[w];
  return r;
}

template <class T>
__global__ void k_weights(int F, const T* beta, const T* e_new, const T* e_ref, T* w, T* sums) {
  __shared__ T sh[kWBlock / 32];
  T m = T(-1e300 > -3e38 ? -3e38 : -1e300);
  for (int k = threadIdx.x; k < F; k += kWBlock) {
    const T x = -beta[k] * (e_new[k] - e_ref[k]);
    m = x > m ? x : m;
  }
  m = block_reduce(m, sh, true);
  T z = 0;
  for (int k = threadIdx.x; k < F; k += kWBlock) z += exp(-beta[k] * (e_new[k] - e_ref[k]) - m);
  z = block_reduce(z, sh, false);
  T h = 0;
  for (int k = threadIdx.x; k < F; k += kWBlock) {
    const T lw = -beta[k] * (e_new[k] - e_ref[k]) - m - log(z);
    const T wk = exp(lw);
    w[k] = wk;
    h += wk * lw;
  }
  h = block_reduce(h, sh, false);
  if (threadIdx.x == 0 && sums) {
    sums[0] = m;
    sums[1] = z;
    sums[2] = h;
    sums[3] = exp(-h) / T(F);
  }
}

template <class T>
static int weights_impl(cudaStream_t s, const mb_weights_args* x) {
  MB_REQUIRE(x && x->n_frames > 0, MB_EINVAL_SHAPE, "weights_neff: bad n_frames");
  MB_REQUIRE(x->beta && x->e_new && x->e_ref && x->weights, MB_EINVAL_SHAPE, "weights_neff: missing buffers");
  k_weights<T><<<1, kWBlock, 0, s>>>(x->n_frames, static_cast<const T*>(x->beta), static_cast<const T*>(x->e_new),
                                    static_cast<const T*>(x->e_ref), static_cast<T*>(x->weights),
                                    static_cast<T*>(x->sums));
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}

}  // namespace mb

extern "C" int mythos_b200_weights_neff_f64(void* stream, const mb_weights_args* a) {
  return mb::weights_impl<double>(static_cast<cudaStream_t>(stream), a);
}
extern "C" int mythos_b200_weights_neff_f32(void* stream, const mb_weights_args* a) {
  return mb::weights_impl<float>(static_cast<cudaStream_t>(stream), a);
}
