// energy_dev.cuh -- device-side argument block and the shared-memory parameter-gradient accumulator shared by
// the pair kernels (energy_kernels.cu) and the frame-resident kernel (frame_kernels.cu).
#pragma once
#include "common.cuh"
#include "oxdna_device.cuh"

namespace mb {

constexpr unsigned kFull = 0xffffffffu;

template <class T>
struct EnergyDev {
  ModelT<T> M;
  int n, n_frames, n_bonded;
  const T* center;
  const T* quat;
  const int32_t* seq;
  const int32_t* nt_type;
  const int32_t* nt_type_stack;
  const int32_t* is_end;
  const int32_t* bonded;
  const int32_t* pairs;
  long long pair_capacity, pair_frame_stride;
  const int32_t* pair_count;  // (F) valid entries at the head of each list, or nullptr
  T all_pairs_cutoff;         // > 0: ignore `pairs`; every non-bonded i<j closer than this (frame-resident kernel only)
  const T* params;
  const T* cot;
  unsigned mask;
  T* terms;
  T* d_center;
  T* d_quat;
  T* d_params;
  long long d_params_frame_stride;
};

// Parameter-gradient accumulator: warp-reduce, then one shared-memory atomic per warp and parameter.
// With several banks (NA1) lanes of one warp may address different banks, so those go out as per-lane
// shared atomics instead.
template <class T>
struct SmemAcc {
  T* sh;
  bool per_lane;
  __device__ __forceinline__ void add(int bank, int idx, T v) {
    if (per_lane) {
      if (v != T(0)) atomicAdd(&sh[bank * MB_P_COUNT + idx], v);
      return;
    }
    if (!__any_sync(kFull, v != T(0))) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sh[idx], v);
  }
  __device__ __forceinline__ void add_scatter(int bank, int idx, T v, bool pred) {
    if (pred && v != T(0)) atomicAdd(&sh[bank * MB_P_COUNT + idx], v);
  }
};


template <class T>
bool frame_kernel_eligible(const EnergyDev<T>& a);
template <class T>
int launch_frame_kernel(cudaStream_t s, const EnergyDev<T>& a, bool want_params);

}  // namespace mb
