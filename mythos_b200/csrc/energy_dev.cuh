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
  T* rec;    // list kernel workspace: (F*N) NucRec rows
  T* gback;  // list kernel workspace: (F*N,3) backbone-site gradients
  unsigned long long* sr_list;  // list kernel workspace: (F, sr_capacity) short-range pairs (i | j << 32)
  int* sr_count;                // (F)
  long long sr_capacity;
  int tagged;  // `pairs` carries support tags (MB_NL_TAG_SUPPORTS)
  const int32_t* pair_split;  // tagged lists: (F) entries before it are short-range pairs, after it Debye pairs (list kernels)
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


// Parameter-gradient accumulator held in registers: a run of COUNT parameters starting at BASE (single bank)
template <class T, int BASE, int COUNT>
struct RegAcc {
  T r[COUNT];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int k = 0; k < COUNT; ++k) r[k] = T(0);
  }
  __device__ __forceinline__ void add(int, int idx, T v) { r[idx - BASE] += v; }
  __device__ __forceinline__ void add_scatter(int, int, T, bool) {}
};

template <class T>
__device__ __forceinline__ Nuc<T> load_nuc(const T* __restrict__ center, const T* __restrict__ quat, long long idx,
                                           T q[4]) {
  Nuc<T> n;
  n.c = v3<T>(center[3 * idx], center[3 * idx + 1], center[3 * idx + 2]);
  q[0] = quat[4 * idx];
  q[1] = quat[4 * idx + 1];
  q[2] = quat[4 * idx + 2];
  q[3] = quat[4 * idx + 3];
  axes_from_quat(q[0], q[1], q[2], q[3], n.a1, n.a2, n.a3);
  return n;
}

// (dE/dc, dE/da1..a3) of one nucleotide -> (dE/dcenter, dE/dquat) atomics; zero components are skipped
template <class T>
__device__ __forceinline__ void scatter_nuc_grad(const EnergyDev<T>& a, long long idx, const NucGrad<T>& g,
                                                 const T q[4]) {
  if (a.d_center) {
    if (g.c.x != T(0)) atomicAdd(&a.d_center[3 * idx], g.c.x);
    if (g.c.y != T(0)) atomicAdd(&a.d_center[3 * idx + 1], g.c.y);
    if (g.c.z != T(0)) atomicAdd(&a.d_center[3 * idx + 2], g.c.z);
  }
  if (a.d_quat) {
    T dq[4];
    quat_grad(g, q[0], q[1], q[2], q[3], dq);
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (dq[c] != T(0)) atomicAdd(&a.d_quat[4 * idx + c], dq[c]);
  }
}


template <class T>
bool frame_kernel_eligible(const EnergyDev<T>& a);
template <class T>
int launch_frame_kernel(cudaStream_t s, const EnergyDev<T>& a, bool want_params);
// unbonded terms of explicit pair lists of any size, phase-queued (list_kernels.cu)
template <class T>
int launch_list_kernel(cudaStream_t s, const EnergyDev<T>& a, void* workspace, bool want_forces, bool want_params);
template <class T>
size_t list_workspace_bytes(long long n, long long n_frames, long long pair_capacity);

}  // namespace mb
