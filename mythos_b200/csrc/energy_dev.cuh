// energy_dev.cuh -- device-side argument block and the shared-memory parameter-gradient accumulator shared by
// the pair kernels (energy_kernels.cu) and the frame-resident kernel (frame_kernels.cu).
#pragma once
#include "common.cuh"
#include "oxdna_device.cuh"

namespace mb {

constexpr unsigned kFull = 0xffffffffu;

// per-frame structural observables (observables_dev.cuh): device-side copy of mb_observable_spec
struct ObsDev {
  const int32_t* base_pairs;
  const int32_t* quartets;
  int n_base_pairs, n_quartets;
  double sigma_backbone;
};
int check_observable_spec(const mb_observable_spec* spec, ObsDev* o);

// Probabilistic sequences (mythos/energy/utils.py:45-132): the sequence weight of a pair is the expectation of the 4x4
// table under the pair's nucleotide distribution.  Nucleotides are independent except the two members of one base pair:
//   different base pairs / unpaired:   w = sum_ab pmarg[i][a] W[a][b] pmarg[j][b]      (pmarg = per-nucleotide marginals)
//   same base pair:                    w = same_w[bp][within_i]  (= sum_t p_bp[t] W[BP[t][within_i]][BP[t][within_j]],
//                                          evaluated on the host where it chains to the table and to the base-pair distribution)
template <class T>
struct PseqDev {
  const T* pmarg;          // (N,4) or nullptr = discrete sequence
  const int32_t* bp_of;    // (N) base-pair index or -1
  const int32_t* within;   // (N) 0 / 1: which member of its base pair
  const T* same_w_stack;   // (n_bp,2) stacking table expectation of a same-base-pair pair, by within_i
  const T* same_w_hb;      // (n_bp,2) hydrogen-bond table
  T* d_pmarg;              // out (N,4) gradient of sum_f sum_t cot E, or nullptr
  T* d_same_w_stack;       // out (n_bp,2)
  T* d_same_w_hb;          // out (n_bp,2)
  unsigned terms;          // bits MB_TERM_STACK / MB_TERM_HB: which terms use the distribution
};
// weight of pair (i,j) for table W (16 reals); `same` <- flat index into the (n_bp,2) same-pair tables or -1
template <class T>
__device__ __forceinline__ T pseq_weight(const PseqDev<T>& ps, const T* W, const T* same_w, int i, int j, int& same) {
  same = -1;
  const int bi = ps.bp_of[i];
  if (bi >= 0 && bi == ps.bp_of[j]) {
    same = 2 * bi + ps.within[i];
    return same_w[same];
  }
  T w = T(0);
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    T row = T(0);
#pragma unroll
    for (int y = 0; y < 4; ++y) row += W[4 * x + y] * ps.pmarg[4 * j + y];
    w += ps.pmarg[4 * i + x] * row;
  }
  return w;
}
// gradient of coef * w(i,j) with respect to the table (-> acc), the marginals and the same-pair table (-> atomics)
template <class T, bool WP, class Acc>
__device__ __forceinline__ void pseq_weight_grad(const PseqDev<T>& ps, const T* W, T* d_same_w, int w00, int i, int j, int same, T coef,
                                                 Acc& acc) {
  // (no early exit: acc.add is warp-convergent, every lane must reach it -- a same-pair lane contributes zeros there)
  if (same >= 0) {
    if (d_same_w && coef != T(0)) atomicAdd(&d_same_w[same], coef);
    coef = T(0);
  }
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    T gi = T(0);
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const T pj = ps.pmarg[4 * j + y];
      gi += W[4 * x + y] * pj;
      if (WP) acc.add(0, w00 + 4 * x + y, coef * ps.pmarg[4 * i + x] * pj);
    }
    if (ps.d_pmarg && coef != T(0)) atomicAdd(&ps.d_pmarg[4 * i + x], coef * gi);
  }
  if (ps.d_pmarg && coef != T(0)) {
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      T gj = T(0);
#pragma unroll
      for (int x = 0; x < 4; ++x) gj += ps.pmarg[4 * i + x] * W[4 * x + y];
      atomicAdd(&ps.d_pmarg[4 * j + y], coef * gj);
    }
  }
}

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
  PseqDev<T> pseq;            // probabilistic sequence weights (generic pair kernel only); pmarg == nullptr: discrete seq
  T* acc_scratch;             // frame-resident kernel: [resident CTAs][MB_P_COUNT][32] parameter-gradient images (workspace)
  ObsDev obs;                 // fused observables epilogue of the frame-resident kernel (obs_out != nullptr)
  T* obs_out;                 // (F, MB_N_OBS) or nullptr
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


// Group form for the warp-reducing accumulator: a reduce-scatter butterfly.  In the first rounds every lane keeps half of
// its values and hands the other half to its partner, so N values cost N - 1 + (5 - log2 N) shuffle-adds instead of
// 5 N, and the N sums end up on N different lanes, which issue their shared-memory atomics in ONE instruction.
template <class T, int N>
__device__ __forceinline__ void acc_add_group(SmemAcc<T>& acc, int bank, const int (&idx)[N], const T (&v)[N]) {
  static_assert(N >= 1 && N <= 8, "group of at most eight values");
  if (acc.per_lane) {
#pragma unroll
    for (int k = 0; k < N; ++k) acc.add(bank, idx[k], v[k]);
    return;
  }
  bool any = false;
#pragma unroll
  for (int k = 0; k < N; ++k) any |= (v[k] != T(0));
  if (!__any_sync(kFull, any)) return;
  const int lane = threadIdx.x & 31;
  constexpr int W = N > 4 ? 8 : (N > 2 ? 4 : (N > 1 ? 2 : 1));  // padded width
  T a[W];
#pragma unroll
  for (int k = 0; k < W; ++k) a[k] = k < N ? v[k] : T(0);
  int p = 0;  // which value this lane ends up holding
  int o = 16;
#pragma unroll
  for (int w = W; w > 1; w >>= 1, o >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int k = 0; k < w / 2; ++k) {
      const T send = hi ? a[k] : a[k + w / 2];
      const T keep = hi ? a[k + w / 2] : a[k];
      a[k] = keep + __shfl_xor_sync(kFull, send, o);
    }
    p = 2 * p + (hi ? 1 : 0);
  }
  T t = a[0];
  for (; o > 0; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
  // the lanes whose low bits (below the rounds that split) are zero hold the finished sums
  constexpr int kLow = 32 / W;  // lanes per value
  int id = idx[0];
#pragma unroll
  for (int k = 1; k < N; ++k) id = (p == k) ? idx[k] : id;
  if ((lane & (kLow - 1)) == 0 && p < N && t != T(0)) atomicAdd(&acc.sh[id], t);
}

// Parameter-gradient accumulator in GLOBAL memory, one slot per (parameter, lane) of this CTA's scratch image: every lane
// adds its own contribution with a fire-and-forget RED.ADD (no warp reduction, no shuffles, no dependent chain); the 32
// lane slots of a parameter are summed once per frame.  The image is L2-resident (232 x 32 reals per resident CTA).
template <class T>
struct GlobAcc {
  T* img;  // [MB_P_COUNT][32]
  __device__ __forceinline__ void add(int, int idx, T v) {
    if (v != T(0)) atomicAdd(img + idx * 32 + (threadIdx.x & 31), v);
  }
  __device__ __forceinline__ void add_scatter(int, int idx, T v, bool pred) {
    if (pred && v != T(0)) atomicAdd(img + idx * 32 + (threadIdx.x & 31), v);
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

// Angular pre-screen of the hydrogen-bonding / cross-stacking queue.  Both terms carry the plain factors
// f4(theta1) f4(theta2) f4(theta3) with theta_k = acos(clamp(x_k)), x1 = -a1i.a1j, x2 = -a1j.dh, x3 = a1i.dh; f4 is non-zero
// only for |theta - theta0| < delta_c, i.e. for x inside a cosine window that depends on the parameters alone.  A pair
// whose x_k miss the windows of BOTH terms contributes exactly zero (value and every gradient) and never needs the six
// acos of the full evaluation.  Windows are widened by 1e-9 so that rounding can only let extra pairs through.
template <class T>
struct CosWin {
  T lo, hi;
};
template <class T>
__device__ __forceinline__ CosWin<T> f4_cos_window(const T* p /* f4 block: theta0, delta_star, delta_c, a, b */) {
  const T pi = Consts<T>::pi();
  const T t_lo = p[0] - p[2], t_hi = p[0] + p[2];
  CosWin<T> w;
  if (t_lo >= pi || t_hi <= T(0)) {  // empty
    w.lo = T(2);
    w.hi = T(-2);
    return w;
  }
  w.hi = (t_lo <= T(0)) ? T(2) : cos(t_lo) + T(1e-9);
  w.lo = (t_hi >= pi) ? T(-2) : cos(t_hi) - T(1e-9);
  return w;
}
// six windows of one bank: [0..2] hydrogen bonding theta1,2,3; [3..5] cross stacking theta1,2,3
template <class T>
__device__ __forceinline__ void bp_windows(const T* P, CosWin<T> w[6]) {
  w[0] = f4_cos_window(P + MB_P_HB_T1_TH0);
  w[1] = f4_cos_window(P + MB_P_HB_T2_TH0);
  w[2] = f4_cos_window(P + MB_P_HB_T3_TH0);
  w[3] = f4_cos_window(P + MB_P_CROSS_T1_TH0);
  w[4] = f4_cos_window(P + MB_P_CROSS_T2_TH0);
  w[5] = f4_cos_window(P + MB_P_CROSS_T3_TH0);
}
// nine windows of one bank (frame kernel): [0..5] as bp_windows; [6..8] hydrogen bonding theta4, theta7, theta8 as windows of
// a3i.a3j, -a3j.dh and -a3i.dh (theta8 = pi - acos(a3i.dh) = acos(-a3i.dh))
template <class T>
__device__ __forceinline__ void bp_windows9(const T* P, CosWin<T> w[9]) {
  bp_windows(P, w);
  w[6] = f4_cos_window(P + MB_P_HB_T4_TH0);
  w[7] = f4_cos_window(P + MB_P_HB_T7_TH0);
  w[8] = f4_cos_window(P + MB_P_HB_T8_TH0);
}
// bit 0: hydrogen bonding of the pair can be non-zero (radial window, non-zero table weight, all SIX angular windows);
// bit 1: cross stacking can be (radial window, the three plain angles).  The two terms are queued separately so that a
// batch of either runs one term's code with dense lanes.
template <class T>
__device__ __forceinline__ unsigned bp_screen2(const T* P, const CosWin<T>* w, unsigned mask, const V3<T>& d, T r2, const V3<T>& a1i,
                                               const V3<T>& a1j, const V3<T>& a3i, const V3<T>& a3j, int tab) {
  // A screen only has to be a SUPERSET of the pairs with a non-zero term (the terms themselves are evaluated exactly in
  // phase 3), so it avoids the double-precision square root and division: radial windows on r^2 against the squared limits
  // (widened by 1e-9 relative), 1/r from the float32 rsqrt (relative error < 3e-7) with the cosine windows widened by 1e-6.
  const T lo_hb = P[MB_P_HB_RCLOW] * P[MB_P_HB_RCLOW] * T(1 - 1e-9), hi_hb = P[MB_P_HB_RCHIGH] * P[MB_P_HB_RCHIGH] * T(1 + 1e-9);
  const T lo_cr = P[MB_P_CROSS_RCLOW] * P[MB_P_CROSS_RCLOW] * T(1 - 1e-9), hi_cr = P[MB_P_CROSS_RCHIGH] * P[MB_P_CROSS_RCHIGH] * T(1 + 1e-9);
  const bool rad_hb = (mask & (1u << MB_TERM_HB)) && lo_hb < r2 && r2 < hi_hb && P[MB_P_HB_W00 + tab] != T(0);
  const bool rad_cr = (mask & (1u << MB_TERM_CROSS)) && lo_cr < r2 && r2 < hi_cr;
  if (!(rad_hb || rad_cr) || !(r2 > T(0))) return 0u;
  const T ir = T(rsqrtf(float(r2)));
  const T eps = T(1e-6);
  const T x1 = -dot(a1i, a1j), x2 = -dot(a1j, d) * ir, x3 = dot(a1i, d) * ir;
  bool hb = rad_hb && w[0].lo - eps < x1 && x1 < w[0].hi + eps && w[1].lo - eps < x2 && x2 < w[1].hi + eps && w[2].lo - eps < x3 &&
            x3 < w[2].hi + eps;
  const bool cr = rad_cr && w[3].lo - eps < x1 && x1 < w[3].hi + eps && w[4].lo - eps < x2 && x2 < w[4].hi + eps && w[5].lo - eps < x3 &&
                  x3 < w[5].hi + eps;
  if (hb) {
    const T x4 = dot(a3i, a3j), x7 = -dot(a3j, d) * ir, x8 = -dot(a3i, d) * ir;
    hb = w[6].lo - eps < x4 && x4 < w[6].hi + eps && w[7].lo - eps < x7 && x7 < w[7].hi + eps && w[8].lo - eps < x8 && x8 < w[8].hi + eps;
  }
  return (hb ? 1u : 0u) | (cr ? 2u : 0u);
}
// true if hydrogen bonding or cross stacking of the pair can be non-zero (d = base_j - base_i, r2 = |d|^2)
template <class T>
__device__ __forceinline__ bool bp_screen(const T* P, const CosWin<T>* w, unsigned mask, const V3<T>& d, T r2, const V3<T>& a1i,
                                          const V3<T>& a1j, int tab) {
  const T r = sqrt(r2);
  const bool rad_hb = (mask & (1u << MB_TERM_HB)) && P[MB_P_HB_RCLOW] < r && r < P[MB_P_HB_RCHIGH] && P[MB_P_HB_W00 + tab] != T(0);
  const bool rad_cr = (mask & (1u << MB_TERM_CROSS)) && P[MB_P_CROSS_RCLOW] < r && r < P[MB_P_CROSS_RCHIGH];
  if (!(rad_hb || rad_cr) || !(r > T(0))) return false;
  const T ir = T(1) / r;
  const T x1 = -dot(a1i, a1j), x2 = -dot(a1j, d) * ir, x3 = dot(a1i, d) * ir;
  const bool hb = rad_hb && w[0].lo < x1 && x1 < w[0].hi && w[1].lo < x2 && x2 < w[1].hi && w[2].lo < x3 && x3 < w[2].hi;
  const bool cr = rad_cr && w[3].lo < x1 && x1 < w[3].hi && w[4].lo < x2 && x2 < w[4].hi && w[5].lo < x3 && x3 < w[5].hi;
  return hb || cr;
}

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
bool frame_kernel_fits(int n, bool want_params);
template <class T>
int launch_frame_kernel(cudaStream_t s, const EnergyDev<T>& a, bool want_params);
size_t frame_scratch_bytes(int real_bytes);  // workspace the frame-resident kernel needs for its parameter-gradient images
// unbonded terms of explicit pair lists of any size, phase-queued (list_kernels.cu)
template <class T>
int launch_list_kernel(cudaStream_t s, const EnergyDev<T>& a, void* workspace, bool want_forces, bool want_params);
template <class T>
size_t list_workspace_bytes(long long n, long long n_frames, long long pair_capacity);

}  // namespace mb
