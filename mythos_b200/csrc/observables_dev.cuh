// observables_dev.cuh -- per-frame structural observables (propeller twist, rise, pitch angle, helical diameter).
//
// What is computed follows the reference expression by expression (paths relative to the reference repo):
//   propeller twist   mythos/observables/propeller.py:18-31,57-71    180 - acos(clamp(n_i . n_j)) in degrees, mean over base pairs
//   local helical axis mythos/observables/base.py:24-44             unit vector between the base-site midpoints of two base pairs
//   rise              mythos/observables/rise.py:19-37,58-70         projection of that displacement on the axis, Angstrom
//   pitch angle       mythos/observables/pitch.py:32-60,76-88        angle between the two backbone-backbone vectors after their
//                                                                    component along the axis is removed, radians
//   diameter          mythos/observables/diameter.py:21-43,63-76     backbone-backbone distance + sigma_backbone, Angstrom
// `displacement_fn` of the reference = disp() with the model's box.  One CTA per frame; `nuc(i)` hands back nucleotide i of
// the frame (from shared memory in the frame-resident energy kernel, from global memory in the standalone kernel).
#pragma once
#include "energy_dev.cuh"

namespace mb {

constexpr double kAngstromPerLength = 8.518;  // mythos/utils/units.py:5

// `red`: shared scratch of 4 * (blockDim.x / 32) reals.  Every thread of the CTA must call this.  The sums are reduced in a
// fixed order (lane tree, then warps in order), so the result is bitwise repeatable.
template <class T, class NucOf, class GeomOf>
__device__ __forceinline__ void frame_observables(const ObsDev& o, const T box[3], NucOf nuc, GeomOf geom, T* red, T* out) {
  const T pi = Consts<T>::pi();
  T s_prop = 0, s_rise = 0, s_pitch = 0, s_diam = 0;
  for (int k = threadIdx.x; k < o.n_base_pairs; k += blockDim.x) {
    const int i = o.base_pairs[2 * k], j = o.base_pairs[2 * k + 1];
    const Nuc<T> ni = nuc(i), nj = nuc(j);
    const Geom<T>& gi = geom(i);
    const Geom<T>& gj = geom(j);
    s_prop += T(180) - acos(clamp1(dot(ni.a3, nj.a3))) * (T(180) / pi);
    const V3<T> d = disp(site(ni, gi.back[0], gi.back[1], gi.back[2]), site(nj, gj.back[0], gj.back[1], gj.back[2]), box);
    s_diam += (sqrt(dot(d, d)) + T(o.sigma_backbone)) * T(kAngstromPerLength);
  }
  for (int k = threadIdx.x; k < o.n_quartets; k += blockDim.x) {
    const int a1 = o.quartets[4 * k], b1 = o.quartets[4 * k + 1], a2 = o.quartets[4 * k + 2], b2 = o.quartets[4 * k + 3];
    const Nuc<T> na1 = nuc(a1), nb1 = nuc(b1), na2 = nuc(a2), nb2 = nuc(b2);
    const Geom<T>&ga1 = geom(a1), &gb1 = geom(b1), &ga2 = geom(a2), &gb2 = geom(b2);
    const V3<T> m1 = T(0.5) * (site(na1, ga1.base, T(0), T(0)) + site(nb1, gb1.base, T(0), T(0)));
    const V3<T> m2 = T(0.5) * (site(na2, ga2.base, T(0), T(0)) + site(nb2, gb2.base, T(0), T(0)));
    const V3<T> dr = disp(m2, m1, box);
    const T norm = sqrt(dot(dr, dr));
    const V3<T> axis = (T(1) / norm) * dr;
    s_rise += dot(dr, axis) * T(kAngstromPerLength);
    const V3<T> bb1 = disp(site(nb1, gb1.back[0], gb1.back[1], gb1.back[2]), site(na1, ga1.back[0], ga1.back[1], ga1.back[2]), box);
    const V3<T> bb2 = disp(site(nb2, gb2.back[0], gb2.back[1], gb2.back[2]), site(na2, ga2.back[0], ga2.back[1], ga2.back[2]), box);
    const V3<T> p1 = disp(bb1, dot(axis, bb1) * axis, box), p2 = disp(bb2, dot(axis, bb2) * axis, box);
    const V3<T> u1 = (T(1) / sqrt(dot(p1, p1))) * p1, u2 = (T(1) / sqrt(dot(p2, p2))) * p2;
    s_pitch += acos(clamp1(dot(u1, u2)));
  }
  T v[4] = {s_prop, s_rise, s_pitch, s_diam};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[c] += __shfl_down_sync(kFull, v[c], off);
    if (lane == 0) red[c * n_warps + warp] = v[c];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    T s = 0;
    for (int w = 0; w < n_warps; ++w) s += red[threadIdx.x * n_warps + w];
    const int cnt = (threadIdx.x == MB_OBS_PROPELLER || threadIdx.x == MB_OBS_DIAMETER) ? o.n_base_pairs : o.n_quartets;
    out[threadIdx.x] = s / T(cnt);  // 0 / 0 = NaN for an empty list, like jnp.mean
  }
  __syncthreads();
}

template <class T>
int launch_observables(cudaStream_t s, const ModelT<T>& M, int n, int n_frames, const T* center, const T* quat,
                       const int32_t* nt_type, const ObsDev& o, T* out);

}  // namespace mb
