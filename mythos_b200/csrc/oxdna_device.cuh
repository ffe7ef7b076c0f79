// oxdna_device.cuh -- per-pair device math of the oxDNA-family energy terms: value, analytic gradient with
// respect to the pair geometry, and analytic gradient with respect to the kernel-level parameter bank.
//
// Everything here is a template over the real type T (float / double) and is written as
// __host__ __device__ so the exact same expressions can be unit-checked by a host-compiled test harness
// (tests/host_check, test infrastructure only -- the shipped library has no host evaluation path).
//
// What each routine evaluates follows the reference term by term (paths relative to the reference repo):
//   f1,f2,f3,f4,f5   mythos/energy/dna1/base_functions.py:13-129        f6  mythos/energy/dna2/base_functions.py:13-17
//   FENE             mythos/energy/dna1/interactions.py:16-41           exc. volume  interactions.py:44-135
//   stacking         dna1/stacking.py:192-289, dna1/interactions.py:142-250, rna2/stacking.py:186-261
//   HB / cross       dna1/hydrogen_bonding.py:232-306, dna1/cross_stacking.py:192-266, rna2/cross_stacking.py:156-223
//   coaxial          dna1/coaxial_stacking.py:181-260, dna2/coaxial_stacking.py:138-201
//   Debye            dna2/debye.py:82-110, dna2/interactions.py:15-28
// Branch tests are strict (<) exactly as the reference's jnp.where conditions.  Derivatives are the analytic
// derivatives of the selected branch (what jax.grad of the where-expression gives), with the limit value 0
// defined where the reference's autodiff would hit 0*inf (acos at |x| = 1).
#pragma once

#include "../../include/mythos_b200.h"

#if defined(__CUDACC__)
#define MB_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define MB_HD inline
#endif

namespace mb {

template <class T>
struct V3 {
  T x, y, z;
};
template <class T>
MB_HD V3<T> v3(T x, T y, T z) {
  V3<T> r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
template <class T>
MB_HD V3<T> operator+(const V3<T>& a, const V3<T>& b) {
  return v3<T>(a.x + b.x, a.y + b.y, a.z + b.z);
}
template <class T>
MB_HD V3<T> operator-(const V3<T>& a, const V3<T>& b) {
  return v3<T>(a.x - b.x, a.y - b.y, a.z - b.z);
}
template <class T>
MB_HD V3<T> operator-(const V3<T>& a) {
  return v3<T>(-a.x, -a.y, -a.z);
}
template <class T>
MB_HD V3<T> operator*(T s, const V3<T>& a) {
  return v3<T>(s * a.x, s * a.y, s * a.z);
}
template <class T>
MB_HD T dot(const V3<T>& a, const V3<T>& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z;
}
template <class T>
MB_HD V3<T> cross(const V3<T>& a, const V3<T>& b) {
  return v3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <class T>
MB_HD void axpy(V3<T>& y, T a, const V3<T>& x) {
  y.x += a * x.x;
  y.y += a * x.y;
  y.z += a * x.z;
}

template <class T>
struct Consts;
template <>
struct Consts<double> {
  static MB_HD double pi() { return 3.14159265358979323846; }
  static MB_HD double fene_eps() { return 1e-10; }
};
template <>
struct Consts<float> {
  static MB_HD float pi() { return 3.14159265358979323846f; }
  static MB_HD float fene_eps() { return 1e-10f; }
};

// ---------------------------------------------------------------------------------------------------------
// Orientation: body axes from the (un-normalised) quaternion, mythos/energy/utils.py:18-36
template <class T>
struct Nuc {
  V3<T> c, a1, a2, a3;
};
template <class T>
MB_HD void axes_from_quat(T q0, T q1, T q2, T q3, V3<T>& a1, V3<T>& a2, V3<T>& a3) {
  const T q00 = q0 * q0, q11 = q1 * q1, q22 = q2 * q2, q33 = q3 * q3;
  a1 = v3<T>(q00 + q11 - q22 - q33, T(2) * (q1 * q2 + q0 * q3), T(2) * (q1 * q3 - q0 * q2));
  a2 = v3<T>(T(2) * (q1 * q2 - q0 * q3), q00 - q11 + q22 - q33, T(2) * (q2 * q3 + q0 * q1));
  a3 = v3<T>(T(2) * (q1 * q3 + q0 * q2), T(2) * (q2 * q3 - q0 * q1), q00 - q11 - q22 + q33);
}

// gradient of the energy with respect to (c, a1, a2, a3) of one nucleotide
template <class T>
struct NucGrad {
  V3<T> c, a1, a2, a3;
  MB_HD void zero() {
    c = a1 = a2 = a3 = v3<T>(T(0), T(0), T(0));
  }
};

// chain (dE/da1, dE/da2, dE/da3) -> dE/dq through the quadratic forms above
template <class T>
MB_HD void quat_grad(const NucGrad<T>& g, T q0, T q1, T q2, T q3, T out[4]) {
  const V3<T>&A = g.a1, &B = g.a2, &C = g.a3;
  out[0] = T(2) * (q0 * A.x + q3 * A.y - q2 * A.z - q3 * B.x + q0 * B.y + q1 * B.z + q2 * C.x - q1 * C.y + q0 * C.z);
  out[1] = T(2) * (q1 * A.x + q2 * A.y + q3 * A.z + q2 * B.x - q1 * B.y + q0 * B.z + q3 * C.x - q0 * C.y - q1 * C.z);
  out[2] = T(2) * (-q2 * A.x + q1 * A.y - q0 * A.z + q1 * B.x + q2 * B.y + q3 * B.z + q0 * C.x + q3 * C.y - q2 * C.z);
  out[3] = T(2) * (-q3 * A.x + q0 * A.y + q1 * A.z - q0 * B.x - q3 * B.y + q2 * B.z + q1 * C.x + q2 * C.y + q3 * C.z);
}

// site = c + k1*a1 + k2*a2 + k3*a3 ; scatter a site gradient back onto the nucleotide
template <class T>
MB_HD V3<T> site(const Nuc<T>& n, T k1, T k2, T k3) {
  return v3<T>(n.c.x + k1 * n.a1.x + k2 * n.a2.x + k3 * n.a3.x, n.c.y + k1 * n.a1.y + k2 * n.a2.y + k3 * n.a3.y,
               n.c.z + k1 * n.a1.z + k2 * n.a2.z + k3 * n.a3.z);
}
template <class T>
MB_HD void site_grad(NucGrad<T>& g, T sgn, const V3<T>& gs, T k1, T k2, T k3) {
  axpy(g.c, sgn, gs);
  if (k1 != T(0)) axpy(g.a1, sgn * k1, gs);
  if (k2 != T(0)) axpy(g.a2, sgn * k2, gs);
  if (k3 != T(0)) axpy(g.a3, sgn * k3, gs);
}

// flavour geometry converted to T once per kernel
template <class T>
struct Geom {
  T back[3], back_stack, stack, base, stack3[2], stack5[2], p3[3], p5[3];
  int use_back_stack;
  MB_HD void load(const mb_flavour_geom& g) {
    #pragma unroll
    for (int k = 0; k < 3; ++k) {
      back[k] = T(g.back[k]);
      p3[k] = T(g.p3[k]);
      p5[k] = T(g.p5[k]);
    }
    #pragma unroll
    for (int k = 0; k < 2; ++k) {
      stack3[k] = T(g.stack3[k]);
      stack5[k] = T(g.stack5[k]);
    }
    back_stack = T(g.back_stack);
    stack = T(g.stack);
    base = T(g.base);
    use_back_stack = g.use_back_stack;
  }
};

// displacement a - b, wrapped like jax_md.space.periodic: mod(d + L/2, L) - L/2 (floor-mod); box 0 = free
template <class T>
MB_HD T wrap1(T d, T L) {
  if (L > T(0)) {
    // floor-mod exactly as XLA lowers jnp.mod: fmod is exact, then one sign fix-up
    T s = fmod(d + T(0.5) * L, L);
    if (s != T(0) && s < T(0)) s += L;
    d = s - T(0.5) * L;
  }
  return d;
}
template <class T>
MB_HD V3<T> disp(const V3<T>& a, const V3<T>& b, const T box[3]) {
  // the box is either all zero (free space) or all positive (ABI contract): one test for the common free case
  if (!(box[0] > T(0))) return v3<T>(a.x - b.x, a.y - b.y, a.z - b.z);
  return v3<T>(wrap1(a.x - b.x, box[0]), wrap1(a.y - b.y, box[1]), wrap1(a.z - b.z, box[2]));
}

// ---------------------------------------------------------------------------------------------------------
// Scalar primitives.  Each *_val returns f and df/dx; each *_par writes coef * df/dparam into the parameter
// accumulator (only the differentiable slots; the branch limits r_low, r_high, delta_star, x_star carry no
// gradient because they only appear in where-conditions).
// Parameter accumulator concept:  acc.add(bank_offset, index, value)  -- warp-convergent on the device.
// Group form: N (<= 8) values with their indices at once.  Accumulators that reduce across the warp overload it so that
// the N values share the shuffle rounds (energy_dev.cuh: SmemAcc); every other accumulator takes N single adds.
template <class Acc, class T, int N>
MB_HD void acc_add_group(Acc& acc, int bank, const int (&idx)[N], const T (&v)[N]) {
#ifdef __CUDA_ARCH__
  #pragma unroll
#endif
  for (int k = 0; k < N; ++k) acc.add(bank, idx[k], v[k]);
}

// 1 / sqrt(x) for x > 0: the device's rsqrt (one special-function seed + Newton steps, ~1 ulp) -- where a distance and its
// reciprocal are both needed this replaces a square root and a division
template <class T>
MB_HD T inv_sqrt(T x) {
#ifdef __CUDA_ARCH__
  // float32 seed (relative error 2^-22) + two Newton steps in the working precision: 1e-13, then rounding level.  The
  // library's rsqrt(double) measures 19 FMA slots (profiles/r02_special_weights.json), this sequence about 10; arguments
  // here are squared distances of order 1e-3 .. 1e2, far from the float range limits.
  T y = T(rsqrtf(float(x)));
  const T hx = T(0.5) * x;
  y = fma(y, fma(-hx * y, y, T(0.5)), y);
  if (sizeof(T) > 4) y = fma(y, fma(-hx * y, y, T(0.5)), y);
  return y;
#else
  return T(1) / sqrt(x);
#endif
}
template <class T>
MB_HD T clamp1(T x) {
  return x >= T(1) ? T(1) : (x <= T(-1) ? T(-1) : x);
}
// theta = acos(clamp(x)); dth = d theta / d x  (0 in the clamped region)
template <class T>
MB_HD T acos_c(T x, T& dth) {
  x = clamp1(x);
  const T s = T(1) - x * x;
  dth = (s > T(0)) ? T(-1) / sqrt(s) : T(0);
  return acos(x);
}

// f1 block layout: [r_low, r_high, r_c_low, r_c_high, a, r0, r_c, b_low, b_high], eps == 1
template <class T>
MB_HD T f1_val(T r, const T* p, T& df) {
  df = T(0);
  if (p[0] < r && r < p[1]) {
    const T u = exp(-p[4] * (r - p[5])), uc = exp(-p[4] * (p[6] - p[5]));
    df = T(2) * (T(1) - u) * p[4] * u;
    return (T(1) - u) * (T(1) - u) - (T(1) - uc) * (T(1) - uc);
  }
  if (p[2] < r && r < p[0]) {
    const T t = p[2] - r;
    df = T(-2) * p[7] * t;
    return p[7] * t * t;
  }
  if (p[1] < r && r < p[3]) {
    const T t = p[3] - r;
    df = T(-2) * p[8] * t;
    return p[8] * t * t;
  }
  return T(0);
}
template <class T, class Acc>
MB_HD void f1_par(T r, const T* p, T coef, int bank, int base, Acc& acc) {
  T g_rcl = 0, g_rch = 0, g_a = 0, g_r0 = 0, g_rc = 0, g_bl = 0, g_bh = 0;
  if (coef != T(0)) {
    if (p[0] < r && r < p[1]) {
      const T u = exp(-p[4] * (r - p[5])), uc = exp(-p[4] * (p[6] - p[5]));
      const T m = T(2) * (T(1) - u) * u, mc = T(2) * (T(1) - uc) * uc;
      g_a = coef * (m * (r - p[5]) - mc * (p[6] - p[5]));
      g_r0 = coef * (-m * p[4] + mc * p[4]);
      g_rc = coef * (-mc * p[4]);
    } else if (p[2] < r && r < p[0]) {
      const T t = p[2] - r;
      g_bl = coef * t * t;
      g_rcl = coef * T(2) * p[7] * t;
    } else if (p[1] < r && r < p[3]) {
      const T t = p[3] - r;
      g_bh = coef * t * t;
      g_rch = coef * T(2) * p[8] * t;
    }
  }
  const int idx[7] = {base + 2, base + 3, base + 4, base + 5, base + 6, base + 7, base + 8};
  const T val[7] = {g_rcl, g_rch, g_a, g_r0, g_rc, g_bl, g_bh};
  acc_add_group(acc, bank, idx, val);
}

// f2 block layout: [r_low, r_high, r_c_low, r_c_high, k, r0, r_c, b_low, b_high]
template <class T>
MB_HD T f2_val(T r, const T* p, T& df) {
  df = T(0);
  if (p[0] < r && r < p[1]) {
    df = p[4] * (r - p[5]);
    return T(0.5) * p[4] * ((r - p[5]) * (r - p[5]) - (p[6] - p[5]) * (p[6] - p[5]));
  }
  if (p[2] < r && r < p[0]) {
    const T t = p[2] - r;
    df = T(-2) * p[4] * p[7] * t;
    return p[4] * p[7] * t * t;
  }
  if (p[1] < r && r < p[3]) {
    const T t = p[3] - r;
    df = T(-2) * p[4] * p[8] * t;
    return p[4] * p[8] * t * t;
  }
  return T(0);
}
template <class T, class Acc>
MB_HD void f2_par(T r, const T* p, T coef, int bank, int base, Acc& acc) {
  T g_rcl = 0, g_rch = 0, g_k = 0, g_r0 = 0, g_rc = 0, g_bl = 0, g_bh = 0;
  if (coef != T(0)) {
    if (p[0] < r && r < p[1]) {
      g_k = coef * T(0.5) * ((r - p[5]) * (r - p[5]) - (p[6] - p[5]) * (p[6] - p[5]));
      g_r0 = coef * p[4] * (p[6] - r);
      g_rc = coef * (-p[4] * (p[6] - p[5]));
    } else if (p[2] < r && r < p[0]) {
      const T t = p[2] - r;
      g_k = coef * p[7] * t * t;
      g_bl = coef * p[4] * t * t;
      g_rcl = coef * T(2) * p[4] * p[7] * t;
    } else if (p[1] < r && r < p[3]) {
      const T t = p[3] - r;
      g_k = coef * p[8] * t * t;
      g_bh = coef * p[4] * t * t;
      g_rch = coef * T(2) * p[4] * p[8] * t;
    }
  }
  const int idx[7] = {base + 2, base + 3, base + 4, base + 5, base + 6, base + 7, base + 8};
  const T val[7] = {g_rcl, g_rch, g_k, g_r0, g_rc, g_bl, g_bh};
  acc_add_group(acc, bank, idx, val);
}

// f3 block layout: [r_star, sigma, b, r_c]; eps separate
template <class T>
MB_HD T f3_val(T r, const T* p, T eps, T& df) {
  df = T(0);
  if (r < p[0]) {
    const T s2 = (p[1] * p[1]) / (r * r), s6 = s2 * s2 * s2;
    df = T(4) * eps * (T(-12) * s6 * s6 + T(6) * s6) / r;
    return T(4) * eps * (s6 * s6 - s6);
  }
  if (p[0] < r && r < p[3]) {
    const T t = p[3] - r;
    df = T(-2) * eps * p[2] * t;
    return eps * p[2] * t * t;
  }
  return T(0);
}
template <class T, class Acc>
MB_HD void f3_par(T r, const T* p, T eps, T coef, int bank, int base, int eps_idx, Acc& acc) {
  T g_eps = 0, g_sig = 0, g_b = 0, g_rc = 0;
  if (coef != T(0)) {
    if (r < p[0]) {
      const T s2 = (p[1] * p[1]) / (r * r), s6 = s2 * s2 * s2;
      g_eps = coef * T(4) * (s6 * s6 - s6);
      g_sig = coef * T(4) * eps * (T(12) * s6 * s6 - T(6) * s6) / p[1];
    } else if (p[0] < r && r < p[3]) {
      const T t = p[3] - r;
      g_eps = coef * p[2] * t * t;
      g_b = coef * eps * t * t;
      g_rc = coef * T(2) * eps * p[2] * t;
    }
  }
  const int idx[4] = {eps_idx, base + 1, base + 2, base + 3};
  const T val[4] = {g_eps, g_sig, g_b, g_rc};
  acc_add_group(acc, bank, idx, val);
}

// f4 block layout: [theta0, delta_star, delta_c, a, b]
// Branch tests are written on theta itself with the breakpoints theta0 -+ delta, exactly as the reference forms them
// (dna1/base_functions.py:82-107): a value that sits ON a breakpoint must fall into the same branch as there.
template <class T>
MB_HD T f4_val(T th, const T* p, T& df) {
  df = T(0);
  const T lo_s = p[0] - p[1], hi_s = p[0] + p[1];
  if (lo_s < th && th < hi_s) {
    const T t = th - p[0];
    df = T(-2) * p[3] * t;
    return T(1) - p[3] * t * t;
  }
  const T lo_c = p[0] - p[2], hi_c = p[0] + p[2];
  if (lo_c < th && th < lo_s) {
    const T u = lo_c - th;
    df = T(-2) * p[4] * u;
    return p[4] * u * u;
  }
  if (hi_s < th && th < hi_c) {
    const T u = hi_c - th;
    df = T(-2) * p[4] * u;
    return p[4] * u * u;
  }
  return T(0);
}
// accumulates into caller-held partial sums (theta0, delta_c, a, b) so that the symmetric forms
// f4(th) + f4(pi - th) cost one accumulator call per slot
template <class T>
MB_HD void f4_par_add(T th, const T* p, T coef, T g[4]) {
  if (coef == T(0)) return;
  const T lo_s = p[0] - p[1], hi_s = p[0] + p[1];
  const T lo_c = p[0] - p[2], hi_c = p[0] + p[2];
  if (lo_s < th && th < hi_s) {
    const T t = th - p[0];
    g[0] += coef * T(2) * p[3] * t;
    g[2] += coef * (-t * t);
  } else if (lo_c < th && th < lo_s) {
    const T u = lo_c - th;
    g[0] += coef * T(2) * p[4] * u;
    g[1] += coef * T(-2) * p[4] * u;
    g[3] += coef * u * u;
  } else if (hi_s < th && th < hi_c) {
    const T u = hi_c - th;
    g[0] += coef * T(2) * p[4] * u;
    g[1] += coef * T(2) * p[4] * u;
    g[3] += coef * u * u;
  }
}
template <class T, class Acc>
MB_HD void f4_par_flush(const T g[4], int bank, int base, Acc& acc) {
  const int idx[4] = {base + 0, base + 2, base + 3, base + 4};
  const T val[4] = {g[0], g[1], g[2], g[3]};
  acc_add_group(acc, bank, idx, val);
}
// two f4 blocks at once (eight values share the reduction rounds)
template <class T, class Acc>
MB_HD void f4_par_flush2(const T ga[4], int base_a, const T gb[4], int base_b, int bank, Acc& acc) {
  const int idx[8] = {base_a + 0, base_a + 2, base_a + 3, base_a + 4, base_b + 0, base_b + 2, base_b + 3, base_b + 4};
  const T val[8] = {ga[0], ga[1], ga[2], ga[3], gb[0], gb[1], gb[2], gb[3]};
  acc_add_group(acc, bank, idx, val);
}

// f5 block layout: [x_star, x_c, a, b]
template <class T>
MB_HD T f5_val(T x, const T* p, T& df) {
  df = T(0);
  if (x > T(0)) return T(1);
  if (p[0] < x && x < T(0)) {
    df = T(-2) * p[2] * x;
    return T(1) - p[2] * x * x;
  }
  if (p[1] < x && x < p[0]) {
    const T t = p[1] - x;
    df = T(-2) * p[3] * t;
    return p[3] * t * t;
  }
  return T(0);
}
template <class T, class Acc>
MB_HD void f5_par(T x, const T* p, T coef, int bank, int base, Acc& acc) {
  T g_xc = 0, g_a = 0, g_b = 0;
  if (coef != T(0) && !(x > T(0))) {
    if (p[0] < x && x < T(0)) {
      g_a = coef * (-x * x);
    } else if (p[1] < x && x < p[0]) {
      const T t = p[1] - x;
      g_b = coef * t * t;
      g_xc = coef * T(2) * p[3] * t;
    }
  }
  const int idx[3] = {base + 1, base + 2, base + 3};
  const T val[3] = {g_xc, g_a, g_b};
  acc_add_group(acc, bank, idx, val);
}

// f6 block layout: [a, b]
template <class T>
MB_HD T f6_val(T th, const T* p, T& df) {
  df = T(0);
  if (th >= p[1]) {
    const T t = th - p[1];
    df = p[0] * t;
    return T(0.5) * p[0] * t * t;
  }
  return T(0);
}

// ---------------------------------------------------------------------------------------------------------
// Helper: x = s * (u . dh) with dh = d / r a unit vector.  Adds gx * dx/du to gu and gx * dx/d(dh) to gdh.
template <class T>
MB_HD void dot_unit_grad(T gx, T s, const V3<T>& u, const V3<T>& dh, V3<T>& gu, V3<T>& gdh) {
  axpy(gu, gx * s, dh);
  axpy(gdh, gx * s, u);
}
// gradient w.r.t. a unit vector dh = d/r -> gradient w.r.t. d
template <class T>
MB_HD V3<T> unit_to_vec_grad(const V3<T>& gdh, const V3<T>& dh, T r) {
  const T pr = dot(gdh, dh);
  const T ir = T(1) / r;
  return v3<T>((gdh.x - pr * dh.x) * ir, (gdh.y - pr * dh.y) * ir, (gdh.z - pr * dh.z) * ir);
}

// ---------------------------------------------------------------------------------------------------------
// FENE (bonded).  d = back_i - back_j, r = |d|.  Returns E; dEdr = dE/dr.
template <class T, bool WP, class Acc>
MB_HD T fene_term(const T* P, int bank, bool act, T r, T cot, T& dEdr, Acc& acc) {
  T e = 0;
  dEdr = 0;
  T g_eps = 0, g_r0 = 0, g_dl = 0, g_fmax = 0, g_finf = 0;
  if (act) {
    const T eps = P[MB_P_FENE_EPS], r0 = P[MB_P_FENE_R0], dl = P[MB_P_FENE_DELTA], fmax = P[MB_P_FENE_FMAX],
            finf = P[MB_P_FENE_FINF];
    const T dr = r - r0;
    const T diff = sqrt(dr * dr + Consts<T>::fene_eps());
    const T S = sqrt(eps * eps + T(4) * fmax * fmax * dl * dl);
    const T xmax = (-eps + S) / (T(2) * fmax);
    if (diff > xmax) {
      const T lg = log(diff / xmax);
      const T om = T(1) - xmax * xmax / (dl * dl);
      e = (fmax - finf) * xmax * lg + finf * (diff - xmax) - T(0.5) * eps * log(om);
      const T dE_ddiff = (fmax - finf) * xmax / diff + finf;
      dEdr = dE_ddiff * dr / diff;
      if (WP) {
        // dE/dxmax = (fmax-finf) ln(diff/xmax) - (fmax-finf) - finf + eps xmax/(dl^2 om)
        const T dE_dx = (fmax - finf) * lg - fmax + eps * xmax / (dl * dl * om);
        const T dx_deps = (-T(1) + eps / S) / (T(2) * fmax);
        const T dx_ddl = T(2) * fmax * dl / S;
        const T dx_dfmax = (T(4) * fmax * dl * dl / S) / (T(2) * fmax) - (-eps + S) / (T(2) * fmax * fmax);
        g_eps = cot * (dE_dx * dx_deps - T(0.5) * log(om));
        g_dl = cot * (dE_dx * dx_ddl - eps * xmax * xmax / (dl * dl * dl * om));
        g_fmax = cot * (xmax * lg + dE_dx * dx_dfmax);
        g_finf = cot * (-xmax * lg + diff - xmax);
        g_r0 = cot * (-dEdr);
      }
    } else {
      const T x = dr * dr / (dl * dl);
      const T om = T(1) - x;
      e = T(-0.5) * eps * log(om);
      dEdr = eps * dr / (dl * dl * om);
      if (WP) {
        g_eps = cot * T(-0.5) * log(om);
        g_r0 = cot * (-dEdr);
        g_dl = cot * (-eps * x / (dl * om));
      }
    }
  }
  if (WP) {
    const int idx[5] = {MB_P_FENE_EPS, MB_P_FENE_R0, MB_P_FENE_DELTA, MB_P_FENE_FMAX, MB_P_FENE_FINF};
    const T val[5] = {g_eps, g_r0, g_dl, g_fmax, g_finf};
    acc_add_group(acc, bank, idx, val);
  }
  return e;
}

// one smoothed-LJ site pair: returns f3(r); adds cot * f3'(r) * dh to gs (gradient w.r.t. the site difference)
template <class T, bool WF, bool WP, class Acc>
MB_HD T exc_site(const T* P, int bank, int base, int eps_idx, bool act, const V3<T>& d, T cot, V3<T>& gs, Acc& acc) {
  T e = 0, r = 0;
  bool in = false;
  if (act) {
    const T r2 = dot(d, d);
    const T rc = P[base + 3];
    if (r2 < rc * rc && r2 > T(0)) {
      r = sqrt(r2);
      T df;
      e = f3_val(r, P + base, P[eps_idx], df);
      in = true;
      if (WF) axpy(gs, cot * df / r, d);
    }
  }
  if (WP) f3_par(r, P + base, P[eps_idx], in ? cot : T(0), bank, base, eps_idx, acc);
  return e;
}

// ---------------------------------------------------------------------------------------------------------
// Stacking (bonded pair i -> j = i's 5' neighbour in the internal order).
//   ds = stack_i - stack_j  (RNA form: stack5_i - stack3_j),  db = backS_i - backS_j
// Geometry gradients are returned w.r.t. ds, db and the axis vectors.
template <class T>
struct StackGrad {
  V3<T> ds, db, a3i, a3j, a2i, a2j, p3j, p5i;
};

template <class T, bool WF, bool WP, class Acc>
MB_HD T stack_term(const T* P, int bank, int form, bool act, const V3<T>& ds, const V3<T>& db, const V3<T>& a3i,
                   const V3<T>& a3j, const V3<T>& a2i, const V3<T>& a2j, const V3<T>& p3j, const V3<T>& p5i,
                   int tab, T cot, StackGrad<T>& G, Acc& acc, bool has_w_ext = false, T w_ext = T(0)) {
  // `has_w_ext`: the pair's sequence weight comes from outside (probabilistic sequences: an expectation over the table,
  // mythos/energy/utils.py:45-132) instead of the table entry `tab`; the table gradient is then the caller's business
  // factor slots: 0 f1(r)  1 f4(th4)|f4(th9)  2 f4(th5)  3 f4(th6)  4 f4(th10) (RNA only)  5 f5(phi1)  6 f5(phi2)
  T e = 0;
  T f[7], df[7], arg[7], dth[7];
  T w = 0, rs = 0, rb = 0;
  V3<T> sh = v3<T>(0, 0, 0), bh = sh;
  bool nz = false;
  const bool rna = (form == MB_STACK_RNA);
  const int b4[5] = {MB_P_STACK_T4_TH0, MB_P_STACK_T5_TH0, MB_P_STACK_T6_TH0, MB_P_STACK_T9_TH0, MB_P_STACK_T10_TH0};
  #pragma unroll
  for (int k = 0; k < 7; ++k) {
    f[k] = T(1);
    df[k] = dth[k] = arg[k] = T(0);
  }
  if (act) {
    w = has_w_ext ? w_ext : P[MB_P_STACK_W00 + tab];
    rs = sqrt(dot(ds, ds));
    f[0] = f1_val(rs, P + MB_P_STACK_RLOW, df[0]);
    arg[0] = rs;
    if (f[0] != T(0) && w != T(0)) {
      sh = (T(1) / rs) * ds;
      rb = sqrt(dot(db, db));
      bh = (T(1) / rb) * db;
      T d;
      // theta5 = pi - acos(sh . a3j), theta6 = pi - acos(a3i . sh)
      arg[2] = Consts<T>::pi() - acos_c(dot(sh, a3j), d);
      dth[2] = -d;
      f[2] = f4_val(arg[2], P + b4[1], df[2]);
      arg[3] = Consts<T>::pi() - acos_c(dot(a3i, sh), d);
      dth[3] = -d;
      f[3] = f4_val(arg[3], P + b4[2], df[3]);
      if (!rna) {
        arg[1] = acos_c(dot(a3i, a3j), d);
        dth[1] = d;
        f[1] = f4_val(arg[1], P + b4[0], df[1]);
      } else {
        arg[1] = acos_c(-dot(p3j, bh), d);  // theta9
        dth[1] = d;
        f[1] = f4_val(arg[1], P + b4[3], df[1]);
        arg[4] = acos_c(-dot(p5i, bh), d);  // theta10
        dth[4] = d;
        f[4] = f4_val(arg[4], P + b4[4], df[4]);
      }
      // f5(-cosphi1) with cosphi1 = -a2i . bh  ->  argument x = a2i . bh
      arg[5] = dot(a2i, bh);
      f[5] = f5_val(arg[5], P + MB_P_STACK_PHI1_XSTAR, df[5]);
      arg[6] = dot(a2j, bh);
      f[6] = f5_val(arg[6], P + MB_P_STACK_PHI2_XSTAR, df[6]);
      T prod = f[0];
      #pragma unroll
      for (int k = 1; k < 7; ++k) prod *= f[k];
      if (prod != T(0)) {
        nz = true;
        e = w * prod;
      }
    }
  }
  if (!(WF || WP)) return e;
  // coefficient of factor k: cot * w * prod_{m != k} f_m
  T oth[7];
  {
    T pre = nz ? cot * w : T(0);
    T suf[8];
    suf[7] = T(1);
    #pragma unroll
    for (int k = 6; k >= 0; --k) suf[k] = suf[k + 1] * f[k];
    #pragma unroll
    for (int k = 0; k < 7; ++k) {
      oth[k] = pre * suf[k + 1];
      pre *= f[k];
    }
  }
  if (WF && nz) {
    V3<T> gsh = v3<T>(0, 0, 0), gbh = gsh;
    axpy(G.ds, oth[0] * df[0], sh);
    T gx;
    gx = oth[2] * df[2] * dth[2];  // x = sh . a3j
    dot_unit_grad(gx, T(1), a3j, sh, G.a3j, gsh);
    gx = oth[3] * df[3] * dth[3];  // x = a3i . sh
    dot_unit_grad(gx, T(1), a3i, sh, G.a3i, gsh);
    if (!rna) {
      gx = oth[1] * df[1] * dth[1];  // x = a3i . a3j
      axpy(G.a3i, gx, a3j);
      axpy(G.a3j, gx, a3i);
    } else {
      gx = oth[1] * df[1] * dth[1];  // x = -p3j . bh
      dot_unit_grad(gx, T(-1), p3j, bh, G.p3j, gbh);
      gx = oth[4] * df[4] * dth[4];  // x = -p5i . bh
      dot_unit_grad(gx, T(-1), p5i, bh, G.p5i, gbh);
    }
    gx = oth[5] * df[5];  // x = a2i . bh
    dot_unit_grad(gx, T(1), a2i, bh, G.a2i, gbh);
    gx = oth[6] * df[6];
    dot_unit_grad(gx, T(1), a2j, bh, G.a2j, gbh);
    G.ds = G.ds + unit_to_vec_grad(gsh, sh, rs);
    G.db = G.db + unit_to_vec_grad(gbh, bh, rb);
  }
  if (WP) {
    f1_par(arg[0], P + MB_P_STACK_RLOW, oth[0], bank, MB_P_STACK_RLOW, acc);
    T g[4];
    // theta4 slot (DNA form only)
    g[0] = g[1] = g[2] = g[3] = T(0);
    if (!rna) f4_par_add(arg[1], P + b4[0], oth[1], g);
    f4_par_flush(g, bank, b4[0], acc);
    g[0] = g[1] = g[2] = g[3] = T(0);
    f4_par_add(arg[2], P + b4[1], oth[2], g);
    f4_par_flush(g, bank, b4[1], acc);
    g[0] = g[1] = g[2] = g[3] = T(0);
    f4_par_add(arg[3], P + b4[2], oth[3], g);
    f4_par_flush(g, bank, b4[2], acc);
    g[0] = g[1] = g[2] = g[3] = T(0);
    if (rna) f4_par_add(arg[1], P + b4[3], oth[1], g);
    f4_par_flush(g, bank, b4[3], acc);
    g[0] = g[1] = g[2] = g[3] = T(0);
    if (rna) f4_par_add(arg[4], P + b4[4], oth[4], g);
    f4_par_flush(g, bank, b4[4], acc);
    f5_par(arg[5], P + MB_P_STACK_PHI1_XSTAR, oth[5], bank, MB_P_STACK_PHI1_XSTAR, acc);
    f5_par(arg[6], P + MB_P_STACK_PHI2_XSTAR, oth[6], bank, MB_P_STACK_PHI2_XSTAR, acc);
    if (!has_w_ext) acc.add_scatter(bank, MB_P_STACK_W00 + tab, nz ? cot * e / w : T(0), nz);
  }
  return e;
}

// ---------------------------------------------------------------------------------------------------------
// Hydrogen bonding and cross stacking share six angles built on d = base_j - base_i (dh = d/r):
//   th1 = acos(-a1i.a1j)  th2 = acos(-a1j.dh)  th3 = acos(a1i.dh)  th4 = acos(a3i.a3j)
//   th7 = acos(-a3j.dh)   th8 = pi - acos(a3i.dh)
template <class T>
struct HbAngles {
  T th[6];   // th1, th2, th3, th4, th7, th8
  T dth[6];  // d th / d x
  bool ready;
};
template <class T>
MB_HD void hb_angles(const V3<T>& dh, const V3<T>& a1i, const V3<T>& a1j, const V3<T>& a3i, const V3<T>& a3j,
                     HbAngles<T>& A) {
  A.th[0] = acos_c(-dot(a1i, a1j), A.dth[0]);
  A.th[1] = acos_c(-dot(a1j, dh), A.dth[1]);
  A.th[2] = acos_c(dot(a1i, dh), A.dth[2]);
  A.th[3] = acos_c(dot(a3i, a3j), A.dth[3]);
  A.th[4] = acos_c(-dot(a3j, dh), A.dth[4]);
  T d;
  A.th[5] = Consts<T>::pi() - acos_c(dot(a3i, dh), d);
  A.dth[5] = -d;
  A.ready = true;
}
template <class T>
struct HbGrad {
  V3<T> d, a1i, a1j, a3i, a3j;
};
// scatter the six dE/dx_k onto (dh, axes)
template <class T>
MB_HD void hb_scatter(const T gx[6], const V3<T>& dh, T r, T gr, const V3<T>& a1i, const V3<T>& a1j,
                      const V3<T>& a3i, const V3<T>& a3j, HbGrad<T>& G) {
  V3<T> gdh = v3<T>(0, 0, 0);
  axpy(G.a1i, -gx[0], a1j);
  axpy(G.a1j, -gx[0], a1i);
  dot_unit_grad(gx[1], T(-1), a1j, dh, G.a1j, gdh);
  dot_unit_grad(gx[2], T(1), a1i, dh, G.a1i, gdh);
  axpy(G.a3i, gx[3], a3j);
  axpy(G.a3j, gx[3], a3i);
  dot_unit_grad(gx[4], T(-1), a3j, dh, G.a3j, gdh);
  dot_unit_grad(gx[5], T(1), a3i, dh, G.a3i, gdh);
  G.d = G.d + unit_to_vec_grad(gdh, dh, r);
  axpy(G.d, gr, dh);
}

template <class T, bool WF, bool WP, class Acc>
MB_HD T hb_term(const T* P, int bank, bool act, T r, const V3<T>& dh, const V3<T>& a1i, const V3<T>& a1j,
                const V3<T>& a3i, const V3<T>& a3j, HbAngles<T>& A, int tab, T cot, HbGrad<T>& G, Acc& acc,
                bool has_w_ext = false, T w_ext = T(0)) {
  const int b4[6] = {MB_P_HB_T1_TH0, MB_P_HB_T2_TH0, MB_P_HB_T3_TH0, MB_P_HB_T4_TH0, MB_P_HB_T7_TH0, MB_P_HB_T8_TH0};
  T e = 0, w = 0, fr = 0, dfr = 0;
  T f[6], df[6];
  bool nz = false;
  #pragma unroll
  for (int k = 0; k < 6; ++k) {
    f[k] = T(1);
    df[k] = T(0);
  }
  if (act) {
    w = has_w_ext ? w_ext : P[MB_P_HB_W00 + tab];
    if (w != T(0)) {
      fr = f1_val(r, P + MB_P_HB_RLOW, dfr);
      if (fr != T(0)) {
        if (!A.ready) hb_angles(dh, a1i, a1j, a3i, a3j, A);
        T prod = fr;
        #pragma unroll
        for (int k = 0; k < 6; ++k) {
          f[k] = f4_val(A.th[k], P + b4[k], df[k]);
          prod *= f[k];
        }
        if (prod != T(0)) {
          nz = true;
          e = w * prod;
        }
      }
    }
  }
  if (!(WF || WP)) return e;
  T oth[6], othr;
  {
    T pre = nz ? cot * w : T(0);
    T suf[7];
    suf[6] = T(1);
    #pragma unroll
    for (int k = 5; k >= 0; --k) suf[k] = suf[k + 1] * f[k];
    othr = pre * suf[0];
    pre *= fr;
    #pragma unroll
    for (int k = 0; k < 6; ++k) {
      oth[k] = pre * suf[k + 1];
      pre *= f[k];
    }
  }
  if (WF && nz) {
    T gx[6];
    #pragma unroll
    for (int k = 0; k < 6; ++k) gx[k] = oth[k] * df[k] * A.dth[k];
    hb_scatter(gx, dh, r, othr * dfr, a1i, a1j, a3i, a3j, G);
  }
  if (WP) {
    f1_par(r, P + MB_P_HB_RLOW, othr, bank, MB_P_HB_RLOW, acc);
    #pragma unroll
    for (int k = 0; k < 6; k += 2) {
      T ga[4] = {T(0), T(0), T(0), T(0)}, gb[4] = {T(0), T(0), T(0), T(0)};
      if (nz) {
        f4_par_add(A.th[k], P + b4[k], oth[k], ga);
        f4_par_add(A.th[k + 1], P + b4[k + 1], oth[k + 1], gb);
      }
      f4_par_flush2(ga, b4[k], gb, b4[k + 1], bank, acc);
    }
    if (!has_w_ext) acc.add_scatter(bank, MB_P_HB_W00 + tab, nz ? cot * e / w : T(0), nz);
  }
  return e;
}

template <class T, bool WF, bool WP, class Acc>
MB_HD T cross_term(const T* P, int bank, int form, bool act, T r, const V3<T>& dh, const V3<T>& a1i,
                   const V3<T>& a1j, const V3<T>& a3i, const V3<T>& a3j, HbAngles<T>& A, T cot, HbGrad<T>& G,
                   Acc& acc) {
  const int b4[6] = {MB_P_CROSS_T1_TH0, MB_P_CROSS_T2_TH0, MB_P_CROSS_T3_TH0,
                     MB_P_CROSS_T4_TH0, MB_P_CROSS_T7_TH0, MB_P_CROSS_T8_TH0};
  const T pi = Consts<T>::pi();
  T e = 0, fr = 0, dfr = 0;
  T f[6], df[6];
  bool nz = false;
  const bool rna = (form == MB_CROSS_RNA2);
  #pragma unroll
  for (int k = 0; k < 6; ++k) {
    f[k] = T(1);
    df[k] = T(0);
  }
  if (act) {
    fr = f2_val(r, P + MB_P_CROSS_RLOW, dfr);
    if (fr != T(0)) {
      if (!A.ready) hb_angles(dh, a1i, a1j, a3i, a3j, A);
      T prod = fr;
      #pragma unroll
      for (int k = 0; k < 6; ++k) {
        if (k < 3) {
          f[k] = f4_val(A.th[k], P + b4[k], df[k]);
        } else if (k == 3 && rna) {
          f[k] = T(1);
          df[k] = T(0);
        } else {
          T d2;
          f[k] = f4_val(A.th[k], P + b4[k], df[k]) + f4_val(pi - A.th[k], P + b4[k], d2);
          df[k] -= d2;
        }
        prod *= f[k];
      }
      if (prod != T(0)) {
        nz = true;
        e = prod;
      }
    }
  }
  if (!(WF || WP)) return e;
  T oth[6], othr;
  {
    T pre = nz ? cot : T(0);
    T suf[7];
    suf[6] = T(1);
    #pragma unroll
    for (int k = 5; k >= 0; --k) suf[k] = suf[k + 1] * f[k];
    othr = pre * suf[0];
    pre *= fr;
    #pragma unroll
    for (int k = 0; k < 6; ++k) {
      oth[k] = pre * suf[k + 1];
      pre *= f[k];
    }
  }
  if (WF && nz) {
    T gx[6];
    #pragma unroll
    for (int k = 0; k < 6; ++k) gx[k] = oth[k] * df[k] * A.dth[k];
    hb_scatter(gx, dh, r, othr * dfr, a1i, a1j, a3i, a3j, G);
  }
  if (WP) {
    f2_par(r, P + MB_P_CROSS_RLOW, othr, bank, MB_P_CROSS_RLOW, acc);
    #pragma unroll
    for (int k = 0; k < 6; k += 2) {
      T g2[2][4] = {{T(0), T(0), T(0), T(0)}, {T(0), T(0), T(0), T(0)}};
      #pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int kk = k + h;
        if (nz && !(kk == 3 && rna)) {
          f4_par_add(A.th[kk], P + b4[kk], oth[kk], g2[h]);
          if (kk >= 3) f4_par_add(pi - A.th[kk], P + b4[kk], oth[kk], g2[h]);
        }
      }
      f4_par_flush2(g2[0], b4[k], g2[1], b4[k + 1], bank, acc);
    }
  }
  return e;
}

// ---------------------------------------------------------------------------------------------------------
// Coaxial stacking.  ds = stack_j - stack_i (sh = ds/|ds|), db = back_j - back_i (bh = db/|db|, dna1 form only)
//   th4 = acos(a3i.a3j)  th1 = acos(-a1i.a1j)  th5 = acos(a3i.sh)  th6 = acos(-a3j.sh)
//   dna1 form: cosphi3 = sh.(bh x a1j), cosphi4 = sh.(bh x a1i), factor f4(th1)+f4(2pi-th1)
//   dna2 form: factor f4(th1)+f6(th1), no phi factors
template <class T>
struct CoaxGrad {
  V3<T> ds, db, a1i, a1j, a3i, a3j;
};
template <class T, bool WF, bool WP, class Acc>
MB_HD T coax_term(const T* P, int bank, int form, bool act, const V3<T>& ds, T rs, const V3<T>& db,
                  const V3<T>& a1i, const V3<T>& a1j, const V3<T>& a3i, const V3<T>& a3j, T cot, CoaxGrad<T>& G,
                  Acc& acc) {
  // slots: 0 f4(th4)  1 F(th1)  2 F(th5)  3 F(th6)  4 f5(phi3)  5 f5(phi4)
  const int b4[4] = {MB_P_COAX_T4_TH0, MB_P_COAX_T1_TH0, MB_P_COAX_T5_TH0, MB_P_COAX_T6_TH0};
  const T pi = Consts<T>::pi();
  const bool d2form = (form == MB_COAX_DNA2);
  T e = 0, fr = 0, dfr = 0, rb = 0;
  T f[6], df[6], th[4], dth[4], cphi[2];
  V3<T> sh = v3<T>(0, 0, 0), bh = sh;
  bool nz = false;
  #pragma unroll
  for (int k = 0; k < 6; ++k) {
    f[k] = T(1);
    df[k] = T(0);
  }
  #pragma unroll
  for (int k = 0; k < 4; ++k) th[k] = dth[k] = T(0);
  cphi[0] = cphi[1] = T(0);
  if (act) {
    fr = f2_val(rs, P + MB_P_COAX_RLOW, dfr);
    if (fr != T(0)) {
      sh = (T(1) / rs) * ds;
      th[0] = acos_c(dot(a3i, a3j), dth[0]);
      f[0] = f4_val(th[0], P + b4[0], df[0]);
      th[1] = acos_c(-dot(a1i, a1j), dth[1]);
      T d2;
      f[1] = f4_val(th[1], P + b4[1], df[1]);
      if (d2form) {
        f[1] += f6_val(th[1], P + MB_P_COAX_F6_A, d2);
        df[1] += d2;
      } else {
        f[1] += f4_val(T(2) * pi - th[1], P + b4[1], d2);
        df[1] -= d2;
      }
      th[2] = acos_c(dot(a3i, sh), dth[2]);
      f[2] = f4_val(th[2], P + b4[2], df[2]) + f4_val(pi - th[2], P + b4[2], d2);
      df[2] -= d2;
      th[3] = acos_c(-dot(a3j, sh), dth[3]);
      f[3] = f4_val(th[3], P + b4[3], df[3]) + f4_val(pi - th[3], P + b4[3], d2);
      df[3] -= d2;
      if (!d2form) {
        rb = sqrt(dot(db, db));
        bh = (T(1) / rb) * db;
        cphi[0] = dot(sh, cross(bh, a1j));
        cphi[1] = dot(sh, cross(bh, a1i));
        f[4] = f5_val(cphi[0], P + MB_P_COAX_PHI3_XSTAR, df[4]);
        f[5] = f5_val(cphi[1], P + MB_P_COAX_PHI4_XSTAR, df[5]);
      }
      T prod = fr;
      #pragma unroll
      for (int k = 0; k < 6; ++k) prod *= f[k];
      if (prod != T(0)) {
        nz = true;
        e = prod;
      }
    }
  }
  if (!(WF || WP)) return e;
  T oth[6], othr;
  {
    T pre = nz ? cot : T(0);
    T suf[7];
    suf[6] = T(1);
    #pragma unroll
    for (int k = 5; k >= 0; --k) suf[k] = suf[k + 1] * f[k];
    othr = pre * suf[0];
    pre *= fr;
    #pragma unroll
    for (int k = 0; k < 6; ++k) {
      oth[k] = pre * suf[k + 1];
      pre *= f[k];
    }
  }
  if (WF && nz) {
    V3<T> gsh = v3<T>(0, 0, 0), gbh = gsh;
    axpy(G.ds, othr * dfr, sh);
    T gx = oth[0] * df[0] * dth[0];  // a3i . a3j
    axpy(G.a3i, gx, a3j);
    axpy(G.a3j, gx, a3i);
    gx = oth[1] * df[1] * dth[1];  // -a1i . a1j
    axpy(G.a1i, -gx, a1j);
    axpy(G.a1j, -gx, a1i);
    gx = oth[2] * df[2] * dth[2];  // a3i . sh
    dot_unit_grad(gx, T(1), a3i, sh, G.a3i, gsh);
    gx = oth[3] * df[3] * dth[3];  // -a3j . sh
    dot_unit_grad(gx, T(-1), a3j, sh, G.a3j, gsh);
    if (!d2form) {
      // x = sh . (bh x a) = a . (sh x bh) = bh . (a x sh)
      const V3<T> sxb = cross(sh, bh);
      gx = oth[4] * df[4];
      axpy(gsh, gx, cross(bh, a1j));
      axpy(gbh, gx, cross(a1j, sh));
      axpy(G.a1j, gx, sxb);
      gx = oth[5] * df[5];
      axpy(gsh, gx, cross(bh, a1i));
      axpy(gbh, gx, cross(a1i, sh));
      axpy(G.a1i, gx, sxb);
      G.db = G.db + unit_to_vec_grad(gbh, bh, rb);
    }
    G.ds = G.ds + unit_to_vec_grad(gsh, sh, rs);
  }
  if (WP) {
    f2_par(rs, P + MB_P_COAX_RLOW, othr, bank, MB_P_COAX_RLOW, acc);
    T g[4] = {T(0), T(0), T(0), T(0)};
    if (nz) f4_par_add(th[0], P + b4[0], oth[0], g);
    f4_par_flush(g, bank, b4[0], acc);
    g[0] = g[1] = g[2] = g[3] = T(0);
    T g6a = 0, g6b = 0;
    if (nz) {
      f4_par_add(th[1], P + b4[1], oth[1], g);
      if (d2form) {
        if (th[1] >= P[MB_P_COAX_F6_B]) {
          const T t = th[1] - P[MB_P_COAX_F6_B];
          g6a = oth[1] * T(0.5) * t * t;
          g6b = -oth[1] * P[MB_P_COAX_F6_A] * t;
        }
      } else {
        f4_par_add(T(2) * pi - th[1], P + b4[1], oth[1], g);
      }
    }
    f4_par_flush(g, bank, b4[1], acc);
    {
      T g2[2][4] = {{T(0), T(0), T(0), T(0)}, {T(0), T(0), T(0), T(0)}};
      #pragma unroll
      for (int k = 2; k < 4; ++k) {
        if (nz) {
          f4_par_add(th[k], P + b4[k], oth[k], g2[k - 2]);
          f4_par_add(pi - th[k], P + b4[k], oth[k], g2[k - 2]);
        }
      }
      f4_par_flush2(g2[0], b4[2], g2[1], b4[3], bank, acc);
    }
    f5_par(cphi[0], P + MB_P_COAX_PHI3_XSTAR, (nz && !d2form) ? oth[4] : T(0), bank, MB_P_COAX_PHI3_XSTAR, acc);
    f5_par(cphi[1], P + MB_P_COAX_PHI4_XSTAR, (nz && !d2form) ? oth[5] : T(0), bank, MB_P_COAX_PHI4_XSTAR, acc);
    acc.add(bank, MB_P_COAX_F6_A, g6a);
    acc.add(bank, MB_P_COAX_F6_B, g6b);
  }
  return e;
}

// ---------------------------------------------------------------------------------------------------------
// Debye-Hueckel on the backbone-site distance, times the end-charge multiplier m.
template <class T, bool WF, bool WP, class Acc>
MB_HD T debye_term(const T* P, int bank, bool act, const V3<T>& d, T m, T cot, V3<T>& gd, Acc& acc) {
  T e = 0;
  T g_k = 0, g_A = 0, g_S = 0, g_rc = 0;
  if (act) {
    const T r2 = dot(d, d);
    const T rc = P[MB_P_DEBYE_RCUT];
    if (r2 < rc * rc && r2 > T(0)) {
      const T ir = inv_sqrt(r2);  // one reciprocal square root instead of a square root and two divisions
      const T r = r2 * ir;
      T dEdr;
      if (r < P[MB_P_DEBYE_RHIGH]) {
        const T ex = exp(-P[MB_P_DEBYE_KAPPA] * r) * ir;
        e = m * P[MB_P_DEBYE_PREF] * ex;
        dEdr = -e * (P[MB_P_DEBYE_KAPPA] + ir);
        if (WP) {
          g_k = cot * (-r * e);
          g_A = cot * m * ex;
        }
      } else {
        const T t = r - rc;
        e = m * P[MB_P_DEBYE_SMOOTH] * t * t;
        dEdr = T(2) * m * P[MB_P_DEBYE_SMOOTH] * t;
        if (WP) {
          g_S = cot * m * t * t;
          g_rc = cot * (-dEdr);
        }
      }
      if (WF) axpy(gd, cot * dEdr * ir, d);
    }
  }
  if (WP) {
    const int idx[4] = {MB_P_DEBYE_KAPPA, MB_P_DEBYE_PREF, MB_P_DEBYE_SMOOTH, MB_P_DEBYE_RCUT};
    const T val[4] = {g_k, g_A, g_S, g_rc};
    acc_add_group(acc, bank, idx, val);
  }
  return e;
}

// no-op accumulator for the passes that do not need dE/dparams
struct NullAcc {
  template <class T>
  MB_HD void add(int, int, T) {}
  template <class T>
  MB_HD void add_scatter(int, int, T, bool) {}
};

// ---------------------------------------------------------------------------------------------------------
// Whole-pair drivers.  `M` carries the model forms and both flavours' geometry; `P` is the full parameter
// array (n_banks * MB_P_COUNT).  e[] accumulates the per-term energies (unweighted); Gi/Gj the gradients
// of sum_t cot[t] * E_t.
template <class T>
struct ModelT {
  Geom<T> geom[2];
  mb_bank_forms forms[MB_MAX_BANKS];
  T box[3];
  int n_banks, half_charged_ends;
  MB_HD void load(const mb_model& m) {
    geom[0].load(m.geom[0]);
    geom[1].load(m.geom[1]);
    for (int b = 0; b < MB_MAX_BANKS; ++b) forms[b] = m.forms[b];
    #pragma unroll
    for (int k = 0; k < 3; ++k) box[k] = T(m.box[k]);
    n_banks = m.n_banks;
    half_charged_ends = m.half_charged_ends;
  }
};

template <class T, bool WF, bool WP, class Acc>
MB_HD void bonded_pair(const ModelT<T>& M, const T* Pall, bool valid, const Nuc<T>& ni, const Nuc<T>& nj, int seq_i,
                       int seq_j, int nt_i, int nt_j, int snt_i, int snt_j, unsigned mask, const T* cot, T e[MB_N_TERMS],
                       NucGrad<T>& Gi, NucGrad<T>& Gj, Acc& acc, bool has_w_ext = false, T w_ext = T(0)) {
  // bank / flavour selection, mythos/energy/na1/fene.py:96, na1/stacking.py:201: RNA bank iff both RNA
  int bank = 0, fl = 0, sbank = 0, sfl = 0;
  if (M.n_banks > 1) {
    if (nt_i == 2 && nt_j == 2) bank = fl = 1;
    if (snt_i == 2 && snt_j == 2) sbank = sfl = 1;
  }
  const Geom<T>& g = M.geom[fl];
  const T* P = Pall + bank * MB_P_COUNT;
  const V3<T> back_i = site(ni, g.back[0], g.back[1], g.back[2]), back_j = site(nj, g.back[0], g.back[1], g.back[2]);
  const V3<T> base_i = site(ni, g.base, T(0), T(0)), base_j = site(nj, g.base, T(0), T(0));

  if (mask & (1u << MB_TERM_FENE)) {
    const V3<T> d = disp(back_i, back_j, M.box);
    const T r = sqrt(dot(d, d));
    T dEdr;
    e[MB_TERM_FENE] += fene_term<T, WP>(P, bank, valid, r, cot[MB_TERM_FENE], dEdr, acc);
    if (WF && valid) {
      const V3<T> gs = (cot[MB_TERM_FENE] * dEdr / r) * d;
      site_grad(Gi, T(1), gs, g.back[0], g.back[1], g.back[2]);
      site_grad(Gj, T(-1), gs, g.back[0], g.back[1], g.back[2]);
    }
  }
  if (mask & (1u << MB_TERM_BEXC)) {
    const T c = cot[MB_TERM_BEXC];
    V3<T> gs = v3<T>(0, 0, 0);
    e[MB_TERM_BEXC] += exc_site<T, WF, WP>(P, bank, MB_P_BEXC_BASE_RSTAR, MB_P_BEXC_EPS, valid,
                                           disp(base_i, base_j, M.box), c, gs, acc);
    if (WF) {
      site_grad(Gi, T(1), gs, g.base, T(0), T(0));
      site_grad(Gj, T(-1), gs, g.base, T(0), T(0));
    }
    gs = v3<T>(0, 0, 0);
    e[MB_TERM_BEXC] += exc_site<T, WF, WP>(P, bank, MB_P_BEXC_BACK_BASE_RSTAR, MB_P_BEXC_EPS, valid,
                                           disp(back_i, base_j, M.box), c, gs, acc);
    if (WF) {
      site_grad(Gi, T(1), gs, g.back[0], g.back[1], g.back[2]);
      site_grad(Gj, T(-1), gs, g.base, T(0), T(0));
    }
    gs = v3<T>(0, 0, 0);
    e[MB_TERM_BEXC] += exc_site<T, WF, WP>(P, bank, MB_P_BEXC_BASE_BACK_RSTAR, MB_P_BEXC_EPS, valid,
                                           disp(base_i, back_j, M.box), c, gs, acc);
    if (WF) {
      site_grad(Gi, T(1), gs, g.base, T(0), T(0));
      site_grad(Gj, T(-1), gs, g.back[0], g.back[1], g.back[2]);
    }
  }
  if (mask & (1u << MB_TERM_STACK)) {
    const Geom<T>& sg = M.geom[sfl];
    const T* SP = Pall + sbank * MB_P_COUNT;
    const int form = M.forms[sbank].stack_form;
    const bool rna = (form == MB_STACK_RNA);
    // site coefficients: RNA form uses stack5_i / stack3_j, DNA form the stacking site on both
    const T si1 = rna ? sg.stack5[0] : sg.stack, si2 = rna ? sg.stack5[1] : T(0);
    const T sj1 = rna ? sg.stack3[0] : sg.stack, sj2 = rna ? sg.stack3[1] : T(0);
    const T b1 = sg.use_back_stack ? sg.back_stack : sg.back[0];
    const T b2 = sg.use_back_stack ? T(0) : sg.back[1];
    const T b3 = sg.use_back_stack ? T(0) : sg.back[2];
    const V3<T> ds = disp(site(ni, si1, si2, T(0)), site(nj, sj1, sj2, T(0)), M.box);
    const V3<T> db = disp(site(ni, b1, b2, b3), site(nj, b1, b2, b3), M.box);
    V3<T> p3j = v3<T>(0, 0, 0), p5i = p3j;
    if (rna) {
      p3j = sg.p3[0] * nj.a1 + sg.p3[1] * nj.a2 + sg.p3[2] * nj.a3;
      p5i = sg.p5[0] * ni.a1 + sg.p5[1] * ni.a2 + sg.p5[2] * ni.a3;
    }
    StackGrad<T> G;
    G.ds = G.db = G.a3i = G.a3j = G.a2i = G.a2j = G.p3j = G.p5i = v3<T>(0, 0, 0);
    e[MB_TERM_STACK] += stack_term<T, WF, WP>(SP, sbank, form, valid, ds, db, ni.a3, nj.a3, ni.a2, nj.a2, p3j, p5i,
                                              seq_i * 4 + seq_j, cot[MB_TERM_STACK], G, acc, has_w_ext, w_ext);
    if (WF && valid) {
      site_grad(Gi, T(1), G.ds, si1, si2, T(0));
      site_grad(Gj, T(-1), G.ds, sj1, sj2, T(0));
      site_grad(Gi, T(1), G.db, b1, b2, b3);
      site_grad(Gj, T(-1), G.db, b1, b2, b3);
      Gi.a3 = Gi.a3 + G.a3i;
      Gj.a3 = Gj.a3 + G.a3j;
      Gi.a2 = Gi.a2 + G.a2i;
      Gj.a2 = Gj.a2 + G.a2j;
      if (rna) {
        axpy(Gj.a1, sg.p3[0], G.p3j);
        axpy(Gj.a2, sg.p3[1], G.p3j);
        axpy(Gj.a3, sg.p3[2], G.p3j);
        axpy(Gi.a1, sg.p5[0], G.p5i);
        axpy(Gi.a2, sg.p5[1], G.p5i);
        axpy(Gi.a3, sg.p5[2], G.p5i);
      }
    }
  }
}

template <class T, bool WF, bool WP, class Acc>
MB_HD void unbonded_pair(const ModelT<T>& M, const T* Pall, bool valid, const Nuc<T>& ni, const Nuc<T>& nj, int seq_i,
                         int seq_j, int nt_i, int nt_j, T end_mult, unsigned mask, const T* cot, T e[MB_N_TERMS],
                         NucGrad<T>& Gi, NucGrad<T>& Gj, Acc& acc, bool has_w_ext = false, T w_ext = T(0)) {
  // bank / flavour selection, mythos/energy/na1/hydrogen_bonding.py:325-359 (same in every unbonded term)
  int bank = 0, fi = 0, fj = 0;
  if (M.n_banks > 1) {
    const bool ri = (nt_i == 2), rj = (nt_j == 2), di = (nt_i == 1), dj = (nt_j == 1);
    if (ri && rj) {
      bank = MB_BANK_RNA;
      fi = fj = 1;
    } else if (di && rj) {
      bank = MB_BANK_DRH;
      fi = 0;
      fj = 1;
    } else if (dj && ri) {
      bank = MB_BANK_DRH;
      fi = 1;
      fj = 0;
    }
  }
  const Geom<T>&gi = M.geom[fi], &gj = M.geom[fj];
  const T* P = Pall + bank * MB_P_COUNT;
  const mb_bank_forms F = M.forms[bank];
  const V3<T> back_i = site(ni, gi.back[0], gi.back[1], gi.back[2]), back_j = site(nj, gj.back[0], gj.back[1], gj.back[2]);
  const V3<T> base_i = site(ni, gi.base, T(0), T(0)), base_j = site(nj, gj.base, T(0), T(0));
  const V3<T> d_base = disp(base_j, base_i, M.box);
  const V3<T> d_back = disp(back_j, back_i, M.box);

  if (mask & (1u << MB_TERM_UEXC)) {
    const T c = cot[MB_TERM_UEXC];
    V3<T> gs = v3<T>(0, 0, 0);
    e[MB_TERM_UEXC] += exc_site<T, WF, WP>(P, bank, MB_P_UEXC_BACKBONE_RSTAR, MB_P_UEXC_EPS, valid, d_back, c, gs, acc);
    if (WF) {
      site_grad(Gj, T(1), gs, gj.back[0], gj.back[1], gj.back[2]);
      site_grad(Gi, T(-1), gs, gi.back[0], gi.back[1], gi.back[2]);
    }
    gs = v3<T>(0, 0, 0);
    e[MB_TERM_UEXC] += exc_site<T, WF, WP>(P, bank, MB_P_UEXC_BASE_RSTAR, MB_P_UEXC_EPS, valid, d_base, c, gs, acc);
    if (WF) {
      site_grad(Gj, T(1), gs, gj.base, T(0), T(0));
      site_grad(Gi, T(-1), gs, gi.base, T(0), T(0));
    }
    gs = v3<T>(0, 0, 0);
    e[MB_TERM_UEXC] += exc_site<T, WF, WP>(P, bank, MB_P_UEXC_BACK_BASE_RSTAR, MB_P_UEXC_EPS, valid,
                                           disp(back_i, base_j, M.box), c, gs, acc);
    if (WF) {
      site_grad(Gi, T(1), gs, gi.back[0], gi.back[1], gi.back[2]);
      site_grad(Gj, T(-1), gs, gj.base, T(0), T(0));
    }
    gs = v3<T>(0, 0, 0);
    e[MB_TERM_UEXC] += exc_site<T, WF, WP>(P, bank, MB_P_UEXC_BASE_BACK_RSTAR, MB_P_UEXC_EPS, valid,
                                           disp(base_i, back_j, M.box), c, gs, acc);
    if (WF) {
      site_grad(Gi, T(1), gs, gi.base, T(0), T(0));
      site_grad(Gj, T(-1), gs, gj.back[0], gj.back[1], gj.back[2]);
    }
  }

  if (mask & ((1u << MB_TERM_HB) | (1u << MB_TERM_CROSS))) {
    const T r2 = dot(d_base, d_base);
    const T ir = (valid && r2 > T(0)) ? inv_sqrt(r2) : T(0);  // one reciprocal square root instead of a square root and a division
    const T r = r2 * ir;
    const bool in_hb = valid && (mask & (1u << MB_TERM_HB)) && P[MB_P_HB_RCLOW] < r && r < P[MB_P_HB_RCHIGH];
    const bool in_cr = valid && (mask & (1u << MB_TERM_CROSS)) && P[MB_P_CROSS_RCLOW] < r && r < P[MB_P_CROSS_RCHIGH];
    V3<T> dh = v3<T>(0, 0, 0);
    if (in_hb || in_cr) dh = ir * d_base;
    HbAngles<T> A;
    A.ready = false;
    HbGrad<T> G;
    G.d = G.a1i = G.a1j = G.a3i = G.a3j = v3<T>(0, 0, 0);
    if (mask & (1u << MB_TERM_HB))
      e[MB_TERM_HB] += hb_term<T, WF, WP>(P, bank, in_hb, r, dh, ni.a1, nj.a1, ni.a3, nj.a3, A, seq_i * 4 + seq_j,
                                          cot[MB_TERM_HB], G, acc, has_w_ext, w_ext);
    if (mask & (1u << MB_TERM_CROSS))
      e[MB_TERM_CROSS] += cross_term<T, WF, WP>(P, bank, F.cross_form, in_cr, r, dh, ni.a1, nj.a1, ni.a3, nj.a3, A,
                                                cot[MB_TERM_CROSS], G, acc);
    if (WF && (in_hb || in_cr)) {
      site_grad(Gj, T(1), G.d, gj.base, T(0), T(0));
      site_grad(Gi, T(-1), G.d, gi.base, T(0), T(0));
      Gi.a1 = Gi.a1 + G.a1i;
      Gj.a1 = Gj.a1 + G.a1j;
      Gi.a3 = Gi.a3 + G.a3i;
      Gj.a3 = Gj.a3 + G.a3j;
    }
  }

  if (mask & (1u << MB_TERM_COAX)) {
    const V3<T> ds = disp(site(nj, gj.stack, T(0), T(0)), site(ni, gi.stack, T(0), T(0)), M.box);
    const T rs = sqrt(dot(ds, ds));
    const bool in = valid && P[MB_P_COAX_RCLOW] < rs && rs < P[MB_P_COAX_RCHIGH];
    CoaxGrad<T> G;
    G.ds = G.db = G.a1i = G.a1j = G.a3i = G.a3j = v3<T>(0, 0, 0);
    e[MB_TERM_COAX] += coax_term<T, WF, WP>(P, bank, F.coax_form, in, ds, rs, d_back, ni.a1, nj.a1, ni.a3, nj.a3,
                                            cot[MB_TERM_COAX], G, acc);
    if (WF && in) {
      site_grad(Gj, T(1), G.ds, gj.stack, T(0), T(0));
      site_grad(Gi, T(-1), G.ds, gi.stack, T(0), T(0));
      site_grad(Gj, T(1), G.db, gj.back[0], gj.back[1], gj.back[2]);
      site_grad(Gi, T(-1), G.db, gi.back[0], gi.back[1], gi.back[2]);
      Gi.a1 = Gi.a1 + G.a1i;
      Gj.a1 = Gj.a1 + G.a1j;
      Gi.a3 = Gi.a3 + G.a3i;
      Gj.a3 = Gj.a3 + G.a3j;
    }
  }

  if ((mask & (1u << MB_TERM_DEBYE)) && F.has_debye) {
    V3<T> gd = v3<T>(0, 0, 0);
    e[MB_TERM_DEBYE] += debye_term<T, WF, WP>(P, bank, valid, d_back, end_mult, cot[MB_TERM_DEBYE], gd, acc);
    if (WF) {
      site_grad(Gj, T(1), gd, gj.back[0], gj.back[1], gj.back[2]);
      site_grad(Gi, T(-1), gd, gi.back[0], gi.back[1], gi.back[2]);
    }
  }
}

}  // namespace mb
