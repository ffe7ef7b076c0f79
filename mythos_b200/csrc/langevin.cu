// langevin.cu -- fused rigid-body Langevin (BAOAB) step, one thread per nucleotide, state kept in registers.
//
// Stands in for the step_fn of jax_md.simulate.nvt_langevin specialised to jax_md.rigid_body.RigidBody
// (third party, jax_md==0.2.28, not vendored in the reference; call site mythos/simulators/jax_md/jaxmd.py:73,82-94).
// The algorithm restated here:
//   B  p_c -= h * dE/dc ;  p_q -= h * dE/dq                       (force = -grad E, h = dt/2)
//   A  c  <- shift(c, h * p_c / m);  quaternion free-rotor splitting (Miller et al. 2002): rotations about the
//      body axes in the order 3,2,1,2,3 with steps h/2,h/2,h,h/2,h/2 and angle zeta_k = step * (p . P_k q)/(4 I_k)
//   O  p_c <- c1 p_c + sqrt(kT (1-c1^2) m) xi;   L = (S(q)^T p)/2 (body-frame angular momentum, 3 components)
//      L_k <- c1' L_k + sqrt(kT (1-c1'^2) I_k) xi_k;  p_q <- 2 S(q) [0, L]
// Noise: counter-based Philox4x32-10 keyed by (seed, step, nucleotide), Box-Muller; or caller-injected normals
// (used by the parity tests, since jax's threefry stream cannot be reproduced).
#include <cmath>

#include "common.cuh"

namespace mb {

struct Philox {
  uint32_t c[4], k[2];
  __device__ __forceinline__ void round_() {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0;
    c[1] = n1;
    c[2] = n2;
    c[3] = n3;
  }
  __device__ __forceinline__ void run() {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round_();
      k[0] += 0x9E3779B9u;
      k[1] += 0xBB67AE85u;
    }
  }
};

// six standard normals for (seed, step, nucleotide)
__device__ __forceinline__ void normals6(uint64_t seed, uint64_t step, uint32_t idx, double z[6]) {
  uint32_t u[8];
  for (int blk = 0; blk < 2; ++blk) {
    Philox p;
    p.c[0] = idx;
    p.c[1] = uint32_t(step);
    p.c[2] = uint32_t(step >> 32);
    p.c[3] = uint32_t(blk);
    p.k[0] = uint32_t(seed);
    p.k[1] = uint32_t(seed >> 32);
    p.run();
    for (int k = 0; k < 4; ++k) u[4 * blk + k] = p.c[k];
  }
  // 3 Box-Muller pairs from 6 of the 8 words (53-bit mantissas are not needed for thermal noise)
  for (int k = 0; k < 3; ++k) {
    const double u1 = (double(u[2 * k]) + 1.0) * (1.0 / 4294967296.0);  // (0,1]
    const double u2 = double(u[2 * k + 1]) * (1.0 / 4294967296.0);
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z[2 * k] = r * c;
    z[2 * k + 1] = r * s;
  }
}

template <class T>
struct LangevinDev {
  int n, phase;
  T* center;
  T* quat;
  T* p_center;
  T* p_quat;
  T* d_center;
  T* d_quat;
  const T* noise;
  int zero_forces;
  T dt, kT, gamma_c, gamma_q, mass, inertia[3], box[3];
  // uniform constants hoisted off the per-nucleotide dependency chain (computed once on the host, in double)
  T inv_mass, inv4I[3], ou_c1, ou_c2, ou_r1, ou_r2[3];
  uint64_t seed, step;
  unsigned long long* step_ptr;        // device-side step counter (CUDA-graph replays), overrides `step`; [1] = block counter
  int advance;                         // add 1 to *step_ptr once every block has read it
  T* traj_center;                      // optional (S,N,3): position after this step is stored at row *step_ptr
  T* traj_quat;
  long long traj_rows;
};

template <class T>
__device__ __forceinline__ void perm(int k, const T q[4], T o[4]) {
  // P_k q = column k of S(q)
  if (k == 1) {
    o[0] = -q[1]; o[1] = q[0]; o[2] = q[3]; o[3] = -q[2];
  } else if (k == 2) {
    o[0] = -q[2]; o[1] = -q[3]; o[2] = q[0]; o[3] = q[1];
  } else {
    o[0] = -q[3]; o[1] = q[2]; o[2] = -q[1]; o[3] = q[0];
  }
}
template <class T>
__device__ __forceinline__ void free_rotor(int k, T step, T inv4Ik, T q[4], T p[4]) {
  T pq[4], pp[4];
  perm(k, q, pq);
  perm(k, p, pp);
  const T zeta = step * (p[0] * pq[0] + p[1] * pq[1] + p[2] * pq[2] + p[3] * pq[3]) * inv4Ik;
  T s, c;
  const T z2 = zeta * zeta;
  if (z2 < T(0.0625)) {
    // |zeta| < 1/4 (a time step rotates a nucleotide by ~1e-2 rad): Taylor polynomials, truncation error < 1e-21 -- a
    // dozen dependent FMAs instead of the library sincos on the step's longest dependency chain (ten rotations per step)
    s = zeta * (T(1) + z2 * (T(-1.0 / 6) + z2 * (T(1.0 / 120) + z2 * (T(-1.0 / 5040) + z2 * (T(1.0 / 362880) +
        z2 * (T(-1.0 / 39916800) + z2 * T(1.0 / 6227020800)))))));
    c = T(1) + z2 * (T(-0.5) + z2 * (T(1.0 / 24) + z2 * (T(-1.0 / 720) + z2 * (T(1.0 / 40320) + z2 * (T(-1.0 / 3628800) +
        z2 * (T(1.0 / 479001600) + z2 * T(-1.0 / 87178291200)))))));
  } else {
    sincos(zeta, &s, &c);
  }
  for (int a = 0; a < 4; ++a) {
    q[a] = c * q[a] + s * pq[a];
    p[a] = c * p[a] + s * pp[a];
  }
}

template <class T>
__device__ __forceinline__ T shift1(T x, T L) {
  if (L > T(0)) {
    x = fmod(x, L);
    if (x < T(0)) x += L;
  }
  return x;
}

template <class T>
__device__ __forceinline__ void langevin_body(const LangevinDev<T>& a, int i, unsigned long long step_now);

template <class T>
__global__ void k_langevin(LangevinDev<T> a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long step_now = a.step_ptr ? *a.step_ptr : a.step;
  if (i < a.n) langevin_body(a, i, step_now);
  if (a.step_ptr && a.advance) {
    // the last block to finish bumps the step counter (every block has read it by then) and re-arms the block counter
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      unsigned* done = reinterpret_cast<unsigned*>(a.step_ptr + 1);
      if (atomicAdd(done, 1u) == gridDim.x - 1) {
        *done = 0u;
        *a.step_ptr = step_now + 1ull;
      }
    }
  }
}

template <class T>
__device__ __forceinline__ void langevin_body(const LangevinDev<T>& a, int i, unsigned long long step_now) {
  T c[3], q[4], pc[3], pq[4];
  for (int d = 0; d < 3; ++d) {
    c[d] = a.center[3 * i + d];
    pc[d] = a.p_center[3 * i + d];
  }
  for (int d = 0; d < 4; ++d) {
    q[d] = a.quat[4 * i + d];
    pq[d] = a.p_quat[4 * i + d];
  }
  const T h = T(0.5) * a.dt;
  // B: phase 0 and 1 kick by dt/2, phase 2 (closing kick of the previous step + opening kick of this one) by dt
  const T kick = (a.phase == 2) ? a.dt : h;
  for (int d = 0; d < 3; ++d) pc[d] -= kick * a.d_center[3 * i + d];
  for (int d = 0; d < 4; ++d) pq[d] -= kick * a.d_quat[4 * i + d];
  if (a.zero_forces) {  // hand the gradient buffers back zeroed: the next force evaluation accumulates into them
    for (int d = 0; d < 3; ++d) a.d_center[3 * i + d] = T(0);
    for (int d = 0; d < 4; ++d) a.d_quat[4 * i + d] = T(0);
  }
  if (a.phase != 1) {
    for (int half = 0; half < 2; ++half) {
      // A(dt/2)
      for (int d = 0; d < 3; ++d) c[d] = shift1(c[d] + h * pc[d] * a.inv_mass, a.box[d]);
      free_rotor(3, T(0.5) * h, a.inv4I[2], q, pq);
      free_rotor(2, T(0.5) * h, a.inv4I[1], q, pq);
      free_rotor(1, h, a.inv4I[0], q, pq);
      free_rotor(2, T(0.5) * h, a.inv4I[1], q, pq);
      free_rotor(3, T(0.5) * h, a.inv4I[2], q, pq);
      if (half == 0) {
        // O(dt)
        double z[6];
        if (a.noise) {
          for (int k = 0; k < 6; ++k) z[k] = double(a.noise[6 * i + k]);
        } else {
          normals6(a.seed, (uint64_t)step_now, uint32_t(i), z);
        }
        const T c1 = a.ou_c1, c2 = a.ou_c2;
        for (int d = 0; d < 3; ++d) pc[d] = c1 * pc[d] + c2 * T(z[d]);
        const T r1 = a.ou_r1;
        T L[3];
        for (int k = 1; k <= 3; ++k) {
          T pk[4];
          perm(k, q, pk);
          const T Lk = T(0.5) * (pq[0] * pk[0] + pq[1] * pk[1] + pq[2] * pk[2] + pq[3] * pk[3]);
          L[k - 1] = r1 * Lk + a.ou_r2[k - 1] * T(z[2 + k]);
        }
        // p = 2 S(q) [0, L] = 2 sum_k L_k P_k q
        for (int d = 0; d < 4; ++d) pq[d] = T(0);
        for (int k = 1; k <= 3; ++k) {
          T pk[4];
          perm(k, q, pk);
          for (int d = 0; d < 4; ++d) pq[d] += T(2) * L[k - 1] * pk[d];
        }
      }
    }
    for (int d = 0; d < 3; ++d) a.center[3 * i + d] = c[d];
    for (int d = 0; d < 4; ++d) a.quat[4 * i + d] = q[d];
    if (a.traj_center && a.step_ptr) {
      const long long row = (long long)step_now;
      if (row < a.traj_rows) {
        for (int d = 0; d < 3; ++d) a.traj_center[(row * a.n + i) * 3 + d] = c[d];
        for (int d = 0; d < 4; ++d) a.traj_quat[(row * a.n + i) * 4 + d] = q[d];
      }
    }
  }
  for (int d = 0; d < 3; ++d) a.p_center[3 * i + d] = pc[d];
  for (int d = 0; d < 4; ++d) a.p_quat[4 * i + d] = pq[d];
}

template <class T>
static int langevin_impl(cudaStream_t s, const mb_langevin_args* x) {
  MB_REQUIRE(x && x->n > 0, MB_EINVAL_SHAPE, "langevin: bad n");
  MB_REQUIRE(x->center && x->quat && x->p_center && x->p_quat && x->d_center && x->d_quat, MB_EINVAL_SHAPE,
             "langevin: missing state buffers");
  MB_REQUIRE(x->phase >= 0 && x->phase <= 2, MB_EINVAL_SHAPE, "langevin: phase must be 0, 1 or 2");
  MB_REQUIRE(x->mass > 0 && x->inertia[0] > 0 && x->inertia[1] > 0 && x->inertia[2] > 0, MB_EINVAL_SHAPE,
             "langevin: mass and inertia must be positive");
  LangevinDev<T> a;
  a.n = x->n;
  a.phase = x->phase;
  a.center = static_cast<T*>(x->center);
  a.quat = static_cast<T*>(x->quat);
  a.p_center = static_cast<T*>(x->p_center);
  a.p_quat = static_cast<T*>(x->p_quat);
  a.d_center = static_cast<T*>(const_cast<void*>(x->d_center));
  a.d_quat = static_cast<T*>(const_cast<void*>(x->d_quat));
  a.zero_forces = x->zero_forces;
  a.noise = static_cast<const T*>(x->noise);
  a.dt = T(x->dt);
  a.kT = T(x->kT);
  a.gamma_c = T(x->gamma_center);
  a.gamma_q = T(x->gamma_quat);
  a.mass = T(x->mass);
  for (int d = 0; d < 3; ++d) {
    a.inertia[d] = T(x->inertia[d]);
    a.box[d] = T(x->box[d]);
  }
  {
    const double c1 = exp(-x->gamma_center * x->dt), r1 = exp(-x->gamma_quat * x->dt);
    a.inv_mass = T(1.0 / x->mass);
    a.ou_c1 = T(c1);
    a.ou_c2 = T(sqrt(x->kT * (1.0 - c1 * c1) * x->mass));
    a.ou_r1 = T(r1);
    for (int d = 0; d < 3; ++d) {
      a.inv4I[d] = T(1.0 / (4.0 * x->inertia[d]));
      a.ou_r2[d] = T(sqrt(x->kT * (1.0 - r1 * r1) * x->inertia[d]));
    }
  }
  a.seed = x->seed;
  a.step = x->step;
  a.step_ptr = static_cast<unsigned long long*>(const_cast<void*>(x->step_ptr));
  a.advance = x->advance_step;
  a.traj_center = static_cast<T*>(x->traj_center);
  a.traj_quat = static_cast<T*>(x->traj_quat);
  a.traj_rows = x->traj_rows;
  MB_REQUIRE(!x->traj_center || (x->traj_quat && x->step_ptr), MB_EINVAL_SHAPE, "langevin: trajectory output needs traj_quat and step_ptr");
  k_langevin<T><<<ceil_div(x->n, 128), 128, 0, s>>>(a);
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// Adjoint (vector-Jacobian product) of one step: SURVEY 8f rank 4 -- what jax.grad through `step_fn` inside
// `checkpoint_scan` computes (mythos/simulators/jax_md/utils.py:174-193, jaxmd.py:54-58,94), for the integrator part.
// The step is   p1 = p0 - kick G ;  A(h) ;  O(dt, xi) ;  A(h)   with G = (dE/dc, dE/dq)(x0) an INPUT here: the kernel
// returns the cotangents of (c0, q0, p_c0, p_q0) through the integrator and the cotangent of G (= -kick * lambda_p1); the
// caller adds the force's own dependence on (x0, theta) with two displaced force evaluations (mythos_b200.simulators.adjoint).
// One thread per nucleotide: the forward sub-steps are recomputed from the saved pre-step state, then undone one by one --
// a free-rotor rotation conserves p.P_k q, so its angle can be recovered from the rotated state and nothing is stored.
template <class T>
struct LangevinAdjDev {
  int n, phase;
  const T* center;
  const T* quat;
  const T* p_center;
  const T* p_quat;
  const T* d_center;
  const T* d_quat;
  const T* noise;
  T* lam_center;   // in: cotangent of the post-step state; out: of the pre-step state
  T* lam_quat;
  T* lam_p_center;
  T* lam_p_quat;
  T* lam_force_center;  // out (N,3): cotangent of dE/dcenter
  T* lam_force_quat;    // out (N,4)
  T dt, inv_mass, inv4I[3], ou_c1, ou_c2, ou_r1, ou_r2[3];
  uint64_t seed, step;
};

template <class T>
__device__ __forceinline__ T rotor_angle(int k, T step, T inv4Ik, const T q[4], const T p[4]) {
  T pq[4];
  perm(k, q, pq);
  return step * (p[0] * pq[0] + p[1] * pq[1] + p[2] * pq[2] + p[3] * pq[3]) * inv4Ik;
}
// (q, p) hold the state AFTER the rotation on entry and BEFORE it on exit; (lq, lp) the cotangents likewise
template <class T>
__device__ __forceinline__ void free_rotor_adjoint(int k, T step, T inv4Ik, T q[4], T p[4], T lq[4], T lp[4]) {
  const T zeta = rotor_angle(k, step, inv4Ik, q, p);  // conserved by the rotation
  T s, c;
  sincos(zeta, &s, &c);
  T Pq[4], Pp[4];
  perm(k, q, Pq);
  perm(k, p, Pp);
  for (int a = 0; a < 4; ++a) {  // undo: rotate by -zeta
    q[a] = c * q[a] - s * Pq[a];
    p[a] = c * p[a] - s * Pp[a];
  }
  perm(k, q, Pq);
  perm(k, p, Pp);
  T gz = T(0);
  for (int a = 0; a < 4; ++a) gz += lq[a] * (-s * q[a] + c * Pq[a]) + lp[a] * (-s * p[a] + c * Pp[a]);
  T Plq[4], Plp[4];
  perm(k, lq, Plq);
  perm(k, lp, Plp);
  const T w = gz * step * inv4Ik;
  for (int a = 0; a < 4; ++a) {  // P_k^T = -P_k
    const T nq = c * lq[a] - s * Plq[a] - w * Pp[a];
    const T np = c * lp[a] - s * Plp[a] + w * Pq[a];
    lq[a] = nq;
    lp[a] = np;
  }
}
template <class T>
__device__ __forceinline__ void drift_rotations(const LangevinAdjDev<T>& a, T h, T q[4], T p[4]) {
  free_rotor(3, T(0.5) * h, a.inv4I[2], q, p);
  free_rotor(2, T(0.5) * h, a.inv4I[1], q, p);
  free_rotor(1, h, a.inv4I[0], q, p);
  free_rotor(2, T(0.5) * h, a.inv4I[1], q, p);
  free_rotor(3, T(0.5) * h, a.inv4I[2], q, p);
}
template <class T>
__device__ __forceinline__ void drift_rotations_adjoint(const LangevinAdjDev<T>& a, T h, T q[4], T p[4], T lq[4], T lp[4]) {
  free_rotor_adjoint(3, T(0.5) * h, a.inv4I[2], q, p, lq, lp);
  free_rotor_adjoint(2, T(0.5) * h, a.inv4I[1], q, p, lq, lp);
  free_rotor_adjoint(1, h, a.inv4I[0], q, p, lq, lp);
  free_rotor_adjoint(2, T(0.5) * h, a.inv4I[1], q, p, lq, lp);
  free_rotor_adjoint(3, T(0.5) * h, a.inv4I[2], q, p, lq, lp);
}

template <class T>
__global__ void k_langevin_adjoint(LangevinAdjDev<T> a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const T h = T(0.5) * a.dt;
  const T kick = (a.phase == 2) ? a.dt : h;
  T q[4], pq[4], pc[3];
  for (int d = 0; d < 4; ++d) {
    q[d] = a.quat[4 * i + d];
    pq[d] = a.p_quat[4 * i + d] - kick * a.d_quat[4 * i + d];
  }
  for (int d = 0; d < 3; ++d) pc[d] = a.p_center[3 * i + d] - kick * a.d_center[3 * i + d];
  // ---- forward recompute: A, O (keeping what O's adjoint needs), A
  drift_rotations(a, h, q, pq);  // (q1, pq2)
  T q1[4], pq2[4], Lp[3];
  for (int d = 0; d < 4; ++d) {
    q1[d] = q[d];
    pq2[d] = pq[d];
  }
  double z[6];
  if (a.noise) {
    for (int k = 0; k < 6; ++k) z[k] = double(a.noise[6 * i + k]);
  } else {
    normals6(a.seed, a.step, uint32_t(i), z);
  }
  for (int k = 1; k <= 3; ++k) {
    T pk[4];
    perm(k, q1, pk);
    const T Lk = T(0.5) * (pq2[0] * pk[0] + pq2[1] * pk[1] + pq2[2] * pk[2] + pq2[3] * pk[3]);
    Lp[k - 1] = a.ou_r1 * Lk + a.ou_r2[k - 1] * T(z[2 + k]);
  }
  for (int d = 0; d < 4; ++d) pq[d] = T(0);
  for (int k = 1; k <= 3; ++k) {
    T pk[4];
    perm(k, q1, pk);
    for (int d = 0; d < 4; ++d) pq[d] += T(2) * Lp[k - 1] * pk[d];
  }
  drift_rotations(a, h, q, pq);  // (q2, pq4): the post-step state
  // ---- backward
  T lc[3], lq[4], lpc[3], lpq[4];
  for (int d = 0; d < 3; ++d) {
    lc[d] = a.lam_center[3 * i + d];
    lpc[d] = a.lam_p_center[3 * i + d];
  }
  for (int d = 0; d < 4; ++d) {
    lq[d] = a.lam_quat[4 * i + d];
    lpq[d] = a.lam_p_quat[4 * i + d];
  }
  // A2^T: c2 = c1 + h pc2 / m (the periodic shift has unit derivative)
  for (int d = 0; d < 3; ++d) lpc[d] += h * a.inv_mass * lc[d];
  drift_rotations_adjoint(a, h, q, pq, lq, lpq);  // -> cotangents of (q1, pq3); (q, pq) are (q1, pq3) again
  // O^T
  for (int d = 0; d < 3; ++d) lpc[d] *= a.ou_c1;
  T lL[3];
  T lq_add[4] = {T(0), T(0), T(0), T(0)}, lpq2[4] = {T(0), T(0), T(0), T(0)};
  for (int k = 1; k <= 3; ++k) {
    T Pq1[4], Plp[4], Ppq2[4];
    perm(k, q1, Pq1);
    perm(k, lpq, Plp);
    perm(k, pq2, Ppq2);
    T dotv = T(0);
    for (int d = 0; d < 4; ++d) dotv += Pq1[d] * lpq[d];
    lL[k - 1] = a.ou_r1 * T(2) * dotv;                       // dL/dL_k through L'_k = r1 L_k + ...
    for (int d = 0; d < 4; ++d) {
      lq_add[d] += -T(2) * Lp[k - 1] * Plp[d];               // pq3 = 2 sum L'_k P_k q1
      lq_add[d] += -T(0.5) * lL[k - 1] * Ppq2[d];            // L_k = pq2 . P_k q1 / 2
      lpq2[d] += T(0.5) * lL[k - 1] * Pq1[d];
    }
  }
  for (int d = 0; d < 4; ++d) {
    lq[d] += lq_add[d];
    lpq[d] = lpq2[d];
    q[d] = q1[d];
    pq[d] = pq2[d];
  }
  // A1^T
  for (int d = 0; d < 3; ++d) lpc[d] += h * a.inv_mass * lc[d];
  drift_rotations_adjoint(a, h, q, pq, lq, lpq);  // -> cotangents of (q0, pq1)
  // kick^T: p1 = p0 - kick G
  for (int d = 0; d < 3; ++d) {
    a.lam_center[3 * i + d] = lc[d];
    a.lam_p_center[3 * i + d] = lpc[d];
    a.lam_force_center[3 * i + d] = -kick * lpc[d];
  }
  for (int d = 0; d < 4; ++d) {
    a.lam_quat[4 * i + d] = lq[d];
    a.lam_p_quat[4 * i + d] = lpq[d];
    a.lam_force_quat[4 * i + d] = -kick * lpq[d];
  }
}

template <class T>
static int langevin_adjoint_impl(cudaStream_t s, const mb_langevin_adjoint_args* x) {
  MB_REQUIRE(x && x->n > 0, MB_EINVAL_SHAPE, "langevin_adjoint: bad n");
  MB_REQUIRE(x->center && x->quat && x->p_center && x->p_quat && x->d_center && x->d_quat, MB_EINVAL_SHAPE,
             "langevin_adjoint: missing pre-step state / forces");
  MB_REQUIRE(x->lam_center && x->lam_quat && x->lam_p_center && x->lam_p_quat && x->lam_force_center && x->lam_force_quat,
             MB_EINVAL_SHAPE, "langevin_adjoint: missing cotangent buffers");
  MB_REQUIRE(x->phase == 0 || x->phase == 2, MB_EINVAL_SHAPE, "langevin_adjoint: phase must be 0 or 2");
  MB_REQUIRE(x->mass > 0 && x->inertia[0] > 0 && x->inertia[1] > 0 && x->inertia[2] > 0, MB_EINVAL_SHAPE,
             "langevin_adjoint: mass and inertia must be positive");
  LangevinAdjDev<T> a;
  a.n = x->n;
  a.phase = x->phase;
  a.center = static_cast<const T*>(x->center);
  a.quat = static_cast<const T*>(x->quat);
  a.p_center = static_cast<const T*>(x->p_center);
  a.p_quat = static_cast<const T*>(x->p_quat);
  a.d_center = static_cast<const T*>(x->d_center);
  a.d_quat = static_cast<const T*>(x->d_quat);
  a.noise = static_cast<const T*>(x->noise);
  a.lam_center = static_cast<T*>(x->lam_center);
  a.lam_quat = static_cast<T*>(x->lam_quat);
  a.lam_p_center = static_cast<T*>(x->lam_p_center);
  a.lam_p_quat = static_cast<T*>(x->lam_p_quat);
  a.lam_force_center = static_cast<T*>(x->lam_force_center);
  a.lam_force_quat = static_cast<T*>(x->lam_force_quat);
  a.dt = T(x->dt);
  const double c1 = exp(-x->gamma_center * x->dt), r1 = exp(-x->gamma_quat * x->dt);
  a.inv_mass = T(1.0 / x->mass);
  a.ou_c1 = T(c1);
  a.ou_c2 = T(sqrt(x->kT * (1.0 - c1 * c1) * x->mass));
  a.ou_r1 = T(r1);
  for (int d = 0; d < 3; ++d) {
    a.inv4I[d] = T(1.0 / (4.0 * x->inertia[d]));
    a.ou_r2[d] = T(sqrt(x->kT * (1.0 - r1 * r1) * x->inertia[d]));
  }
  a.seed = x->seed;
  a.step = x->step;
  k_langevin_adjoint<T><<<ceil_div(x->n, 128), 128, 0, s>>>(a);
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}

}  // namespace mb

extern "C" int mythos_b200_langevin_adjoint_f64(void* stream, const mb_langevin_adjoint_args* a) {
  return mb::langevin_adjoint_impl<double>(static_cast<cudaStream_t>(stream), a);
}
extern "C" int mythos_b200_langevin_adjoint_f32(void* stream, const mb_langevin_adjoint_args* a) {
  return mb::langevin_adjoint_impl<float>(static_cast<cudaStream_t>(stream), a);
}
extern "C" int mythos_b200_langevin_f64(void* stream, const mb_langevin_args* a) {
  return mb::langevin_impl<double>(static_cast<cudaStream_t>(stream), a);
}
extern "C" int mythos_b200_langevin_f32(void* stream, const mb_langevin_args* a) {
  return mb::langevin_impl<float>(static_cast<cudaStream_t>(stream), a);
}
