// frame_kernels.cu -- frame-resident energy + dE/dparams kernel (the DiffTRe / EnergyFunction.map shape).
//
// One CTA owns one stored frame.  The frame's (center, quaternion) rows are staged once in shared memory
// (56 B per nucleotide in float64 -> 114 KB at N = 2040, inside the 227 KB a CTA may use on sm_100a), and every
// pair of the frame is evaluated from there; nothing per-pair is ever written back to HBM.  Work is regrouped
// so that each code region runs with (nearly) full warps -- the generic one-thread-per-pair kernel spent 93 % of
// its issue slots waiting on instruction fetch because every lane wandered through a 229 KB instruction stream:
//
//   phase B  bonded pairs (FENE, bonded excluded volume, stacking), one thread per bond
//   phase 1  every listed pair: centre distance, Debye-Hueckel on the backbone sites; parameter gradients in
//            REGISTERS; pairs inside the short-range cutoff are compacted into queue SR (shared memory)
//   phase 2  when SR holds a CTA-full: excluded volume (4 site pairs), gradients in registers; pairs inside the
//            hydrogen-bond / cross-stacking radial window go to queue BP, inside the coaxial window to queue CX
//   phase 3  when BP / CX hold a CTA-full: the six-angle products with dense lanes; their parameter gradients are
//            warp-reduced into the shared-memory bank image
//   flush    partial queues, register accumulators -> bank image -> one J row, per-term energies -> one terms row
//
// Queue order is made deterministic (block prefix over warp ballots), so results are bitwise repeatable.
#include "energy_dev.cuh"

namespace mb {

#ifndef MB_FRAME_THREADS
#define MB_FRAME_THREADS 512
#endif
constexpr int kFB = MB_FRAME_THREADS;  // threads per CTA (one CTA per SM: the frame fills most of shared memory)
constexpr int kFWarps = kFB / 32;
constexpr int kQCap = 2 * kFB;

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

struct FrameSmem {
  // byte offsets into dynamic shared memory
  size_t c, q, p, acc, e, flags, q_sr, q_bp, q_cx, wcnt, ctr, total;
};
template <class T>
__host__ __device__ inline FrameSmem frame_smem_layout(int n, bool wp) {
  FrameSmem L;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 15) & ~size_t(15);
    return o;
  };
  L.c = take(sizeof(T) * 3 * n);
  L.q = take(sizeof(T) * 4 * n);
  L.p = take(sizeof(T) * MB_P_COUNT);
  L.acc = take(wp ? sizeof(T) * MB_P_COUNT : 0);
  L.e = take(sizeof(T) * MB_N_TERMS * kFWarps);
  L.flags = take(n);
  L.q_sr = take(sizeof(uint32_t) * kQCap);
  L.q_bp = take(sizeof(uint32_t) * kQCap);
  L.q_cx = take(sizeof(uint32_t) * kQCap);
  L.wcnt = take(sizeof(int) * (kFWarps + 1));
  L.ctr = take(sizeof(int) * 4);
  L.total = off;
  return L;
}

// deterministic block-wide append: entries keep (warp, lane) order
__device__ __forceinline__ void q_push(uint32_t* q, int* n, int* wcnt, bool pred, uint32_t val) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(kFull, pred);
  if (lane == 0) wcnt[warp] = __popc(m);
  __syncthreads();
  int base = *n;
  for (int w = 0; w < warp; ++w) base += wcnt[w];
  if (pred) q[base + __popc(m & ((1u << lane) - 1u))] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kFWarps; ++w) t += wcnt[w];
    *n += t;
  }
  __syncthreads();
}

template <class T>
__device__ __forceinline__ Nuc<T> smem_nuc(const T* sC, const T* sQ, int i) {
  Nuc<T> n;
  n.c = v3<T>(sC[3 * i], sC[3 * i + 1], sC[3 * i + 2]);
  axes_from_quat(sQ[4 * i], sQ[4 * i + 1], sQ[4 * i + 2], sQ[4 * i + 3], n.a1, n.a2, n.a3);
  return n;
}

template <class T>
__device__ __forceinline__ T block_sum_to(T v, T* dst) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  if ((threadIdx.x & 31) == 0 && v != T(0)) atomicAdd(dst, v);
  return v;
}

template <class T, bool WP>
__global__ void __launch_bounds__(kFB, 1) k_frame_energy(const EnergyDev<T> a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const FrameSmem L = frame_smem_layout<T>(a.n, WP);
  T* sC = reinterpret_cast<T*>(smem + L.c);
  T* sQ = reinterpret_cast<T*>(smem + L.q);
  T* sP = reinterpret_cast<T*>(smem + L.p);
  T* sAcc = reinterpret_cast<T*>(smem + L.acc);
  T* sE = reinterpret_cast<T*>(smem + L.e);
  unsigned char* sF = smem + L.flags;  // bits 0-1 seq, bit 2 is_end
  uint32_t* qSR = reinterpret_cast<uint32_t*>(smem + L.q_sr);
  uint32_t* qBP = reinterpret_cast<uint32_t*>(smem + L.q_bp);
  uint32_t* qCX = reinterpret_cast<uint32_t*>(smem + L.q_cx);
  int* wcnt = reinterpret_cast<int*>(smem + L.wcnt);
  int* ctr = reinterpret_cast<int*>(smem + L.ctr);  // [0] n_sr [1] n_bp [2] n_cx

  const int frame = blockIdx.x;
  const int n = a.n;
  const long long fbase = (long long)frame * n;
  for (int k = threadIdx.x; k < 3 * n; k += kFB) sC[k] = a.center[3 * fbase + k];
  for (int k = threadIdx.x; k < 4 * n; k += kFB) sQ[k] = a.quat[4 * fbase + k];
  for (int k = threadIdx.x; k < MB_P_COUNT; k += kFB) {
    sP[k] = a.params[k];
    if (WP) sAcc[k] = T(0);
  }
  for (int k = threadIdx.x; k < n; k += kFB)
    sF[k] = (unsigned char)((a.seq[k] & 3) | ((a.is_end && a.is_end[k]) ? 4 : 0));
  if (threadIdx.x < 4) ctr[threadIdx.x] = 0;
  __syncthreads();

  const ModelT<T>& M = a.M;
  const Geom<T>& g = M.geom[0];
  const mb_bank_forms F = M.forms[0];
  const unsigned mask = a.mask;
  T cot[MB_N_TERMS];
#pragma unroll
  for (int t = 0; t < MB_N_TERMS; ++t) cot[t] = a.cot ? a.cot[(long long)frame * MB_N_TERMS + t] : T(1);
  T e[MB_N_TERMS];
#pragma unroll
  for (int t = 0; t < MB_N_TERMS; ++t) e[t] = T(0);
  SmemAcc<T> sacc{sAcc, false};
  NullAcc nacc;
  NucGrad<T> G0, G1;  // unused (WF = false) but required by the pair drivers' signatures

  // ---------------------------------------------------------------- phase B: bonded pairs
  if (mask & MB_BONDED_TERMS) {
    for (int base = 0; base < a.n_bonded; base += kFB) {
      const int k = base + threadIdx.x;
      const bool valid = k < a.n_bonded;
      int i = 0, j = 0;
      if (valid) {
        i = a.bonded[2 * k];
        j = a.bonded[2 * k + 1];
      }
      const Nuc<T> ni = smem_nuc(sC, sQ, i), nj = smem_nuc(sC, sQ, j);
      if (WP)
        bonded_pair<T, false, true>(M, sP, valid, ni, nj, sF[i] & 3, sF[j] & 3, 1, 1, 1, 1, mask, cot, e, G0, G1, sacc);
      else
        bonded_pair<T, false, false>(M, sP, valid, ni, nj, sF[i] & 3, sF[j] & 3, 1, 1, 1, 1, mask, cot, e, G0, G1, nacc);
    }
  }

  // ---------------------------------------------------------------- unbonded pairs
  RegAcc<T, MB_P_DEBYE_KAPPA, 5> dacc;
  RegAcc<T, MB_P_UEXC_EPS, 17> xacc;
  dacc.zero();
  xacc.zero();
  const bool want_debye = (mask & (1u << MB_TERM_DEBYE)) && F.has_debye;
  const bool want_sr = (mask & ((1u << MB_TERM_UEXC) | (1u << MB_TERM_HB) | (1u << MB_TERM_CROSS) | (1u << MB_TERM_COAX))) != 0;
  // short-range centre cutoff: the widest site-pair cutoff plus both site offsets
  T sr_cut;
  {
    const T ob = sqrt(g.back[0] * g.back[0] + g.back[1] * g.back[1] + g.back[2] * g.back[2]);
    const T oh = fabs(g.base), os = fabs(g.stack);
    T r = T(0);
    if (mask & (1u << MB_TERM_UEXC)) {
      r = fmax(r, sP[MB_P_UEXC_BACKBONE_RC] + 2 * ob);
      r = fmax(r, sP[MB_P_UEXC_BASE_RC] + 2 * oh);
      r = fmax(r, fmax(sP[MB_P_UEXC_BACK_BASE_RC], sP[MB_P_UEXC_BASE_BACK_RC]) + ob + oh);
    }
    if (mask & (1u << MB_TERM_HB)) r = fmax(r, sP[MB_P_HB_RCHIGH] + 2 * oh);
    if (mask & (1u << MB_TERM_CROSS)) r = fmax(r, sP[MB_P_CROSS_RCHIGH] + 2 * oh);
    if (mask & (1u << MB_TERM_COAX)) r = fmax(r, sP[MB_P_COAX_RCHIGH] + 2 * os);
    sr_cut = r * T(1.000001);
  }
  const T sr_cut2 = sr_cut * sr_cut;
  T bp_lo = T(1e30), bp_hi = T(0);
  if (mask & (1u << MB_TERM_HB)) {
    bp_lo = fmin(bp_lo, sP[MB_P_HB_RCLOW]);
    bp_hi = fmax(bp_hi, sP[MB_P_HB_RCHIGH]);
  }
  if (mask & (1u << MB_TERM_CROSS)) {
    bp_lo = fmin(bp_lo, sP[MB_P_CROSS_RCLOW]);
    bp_hi = fmax(bp_hi, sP[MB_P_CROSS_RCHIGH]);
  }

  // ---- one scheduler loop; every phase body appears exactly once so it is inlined and its accumulators stay
  // in registers.  All conditions are CTA-uniform (queue counters live in shared memory, read after barriers).
  if ((mask & MB_UNBONDED_TERMS) && a.pair_capacity > 0) {
    const int32_t* pl = a.pairs + (long long)frame * a.pair_frame_stride;
    long long count = a.pair_capacity;
    if (a.pair_count) {
      const long long c = a.pair_count[frame];
      count = c < count ? c : count;
    }
    long long base = 0;
    bool flush = false;
    while (true) {
      const int n_sr = ctr[0], n_bp = ctr[1], n_cx = ctr[2];
      __syncthreads();  // everyone has read the counters before anyone updates them
      if (n_bp >= kFB || (flush && n_bp > 0)) {
        // ---------------- phase 3a: hydrogen bonding + cross stacking on the BP queue
        const int cnt = n_bp >= kFB ? kFB : n_bp;
        const int t = threadIdx.x;
        const bool valid = t < cnt;
        const uint32_t pk = valid ? qBP[n_bp - cnt + t] : 0u;
        const int i = pk & 0xffff, j = pk >> 16;
        const Nuc<T> ni = smem_nuc(sC, sQ, i), nj = smem_nuc(sC, sQ, j);
        const V3<T> d = disp(site(nj, g.base, T(0), T(0)), site(ni, g.base, T(0), T(0)), M.box);
        const T r = sqrt(dot(d, d));
        const bool in_hb = valid && (mask & (1u << MB_TERM_HB)) && sP[MB_P_HB_RCLOW] < r && r < sP[MB_P_HB_RCHIGH];
        const bool in_cr = valid && (mask & (1u << MB_TERM_CROSS)) && sP[MB_P_CROSS_RCLOW] < r && r < sP[MB_P_CROSS_RCHIGH];
        const V3<T> dh = (valid && r > T(0)) ? (T(1) / r) * d : v3<T>(0, 0, 0);
        HbAngles<T> A;
        A.ready = false;
        HbGrad<T> HG;
        const int tab = (sF[i] & 3) * 4 + (sF[j] & 3);
        if (mask & (1u << MB_TERM_HB)) {
          if (WP)
            e[MB_TERM_HB] += hb_term<T, false, true>(sP, 0, in_hb, r, dh, ni.a1, nj.a1, ni.a3, nj.a3, A, tab, cot[MB_TERM_HB], HG, sacc);
          else
            e[MB_TERM_HB] += hb_term<T, false, false>(sP, 0, in_hb, r, dh, ni.a1, nj.a1, ni.a3, nj.a3, A, tab, cot[MB_TERM_HB], HG, nacc);
        }
        if (mask & (1u << MB_TERM_CROSS)) {
          if (WP)
            e[MB_TERM_CROSS] += cross_term<T, false, true>(sP, 0, F.cross_form, in_cr, r, dh, ni.a1, nj.a1, ni.a3, nj.a3, A, cot[MB_TERM_CROSS], HG, sacc);
          else
            e[MB_TERM_CROSS] += cross_term<T, false, false>(sP, 0, F.cross_form, in_cr, r, dh, ni.a1, nj.a1, ni.a3, nj.a3, A, cot[MB_TERM_CROSS], HG, nacc);
        }
        if (threadIdx.x == 0) ctr[1] = n_bp - cnt;
        __syncthreads();
        continue;
      }
      if (n_cx >= kFB || (flush && n_cx > 0)) {
        // ---------------- phase 3b: coaxial stacking on the CX queue
        const int cnt = n_cx >= kFB ? kFB : n_cx;
        const int t = threadIdx.x;
        const bool valid = t < cnt;
        const uint32_t pk = valid ? qCX[n_cx - cnt + t] : 0u;
        const int i = pk & 0xffff, j = pk >> 16;
        const Nuc<T> ni = smem_nuc(sC, sQ, i), nj = smem_nuc(sC, sQ, j);
        const V3<T> ds = disp(site(nj, g.stack, T(0), T(0)), site(ni, g.stack, T(0), T(0)), M.box);
        const T rs = sqrt(dot(ds, ds));
        const bool in = valid && sP[MB_P_COAX_RCLOW] < rs && rs < sP[MB_P_COAX_RCHIGH];
        const V3<T> db = disp(site(nj, g.back[0], g.back[1], g.back[2]), site(ni, g.back[0], g.back[1], g.back[2]), M.box);
        CoaxGrad<T> CG;
        if (WP)
          e[MB_TERM_COAX] += coax_term<T, false, true>(sP, 0, F.coax_form, in, ds, rs, db, ni.a1, nj.a1, ni.a3, nj.a3, cot[MB_TERM_COAX], CG, sacc);
        else
          e[MB_TERM_COAX] += coax_term<T, false, false>(sP, 0, F.coax_form, in, ds, rs, db, ni.a1, nj.a1, ni.a3, nj.a3, cot[MB_TERM_COAX], CG, nacc);
        if (threadIdx.x == 0) ctr[2] = n_cx - cnt;
        __syncthreads();
        continue;
      }
      if (n_sr >= kFB || (flush && n_sr > 0)) {
        // ---------------- phase 2: excluded volume on the SR queue, radial windows feed BP / CX
        const int cnt = n_sr >= kFB ? kFB : n_sr;
        const int t = threadIdx.x;
        const bool valid = t < cnt;
        const uint32_t pk = valid ? qSR[n_sr - cnt + t] : 0u;
        const int i = pk & 0xffff, j = pk >> 16;
        const Nuc<T> ni = smem_nuc(sC, sQ, i), nj = smem_nuc(sC, sQ, j);
        const V3<T> back_i = site(ni, g.back[0], g.back[1], g.back[2]), back_j = site(nj, g.back[0], g.back[1], g.back[2]);
        const V3<T> base_i = site(ni, g.base, T(0), T(0)), base_j = site(nj, g.base, T(0), T(0));
        const V3<T> d_base = disp(base_j, base_i, M.box);
        if (mask & (1u << MB_TERM_UEXC)) {
          const T c = cot[MB_TERM_UEXC];
          V3<T> gs;
          T ex = T(0);
          if (WP) {
            ex += exc_site<T, false, true>(sP, 0, MB_P_UEXC_BACKBONE_RSTAR, MB_P_UEXC_EPS, valid, disp(back_j, back_i, M.box), c, gs, xacc);
            ex += exc_site<T, false, true>(sP, 0, MB_P_UEXC_BASE_RSTAR, MB_P_UEXC_EPS, valid, d_base, c, gs, xacc);
            ex += exc_site<T, false, true>(sP, 0, MB_P_UEXC_BACK_BASE_RSTAR, MB_P_UEXC_EPS, valid, disp(back_i, base_j, M.box), c, gs, xacc);
            ex += exc_site<T, false, true>(sP, 0, MB_P_UEXC_BASE_BACK_RSTAR, MB_P_UEXC_EPS, valid, disp(base_i, back_j, M.box), c, gs, xacc);
          } else {
            ex += exc_site<T, false, false>(sP, 0, MB_P_UEXC_BACKBONE_RSTAR, MB_P_UEXC_EPS, valid, disp(back_j, back_i, M.box), c, gs, nacc);
            ex += exc_site<T, false, false>(sP, 0, MB_P_UEXC_BASE_RSTAR, MB_P_UEXC_EPS, valid, d_base, c, gs, nacc);
            ex += exc_site<T, false, false>(sP, 0, MB_P_UEXC_BACK_BASE_RSTAR, MB_P_UEXC_EPS, valid, disp(back_i, base_j, M.box), c, gs, nacc);
            ex += exc_site<T, false, false>(sP, 0, MB_P_UEXC_BASE_BACK_RSTAR, MB_P_UEXC_EPS, valid, disp(base_i, back_j, M.box), c, gs, nacc);
          }
          e[MB_TERM_UEXC] += ex;
        }
        const T r2 = dot(d_base, d_base);
        const bool to_bp = valid && bp_hi > T(0) && r2 > bp_lo * bp_lo && r2 < bp_hi * bp_hi;
        bool to_cx = false;
        if (mask & (1u << MB_TERM_COAX)) {
          const V3<T> ds = disp(site(nj, g.stack, T(0), T(0)), site(ni, g.stack, T(0), T(0)), M.box);
          const T s2 = dot(ds, ds);
          to_cx = valid && s2 > sP[MB_P_COAX_RCLOW] * sP[MB_P_COAX_RCLOW] && s2 < sP[MB_P_COAX_RCHIGH] * sP[MB_P_COAX_RCHIGH];
        }
        if (threadIdx.x == 0) ctr[0] = n_sr - cnt;
        q_push(qBP, &ctr[1], wcnt, to_bp, pk);
        q_push(qCX, &ctr[2], wcnt, to_cx, pk);
        continue;
      }
      if (flush) break;
      if (base >= count) {
        flush = true;
        continue;
      }
      // ---------------- phase 1: one tile of the frame's pair list: Debye-Hueckel, short-range filter
      {
        const long long k = base + threadIdx.x;
        base += kFB;
        int i = 0, j = 0;
        bool valid = k < count;
        if (valid) {
          i = pl[k];
          j = pl[a.pair_capacity + k];
          valid = (i >= 0 && j >= 0 && i < n && j < n);
          if (!valid) i = j = 0;
        }
        const V3<T> ci = v3<T>(sC[3 * i], sC[3 * i + 1], sC[3 * i + 2]), cj = v3<T>(sC[3 * j], sC[3 * j + 1], sC[3 * j + 2]);
        const V3<T> dc = disp(cj, ci, M.box);
        const T d2 = dot(dc, dc);
        if (want_debye) {
          const Nuc<T> ni = smem_nuc(sC, sQ, i), nj = smem_nuc(sC, sQ, j);
          const V3<T> db = disp(site(nj, g.back[0], g.back[1], g.back[2]), site(ni, g.back[0], g.back[1], g.back[2]), M.box);
          T m = T(1);
          if (M.half_charged_ends) m = ((sF[i] & 4) ? T(0.5) : T(1)) * ((sF[j] & 4) ? T(0.5) : T(1));
          V3<T> gd;
          if (WP)
            e[MB_TERM_DEBYE] += debye_term<T, false, true>(sP, 0, valid, db, m, cot[MB_TERM_DEBYE], gd, dacc);
          else
            e[MB_TERM_DEBYE] += debye_term<T, false, false>(sP, 0, valid, db, m, cot[MB_TERM_DEBYE], gd, nacc);
        }
        if (want_sr) q_push(qSR, &ctr[0], wcnt, valid && d2 < sr_cut2, uint32_t(i) | (uint32_t(j) << 16));
      }
    }
  }

  // ---------------------------------------------------------------- flush
  if (WP) {
#pragma unroll
    for (int k = 0; k < 5; ++k) block_sum_to(dacc.r[k], &sAcc[MB_P_DEBYE_KAPPA + k]);
#pragma unroll
    for (int k = 0; k < 17; ++k) block_sum_to(xacc.r[k], &sAcc[MB_P_UEXC_EPS + k]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < MB_N_TERMS; ++t) {
    T v = e[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if (lane == 0) sE[warp * MB_N_TERMS + t] = v;
  }
  __syncthreads();
  if (threadIdx.x < MB_N_TERMS && a.terms) {
    T v = 0;
    for (int w = 0; w < kFWarps; ++w) v += sE[w * MB_N_TERMS + threadIdx.x];
    if (v != T(0)) atomicAdd(&a.terms[(long long)frame * MB_N_TERMS + threadIdx.x], v);
  }
  if (WP) {
    T* out = a.d_params + (long long)frame * a.d_params_frame_stride;
    for (int p = threadIdx.x; p < MB_P_COUNT; p += kFB) {
      const T v = sAcc[p];
      if (v != T(0)) atomicAdd(&out[p], v);
    }
  }
}

template <class T>
bool frame_kernel_eligible(const EnergyDev<T>& a) {
  if (a.M.n_banks != 1 || a.n > 65535) return false;
  return frame_smem_layout<T>(a.n, true).total <= 227 * 1024;
}

template <class T>
int launch_frame_kernel(cudaStream_t s, const EnergyDev<T>& a, bool wp) {
  const size_t smem = frame_smem_layout<T>(a.n, wp).total;
  if (wp) {
    MB_CUDA_CHECK(cudaFuncSetAttribute(k_frame_energy<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_frame_energy<T, true><<<a.n_frames, kFB, smem, s>>>(a);
  } else {
    MB_CUDA_CHECK(cudaFuncSetAttribute(k_frame_energy<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_frame_energy<T, false><<<a.n_frames, kFB, smem, s>>>(a);
  }
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}

template bool frame_kernel_eligible<float>(const EnergyDev<float>&);
template bool frame_kernel_eligible<double>(const EnergyDev<double>&);
template int launch_frame_kernel<float>(cudaStream_t, const EnergyDev<float>&, bool);
template int launch_frame_kernel<double>(cudaStream_t, const EnergyDev<double>&, bool);

}  // namespace mb
