// list_kernels.cu -- unbonded terms of an explicit pair list of any size: energies, forces (dE/dcenter, dE/dquat)
// and dE/dparams.  This is the large-system shape (8k .. 100k nucleotides, neighbour list from
// mythos_b200_nl_build_*), where a frame no longer fits in shared memory and the frame-resident kernel
// (frame_kernels.cu) does not apply; it also serves the NA1 three-bank model at any size.
//
// The generic one-thread-per-pair kernel (energy_kernels.cu) evaluated every term of a pair in one lane, so 85 % of the
// lanes idled through the short-range code (only ~1 listed pair in 6 is inside the short-range cutoff, ~1 in 30 inside
// the hydrogen-bond window) and the kernel was instruction-fetch bound (profiles/r01_v1_*).  Here the list is regrouped
// so that every code region runs with dense warps, in two kernels with very different shapes:
//
//   k_list_prep    one 64-byte record per nucleotide (centre, backbone site, flags) in the caller's workspace
//   k_list_debye   streaming pass over the whole list at high occupancy (few registers, latency hidden by warps):
//                  centre distance, Debye-Hueckel on the recorded backbone sites (banks that have it), its force as
//                  3+3 reductions into a backbone-site gradient buffer; pairs inside the short-range centre cutoff of
//                  their bank are appended to the short-range list SR in the workspace (~1 listed pair in 6)
//   k_list_sr      one CTA per SM over SR, phase-queued in shared memory:
//                    phase 2  a CTA-full of SR: excluded volume (4 site pairs); pairs inside the hydrogen-bond /
//                             cross-stacking radial window go to queue BP, inside the coaxial-stacking window to queue CX
//                    phase 3  a CTA-full of BP: the six-angle hydrogen-bond + cross-stacking products; a CTA-full of CX:
//                             coaxial stacking
//                    flush    partial queues, per-term energies and the parameter-gradient image -> one atomic per CTA
//   k_list_post    backbone-site gradients -> (dE/dcenter, dE/dquat) of each nucleotide
//
// Each phase calls the same per-pair driver as the generic kernel (oxdna_device.cuh: unbonded_pair) with a
// compile-time term mask, so a phase's instruction stream holds only its own terms and the arithmetic is identical.
// Nucleotide gradients leave the short-range phases as (dE/dcenter, dE/dquat) reductions (RED.ADD).
#include "energy_dev.cuh"

namespace mb {

#ifndef MB_LIST_THREADS
#define MB_LIST_THREADS 256
#endif
#ifndef MB_LIST_MINBLOCKS
#define MB_LIST_MINBLOCKS 2
#endif
constexpr int kLB = MB_LIST_THREADS;  // threads per CTA
constexpr int kLWarps = kLB / 32;
constexpr int kLQCap = 2 * kLB;
#ifndef MB_DEBYE_MINBLOCKS
#define MB_DEBYE_MINBLOCKS 1
#endif
#ifndef MB_DEBYE_UNROLL
#define MB_DEBYE_UNROLL 2
#endif
constexpr int kDB = 256;            // threads per CTA of the Debye / filter pass
constexpr int kDU = MB_DEBYE_UNROLL;  // list entries per thread and step (independent chains)

typedef unsigned long long pk_t;
__device__ __forceinline__ pk_t pk_make(int i, int j) { return (pk_t)(unsigned)i | ((pk_t)(unsigned)j << 32); }

// unordered block-wide append (warp-aggregated shared-memory atomic)
__device__ __forceinline__ void lq_push(pk_t* q, int* n, bool pred, pk_t val) {
  const unsigned m = __ballot_sync(kFull, pred);
  if (m == 0u) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(n, __popc(m));
  base = __shfl_sync(kFull, base, __ffs(m) - 1);
  if (pred) q[base + __popc(m & ((1u << lane) - 1u))] = val;
}

template <class T>
struct BankCuts {
  T sr2, bp_lo2, bp_hi2, cx_lo2, cx_hi2;
  T ev2[4];  // squared excluded-volume site cutoffs (backbone, base, back-base, base-back), formed as exc_site forms them
};

// bank / flavour selection of an unbonded pair, mythos/energy/na1/hydrogen_bonding.py:325-359
__device__ __forceinline__ void pair_bank(int n_banks, int nti, int ntj, int& bank, int& fi, int& fj) {
  bank = fi = fj = 0;
  if (n_banks > 1) {
    const bool ri = (nti == 2), rj = (ntj == 2), di = (nti == 1), dj = (ntj == 1);
    if (ri && rj) {
      bank = MB_BANK_RNA;
      fi = fj = 1;
    } else if (di && rj) {
      bank = MB_BANK_DRH;
      fj = 1;
    } else if (dj && ri) {
      bank = MB_BANK_DRH;
      fi = 1;
    }
  }
}

// Per-nucleotide record of phase 1 (workspace): centre, backbone site of the nucleotide's own flavour, flags
// (bit 2 strand end, bit 3 RNA).  One 16-byte-aligned row -> 128-bit loads.
template <class T>
struct alignas(16) NucRec {
  T c[3], b[3], flags, pad;
};

template <class T>
__global__ void k_list_prep(const EnergyDev<T> a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < a.n_frames) a.sr_count[idx] = 0;
  if (idx >= (long long)a.n * a.n_frames) return;
  const int i = int(idx % a.n);
  T q[4];
  const Nuc<T> nu = load_nuc(a.center, a.quat, idx, q);
  const int rna = (a.M.n_banks > 1 && a.nt_type && a.nt_type[i] == 2) ? 1 : 0;
  const Geom<T>& g = a.M.geom[rna];
  const V3<T> b = site(nu, g.back[0], g.back[1], g.back[2]);
  NucRec<T> r;
  r.c[0] = nu.c.x;
  r.c[1] = nu.c.y;
  r.c[2] = nu.c.z;
  r.b[0] = b.x;
  r.b[1] = b.y;
  r.b[2] = b.z;
  r.flags = T((a.is_end && a.is_end[i] ? 4 : 0) | (rna ? 8 : 0));
  r.pad = T(0);
  reinterpret_cast<NucRec<T>*>(a.rec)[idx] = r;
  if (a.gback) a.gback[3 * idx] = a.gback[3 * idx + 1] = a.gback[3 * idx + 2] = T(0);
}

// backbone-site gradients gathered by phase 1 -> (dE/dcenter, dE/dquat) of the nucleotide
template <class T>
__global__ void k_list_post(const EnergyDev<T> a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.n * a.n_frames) return;
  const V3<T> gb = v3<T>(a.gback[3 * idx], a.gback[3 * idx + 1], a.gback[3 * idx + 2]);
  if (gb.x == T(0) && gb.y == T(0) && gb.z == T(0)) return;
  const int i = int(idx % a.n);
  const int rna = (a.M.n_banks > 1 && a.nt_type && a.nt_type[i] == 2) ? 1 : 0;
  const Geom<T>& g = a.M.geom[rna];
  NucGrad<T> G;
  G.zero();
  site_grad(G, T(1), gb, g.back[0], g.back[1], g.back[2]);
  if (a.d_center) {
    a.d_center[3 * idx] += G.c.x;
    a.d_center[3 * idx + 1] += G.c.y;
    a.d_center[3 * idx + 2] += G.c.z;
  }
  if (a.d_quat) {
    T dq[4];
    quat_grad(G, a.quat[4 * idx], a.quat[4 * idx + 1], a.quat[4 * idx + 2], a.quat[4 * idx + 3], dq);
#pragma unroll
    for (int c = 0; c < 4; ++c) a.d_quat[4 * idx + c] += dq[c];
  }
}

// Debye parameter gradients of up to three banks in registers (bank selected per pair by predication)
template <class T, int NB>
struct DebyeAcc {
  T r[NB][5];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int k = 0; k < 5; ++k) r[b][k] = T(0);
  }
  __device__ __forceinline__ void add(int bank, int idx, T v) {
#pragma unroll
    for (int b = 0; b < NB; ++b) r[b][idx - MB_P_DEBYE_KAPPA] += (NB == 1 || bank == b) ? v : T(0);
  }
  __device__ __forceinline__ void add_scatter(int, int, T, bool) {}
};

template <class T>
__device__ __forceinline__ T list_sr_cut2(const EnergyDev<T>& a, const T* P, bool multi) {
  // short-range centre cutoff of a bank: the widest site-pair cutoff plus both site offsets (the larger flavour)
  const unsigned mask = a.mask;
  T ob = T(0), oh = T(0), os = T(0);
  for (int f = 0; f < (multi ? 2 : 1); ++f) {
    const Geom<T>& g = a.M.geom[f];
    ob = fmax(ob, sqrt(g.back[0] * g.back[0] + g.back[1] * g.back[1] + g.back[2] * g.back[2]));
    oh = fmax(oh, fabs(g.base));
    os = fmax(os, fabs(g.stack));
  }
  T r = T(0);
  if (mask & (1u << MB_TERM_UEXC)) {
    r = fmax(r, P[MB_P_UEXC_BACKBONE_RC] + 2 * ob);
    r = fmax(r, P[MB_P_UEXC_BASE_RC] + 2 * oh);
    r = fmax(r, fmax(P[MB_P_UEXC_BACK_BASE_RC], P[MB_P_UEXC_BASE_BACK_RC]) + ob + oh);
  }
  if (mask & (1u << MB_TERM_HB)) r = fmax(r, P[MB_P_HB_RCHIGH] + 2 * oh);
  if (mask & (1u << MB_TERM_CROSS)) r = fmax(r, P[MB_P_CROSS_RCHIGH] + 2 * oh);
  if (mask & (1u << MB_TERM_COAX)) r = fmax(r, P[MB_P_COAX_RCHIGH] + 2 * os);
  r *= T(1.000001);
  return r * r;
}

// ------------------------------------------------------------------------------------------------------------
// Streaming pass: Debye-Hueckel + short-range filter.  grid = (CTAs, frames), grid-stride over list tiles.
template <class T, bool WF, bool WP, bool MULTI>
__global__ void __launch_bounds__(kDB, MB_DEBYE_MINBLOCKS) k_list_debye(const EnergyDev<T> a) {
  constexpr int NB = MULTI ? MB_MAX_BANKS : 1;
  __shared__ T sD[NB][6];  // kappa, prefactor, smoothing, r_cut, r_high of each bank; [5] = short-range cutoff^2
  __shared__ T sRed[kDB / 32][1 + (WP ? 5 * NB : 0)];
  const int frame = blockIdx.y;
  const int n = a.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long fbase = (long long)frame * n;
  const unsigned mask = a.mask;
  if (threadIdx.x < NB) {
    const T* P = a.params + threadIdx.x * MB_P_COUNT;
#pragma unroll
    for (int k = 0; k < 5; ++k) sD[threadIdx.x][k] = P[MB_P_DEBYE_KAPPA + k];
    sD[threadIdx.x][5] = list_sr_cut2(a, P, MULTI);
  }
  __syncthreads();
  const ModelT<T>& M = a.M;
  bool any_debye = false;
  if (mask & (1u << MB_TERM_DEBYE))
    for (int b = 0; b < NB; ++b) any_debye = any_debye || M.forms[b].has_debye;
  const bool want_sr = (mask & ((1u << MB_TERM_UEXC) | (1u << MB_TERM_HB) | (1u << MB_TERM_CROSS) | (1u << MB_TERM_COAX))) != 0;
  const T cotd = a.cot ? a.cot[(long long)frame * MB_N_TERMS + MB_TERM_DEBYE] : T(1);
  T e = T(0);
  DebyeAcc<T, NB> dacc;
  dacc.zero();
  NullAcc nacc;

  const int32_t* pl = a.pairs + (long long)frame * a.pair_frame_stride;
  long long count = a.pair_capacity;
  if (a.pair_count) {
    const long long c = a.pair_count[frame];
    count = c < count ? c : count;
  }
  const NucRec<T>* rec = reinterpret_cast<const NucRec<T>*>(a.rec) + fbase;
  pk_t* sr_list = a.sr_list + (long long)frame * a.sr_capacity;
  int* sr_count = a.sr_count + frame;
  // support-tagged list: entries [0, split) are the short-range pairs (k_list_sr reads them straight from the list),
  // entries [split, count) the pairs inside the Debye support: this pass evaluates just those, and filters nothing
  const bool tagged = a.tagged != 0;
  const long long first = tagged ? (long long)a.pair_split[frame] : 0;
  constexpr int kTile = kDB * kDU;
  // the index loads of the NEXT tile are issued before the current tile is evaluated: the list streams from HBM
  // (~40 MB at 100k nucleotides), and index -> record -> arithmetic would otherwise be one dependent chain per step
  const long long stride = (long long)gridDim.x * kTile;
  int ni[kDU], nj[kDU];
#pragma unroll
  for (int u = 0; u < kDU; ++u) {
    const long long k = first + (long long)blockIdx.x * kTile + threadIdx.x + u * kDB;
    ni[u] = k < count ? pl[k] : -1;
    nj[u] = k < count ? (pl[a.pair_capacity + k] & 0x1fffffff) : -1;
  }
  for (long long base = first + (long long)blockIdx.x * kTile; base < count; base += stride) {
    int pi[kDU], pj[kDU];
    bool pv[kDU];
#pragma unroll
    for (int u = 0; u < kDU; ++u) {
      pi[u] = ni[u];
      pj[u] = nj[u];
      pv[u] = base + threadIdx.x + u * kDB < count;
      const long long k = base + stride + threadIdx.x + u * kDB;
      ni[u] = k < count ? pl[k] : -1;
      nj[u] = k < count ? (pl[a.pair_capacity + k] & 0x1fffffff) : -1;
    }
    bool sr[kDU];
#pragma unroll
    for (int u = 0; u < kDU; ++u) {
      bool valid = pv[u] && pi[u] >= 0 && pj[u] >= 0 && pi[u] < n && pj[u] < n;
      const int i = valid ? pi[u] : 0, j = valid ? pj[u] : 0;
      pi[u] = i;
      pj[u] = j;
      const NucRec<T> ri = rec[i];
      const NucRec<T> rj = rec[j];
      const int fli = int(ri.flags), flj = int(rj.flags);
      int bank = 0;
      if (MULTI) {
        const bool rna_i = (fli & 8) != 0, rna_j = (flj & 8) != 0;
        bank = (rna_i && rna_j) ? MB_BANK_RNA : ((rna_i != rna_j) ? MB_BANK_DRH : MB_BANK_DNA);
      }
      const V3<T> dc = disp(v3<T>(rj.c[0], rj.c[1], rj.c[2]), v3<T>(ri.c[0], ri.c[1], ri.c[2]), M.box);
      sr[u] = !tagged && valid && dot(dc, dc) < sD[bank][5];
      if (any_debye) {
        const bool act = valid && (!MULTI || M.forms[bank].has_debye);
        const V3<T> db = disp(v3<T>(rj.b[0], rj.b[1], rj.b[2]), v3<T>(ri.b[0], ri.b[1], ri.b[2]), M.box);
        T m = T(1);
        if (M.half_charged_ends) m = ((fli & 4) ? T(0.5) : T(1)) * ((flj & 4) ? T(0.5) : T(1));
        const T* P = &sD[bank][0] - MB_P_DEBYE_KAPPA;  // debye_term indexes the bank by MB_P_DEBYE_*
        V3<T> gd = v3<T>(0, 0, 0);
        if (WP)
          e += debye_term<T, WF, true>(P, bank, act, db, m, cotd, gd, dacc);
        else
          e += debye_term<T, WF, false>(P, bank, act, db, m, cotd, gd, nacc);
        if (WF && (gd.x != T(0) || gd.y != T(0) || gd.z != T(0))) {
          T* gi = a.gback + 3 * (fbase + i);
          T* gj = a.gback + 3 * (fbase + j);
          atomicAdd(gj, gd.x);
          atomicAdd(gj + 1, gd.y);
          atomicAdd(gj + 2, gd.z);
          atomicAdd(gi, -gd.x);
          atomicAdd(gi + 1, -gd.y);
          atomicAdd(gi + 2, -gd.z);
        }
      }
    }
    if (want_sr && !tagged) {
#pragma unroll
      for (int u = 0; u < kDU; ++u) {  // warp-aggregated append to the short-range list
        const unsigned m = __ballot_sync(kFull, sr[u]);
        if (m) {
          const int leader = __ffs(m) - 1;
          int at = 0;
          if (lane == leader) at = atomicAdd(sr_count, __popc(m));
          at = __shfl_sync(kFull, at, leader) + __popc(m & ((1u << lane) - 1u));
          if (sr[u] && at < a.sr_capacity) sr_list[at] = pk_make(pi[u], pj[u]);
        }
      }
    }
  }
  // ---- flush: energy and Debye parameter gradients, one atomic per CTA and slot
  {
    T v = e;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if (lane == 0) sRed[warp][0] = v;
    if (WP) {
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          T w = dacc.r[b][k];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(kFull, w, o);
          if (lane == 0) sRed[warp][1 + b * 5 + k] = w;
        }
    }
  }
  __syncthreads();
  const int nslots = 1 + (WP ? 5 * NB : 0);
  if (threadIdx.x < nslots) {
    T v = T(0);
    for (int w = 0; w < kDB / 32; ++w) v += sRed[w][threadIdx.x];
    if (v != T(0)) {
      if (threadIdx.x == 0) {
        if (a.terms) atomicAdd(&a.terms[(long long)frame * MB_N_TERMS + MB_TERM_DEBYE], v);
      } else {
        const int b = (threadIdx.x - 1) / 5, k = (threadIdx.x - 1) % 5;
        atomicAdd(&a.d_params[(long long)frame * a.d_params_frame_stride + b * MB_P_COUNT + MB_P_DEBYE_KAPPA + k], v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Short-range pass over SR: phase-queued, one CTA per SM.
template <class T, bool WF, bool WP, bool MULTI>
__global__ void __launch_bounds__(kLB, MB_LIST_MINBLOCKS) k_list_sr(const EnergyDev<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nb = MULTI ? a.M.n_banks : 1;
  const int np = nb * MB_P_COUNT;
  pk_t* qBP = reinterpret_cast<pk_t*>(smem_raw);  // queues first: 8-byte entries at 16-byte aligned offsets; hydrogen-bonding candidates
  pk_t* qCX = qBP + kLQCap;                        // coaxial-stacking candidates
  pk_t* qCR = qCX + kLQCap;                        // cross-stacking candidates
  pk_t* qEV = qCR + kLQCap;                        // pairs with an excluded-volume site pair inside its cutoff
  T* sP = reinterpret_cast<T*>(qEV + kLQCap);
  T* sAcc = sP + np;                            // np reals when WP
  T* sE = sAcc + (WP ? np : 0);                 // kLWarps x 8
  BankCuts<T>* sCut = reinterpret_cast<BankCuts<T>*>(sE + kLWarps * MB_N_TERMS);  // MB_MAX_BANKS
  CosWin<T>* sWin = reinterpret_cast<CosWin<T>*>(sCut + MB_MAX_BANKS);             // 9 per bank: angular pre-screens of queues BP / CR
  int* ctr = reinterpret_cast<int*>(sWin + 9 * MB_MAX_BANKS);  // [1] n_bp (hydrogen bonding) [2] n_cx [3] n_cr (cross stacking) [4] n_ev

  const int frame = blockIdx.y;
  const int n = a.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long fbase = (long long)frame * n;
  const unsigned mask = a.mask;
  for (int k = threadIdx.x; k < np; k += kLB) {
    sP[k] = a.params[k];
    if (WP) sAcc[k] = T(0);
  }
  if (threadIdx.x < 8) ctr[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x < nb) {
    const T* P = sP + threadIdx.x * MB_P_COUNT;
    BankCuts<T> c;
    c.sr2 = T(0);
    T lo = T(1e30), hi = T(0);
    if (mask & (1u << MB_TERM_HB)) {
      lo = fmin(lo, P[MB_P_HB_RCLOW]);
      hi = fmax(hi, P[MB_P_HB_RCHIGH]);
    }
    if (mask & (1u << MB_TERM_CROSS)) {
      lo = fmin(lo, P[MB_P_CROSS_RCLOW]);
      hi = fmax(hi, P[MB_P_CROSS_RCHIGH]);
    }
    c.bp_lo2 = lo * lo;
    c.bp_hi2 = hi * hi;
    c.cx_lo2 = c.cx_hi2 = T(0);
    if (mask & (1u << MB_TERM_COAX)) {
      c.cx_lo2 = P[MB_P_COAX_RCLOW] * P[MB_P_COAX_RCLOW];
      c.cx_hi2 = P[MB_P_COAX_RCHIGH] * P[MB_P_COAX_RCHIGH];
    }
    c.ev2[0] = P[MB_P_UEXC_BACKBONE_RSTAR + 3] * P[MB_P_UEXC_BACKBONE_RSTAR + 3];
    c.ev2[1] = P[MB_P_UEXC_BASE_RSTAR + 3] * P[MB_P_UEXC_BASE_RSTAR + 3];
    c.ev2[2] = P[MB_P_UEXC_BACK_BASE_RSTAR + 3] * P[MB_P_UEXC_BACK_BASE_RSTAR + 3];
    c.ev2[3] = P[MB_P_UEXC_BASE_BACK_RSTAR + 3] * P[MB_P_UEXC_BASE_BACK_RSTAR + 3];
    sCut[threadIdx.x] = c;
    bp_windows9(P, sWin + 9 * threadIdx.x);
  }
  __syncthreads();

  const ModelT<T>& M = a.M;
  T cot[MB_N_TERMS];
#pragma unroll
  for (int t = 0; t < MB_N_TERMS; ++t) cot[t] = a.cot ? a.cot[(long long)frame * MB_N_TERMS + t] : T(1);
  T e[MB_N_TERMS];
#pragma unroll
  for (int t = 0; t < MB_N_TERMS; ++t) e[t] = T(0);
  SmemAcc<T> sacc{sAcc, MULTI};
  NullAcc nacc;

  const pk_t* sr_list = a.sr_list + (long long)frame * a.sr_capacity;
  const bool tagged = a.tagged != 0;
  const int32_t* pl = a.pairs + (long long)frame * a.pair_frame_stride;
  long long count;
  if (tagged) {  // the short-range pairs are the head of the support-tagged list
    count = a.pair_split[frame];
    count = count < a.pair_capacity ? count : a.pair_capacity;
  } else {
    count = a.sr_count[frame];
    count = count < a.sr_capacity ? count : a.sr_capacity;
  }
  long long base = (long long)blockIdx.x * kLB;
  bool flush = false;

  while (true) {
    __syncthreads();
    const int n_bp = ctr[1], n_cx = ctr[2], n_cr = ctr[3], n_ev = ctr[4];
    __syncthreads();  // everyone has read the counters before anyone updates them
    if (n_bp >= kLB || (flush && n_bp > 0)) {
      // ---------------- phase 3a: hydrogen bonding (entries passed the term's radial and six angular windows: dense lanes)
      const int cnt = n_bp >= kLB ? kLB : n_bp;
      const bool valid = threadIdx.x < cnt;
      const pk_t pk = valid ? qBP[n_bp - cnt + threadIdx.x] : 0ull;
      const int i = int(pk & 0xffffffffu), j = int(pk >> 32);
      T qi[4], qj[4];
      const Nuc<T> ni = load_nuc(a.center, a.quat, fbase + i, qi), nj = load_nuc(a.center, a.quat, fbase + j, qj);
      const int nti = MULTI ? a.nt_type[i] : 1, ntj = MULTI ? a.nt_type[j] : 1;
      NucGrad<T> Gi, Gj;
      Gi.zero();
      Gj.zero();
      const unsigned m3 = mask & (1u << MB_TERM_HB);
      if (WP)
        unbonded_pair<T, WF, true>(M, sP, valid, ni, nj, a.seq[i], a.seq[j], nti, ntj, T(1), m3, cot, e, Gi, Gj, sacc);
      else
        unbonded_pair<T, WF, false>(M, sP, valid, ni, nj, a.seq[i], a.seq[j], nti, ntj, T(1), m3, cot, e, Gi, Gj, nacc);
      if (WF && valid) {
        scatter_nuc_grad(a, fbase + i, Gi, qi);
        scatter_nuc_grad(a, fbase + j, Gj, qj);
      }
      if (threadIdx.x == 0) ctr[1] = n_bp - cnt;
      continue;
    }
    if (n_cr >= kLB || (flush && n_cr > 0)) {
      // ---------------- phase 3b: cross stacking
      const int cnt = n_cr >= kLB ? kLB : n_cr;
      const bool valid = threadIdx.x < cnt;
      const pk_t pk = valid ? qCR[n_cr - cnt + threadIdx.x] : 0ull;
      const int i = int(pk & 0xffffffffu), j = int(pk >> 32);
      T qi[4], qj[4];
      const Nuc<T> ni = load_nuc(a.center, a.quat, fbase + i, qi), nj = load_nuc(a.center, a.quat, fbase + j, qj);
      const int nti = MULTI ? a.nt_type[i] : 1, ntj = MULTI ? a.nt_type[j] : 1;
      NucGrad<T> Gi, Gj;
      Gi.zero();
      Gj.zero();
      const unsigned m3 = mask & (1u << MB_TERM_CROSS);
      if (WP)
        unbonded_pair<T, WF, true>(M, sP, valid, ni, nj, a.seq[i], a.seq[j], nti, ntj, T(1), m3, cot, e, Gi, Gj, sacc);
      else
        unbonded_pair<T, WF, false>(M, sP, valid, ni, nj, a.seq[i], a.seq[j], nti, ntj, T(1), m3, cot, e, Gi, Gj, nacc);
      if (WF && valid) {
        scatter_nuc_grad(a, fbase + i, Gi, qi);
        scatter_nuc_grad(a, fbase + j, Gj, qj);
      }
      if (threadIdx.x == 0) ctr[3] = n_cr - cnt;
      continue;
    }
    if (n_ev >= kLB || (flush && n_ev > 0)) {
      // ---------------- phase 2b: excluded volume of the pairs with a site pair inside its cutoff (about 1 short-range pair in 100)
      const int cnt = n_ev >= kLB ? kLB : n_ev;
      const bool valid = threadIdx.x < cnt;
      const pk_t pk = valid ? qEV[n_ev - cnt + threadIdx.x] : 0ull;
      const int i = int(pk & 0xffffffffu), j = int(pk >> 32);
      T qi[4], qj[4];
      const Nuc<T> ni = load_nuc(a.center, a.quat, fbase + i, qi), nj = load_nuc(a.center, a.quat, fbase + j, qj);
      const int nti = MULTI ? a.nt_type[i] : 1, ntj = MULTI ? a.nt_type[j] : 1;
      NucGrad<T> Gi, Gj;
      Gi.zero();
      Gj.zero();
      const unsigned m3 = mask & (1u << MB_TERM_UEXC);
      if (WP)
        unbonded_pair<T, WF, true>(M, sP, valid, ni, nj, 0, 0, nti, ntj, T(1), m3, cot, e, Gi, Gj, sacc);
      else
        unbonded_pair<T, WF, false>(M, sP, valid, ni, nj, 0, 0, nti, ntj, T(1), m3, cot, e, Gi, Gj, nacc);
      if (WF && valid) {
        scatter_nuc_grad(a, fbase + i, Gi, qi);
        scatter_nuc_grad(a, fbase + j, Gj, qj);
      }
      if (threadIdx.x == 0) ctr[4] = n_ev - cnt;
      continue;
    }
    if (n_cx >= kLB || (flush && n_cx > 0)) {
      // ---------------- phase 3c: coaxial stacking
      const int cnt = n_cx >= kLB ? kLB : n_cx;
      const bool valid = threadIdx.x < cnt;
      const pk_t pk = valid ? qCX[n_cx - cnt + threadIdx.x] : 0ull;
      const int i = int(pk & 0xffffffffu), j = int(pk >> 32);
      T qi[4], qj[4];
      const Nuc<T> ni = load_nuc(a.center, a.quat, fbase + i, qi), nj = load_nuc(a.center, a.quat, fbase + j, qj);
      const int nti = MULTI ? a.nt_type[i] : 1, ntj = MULTI ? a.nt_type[j] : 1;
      NucGrad<T> Gi, Gj;
      Gi.zero();
      Gj.zero();
      const unsigned m3 = mask & (1u << MB_TERM_COAX);
      if (WP)
        unbonded_pair<T, WF, true>(M, sP, valid, ni, nj, 0, 0, nti, ntj, T(1), m3, cot, e, Gi, Gj, sacc);
      else
        unbonded_pair<T, WF, false>(M, sP, valid, ni, nj, 0, 0, nti, ntj, T(1), m3, cot, e, Gi, Gj, nacc);
      if (WF && valid) {
        scatter_nuc_grad(a, fbase + i, Gi, qi);
        scatter_nuc_grad(a, fbase + j, Gj, qj);
      }
      if (threadIdx.x == 0) ctr[2] = n_cx - cnt;
      continue;
    }
    if (flush) break;
    if (base >= count) {
      flush = true;
      continue;
    }
    {
      // ---------------- phase 2: SCREEN of a CTA-full of SR -- squared site distances against squared cutoffs / radial windows,
      // cosine windows with a float32 rsqrt; no double-precision special function.  Survivors go to the queue of each term
      // that can be non-zero for them (EV / HB / CR / CX), where they are evaluated with dense lanes.
      const long long k = base + threadIdx.x;
      base += (long long)gridDim.x * kLB;
      const bool valid = k < count;
      pk_t pk = 0ull;
      if (valid) pk = tagged ? pk_make(pl[k], pl[a.pair_capacity + k] & 0x1fffffff) : sr_list[k];
      const int i = int(pk & 0xffffffffu), j = int(pk >> 32);
      T qi[4], qj[4];
      const Nuc<T> ni = load_nuc(a.center, a.quat, fbase + i, qi), nj = load_nuc(a.center, a.quat, fbase + j, qj);
      const int nti = MULTI ? a.nt_type[i] : 1, ntj = MULTI ? a.nt_type[j] : 1;
      int bank, fi, fj;
      pair_bank(nb, nti, ntj, bank, fi, fj);
      const BankCuts<T> cut = sCut[bank];
      const Geom<T>&gi = M.geom[fi], &gj = M.geom[fj];
      const V3<T> base_i = site(ni, gi.base, T(0), T(0)), base_j = site(nj, gj.base, T(0), T(0));
      const V3<T> d_base = disp(base_j, base_i, M.box);
      const T r2 = dot(d_base, d_base);
      bool to_ev = false;
      if (mask & (1u << MB_TERM_UEXC)) {
        const V3<T> back_i = site(ni, gi.back[0], gi.back[1], gi.back[2]), back_j = site(nj, gj.back[0], gj.back[1], gj.back[2]);
        const V3<T> d_bb = disp(back_j, back_i, M.box), d_bh = disp(back_i, base_j, M.box), d_hb = disp(base_i, back_j, M.box);
        to_ev = valid && (dot(d_bb, d_bb) < cut.ev2[0] || r2 < cut.ev2[1] || dot(d_bh, d_bh) < cut.ev2[2] || dot(d_hb, d_hb) < cut.ev2[3]);
      }
      unsigned to_bp = (valid && r2 > cut.bp_lo2 && r2 < cut.bp_hi2) ? 1u : 0u;
      if (to_bp)  // bit 0: hydrogen bonding can be non-zero (six windows), bit 1: cross stacking (three)
        to_bp = bp_screen2(sP + bank * MB_P_COUNT, sWin + 9 * bank, mask, d_base, r2, ni.a1, nj.a1, ni.a3, nj.a3, (a.seq[i] & 3) * 4 + (a.seq[j] & 3));
      const V3<T> ds = disp(site(nj, gj.stack, T(0), T(0)), site(ni, gi.stack, T(0), T(0)), M.box);
      const T s2 = dot(ds, ds);
      const bool to_cx = valid && s2 > cut.cx_lo2 && s2 < cut.cx_hi2;
      lq_push(qBP, &ctr[1], (to_bp & 1u) != 0u, pk);
      lq_push(qCR, &ctr[3], (to_bp & 2u) != 0u, pk);
      lq_push(qCX, &ctr[2], to_cx, pk);
      lq_push(qEV, &ctr[4], to_ev, pk);
    }
  }

  // ---------------------------------------------------------------- flush
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
    for (int w = 0; w < kLWarps; ++w) v += sE[w * MB_N_TERMS + threadIdx.x];
    if (v != T(0)) atomicAdd(&a.terms[(long long)frame * MB_N_TERMS + threadIdx.x], v);
  }
  if (WP) {
    T* out = a.d_params + (long long)frame * a.d_params_frame_stride;
    for (int p = threadIdx.x; p < np; p += kLB) {
      const T v = sAcc[p];
      if (v != T(0)) atomicAdd(&out[p], v);
    }
  }
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

template <class T>
size_t list_workspace_bytes(long long n, long long n_frames, long long pair_capacity) {
  const size_t rows = (size_t)(n * n_frames);
  return align256(rows * sizeof(NucRec<T>)) + align256(rows * 3 * sizeof(T)) + align256(sizeof(int) * (size_t)n_frames) +
         align256(sizeof(pk_t) * (size_t)n_frames * (size_t)pair_capacity);
}
template size_t list_workspace_bytes<float>(long long, long long, long long);
template size_t list_workspace_bytes<double>(long long, long long, long long);

template <class T, bool WF, bool WP, bool MULTI>
static int launch_list(cudaStream_t s, const EnergyDev<T>& a) {
  const int nb = MULTI ? a.M.n_banks : 1;
  const size_t np = (size_t)nb * MB_P_COUNT;
  const size_t smem = sizeof(pk_t) * (4 * kLQCap) + sizeof(T) * (np * (WP ? 2 : 1) + kLWarps * MB_N_TERMS) +
                      sizeof(BankCuts<T>) * MB_MAX_BANKS + sizeof(CosWin<T>) * 9 * MB_MAX_BANKS + sizeof(int) * 8;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    MB_CUDA_CHECK(cudaGetDevice(&dev));
    MB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long rows = (long long)a.n * a.n_frames;
  k_list_prep<T><<<ceil_div(rows, 256), 256, 0, s>>>(a);
  {
    const long long ntiles = (a.pair_capacity + kDB * kDU - 1) / (kDB * kDU);
    long long per_frame = 8ll * sms;
    if (a.n_frames > 1) per_frame = (per_frame + a.n_frames - 1) / a.n_frames;
    per_frame = per_frame > ntiles ? ntiles : per_frame;
    dim3 grid((unsigned)(per_frame < 1 ? 1 : per_frame), a.n_frames);
    k_list_debye<T, WF, WP, MULTI><<<grid, kDB, 0, s>>>(a);
  }
  if (a.mask & (MB_UNBONDED_TERMS & ~(1u << MB_TERM_DEBYE))) {
    long long per_frame = (long long)MB_LIST_MINBLOCKS * sms;  // resident CTAs; the list length is known only on the device
    if (a.n_frames > 1) per_frame = (per_frame + a.n_frames - 1) / a.n_frames;
    const long long ntiles = (a.pair_capacity + kLB - 1) / kLB;
    per_frame = per_frame > ntiles ? ntiles : per_frame;
    dim3 grid((unsigned)(per_frame < 1 ? 1 : per_frame), a.n_frames);
    MB_CUDA_CHECK(cudaFuncSetAttribute(k_list_sr<T, WF, WP, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_list_sr<T, WF, WP, MULTI><<<grid, kLB, smem, s>>>(a);
  }
  if (WF && a.gback) k_list_post<T><<<ceil_div(rows, 256), 256, 0, s>>>(a);
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}

template <class T>
int launch_list_kernel(cudaStream_t s, const EnergyDev<T>& a0, void* workspace, bool wf, bool wp) {
  EnergyDev<T> a = a0;
  const size_t rows = (size_t)a.n * a.n_frames;
  char* w = static_cast<char*>(workspace);
  a.rec = reinterpret_cast<T*>(w);
  w += align256(rows * sizeof(NucRec<T>));
  a.gback = wf ? reinterpret_cast<T*>(w) : nullptr;
  w += align256(rows * 3 * sizeof(T));
  a.sr_count = reinterpret_cast<int*>(w);
  w += align256(sizeof(int) * (size_t)a.n_frames);
  a.sr_list = reinterpret_cast<pk_t*>(w);
  a.sr_capacity = a.pair_capacity;
  const bool multi = a.M.n_banks > 1;
  if (multi) {
    if (wf && wp) return launch_list<T, true, true, true>(s, a);
    if (wf) return launch_list<T, true, false, true>(s, a);
    if (wp) return launch_list<T, false, true, true>(s, a);
    return launch_list<T, false, false, true>(s, a);
  }
  if (wf && wp) return launch_list<T, true, true, false>(s, a);
  if (wf) return launch_list<T, true, false, false>(s, a);
  if (wp) return launch_list<T, false, true, false>(s, a);
  return launch_list<T, false, false, false>(s, a);
}

template int launch_list_kernel<float>(cudaStream_t, const EnergyDev<float>&, void*, bool, bool);
template int launch_list_kernel<double>(cudaStream_t, const EnergyDev<double>&, void*, bool, bool);

}  // namespace mb
