// energy_kernels.cu -- fused per-pair energy / force / parameter-gradient kernels and their C-ABI entry.
//
// Launch shape: grid = (pair blocks, frames).  One thread owns one listed pair and evaluates every enabled term
// of that pair in one pass from the two nucleotides' (center, quaternion) -- the body axes and interaction
// sites are rebuilt in registers, never materialised in HBM (the reference re-derives and stores them once
// per term, mythos/energy/base.py:205-208).  Per-term energies are reduced warp -> block -> one atomic per
// block and term; nucleotide gradients go out as (dE/dcenter, dE/dquat) atomics; parameter gradients are
// warp-reduced into a shared-memory bank image and flushed once per block.
#include "common.cuh"
#include "energy_dev.cuh"
#include "observables_dev.cuh"

namespace mb {

constexpr int kBlock = 128;
constexpr long long kListKernelMinPairs = 65536;  // below this the list kernels' extra launches cost more than they save

// block-reduce the 8 per-term energies and add them to terms[frame]
template <class T>
__device__ __forceinline__ void reduce_terms(T e[MB_N_TERMS], T* sE, T* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < MB_N_TERMS; ++t) {
    T v = e[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if (lane == 0) sE[warp * MB_N_TERMS + t] = v;
  }
  __syncthreads();
  if (threadIdx.x < MB_N_TERMS && out) {
    T v = 0;
    for (int w = 0; w < kBlock / 32; ++w) v += sE[w * MB_N_TERMS + threadIdx.x];
    if (v != T(0)) atomicAdd(&out[threadIdx.x], v);
  }
}

// one block of pairs: BONDED -> (FENE | bonded excluded volume | stacking, selected by `mask`) of kBlock bonds,
// else every enabled unbonded term of kBlock listed pairs
template <class T, bool WF, bool WP, bool BONDED>
__device__ __forceinline__ void pairs_body(const EnergyDev<T>& a, long long k, unsigned mask, T* sP, T* sE, T* sAcc, int np) {
  const int frame = blockIdx.y;
  const long long fbase = (long long)frame * a.n;
  int i = 0, j = 0;
  bool valid;
  if (BONDED) {
    valid = k < a.n_bonded;
    if (valid) {
      i = a.bonded[2 * k];
      j = a.bonded[2 * k + 1];
    }
  } else {
    valid = k < a.pair_capacity;
    if (valid) {
      const int32_t* pl = a.pairs + (long long)frame * a.pair_frame_stride;
      i = pl[k];
      j = pl[a.pair_capacity + k];
      valid = (i >= 0 && j >= 0 && i < a.n && j < a.n);
      if (!valid) i = j = 0;
    }
  }
  T cot[MB_N_TERMS];
#pragma unroll
  for (int t = 0; t < MB_N_TERMS; ++t) cot[t] = a.cot ? a.cot[(long long)frame * MB_N_TERMS + t] : T(1);

  T e[MB_N_TERMS];
#pragma unroll
  for (int t = 0; t < MB_N_TERMS; ++t) e[t] = T(0);
  T qi[4], qj[4];
  const Nuc<T> ni = load_nuc(a.center, a.quat, fbase + i, qi);
  const Nuc<T> nj = load_nuc(a.center, a.quat, fbase + j, qj);
  NucGrad<T> Gi, Gj;
  Gi.zero();
  Gj.zero();
  SmemAcc<T> acc{sAcc, a.M.n_banks > 1};
  const int nti = a.nt_type ? a.nt_type[i] : 1, ntj = a.nt_type ? a.nt_type[j] : 1;
  if (BONDED) {
    const int32_t* snt = a.nt_type_stack ? a.nt_type_stack : a.nt_type;
    const int si = snt ? snt[i] : 1, sj = snt ? snt[j] : 1;
    const bool ps = a.pseq.pmarg != nullptr && (mask & a.pseq.terms & (1u << MB_TERM_STACK));
    int same = -1;
    const T wx = (ps && valid) ? pseq_weight(a.pseq, sP + MB_P_STACK_W00, a.pseq.same_w_stack, i, j, same) : T(0);
    bonded_pair<T, WF, WP>(a.M, sP, valid, ni, nj, a.seq[i], a.seq[j], nti, ntj, si, sj, mask, cot, e, Gi, Gj, acc, ps, wx);
    if (ps) {  // (warp-uniform branch: acc.add is warp-convergent; invalid / zero pairs contribute coef = 0)
      const T coef = (valid && wx != T(0)) ? cot[MB_TERM_STACK] * e[MB_TERM_STACK] / wx : T(0);
      pseq_weight_grad<T, WP>(a.pseq, sP + MB_P_STACK_W00, a.pseq.d_same_w_stack, MB_P_STACK_W00, i, j, same, coef, acc);
    }
  } else {
    T m = T(1);
    if (a.M.half_charged_ends && a.is_end) m = (a.is_end[i] ? T(0.5) : T(1)) * (a.is_end[j] ? T(0.5) : T(1));
    const bool ps = a.pseq.pmarg != nullptr && (mask & a.pseq.terms & (1u << MB_TERM_HB));
    int same = -1;
    const T wx = (ps && valid) ? pseq_weight(a.pseq, sP + MB_P_HB_W00, a.pseq.same_w_hb, i, j, same) : T(0);
    unbonded_pair<T, WF, WP>(a.M, sP, valid, ni, nj, a.seq[i], a.seq[j], nti, ntj, m, mask, cot, e, Gi, Gj, acc, ps, wx);
    if (ps) {
      const T coef = (valid && wx != T(0)) ? cot[MB_TERM_HB] * e[MB_TERM_HB] / wx : T(0);
      pseq_weight_grad<T, WP>(a.pseq, sP + MB_P_HB_W00, a.pseq.d_same_w_hb, MB_P_HB_W00, i, j, same, coef, acc);
    }
  }
  if (WF && valid) {
    scatter_nuc_grad(a, fbase + i, Gi, qi);
    scatter_nuc_grad(a, fbase + j, Gj, qj);
  }
  reduce_terms(e, sE, a.terms ? a.terms + (long long)frame * MB_N_TERMS : nullptr);
  if (WP) {
    __syncthreads();
    T* out = a.d_params + (long long)frame * a.d_params_frame_stride;
    for (int p = threadIdx.x; p < np; p += kBlock) {
      const T v = sAcc[p];
      if (v != T(0)) atomicAdd(&out[p], v);
    }
  }
}

// ONE launch for bonded and unbonded pairs: blocks [0, 3 * bonded_chunks) evaluate one bonded term each for one chunk of
// bonds (the three terms of a bond run in different blocks: the per-thread dependency chain -- what bounds small
// systems such as the 60-bp MD duplex -- is a single term long), the remaining blocks one chunk of listed pairs each.
// `split_unbonded`: short lists (the MD duplex) also spread the unbonded terms of a chunk of pairs over kSplit blocks
// (excluded volume + Debye | hydrogen bonding | cross stacking | coaxial stacking): the longest per-thread chain drops
// from the sum of the terms to the longest single term, at the price of loading each pair kSplit times (and of
// computing the six shared angles in both the hydrogen-bonding and the cross-stacking block).
constexpr int kSplit = 4;
__constant__ unsigned kUnbondedGroups[kSplit] = {(1u << MB_TERM_UEXC) | (1u << MB_TERM_DEBYE), 1u << MB_TERM_HB, 1u << MB_TERM_CROSS,
                                                 1u << MB_TERM_COAX};
template <class T, bool WF, bool WP>
__global__ void __launch_bounds__(kBlock) k_pairs(const EnergyDev<T> a, int bonded_chunks, int split_unbonded) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int np = a.M.n_banks * MB_P_COUNT;
  T* sP = reinterpret_cast<T*>(smem_raw);
  T* sE = sP + np;
  T* sAcc = sE + (kBlock / 32) * MB_N_TERMS;
  unsigned mask = a.mask;
  const bool bonded = blockIdx.x < 3 * bonded_chunks;
  if (bonded) {
    mask &= 1u << (blockIdx.x % 3);  // MB_TERM_FENE, MB_TERM_BEXC, MB_TERM_STACK = 0, 1, 2
    if (!mask) return;
  } else if (split_unbonded) {
    if (!(mask & kUnbondedGroups[(blockIdx.x - 3 * bonded_chunks) % kSplit])) return;  // this block's term group is not enabled
  }
  for (int k = threadIdx.x; k < np; k += kBlock) {
    sP[k] = a.params[k];
    if (WP) sAcc[k] = T(0);
  }
  __syncthreads();
  if (bonded) {
    pairs_body<T, WF, WP, true>(a, (long long)(blockIdx.x / 3) * kBlock + threadIdx.x, mask, sP, sE, sAcc, np);
  } else {
    long long chunk = blockIdx.x - 3 * bonded_chunks;
    if (split_unbonded) {
      mask &= kUnbondedGroups[chunk % kSplit];
      chunk /= kSplit;
    }
    pairs_body<T, WF, WP, false>(a, chunk * kBlock + threadIdx.x, mask, sP, sE, sAcc, np);
  }
}

template <class T, bool WF, bool WP>
static int launch_pairs(cudaStream_t s, const EnergyDev<T>& a, void* list_ws) {
  const size_t smem = sizeof(T) * (size_t)(a.M.n_banks * MB_P_COUNT * (WP ? 2 : 1) + (kBlock / 32) * MB_N_TERMS);
  const int bonded_chunks = ((a.mask & MB_BONDED_TERMS) && a.n_bonded > 0) ? ceil_div(a.n_bonded, kBlock) : 0;
  const bool unbonded = (a.mask & MB_UNBONDED_TERMS) && a.pair_capacity > 0;
  const bool lists = unbonded && list_ws != nullptr;  // phase-queued list kernels (list_kernels.cu) take the unbonded part
  long long pair_chunks = (unbonded && !lists) ? ceil_div(a.pair_capacity, kBlock) : 0;
  const int split = (pair_chunks > 0 && a.pair_capacity * a.n_frames < kListKernelMinPairs) ? 1 : 0;  // latency-bound sizes only
  if (split) pair_chunks *= kSplit;
  if (3 * bonded_chunks + pair_chunks > 0) {
    dim3 grid((unsigned)(3 * bonded_chunks + pair_chunks), a.n_frames);
    k_pairs<T, WF, WP><<<grid, kBlock, smem, s>>>(a, bonded_chunks, split);
    MB_CUDA_CHECK(cudaGetLastError());
  }
  if (lists) return launch_list_kernel<T>(s, a, list_ws, WF, WP);
  return MB_OK;
}

template <class T>
static int energy_impl(cudaStream_t s, const mb_energy_args* x) {
  MB_REQUIRE(x && x->model, MB_EINVAL_SHAPE, "energy: null args / model");
  const mb_model& m = *x->model;
  MB_REQUIRE(m.n_banks == 1 || m.n_banks == 3, MB_EINVAL_MODEL, "energy: n_banks must be 1 or 3");
  MB_REQUIRE(x->n > 0 && x->n_frames > 0, MB_EINVAL_SHAPE, "energy: n and n_frames must be positive");
  MB_REQUIRE(x->n_frames <= 65535, MB_EINVAL_SHAPE, "energy: at most 65535 frames per call");
  MB_REQUIRE(x->center && x->quat && x->seq && x->params, MB_EINVAL_SHAPE, "energy: center/quat/seq/params required");
  MB_REQUIRE(m.n_banks == 1 || x->nt_type, MB_EINVAL_SHAPE, "energy: nt_type required for the 3-bank (NA1) model");
  MB_REQUIRE(x->n_bonded == 0 || x->bonded, MB_EINVAL_SHAPE, "energy: bonded list missing");
  MB_REQUIRE(x->pair_capacity == 0 || x->pairs, MB_EINVAL_SHAPE, "energy: pair list missing");
  MB_REQUIRE(!(x->all_pairs_cutoff < 0), MB_EINVAL_SHAPE, "energy: all_pairs_cutoff must be >= 0");
  {
    const bool any = m.box[0] > 0 || m.box[1] > 0 || m.box[2] > 0, all = m.box[0] > 0 && m.box[1] > 0 && m.box[2] > 0;
    MB_REQUIRE(!any || all, MB_EINVAL_MODEL, "energy: box must be all zero (free space) or all positive");
  }
  MB_REQUIRE(x->terms || x->d_center || x->d_quat || x->d_params, MB_EINVAL_SHAPE, "energy: no output requested");
  for (int b = 0; b < m.n_banks; ++b) {
    MB_REQUIRE(m.forms[b].stack_form >= 0 && m.forms[b].stack_form <= 1 && m.forms[b].cross_form >= 0 &&
                   m.forms[b].cross_form <= 1 && m.forms[b].coax_form >= 0 && m.forms[b].coax_form <= 1,
               MB_EINVAL_MODEL, "energy: unknown term form");
  }
  EnergyDev<T> a;
  a.M.load(m);
  a.n = x->n;
  a.n_frames = x->n_frames;
  a.n_bonded = x->n_bonded;
  a.center = static_cast<const T*>(x->center);
  a.quat = static_cast<const T*>(x->quat);
  a.seq = x->seq;
  a.nt_type = x->nt_type;
  a.nt_type_stack = x->nt_type_stack;
  a.is_end = x->is_end;
  a.bonded = x->bonded;
  a.pairs = x->pairs;
  a.pair_capacity = x->pair_capacity;
  a.pair_frame_stride = x->pair_frame_stride;
  a.pair_count = x->pair_count;
  a.all_pairs_cutoff = T(x->all_pairs_cutoff);
  a.params = static_cast<const T*>(x->params);
  a.cot = static_cast<const T*>(x->cot);
  a.mask = x->term_mask;
  a.terms = static_cast<T*>(x->terms);
  a.d_center = static_cast<T*>(x->d_center);
  a.d_quat = static_cast<T*>(x->d_quat);
  a.d_params = static_cast<T*>(x->d_params);
  a.d_params_frame_stride = x->d_params_frame_stride;
  a.rec = nullptr;
  a.gback = nullptr;
  a.sr_list = nullptr;
  a.sr_count = nullptr;
  a.sr_capacity = 0;
  a.tagged = (x->flags & MB_FLAG_TAGGED_PAIRS) ? 1 : 0;
  a.pair_split = x->pair_split;
  a.obs_out = nullptr;
  a.acc_scratch = nullptr;
  a.pseq = PseqDev<T>{};
  if (x->pseq) {
    const mb_pseq& q = *x->pseq;
    MB_REQUIRE(m.n_banks == 1, MB_EINVAL_MODEL, "energy: probabilistic sequences are supported for single-bank models");
    MB_REQUIRE(q.pmarg && q.bp_of && q.within, MB_EINVAL_SHAPE, "energy: pseq needs pmarg, bp_of and within");
    MB_REQUIRE((x->flags & MB_FLAG_GENERIC_KERNEL) && x->all_pairs_cutoff == 0, MB_EINVAL_SHAPE,
               "energy: probabilistic sequences run through the generic pair kernel (set MB_FLAG_GENERIC_KERNEL, explicit pair list)");
    a.pseq.pmarg = static_cast<const T*>(q.pmarg);
    a.pseq.bp_of = q.bp_of;
    a.pseq.within = q.within;
    a.pseq.same_w_stack = static_cast<const T*>(q.same_w_stack);
    a.pseq.same_w_hb = static_cast<const T*>(q.same_w_hb);
    a.pseq.d_pmarg = static_cast<T*>(q.d_pmarg);
    a.pseq.d_same_w_stack = static_cast<T*>(q.d_same_w_stack);
    a.pseq.d_same_w_hb = static_cast<T*>(q.d_same_w_hb);
    a.pseq.terms = q.terms;
    MB_REQUIRE(!(q.terms & (1u << MB_TERM_STACK)) || q.same_w_stack || true, MB_EINVAL_SHAPE, "energy: pseq stacking table missing");
  }
  ObsDev obs{};
  T* obs_out = nullptr;
  if (x->observables) {
    MB_REQUIRE(x->observables_out, MB_EINVAL_SHAPE, "energy: observables requested without observables_out");
    const int st = check_observable_spec(x->observables, &obs);
    if (st != MB_OK) return st;
    obs_out = static_cast<T*>(x->observables_out);
  }

  const size_t np = (size_t)m.n_banks * MB_P_COUNT;
  if (!(x->flags & MB_FLAG_ACCUMULATE)) {
    const size_t F = x->n_frames, N = x->n;
    if (a.terms) MB_CUDA_CHECK(cudaMemsetAsync(a.terms, 0, sizeof(T) * F * MB_N_TERMS, s));
    if (a.d_center) MB_CUDA_CHECK(cudaMemsetAsync(a.d_center, 0, sizeof(T) * F * N * 3, s));
    if (a.d_quat) MB_CUDA_CHECK(cudaMemsetAsync(a.d_quat, 0, sizeof(T) * F * N * 4, s));
    if (a.d_params) {
      const size_t rows = x->d_params_frame_stride ? F : 1;
      const size_t stride = x->d_params_frame_stride ? (size_t)x->d_params_frame_stride : np;
      MB_CUDA_CHECK(cudaMemsetAsync(a.d_params, 0, sizeof(T) * ((rows - 1) * stride + np), s));
    }
  }
  const bool wf = a.d_center || a.d_quat, wp = a.d_params != nullptr;
  // frame-resident path: one block per frame with the frame staged in shared memory (energies and dE/dparams only)
  if (!wf && !(x->flags & MB_FLAG_GENERIC_KERNEL) && frame_kernel_eligible<T>(a)) {
    a.obs = obs;
    a.obs_out = obs_out;  // epilogue of the same kernel: the frame is already in shared memory
    a.acc_scratch = nullptr;
    if (wp && frame_scratch_bytes(sizeof(T)) > 0) {
      MB_REQUIRE(x->workspace && x->workspace_bytes >= frame_scratch_bytes(sizeof(T)), MB_ECAPACITY,
                 "energy: dE/dparams through the frame-resident kernel needs a workspace of mythos_b200_energy_workspace_bytes() bytes");
      a.acc_scratch = static_cast<T*>(x->workspace);
    }
    return launch_frame_kernel<T>(s, a, wp);
  }
  if (obs_out) {  // every other route: the standalone kernel, enqueued on the same stream
    const int st = launch_observables<T>(s, a.M, a.n, a.n_frames, a.center, a.quat, m.n_banks == 1 ? nullptr : a.nt_type, obs, obs_out);
    if (st != MB_OK) return st;
  }
  MB_REQUIRE(!a.tagged || (x->pair_split && x->pair_count && x->workspace), MB_ECAPACITY,
             "energy: tagged pair lists need the frame-resident kernel (single bank, no position gradients, frame fits in "
             "shared memory, n < 16384) or the list kernels (workspace, pair_count and pair_split)");
  MB_REQUIRE(!(x->all_pairs_cutoff > 0), MB_ECAPACITY,
             "energy: all_pairs_cutoff needs the frame-resident kernel (single bank, no position gradients, frame fits in "
             "shared memory); build a neighbour list with mythos_b200_nl_build_* instead");
  // phase-queued list kernel when the caller lends the workspace it needs; otherwise one thread per pair
  // (short lists -- the 60-bp MD duplex has 7 021 pairs -- are launch-latency bound: one generic launch beats four)
  void* lk = nullptr;
  if (!(x->flags & MB_FLAG_GENERIC_KERNEL) && x->workspace && (a.tagged || (x->flags & MB_FLAG_LIST_KERNEL) || (long long)x->pair_capacity * x->n_frames >= kListKernelMinPairs) &&
      x->workspace_bytes >= list_workspace_bytes<T>(x->n, x->n_frames, x->pair_capacity))
    lk = x->workspace;
  if (wf && wp) return launch_pairs<T, true, true>(s, a, lk);
  if (wf) return launch_pairs<T, true, false>(s, a, lk);
  if (wp) return launch_pairs<T, false, true>(s, a, lk);
  return launch_pairs<T, false, false>(s, a, lk);
}

}  // namespace mb

namespace mb {
template <class T>
__global__ void k_backbone_sites(Geom<T> g0, Geom<T> g1, const int32_t* __restrict__ nt_type, int n, long long n_total,
                                 const T* __restrict__ center, const T* __restrict__ quat, T* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_total) return;
  T q[4];
  const Nuc<T> nu = load_nuc(center, quat, idx, q);
  const Geom<T>& g = (nt_type && nt_type[idx % n] == 2) ? g1 : g0;
  const V3<T> b = site(nu, g.back[0], g.back[1], g.back[2]);
  out[3 * idx] = b.x;
  out[3 * idx + 1] = b.y;
  out[3 * idx + 2] = b.z;
}
template <class T>
static int backbone_sites_impl(cudaStream_t s, const mb_model* m, int64_t n_total, const void* center, const void* quat, void* out,
                               const int32_t* nt_type, int32_t n) {
  MB_REQUIRE(m && center && quat && out && n_total > 0, MB_EINVAL_SHAPE, "backbone_sites: missing arguments");
  MB_REQUIRE(m->n_banks == 1 || (nt_type && n > 0), MB_EINVAL_SHAPE, "backbone_sites: nt_type (N) required for the 3-bank model");
  Geom<T> g0, g1;
  g0.load(m->geom[0]);
  g1.load(m->geom[1]);
  k_backbone_sites<T><<<ceil_div(n_total, 256), 256, 0, s>>>(g0, g1, m->n_banks > 1 ? nt_type : nullptr, n > 0 ? n : 1, n_total,
                                                              static_cast<const T*>(center), static_cast<const T*>(quat), static_cast<T*>(out));
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}
}  // namespace mb
namespace mb {
// centres and backbone sites of all frames as float32, brought near the origin (periodic: primary image; free space:
// relative to the frame's first nucleotide / its backbone site), plus the largest coordinate magnitude written
template <class T>
__global__ void k_support_points(Geom<T> g0, Geom<T> g1, const int32_t* __restrict__ nt_type, int n, long long n_total,
                                 const T* __restrict__ center, const T* __restrict__ quat, T bx, T by, T bz,
                                 float* __restrict__ out_c, float* __restrict__ out_s, float* __restrict__ extent) {
  // grid = (blocks over the nucleotides of a frame, frames): no 64-bit division per thread
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const long long first = (long long)blockIdx.y * n, idx = first + i;
  float m = 0.f;
  if (i < n && idx < n_total) {
    T q[4], q0[4];
    const Nuc<T> nu = load_nuc(center, quat, idx, q);
    const Geom<T>& g = (nt_type && nt_type[i] == 2) ? g1 : g0;
    V3<T> c = nu.c, b = site(nu, g.back[0], g.back[1], g.back[2]);
    const T box[3] = {bx, by, bz};
    if (bx > T(0)) {
      T* pc[3] = {&c.x, &c.y, &c.z};
      T* pb[3] = {&b.x, &b.y, &b.z};
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        T v = fmod(*pc[d], box[d]);
        *pc[d] = v < T(0) ? v + box[d] : v;
        v = fmod(*pb[d], box[d]);
        *pb[d] = v < T(0) ? v + box[d] : v;
      }
    } else {
      const Nuc<T> n0 = load_nuc(center, quat, first, q0);
      const Geom<T>& gf = (nt_type && nt_type[0] == 2) ? g1 : g0;
      const V3<T> b0 = site(n0, gf.back[0], gf.back[1], gf.back[2]);
      c = c - n0.c;
      b = b - b0;
    }
    out_c[3 * idx] = float(c.x);
    out_c[3 * idx + 1] = float(c.y);
    out_c[3 * idx + 2] = float(c.z);
    out_s[3 * idx] = float(b.x);
    out_s[3 * idx + 1] = float(b.y);
    out_s[3 * idx + 2] = float(b.z);
    m = fmaxf(fmaxf(fmaxf(fabsf(float(c.x)), fabsf(float(c.y))), fabsf(float(c.z))),
              fmaxf(fmaxf(fabsf(float(b.x)), fabsf(float(b.y))), fabsf(float(b.z))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(extent), __float_as_int(m));  // non-negative floats order like ints
}
template <class T>
static int support_points_impl(cudaStream_t s, const mb_model* m, int32_t n, int32_t n_frames, const void* center, const void* quat,
                               const int32_t* nt_type, void* out_center, void* out_site, void* extent) {
  MB_REQUIRE(m && center && quat && out_center && out_site && extent && n > 0 && n_frames > 0, MB_EINVAL_SHAPE,
             "support_points: missing arguments");
  MB_REQUIRE(m->n_banks == 1 || nt_type, MB_EINVAL_SHAPE, "support_points: nt_type (N) required for the 3-bank model");
  Geom<T> g0, g1;
  g0.load(m->geom[0]);
  g1.load(m->geom[1]);
  const long long total = (long long)n * n_frames;
  MB_REQUIRE(n_frames <= 65535, MB_EINVAL_SHAPE, "support_points: at most 65535 frames per call");
  k_support_points<T><<<dim3((unsigned)ceil_div(n, 256), (unsigned)n_frames), 256, 0, s>>>(g0, g1, m->n_banks > 1 ? nt_type : nullptr, n, total,
                                                            static_cast<const T*>(center), static_cast<const T*>(quat), T(m->box[0]),
                                                            T(m->box[1]), T(m->box[2]), static_cast<float*>(out_center),
                                                            static_cast<float*>(out_site), static_cast<float*>(extent));
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}
}  // namespace mb
extern "C" int mythos_b200_support_points_f64(void* stream, const mb_model* m, int32_t n, int32_t n_frames, const void* center,
                                              const void* quat, const int32_t* nt_type, void* out_center, void* out_site, void* extent) {
  return mb::support_points_impl<double>(static_cast<cudaStream_t>(stream), m, n, n_frames, center, quat, nt_type, out_center, out_site, extent);
}
extern "C" int mythos_b200_support_points_f32(void* stream, const mb_model* m, int32_t n, int32_t n_frames, const void* center,
                                              const void* quat, const int32_t* nt_type, void* out_center, void* out_site, void* extent) {
  return mb::support_points_impl<float>(static_cast<cudaStream_t>(stream), m, n, n_frames, center, quat, nt_type, out_center, out_site, extent);
}

extern "C" int mythos_b200_backbone_sites_f64(void* stream, const mb_model* m, int64_t n_total, const void* center, const void* quat, void* out,
                                              const int32_t* nt_type, int32_t n) {
  return mb::backbone_sites_impl<double>(static_cast<cudaStream_t>(stream), m, n_total, center, quat, out, nt_type, n);
}
extern "C" int mythos_b200_backbone_sites_f32(void* stream, const mb_model* m, int64_t n_total, const void* center, const void* quat, void* out,
                                              const int32_t* nt_type, int32_t n) {
  return mb::backbone_sites_impl<float>(static_cast<cudaStream_t>(stream), m, n_total, center, quat, out, nt_type, n);
}

extern "C" int mythos_b200_frame_kernel_fits(int32_t n, int32_t real_bytes, int32_t want_params) {
  if (n <= 0) return 0;
  return (real_bytes == 4 ? mb::frame_kernel_fits<float>(n, want_params != 0) : mb::frame_kernel_fits<double>(n, want_params != 0)) ? 1 : 0;
}
extern "C" size_t mythos_b200_energy_workspace_bytes(int32_t n, int32_t n_frames, int64_t pair_capacity, int32_t real_bytes) {
  if (n <= 0 || n_frames <= 0) return 0;
  // the larger of: the list kernels' scratch (explicit lists) and the frame-resident kernel's parameter-gradient images
  const size_t frame_ws = mb::frame_scratch_bytes(real_bytes == 4 ? 4 : 8);
  if (pair_capacity <= 0) return frame_ws;
  const size_t list_ws = real_bytes == 4 ? mb::list_workspace_bytes<float>(n, n_frames, pair_capacity)
                                         : mb::list_workspace_bytes<double>(n, n_frames, pair_capacity);
  return list_ws > frame_ws ? list_ws : frame_ws;
}
extern "C" int mythos_b200_energy_f64(void* stream, const mb_energy_args* a) {
  return mb::energy_impl<double>(static_cast<cudaStream_t>(stream), a);
}
extern "C" int mythos_b200_energy_f32(void* stream, const mb_energy_args* a) {
  return mb::energy_impl<float>(static_cast<cudaStream_t>(stream), a);
}
