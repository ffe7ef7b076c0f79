// frame_kernels.cu -- frame-resident energy + dE/dparams kernel (the DiffTRe / EnergyFunction.map shape).
//
// One CTA owns one stored frame.  The frame's (center, quaternion) rows -- and, when they fit, the backbone sites
// -- are staged once in shared memory (56 + 24 B per nucleotide in float64: 163 KB at N = 2040, inside the 227 KB a
// CTA may use on sm_100a), and every pair of the frame is found AND evaluated from there: in all-pairs mode
// (the reference's `topology.unbonded_neighbors` semantics) the CTA bins its nucleotides into a shared-memory cell
// list and walks it itself, so no pair list ever exists in HBM; with an explicit list the list is streamed.
//
// Work is regrouped so that each code region runs with (nearly) full warps -- the generic one-thread-per-pair
// kernel spent 93 % of its issue slots waiting on instruction fetch because every lane wandered through a 229 KB
// instruction stream (profiles/r01_v1_*):
//
//   phase B   bonded pairs (FENE, bonded excluded volume, stacking), one thread per bond
//   producer  all-pairs mode: cells of HALF the cutoff (5x5x5 stencil: 58 % of the candidates a 3x3x3 stencil of
//             full-size cells has to test).  Thread t walks the forward half shell of the nucleotides at cell-order
//             positions t, t+512, ...: 13 (dy,dz) rows, each one contiguous run of the cell-ordered list.  A step tests
//             kSlice candidates per thread.  list mode: kSlice list entries per thread.
//             Either way a candidate inside the caller's centre cutoff is split by the exact supports of the terms:
//             backbone-site distance inside the Debye-Hueckel cutoff -> queue DB (about 1 listed pair in 3), centre
//             distance inside the short-range cutoff -> queue SR (1 in 6); the rest is dropped right there.
//   phase 1   a CTA-full of DB: Debye-Hueckel on the cached backbone sites, parameter gradients in REGISTERS
//   phase 2   a CTA-full of SR: excluded volume (4 site pairs), gradients in registers; pairs inside the hydrogen-bond /
//             cross-stacking radial window go to queue BP, inside the coaxial window to queue CX
//   phase 3   a CTA-full of BP / CX: the six-angle products with dense lanes; their parameter gradients are
//             warp-reduced into the shared-memory bank image
//   flush     partial queues, register accumulators -> bank image -> one J row, per-term energies -> one terms row
//
// Every queue append is ordered (block prefix over warp counts), so results are bitwise repeatable run to run.
#include "energy_dev.cuh"
#include "observables_dev.cuh"

namespace mb {

#ifndef MB_FRAME_GLOBACC
#define MB_FRAME_GLOBACC 0  // parameter gradients: lane-private RED.ADD into an L2-resident image (1) or warp reductions + shared atomics (0)
#endif
constexpr int kMaxResidentCtas = 148 * 4;  // scratch images are sized for this many persistent CTAs
#ifndef MB_FRAME_THREADS
#define MB_FRAME_THREADS 512
#endif
constexpr int kFB = MB_FRAME_THREADS;  // threads per CTA (one CTA per SM: the frame fills most of shared memory)
constexpr int kFWarps = kFB / 32;
constexpr int kQCap = 2 * kFB;          // SR / BP / CX queues
#ifndef MB_FRAME_STENCIL
#define MB_FRAME_STENCIL 2  // cells of cutoff / 2 (falls back to 1 when the grid would not fit)
#endif
#ifndef MB_FRAME_SLICE
#define MB_FRAME_SLICE 6
#endif
constexpr int kSlice = MB_FRAME_SLICE;  // candidates one thread examines per producer step (independent, unrolled)
constexpr int kU1 = 2;                  // pairs per thread in one phase-1 batch (instruction-level parallelism)
constexpr int kNlCap = kU1 * kFB + kFB * kSlice;  // queue DB: a phase-1 backlog + one producer step
constexpr int kSrCap = kFB + kFB * kSlice;        // queue SR: a phase-2 backlog + one producer step
constexpr int kMaxCells = 2048;
constexpr int kExcl = 2;                // bonded partners per nucleotide (as the reference's (N,2) dense mask)

// Shared-memory layout.  Every fixed-size array sits at a COMPILE-TIME offset in front (FixedSmem: the kernel addresses them
// with immediates -- with run-time offsets the compiler, short of registers, re-derived these pointers from the constant
// bank all over the phases: 7 % of the issued instructions); the arrays whose size depends on n or on the number of
// parameter-gradient image copies follow at run-time offsets (FrameSmem, a kernel parameter).
constexpr unsigned smem_up(size_t bytes) { return (unsigned(bytes) + 15u) & ~15u; }
template <class T>
struct FixedSmem {
  static constexpr unsigned p = 0;
  static constexpr unsigned e = p + smem_up(sizeof(T) * MB_P_COUNT);
  static constexpr unsigned q_nl = e + smem_up(sizeof(T) * MB_N_TERMS * kFWarps);
  static constexpr unsigned q_sr = q_nl + smem_up(sizeof(uint32_t) * kNlCap);
  static constexpr unsigned q_bp = q_sr + smem_up(sizeof(uint32_t) * kSrCap);
  static constexpr unsigned q_cr = q_bp + smem_up(sizeof(uint32_t) * kQCap);
  static constexpr unsigned q_cx = q_cr + smem_up(sizeof(uint32_t) * kQCap);
  static constexpr unsigned q_ev = q_cx + smem_up(sizeof(uint32_t) * kQCap);
  static constexpr unsigned wcnt = q_ev + smem_up(sizeof(uint32_t) * kQCap);
  static constexpr unsigned ctr = wcnt + smem_up(sizeof(int) * 2 * (kFWarps + 1));
  static constexpr unsigned grid = ctr + smem_up(sizeof(int) * 8);
  static constexpr unsigned win = grid + smem_up(sizeof(T) * 8 + sizeof(int) * 8);
  static constexpr unsigned cot = win + smem_up(sizeof(T) * 2 * 9);  // nine CosWin<T>
  static constexpr unsigned cst = cot + smem_up(sizeof(T) * MB_N_TERMS);  // loop-invariant scalars (squared cutoffs / windows)
  static constexpr unsigned bar = cst + smem_up(sizeof(T) * 16);          // mbarrier of the bulk (TMA) frame staging
  static constexpr unsigned c = bar + smem_up(sizeof(uint64_t));  // the frame's centres: first of the run-time sized arrays, so its offset is still a constant
};
struct FrameSmem {
  // byte offsets into dynamic shared memory of the run-time sized arrays; computed on the host, passed as a kernel parameter
  unsigned q, back, flags, cstart, corder, excl, acc, total;
  int acc_rows;  // copies of the parameter-gradient image: warps are spread over them so that their shared-memory atomics do not collide
};
template <class T>
inline FrameSmem frame_smem_layout(int n, bool wp, bool cache_back, bool cells, int acc_rows) {
  FrameSmem L;
  unsigned off = FixedSmem<T>::c;
  auto take = [&](size_t bytes) {
    unsigned o = off;
    off += smem_up(bytes);
    return o;
  };
  L.acc_rows = acc_rows;
  take(sizeof(T) * 3 * n);  // centres at FixedSmem<T>::c
  L.q = take(sizeof(T) * 4 * n);
  L.back = take(cache_back ? sizeof(T) * 3 * n : 0);
  L.flags = take(n);
  L.cstart = take(cells ? sizeof(int) * (kMaxCells + 1) : 0);
  L.corder = take(cells ? sizeof(uint16_t) * n : 0);
  L.excl = take(cells ? sizeof(uint16_t) * kExcl * n : 0);
  L.acc = take(wp ? sizeof(T) * MB_P_COUNT * acc_rows : 0);  // acc_rows copies of the parameter-gradient image
  L.total = off;
  return L;
}

// exclusive prefix of the per-warp counts for `warp` (every lane gets it): one load + one shuffle scan
__device__ __forceinline__ int warp_prefix(const int* wcnt, int warp, int lane) {
  int x = (lane < kFWarps) ? wcnt[lane] : 0;
#pragma unroll
  for (int o = 1; o < kFWarps; o <<= 1) {
    const int y = __shfl_up_sync(kFull, x, o);
    if (lane >= o) x += y;
  }
  const int incl = __shfl_sync(kFull, x, warp);
  const int own = __shfl_sync(kFull, (lane < kFWarps) ? wcnt[lane] : 0, warp);
  return incl - own;
}

// deterministic block-wide append: entries keep (warp, lane) order
__device__ __forceinline__ void q_push(uint32_t* q, int* n, int* wcnt, bool pred, uint32_t val) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(kFull, pred);
  if (lane == 0) wcnt[warp] = __popc(m);
  __syncthreads();
  const int old = *n;
  const int before = warp_prefix(wcnt, warp, lane);
  if (pred) q[old + before + __popc(m & ((1u << lane) - 1u))] = val;
  __syncthreads();
  if (threadIdx.x == kFB - 1) *n = old + before + __popc(m);  // last warp: its prefix + own count = total
  __syncthreads();
}

// ordered block-wide append of up to U entries per thread (bit u of `bits` selects vals[u]); order = (thread, u)
template <int U>
__device__ __forceinline__ void q_push_multi(uint32_t* q, int* n, int* wcnt, unsigned bits, const uint32_t vals[U]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mine = __popc(bits);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) wcnt[warp] = incl;
  __syncthreads();
  const int old = *n;
  const int before = warp_prefix(wcnt, warp, lane);
  int off = old + before + incl - mine;
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (bits & (1u << u)) q[off++] = vals[u];
  __syncthreads();
  if (threadIdx.x == kFB - 1) *n = old + before + incl;
  __syncthreads();
}

// ordered block-wide append of up to U entries per thread into TWO queues at once (bit u of bits_a / bits_b selects
// vals[u] for queue a / b); one scan serves both (counts packed 16 + 16 bits); order = (thread, u)
template <int U>
__device__ __forceinline__ void q_push2_multi(uint32_t* qa, int* na, unsigned bits_a, uint32_t* qb, int* nb, unsigned bits_b,
                                              int* wcnt, const uint32_t vals[U]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mine = __popc(bits_a) | (__popc(bits_b) << 16);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) wcnt[warp] = incl;
  __syncthreads();
  const int old_a = *na, old_b = *nb;
  const int before = warp_prefix(wcnt, warp, lane);
  const int excl = before + incl - mine;
  int off_a = old_a + (excl & 0xffff), off_b = old_b + (excl >> 16);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (bits_a & (1u << u)) qa[off_a++] = vals[u];
    if (bits_b & (1u << u)) qb[off_b++] = vals[u];
  }
  __syncthreads();
  if (threadIdx.x == kFB - 1) {
    *na = old_a + ((before + incl) & 0xffff);
    *nb = old_b + ((before + incl) >> 16);
  }
  __syncthreads();
}

// ordered block-wide append of one entry per thread into up to THREE queues at once (one barrier set, one scan: the three
// per-warp counts are packed 10 bits each -- a CTA appends at most kFB <= 1023 entries per queue and call)
__device__ __forceinline__ void q_push3(uint32_t* qa, int* na, bool pa, uint32_t* qb, int* nb, bool pb, uint32_t* qc, int* nc, bool pc,
                                        int* wcnt, uint32_t val) {
  static_assert(kFB < 1024, "10-bit packed counts");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned ma = __ballot_sync(kFull, pa), mb_ = __ballot_sync(kFull, pb), mc = __ballot_sync(kFull, pc);
  if (lane == 0) wcnt[warp] = __popc(ma) | (__popc(mb_) << 10) | (__popc(mc) << 20);
  __syncthreads();
  const int old_a = *na, old_b = *nb, old_c = *nc;
  const int before = warp_prefix(wcnt, warp, lane);
  const unsigned below = (1u << lane) - 1u;
  if (pa) qa[old_a + (before & 1023) + __popc(ma & below)] = val;
  if (pb) qb[old_b + ((before >> 10) & 1023) + __popc(mb_ & below)] = val;
  if (pc) qc[old_c + (before >> 20) + __popc(mc & below)] = val;
  __syncthreads();
  if (threadIdx.x == kFB - 1) {  // last warp: its prefix + own count = total
    *na = old_a + (before & 1023) + __popc(ma);
    *nb = old_b + ((before >> 10) & 1023) + __popc(mb_);
    *nc = old_c + (before >> 20) + __popc(mc);
  }
  __syncthreads();
}

// the same for FOUR queues: two packed words of per-warp counts (16 bits each), one barrier set
__device__ __forceinline__ void q_push4(uint32_t* qa, int* na, bool pa, uint32_t* qb, int* nb, bool pb, uint32_t* qc, int* nc, bool pc,
                                        uint32_t* qd, int* nd, bool pd, int* wcnt, uint32_t val) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned ma = __ballot_sync(kFull, pa), mb_ = __ballot_sync(kFull, pb), mc = __ballot_sync(kFull, pc), md = __ballot_sync(kFull, pd);
  if (lane == 0) {
    wcnt[warp] = __popc(ma) | (__popc(mb_) << 16);
    wcnt[kFWarps + 1 + warp] = __popc(mc) | (__popc(md) << 16);
  }
  __syncthreads();
  const int old_a = *na, old_b = *nb, old_c = *nc, old_d = *nd;
  const int ab = warp_prefix(wcnt, warp, lane), cd = warp_prefix(wcnt + kFWarps + 1, warp, lane);
  const unsigned below = (1u << lane) - 1u;
  if (pa) qa[old_a + (ab & 0xffff) + __popc(ma & below)] = val;
  if (pb) qb[old_b + (ab >> 16) + __popc(mb_ & below)] = val;
  if (pc) qc[old_c + (cd & 0xffff) + __popc(mc & below)] = val;
  if (pd) qd[old_d + (cd >> 16) + __popc(md & below)] = val;
  __syncthreads();
  if (threadIdx.x == kFB - 1) {  // last warp: its prefix + own count = total
    *na = old_a + (ab & 0xffff) + __popc(ma);
    *nb = old_b + (ab >> 16) + __popc(mb_);
    *nc = old_c + (cd & 0xffff) + __popc(mc);
    *nd = old_d + (cd >> 16) + __popc(md);
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

// a phase's energy of this warp's pairs -> this warp's row of the per-warp energy table (the warp owns the row: plain add)
template <class T>
__device__ __forceinline__ void warp_energy_to(T v, T* row_entry) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  if ((threadIdx.x & 31) == 0) *row_entry += v;
}

template <class T>
__device__ __forceinline__ void block_sum_to(T v, T* dst) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  if ((threadIdx.x & 31) == 0 && v != T(0)) atomicAdd(dst, v);
}


// cell-grid description shared by the CTA (written by thread 0)
template <class T>
struct CellGrid {
  T origin[3], inv[3];  // cell = (x - origin) * inv
  int n[3], ncell, S;   // S = stencil half-width in cells
};

// Cell grid with stencil half-width S (cell edge >= cutoff / S).  The forward half shell of a cell is 1 + S + S(2S+1)
// (dy,dz) rows; cells that are neighbours along x have consecutive ids, so a row is one contiguous run of the
// cell-ordered list (two runs when a periodic x wraps).  Row 0 = own cell (partners after p only) and the +x cells.
struct RowRun {
  int s0, len0, s1, len1;
};
__device__ __forceinline__ int half_shell_rows(int S) { return 1 + S + S * (2 * S + 1); }

template <class T>
__device__ __forceinline__ RowRun row_run(const CellGrid<T>* grid, const int* sCstart, bool periodic, int ax, int ay, int az,
                                          int row, int own_skip) {
  RowRun r{0, 0, 0, 0};
  const int S = grid->S;
  const int n0 = grid->n[0], n1 = grid->n[1], n2 = grid->n[2];
  int dy = 0, dz = 0;
  if (row > S) {
    const int k = row - S - 1;
    dz = 1 + k / (2 * S + 1);
    dy = k % (2 * S + 1) - S;
  } else {
    dy = row;
  }
  int by = ay + dy, bz = az + dz;
  if (periodic) {
    if ((n1 == 1 && dy) || (n2 == 1 && dz)) return r;
    by = by < 0 ? by + n1 : (by >= n1 ? by - n1 : by);
    bz = bz < 0 ? bz + n2 : (bz >= n2 ? bz - n2 : bz);
  } else if (by < 0 || bz < 0 || by >= n1 || bz >= n2) {
    return r;
  }
  const int rowbase = n0 * (by + n1 * bz);
  int xlo = (row == 0) ? ax : ax - S, xhi = ax + S;  // inclusive cell range along x
  if (periodic && n0 > 1) {
    if (xlo < 0) {  // wraps on the low side: [n0+xlo .. n0-1] and [0 .. xhi]
      r.s1 = sCstart[rowbase + n0 + xlo];
      r.len1 = sCstart[rowbase + n0] - r.s1;
      xlo = 0;
    } else if (xhi >= n0) {  // wraps on the high side: [xlo .. n0-1] and [0 .. xhi-n0]
      r.s1 = sCstart[rowbase];
      r.len1 = sCstart[rowbase + xhi - n0 + 1] - r.s1;
      xhi = n0 - 1;
    }
  } else {
    xlo = xlo < 0 ? 0 : xlo;
    xhi = xhi >= n0 ? n0 - 1 : xhi;
    if (periodic) xlo = xhi = ax;  // a single cell spans a periodic x: only dx = 0
  }
  r.s0 = sCstart[rowbase + xlo];
  r.len0 = sCstart[rowbase + xhi + 1] - r.s0;
  if (row == 0) {  // own cell: members are in ascending id order, partners of p are the ones after it
    r.s0 += own_skip;
    r.len0 -= own_skip;
  }
  return r;
}

// integer cell coordinates of a position (the same arithmetic wherever it is needed, so results agree bit for bit)
template <class T>
__device__ __forceinline__ void cell_of(const CellGrid<T>* grid, const T* box, bool periodic, T px, T py, T pz, int cc[3]) {
  const T pos[3] = {px, py, pz};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    T x = pos[d] - grid->origin[d];
    if (periodic) {
      x = fmod(x, box[d]);
      if (x < T(0)) x += box[d];
    }
    int ci = int(x * grid->inv[d]);
    ci = ci < 0 ? 0 : (ci >= grid->n[d] ? grid->n[d] - 1 : ci);
    cc[d] = ci;
  }
}

// ---- bulk asynchronous copy (TMA, cp.async.bulk) global -> shared with mbarrier completion: one thread stages a whole frame
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
               "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done)
               : "r"(smem_addr(bar)), "r"(parity)
               : "memory");
  return done != 0;
}

// the observables epilogue as a real call with scalar arguments: its registers are allocated on their own and nothing of the
// energy phases' state has its address taken (which would push it into local memory)
template <class T>
__device__ __noinline__ void frame_observables_smem(const int32_t* base_pairs, const int32_t* quartets, int n_bp, int n_q, double sigma,
                                                    T box0, T box1, T box2, T back0, T back1, T back2, T base, const T* sC, const T* sQ,
                                                    T* red, T* out) {
  ObsDev o{base_pairs, quartets, n_bp, n_q, sigma};
  const T box[3] = {box0, box1, box2};
  Geom<T> g{};
  g.back[0] = back0;
  g.back[1] = back1;
  g.back[2] = back2;
  g.base = base;
  frame_observables<T>(
      o, box, [&](int i) { return smem_nuc(sC, sQ, i); }, [&](int) -> const Geom<T>& { return g; }, red, out);
}

#ifdef MB_FRAME_PROFILE
#define MB_TICK(slot) { const long long _now = clock64(); prof_t[slot] += _now - prof_last; prof_n[slot] += 1; prof_last = _now; }
#else
#define MB_TICK(slot)
#endif

template <class T, bool WP, bool CACHE_BACK, bool CELLS>
__global__ void __launch_bounds__(kFB, 1) k_frame_energy(const EnergyDev<T> a, const FrameSmem L) {
  extern __shared__ __align__(16) unsigned char smem[];
  using K = FixedSmem<T>;
  constexpr bool cells = CELLS;  // all-pairs mode (in-kernel cell list) vs an explicit pair list
  T* sC = reinterpret_cast<T*>(smem + K::c);
  T* sQ = reinterpret_cast<T*>(smem + L.q);
  T* sB = reinterpret_cast<T*>(smem + L.back);
  T* sP = reinterpret_cast<T*>(smem + K::p);
  T* sAcc = reinterpret_cast<T*>(smem + L.acc);
  T* sE = reinterpret_cast<T*>(smem + K::e);
  unsigned char* sF = smem + L.flags;  // bits 0-1 seq, bit 2 is_end
  uint32_t* qNL = reinterpret_cast<uint32_t*>(smem + K::q_nl);
  uint32_t* qSR = reinterpret_cast<uint32_t*>(smem + K::q_sr);
  uint32_t* qBP = reinterpret_cast<uint32_t*>(smem + K::q_bp);  // hydrogen-bonding candidates
  uint32_t* qCR = reinterpret_cast<uint32_t*>(smem + K::q_cr);  // cross-stacking candidates
  uint32_t* qCX = reinterpret_cast<uint32_t*>(smem + K::q_cx);
  uint32_t* qEV = reinterpret_cast<uint32_t*>(smem + K::q_ev);  // pairs with an excluded-volume site pair in range
  int* wcnt = reinterpret_cast<int*>(smem + K::wcnt);
  int* ctr = reinterpret_cast<int*>(smem + K::ctr);  // [0] n_sr [1] n_bp (hydrogen bonding) [2] n_cx [3] n_nl [4] n_cr (cross stacking) [5] n_ev (excluded volume) [6] too many bonds
  // packed cell coordinates (10 bits per axis): needed only while the cell list is built, aliases queues SR/BP/CX
  uint32_t* sCell = reinterpret_cast<uint32_t*>(smem + K::q_sr);
  int* sCstart = reinterpret_cast<int*>(smem + L.cstart);
  uint16_t* sCorder = reinterpret_cast<uint16_t*>(smem + L.corder);
  uint16_t* sExcl = reinterpret_cast<uint16_t*>(smem + L.excl);
  CellGrid<T>* grid = reinterpret_cast<CellGrid<T>*>(smem + K::grid);
  T* sCot = reinterpret_cast<T*>(smem + K::cot);
  T* sK = reinterpret_cast<T*>(smem + K::cst);  // [0] short-range centre cutoff^2 [1] [2] base-site window (low^2, high^2) [3] Debye cutoff^2 [4] all-pairs cutoff^2
  CosWin<T>* sWin = reinterpret_cast<CosWin<T>*>(smem + K::win);  // angular pre-screen of the hydrogen-bond / cross queue
  int* sCursor = reinterpret_cast<int*>(smem + K::q_nl);  // per-cell fill cursors; aliases queue NL, used only during the cell build

#ifdef MB_FRAME_PROFILE
  long long prof_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, prof_last = clock64();
  int prof_n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
  const int n = a.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // frame-independent staging: parameters, sequence / end flags, angular windows
  for (int k = threadIdx.x; k < MB_P_COUNT; k += kFB) sP[k] = a.params[k];
  for (int k = threadIdx.x; k < n; k += kFB)
    sF[k] = (unsigned char)((a.seq[k] & 3) | ((a.is_end && a.is_end[k]) ? 4 : 0));
#if MB_FRAME_GLOBACC
  // this CTA's parameter-gradient image in the workspace: zeroed here, left zeroed by every frame's flush
  T* const gimg = WP ? a.acc_scratch + (size_t)blockIdx.x * MB_P_COUNT * 32 : nullptr;
  if (WP)
    for (int k = threadIdx.x; k < MB_P_COUNT * 32; k += kFB) gimg[k] = T(0);
  GlobAcc<T> pacc{gimg};
#endif
  // bulk staging needs 16-byte aligned rows: every frame's centre block starts at a multiple of 24 n bytes
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + K::bar);
  const bool bulk = ((sizeof(T) * 3 * n) % 16 == 0) && ((reinterpret_cast<uintptr_t>(a.center) | reinterpret_cast<uintptr_t>(a.quat)) % 16 == 0) &&
                    sizeof(T) * 7 * n < (1u << 20);
  uint32_t bar_parity = 0u;
  if (bulk && threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    bp_windows9(sP, sWin);  // visible after the next barrier
    // loop-invariant scalars of the unbonded phases, kept in shared memory instead of registers that live across the loop
    // short-range centre cutoff: the widest site-pair cutoff plus both site offsets
    const Geom<T>& g = a.M.geom[0];
    const unsigned mask = a.mask;
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
    T bp_lo = T(1e30), bp_hi = T(0);
    if (mask & (1u << MB_TERM_HB)) {
      bp_lo = fmin(bp_lo, sP[MB_P_HB_RCLOW]);
      bp_hi = fmax(bp_hi, sP[MB_P_HB_RCHIGH]);
    }
    if (mask & (1u << MB_TERM_CROSS)) {
      bp_lo = fmin(bp_lo, sP[MB_P_CROSS_RCLOW]);
      bp_hi = fmax(bp_hi, sP[MB_P_CROSS_RCHIGH]);
    }
    sK[0] = sr_cut * sr_cut;
    sK[1] = bp_lo * bp_lo;
    sK[2] = bp_hi * bp_hi;  // 0: neither term requested
    sK[3] = ((mask & (1u << MB_TERM_DEBYE)) && a.M.forms[0].has_debye) ? sP[MB_P_DEBYE_RCUT] * sP[MB_P_DEBYE_RCUT] : T(-1);
    sK[4] = a.all_pairs_cutoff * a.all_pairs_cutoff;
    // squared excluded-volume site cutoffs, formed as exc_site forms them (rc * rc): the screen of phase 2 and the evaluation agree bit for bit
    sK[5] = sP[MB_P_UEXC_BACKBONE_RSTAR + 3] * sP[MB_P_UEXC_BACKBONE_RSTAR + 3];
    sK[6] = sP[MB_P_UEXC_BASE_RSTAR + 3] * sP[MB_P_UEXC_BASE_RSTAR + 3];
    sK[7] = sP[MB_P_UEXC_BACK_BASE_RSTAR + 3] * sP[MB_P_UEXC_BACK_BASE_RSTAR + 3];
    sK[8] = sP[MB_P_UEXC_BASE_BACK_RSTAR + 3] * sP[MB_P_UEXC_BASE_BACK_RSTAR + 3];
  }
  const ModelT<T>& M = a.M;
  const Geom<T>& g = M.geom[0];
  const mb_bank_forms F = M.forms[0];
  const unsigned mask = a.mask;

  // persistent CTA: frames blockIdx.x, blockIdx.x + gridDim.x, ...
  for (int frame = blockIdx.x; frame < a.n_frames; frame += gridDim.x) {
  __syncthreads();  // the previous frame is finished with shared memory
  const long long fbase = (long long)frame * n;
  if (bulk) {
    // the frame's (center, quat) rows arrive as two bulk asynchronous copies (TMA engine, 114 KB at N = 2040, float64) issued
    // by one thread; everyone waits on the mbarrier's transaction count instead of each thread moving 28 reals by hand
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy accesses of sC / sQ are ordered first
      mbar_expect_tx(bar, uint32_t(sizeof(T) * 7 * n));
      bulk_load(sC, a.center + 3 * fbase, uint32_t(sizeof(T) * 3 * n), bar);
      bulk_load(sQ, a.quat + 4 * fbase, uint32_t(sizeof(T) * 4 * n), bar);
    }
    while (!mbar_try_wait(bar, bar_parity)) {
    }
    bar_parity ^= 1u;
  } else {
    for (int k = threadIdx.x; k < 3 * n; k += kFB) sC[k] = a.center[3 * fbase + k];
    for (int k = threadIdx.x; k < 4 * n; k += kFB) sQ[k] = a.quat[4 * fbase + k];
  }
  if (WP)
    for (int k = threadIdx.x; k < MB_P_COUNT * L.acc_rows; k += kFB) sAcc[k] = T(0);
  if (threadIdx.x < 8) ctr[threadIdx.x] = 0;
  if (threadIdx.x < MB_N_TERMS * kFWarps) sE[threadIdx.x] = T(0);
  if (threadIdx.x < MB_N_TERMS) sCot[threadIdx.x] = a.cot ? a.cot[(long long)frame * MB_N_TERMS + threadIdx.x] : T(1);
  __syncthreads();

  if (CACHE_BACK) {
    for (int i = threadIdx.x; i < n; i += kFB) {
      const Nuc<T> ni = smem_nuc(sC, sQ, i);
      const V3<T> b = site(ni, g.back[0], g.back[1], g.back[2]);
      sB[3 * i] = b.x;
      sB[3 * i + 1] = b.y;
      sB[3 * i + 2] = b.z;
    }
  }
  const T* cot = sCot;
  T* const eW = sE + warp * MB_N_TERMS;  // this warp's energy row: every phase adds its batch's energy here (no registers held)
  // warp-reduced shared-memory image: the bonded phase (a burst of ~60 parameters x 4 batches from every thread at once
  // throttles the load/store unit when it goes out as RED.ADDs: measured) -- and every phase when MB_FRAME_GLOBACC is 0
  SmemAcc<T> sacc{sAcc + ((warp * L.acc_rows) / kFWarps) * MB_P_COUNT, false};  // this warp's copy of the image
#if !MB_FRAME_GLOBACC
  SmemAcc<T>& pacc = sacc;
#endif
  NullAcc nacc;
  NucGrad<T> G0, G1;  // unused (WF = false) but required by the pair drivers' signatures

  // ---------------------------------------------------------------- phase B: bonded pairs
  if (mask & MB_BONDED_TERMS) {
    T e[MB_N_TERMS];
#pragma unroll
    for (int t = 0; t < MB_N_TERMS; ++t) e[t] = T(0);
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
    warp_energy_to(e[MB_TERM_FENE], &eW[MB_TERM_FENE]);
    warp_energy_to(e[MB_TERM_BEXC], &eW[MB_TERM_BEXC]);
    warp_energy_to(e[MB_TERM_STACK], &eW[MB_TERM_STACK]);
  }

  MB_TICK(0)  // staging + bonded
  // ---------------------------------------------------------------- unbonded pairs
  // (parameter gradients of every phase leave per batch through the grouped warp reduction of SmemAcc: accumulators held in
  // registers across the whole scheduler loop -- 17 + 5 doubles -- were what pushed the loop's state into local memory)
  const bool want_debye = (mask & (1u << MB_TERM_DEBYE)) && F.has_debye;
  const bool want_sr = (mask & ((1u << MB_TERM_UEXC) | (1u << MB_TERM_HB) | (1u << MB_TERM_CROSS) | (1u << MB_TERM_COAX))) != 0;
  const bool have_unbonded = (mask & MB_UNBONDED_TERMS) && (cells || a.pair_capacity > 0);
  const bool periodic = M.box[0] > T(0);

  // ---------------------------------------------------------------- all-pairs mode: shared-memory cell list
  if (have_unbonded && cells) {
    // exclusion table from the bonded list (up to kExcl partners per nucleotide; more -> NaN energies)
    for (int k = threadIdx.x; k < kExcl * n; k += kFB) sExcl[k] = 0xffff;
    // bounding box (free space) -> grid
    T lo[3] = {T(1e30), T(1e30), T(1e30)}, hi[3] = {T(-1e30), T(-1e30), T(-1e30)};
    if (!periodic) {
      for (int i = threadIdx.x; i < n; i += kFB)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const T x = sC[3 * i + d];
          lo[d] = fmin(lo[d], x);
          hi[d] = fmax(hi[d], x);
        }
    }
    T* red = reinterpret_cast<T*>(smem + K::q_nl);  // scratch: 6 x kFWarps reals in the (still empty) DB queue; read back before the cell cursors, which alias it, are written
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      T l = lo[d], h = hi[d];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        l = fmin(l, __shfl_xor_sync(kFull, l, o));
        h = fmax(h, __shfl_xor_sync(kFull, h, o));
      }
      if (lane == 0) {
        red[(2 * d) * kFWarps + warp] = l;
        red[(2 * d + 1) * kFWarps + warp] = h;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      T ext[3];
      for (int d = 0; d < 3; ++d) {
        if (periodic) {
          grid->origin[d] = T(0);
          ext[d] = M.box[d];
        } else {
          T l = red[(2 * d) * kFWarps], h = red[(2 * d + 1) * kFWarps];
          for (int w = 1; w < kFWarps; ++w) {
            l = fmin(l, red[(2 * d) * kFWarps + w]);
            h = fmax(h, red[(2 * d + 1) * kFWarps + w]);
          }
          grid->origin[d] = l;
          ext[d] = h - l;
        }
      }
      // cells of half the cutoff (stencil +-2) when the grid fits, else of the full cutoff (stencil +-1), else coarser
      int S = MB_FRAME_STENCIL;
      T cs = a.all_pairs_cutoff * T(1.0001) / T(S);
      while (true) {
        long long tot = 1;
        for (int d = 0; d < 3; ++d) {
          int nd = periodic ? int(ext[d] / cs) : int(ext[d] / cs) + 1;
          if (nd < 1) nd = 1;
          if (periodic && nd < 2 * S + 1) nd = 1;  // too few cells along a periodic axis for the stencil: one cell spanning it
          grid->n[d] = nd;
          tot *= nd;
        }
        if (tot <= kMaxCells) {
          grid->ncell = int(tot);
          break;
        }
        if (S > 1) {
          S = 1;
          cs = a.all_pairs_cutoff * T(1.0001);
        } else {
          cs *= T(1.26);
        }
      }
      grid->S = S;
      for (int d = 0; d < 3; ++d) grid->inv[d] = periodic ? T(grid->n[d]) / ext[d] : T(1) / cs;
    }
    __syncthreads();
    const int ncell = grid->ncell;
    for (int k = threadIdx.x; k <= ncell; k += kFB) {
      sCstart[k] = 0;
      sCursor[k] = 0;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < a.n_bonded; k += kFB) {
      const int p = a.bonded[2 * k], r = a.bonded[2 * k + 1];
      for (int side = 0; side < 2; ++side) {
        const int me = side ? r : p, other = side ? p : r;
        bool placed = false;
        for (int s = 0; s < kExcl && !placed; ++s) {
          // 16-bit compare-and-swap through the containing 32-bit word
          unsigned* word = reinterpret_cast<unsigned*>(sExcl) + ((me * kExcl + s) >> 1);
          const int sh = ((me * kExcl + s) & 1) * 16;
          unsigned old = *word;
          while (((old >> sh) & 0xffffu) == 0xffffu) {
            const unsigned want = (old & ~(0xffffu << sh)) | (unsigned(other) << sh);
            const unsigned prev = atomicCAS(word, old, want);
            if (prev == old) {
              placed = true;
              break;
            }
            old = prev;
          }
        }
        if (!placed) ctr[6] = 1;  // more than kExcl bonded partners
      }
    }
    for (int i = threadIdx.x; i < n; i += kFB) {
      int cc[3];
      cell_of(grid, M.box, periodic, sC[3 * i], sC[3 * i + 1], sC[3 * i + 2], cc);
      const int cid = cc[0] + grid->n[0] * (cc[1] + grid->n[1] * cc[2]);
      sCell[i] = uint32_t(cc[0]) | (uint32_t(cc[1]) << 10) | (uint32_t(cc[2]) << 20);
      atomicAdd(&sCstart[cid + 1], 1);
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the (<= kMaxCells) cell counts by one warp
      int carry = 0;
      for (int base = 0; base < ncell; base += 32) {
        const int k = base + lane;
        const int v = (k < ncell) ? sCstart[k + 1] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(kFull, x, o);
          if (lane >= o) x += y;
        }
        if (k < ncell) sCstart[k + 1] = carry + x;
        carry += __shfl_sync(kFull, x, 31);
      }
    }
    __syncthreads();
    // members of each cell in ascending nucleotide order (deterministic): warp 0 walks the nucleotides in chunks of
    // 32; lanes that share a cell rank themselves with match_any, the lowest such lane bumps the cell's cursor
    if (warp == 0) {
      const int n0 = grid->n[0], n1 = grid->n[1];
      for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const bool valid = i < n;
        int cid = -1 - lane;  // distinct dummy keys for idle lanes
        if (valid) {
          const uint32_t cc = sCell[i];
          cid = int(cc & 1023) + n0 * (int((cc >> 10) & 1023) + n1 * int(cc >> 20));
        }
        const unsigned peers = __match_any_sync(kFull, cid);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const int leader = __ffs(peers) - 1;
        int slot = 0;
        if (valid && lane == leader) {
          slot = sCursor[cid];
          sCursor[cid] = slot + __popc(peers);
        }
        slot = __shfl_sync(kFull, slot, leader);
        if (valid) sCorder[sCstart[cid] + slot + rank] = (uint16_t)i;
        __syncwarp();
      }
    }
    __syncthreads();
  }

  MB_TICK(1)  // cell list build
  // ---------------------------------------------------------------- scheduler loop
  // Every phase body appears exactly once so it is inlined and its accumulators stay in registers.  All conditions
  // are CTA-uniform (queue counters live in shared memory, read after barriers).
  if (have_unbonded) {
    const int32_t* pl = cells ? nullptr : a.pairs + (long long)frame * a.pair_frame_stride;
    long long count = 0;
    if (!cells) {
      count = a.pair_capacity;
      if (a.pair_count) {
        const long long c = a.pair_count[frame];
        count = c < count ? c : count;
      }
    }
    long long base = 0;  // list mode: next list entry
    // all-pairs mode producer state of this thread: it walks the forward half shells of the nucleotides at cell-order
    // positions threadIdx.x, +kFB, ... as one candidate stream, kSlice candidates per step
    // (packed into five registers so that the consumer phases keep theirs: position | row << 16 | fresh << 24, k | len << 16,
    // the row's two runs as start | length << 16, cell coordinates 10 bits each)
    unsigned st_a = unsigned(threadIdx.x) | (1u << 24), st_kl = 0u, st_r0 = 0u, st_r1 = 0u, st_c = 0u;
    // list mode: the NEXT producer step's entries (i | j << 16, all ones = none) are loaded one step ahead, so that the
    // list, which streams from HBM, is never waited for.  Tagged lists (MB_NL_TAG_SUPPORTS: the neighbour build already
    // split the pairs by the terms' supports) pack i | j << 14 | Debye tag << 28 | short-range tag << 29.
    const bool tagged = a.tagged != 0;
    // (the RAW indices are what is kept across steps: decoding a value in the step that loads it would wait for the load)
    int pfi[kSlice], pfj[kSlice];
    auto issue_loads = [&](long long b0) {
#pragma unroll
      for (int u = 0; u < kSlice; ++u) {
        const long long k = b0 + threadIdx.x + (long long)u * kFB;
        pfi[u] = k < count ? pl[k] : -1;
        pfj[u] = k < count ? pl[a.pair_capacity + k] : -1;
      }
    };
    auto decode = [&](int i, int j) -> uint32_t {
      uint32_t tg = 0u;
      if (tagged && j >= 0) {
        tg = (uint32_t(j) >> 29) & 3u;  // bit 0: second sites inside the Debye cutoff, bit 1: centres inside the short-range cutoff
        j &= 0x1fffffff;
      }
      if (!(i >= 0 && j >= 0 && i < n && j < n)) return 0xffffffffu;
      return tagged ? (uint32_t(i) | (uint32_t(j) << 14) | (tg << 28)) : (uint32_t(i) | (uint32_t(j) << 16));
    };
    if (!cells) issue_loads(0);
    const int n_rows = cells ? half_shell_rows(grid->S) : 0;
    bool flush = false;
    while (true) {
      // (the six queue counters in two vector loads: this head runs ~90 times per frame with every warp waiting on it)
      const int4 c0 = *reinterpret_cast<const int4*>(ctr);
      const int2 c1 = *reinterpret_cast<const int2*>(ctr + 4);
      const int n_sr = c0.x, n_bp = c0.y, n_cx = c0.z, n_nl = c0.w, n_cr = c1.x, n_ev = c1.y;
      __syncthreads();  // everyone has read the counters before anyone updates them
      // at the end of the list the screening phase is drained FIRST: a partial batch of a downstream queue waits until nothing
      // can be appended to it any more (otherwise every downstream queue pays for two partial batches per frame)
      const bool drain = flush && n_sr == 0;
      if (n_bp >= kFB || (drain && n_bp > 0)) {
        // ---------------- phase 3a: hydrogen bonding on the HB queue (every entry passed the term's radial window and all six
        // angular windows in phase 2, so the lanes of a batch are dense in this term's code)
        const int cnt = n_bp >= kFB ? kFB : n_bp;
        const int t = threadIdx.x;
        const bool valid = t < cnt;
        const uint32_t pk = valid ? qBP[n_bp - cnt + t] : 0u;
        const int i = pk & 0xffff, j = pk >> 16;
        const Nuc<T> ni = smem_nuc(sC, sQ, i), nj = smem_nuc(sC, sQ, j);
        const V3<T> d = disp(site(nj, g.base, T(0), T(0)), site(ni, g.base, T(0), T(0)), M.box);
        const T r2 = dot(d, d);
        const T ir = (valid && r2 > T(0)) ? inv_sqrt(r2) : T(0);  // (one reciprocal square root instead of a square root and a division)
        const T r = r2 * ir;
        const bool in_hb = valid && sP[MB_P_HB_RCLOW] < r && r < sP[MB_P_HB_RCHIGH];
        const V3<T> dh = ir * d;
        HbAngles<T> A;
        A.ready = false;
        HbGrad<T> HG;
        const int tab = (sF[i] & 3) * 4 + (sF[j] & 3);
        T ev;
        if (WP)
          ev = hb_term<T, false, true>(sP, 0, in_hb, r, dh, ni.a1, nj.a1, ni.a3, nj.a3, A, tab, cot[MB_TERM_HB], HG, pacc);
        else
          ev = hb_term<T, false, false>(sP, 0, in_hb, r, dh, ni.a1, nj.a1, ni.a3, nj.a3, A, tab, cot[MB_TERM_HB], HG, nacc);
        warp_energy_to(ev, &eW[MB_TERM_HB]);
        if (threadIdx.x == 0) ctr[1] = n_bp - cnt;
        __syncthreads();
        MB_TICK(5)
        continue;
      }
      if (n_cr >= kFB || (drain && n_cr > 0)) {
        // ---------------- phase 3b: cross stacking on the CR queue
        const int cnt = n_cr >= kFB ? kFB : n_cr;
        const int t = threadIdx.x;
        const bool valid = t < cnt;
        const uint32_t pk = valid ? qCR[n_cr - cnt + t] : 0u;
        const int i = pk & 0xffff, j = pk >> 16;
        const Nuc<T> ni = smem_nuc(sC, sQ, i), nj = smem_nuc(sC, sQ, j);
        const V3<T> d = disp(site(nj, g.base, T(0), T(0)), site(ni, g.base, T(0), T(0)), M.box);
        const T r2 = dot(d, d);
        const T ir = (valid && r2 > T(0)) ? inv_sqrt(r2) : T(0);
        const T r = r2 * ir;
        const bool in_cr = valid && sP[MB_P_CROSS_RCLOW] < r && r < sP[MB_P_CROSS_RCHIGH];
        const V3<T> dh = ir * d;
        HbAngles<T> A;
        A.ready = false;
        HbGrad<T> HG;
        T ev;
        if (WP)
          ev = cross_term<T, false, true>(sP, 0, F.cross_form, in_cr, r, dh, ni.a1, nj.a1, ni.a3, nj.a3, A, cot[MB_TERM_CROSS], HG, pacc);
        else
          ev = cross_term<T, false, false>(sP, 0, F.cross_form, in_cr, r, dh, ni.a1, nj.a1, ni.a3, nj.a3, A, cot[MB_TERM_CROSS], HG, nacc);
        warp_energy_to(ev, &eW[MB_TERM_CROSS]);
        if (threadIdx.x == 0) ctr[4] = n_cr - cnt;
        __syncthreads();
        MB_TICK(7)
        continue;
      }
      if (n_cx >= kFB || (drain && n_cx > 0)) {
        // ---------------- phase 3c: coaxial stacking on the CX queue
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
        T ev;
        if (WP)
          ev = coax_term<T, false, true>(sP, 0, F.coax_form, in, ds, rs, db, ni.a1, nj.a1, ni.a3, nj.a3, cot[MB_TERM_COAX], CG, pacc);
        else
          ev = coax_term<T, false, false>(sP, 0, F.coax_form, in, ds, rs, db, ni.a1, nj.a1, ni.a3, nj.a3, cot[MB_TERM_COAX], CG, nacc);
        warp_energy_to(ev, &eW[MB_TERM_COAX]);
        if (threadIdx.x == 0) ctr[2] = n_cx - cnt;
        __syncthreads();
        MB_TICK(6)
        continue;
      }
      if (n_ev >= kFB || (drain && n_ev > 0)) {
        // ---------------- phase 2b: excluded volume on the EV queue (pairs with at least one site pair inside its cutoff:
        // about one short-range pair in a hundred, so it is evaluated with dense lanes here instead of by one or two lanes of
        // every warp of the screening phase)
        const int cnt = n_ev >= kFB ? kFB : n_ev;
        const int t = threadIdx.x;
        const bool valid = t < cnt;
        const uint32_t pk = valid ? qEV[n_ev - cnt + t] : 0u;
        const int i = pk & 0xffff, j = pk >> 16;
        const Nuc<T> ni = smem_nuc(sC, sQ, i), nj = smem_nuc(sC, sQ, j);
        const V3<T> back_i = site(ni, g.back[0], g.back[1], g.back[2]), back_j = site(nj, g.back[0], g.back[1], g.back[2]);
        const V3<T> base_i = site(ni, g.base, T(0), T(0)), base_j = site(nj, g.base, T(0), T(0));
        const T c = cot[MB_TERM_UEXC];
        V3<T> gs;
        T ex = T(0);
        if (WP) {
          ex += exc_site<T, false, true>(sP, 0, MB_P_UEXC_BACKBONE_RSTAR, MB_P_UEXC_EPS, valid, disp(back_j, back_i, M.box), c, gs, pacc);
          ex += exc_site<T, false, true>(sP, 0, MB_P_UEXC_BASE_RSTAR, MB_P_UEXC_EPS, valid, disp(base_j, base_i, M.box), c, gs, pacc);
          ex += exc_site<T, false, true>(sP, 0, MB_P_UEXC_BACK_BASE_RSTAR, MB_P_UEXC_EPS, valid, disp(back_i, base_j, M.box), c, gs, pacc);
          ex += exc_site<T, false, true>(sP, 0, MB_P_UEXC_BASE_BACK_RSTAR, MB_P_UEXC_EPS, valid, disp(base_i, back_j, M.box), c, gs, pacc);
        } else {
          ex += exc_site<T, false, false>(sP, 0, MB_P_UEXC_BACKBONE_RSTAR, MB_P_UEXC_EPS, valid, disp(back_j, back_i, M.box), c, gs, nacc);
          ex += exc_site<T, false, false>(sP, 0, MB_P_UEXC_BASE_RSTAR, MB_P_UEXC_EPS, valid, disp(base_j, base_i, M.box), c, gs, nacc);
          ex += exc_site<T, false, false>(sP, 0, MB_P_UEXC_BACK_BASE_RSTAR, MB_P_UEXC_EPS, valid, disp(back_i, base_j, M.box), c, gs, nacc);
          ex += exc_site<T, false, false>(sP, 0, MB_P_UEXC_BASE_BACK_RSTAR, MB_P_UEXC_EPS, valid, disp(base_i, back_j, M.box), c, gs, nacc);
        }
        warp_energy_to(ex, &eW[MB_TERM_UEXC]);
        if (threadIdx.x == 0) ctr[5] = n_ev - cnt;
        __syncthreads();
        MB_TICK(1)
        continue;
      }
      if (n_sr >= kFB || (flush && n_sr > 0)) {
        // ---------------- phase 2: SCREEN of the SR queue -- squared site distances against the squared cutoffs / radial windows
        // and cosine windows only (no square root, division or branch on a term's value): every lane does the same work, and
        // the survivors go to the queues of the terms that can be non-zero for them (EV / HB / CR / CX)
        const int cnt = n_sr >= kFB ? kFB : n_sr;
        const int t = threadIdx.x;
        const bool valid = t < cnt;
        const uint32_t pk = valid ? qSR[n_sr - cnt + t] : 0u;
        const int i = pk & 0xffff, j = pk >> 16;
        const Nuc<T> ni = smem_nuc(sC, sQ, i), nj = smem_nuc(sC, sQ, j);
        const V3<T> back_i = site(ni, g.back[0], g.back[1], g.back[2]), back_j = site(nj, g.back[0], g.back[1], g.back[2]);
        const V3<T> base_i = site(ni, g.base, T(0), T(0)), base_j = site(nj, g.base, T(0), T(0));
        const V3<T> d_base = disp(base_j, base_i, M.box);
        const T r2 = dot(d_base, d_base);
        bool to_ev = false;
        if (mask & (1u << MB_TERM_UEXC)) {
          const V3<T> d_bb = disp(back_j, back_i, M.box), d_bh = disp(back_i, base_j, M.box), d_hb = disp(base_i, back_j, M.box);
          to_ev = valid && (dot(d_bb, d_bb) < sK[5] || r2 < sK[6] || dot(d_bh, d_bh) < sK[7] || dot(d_hb, d_hb) < sK[8]);
        }
        unsigned to_bp = (valid && r2 > sK[1] && r2 < sK[2]) ? 1u : 0u;
        // radial window passed: cheap cosine tests of the angles decide whether the six-acos evaluations can be non-zero at
        // all (most pairs inside the window have the wrong orientation) -- bit 0 hydrogen bonding, bit 1 cross stacking
        if (to_bp) to_bp = bp_screen2(sP, sWin, mask, d_base, r2, ni.a1, nj.a1, ni.a3, nj.a3, (sF[i] & 3) * 4 + (sF[j] & 3));
        bool to_cx = false;
        if (mask & (1u << MB_TERM_COAX)) {
          const V3<T> ds = disp(site(nj, g.stack, T(0), T(0)), site(ni, g.stack, T(0), T(0)), M.box);
          const T s2 = dot(ds, ds);
          to_cx = valid && s2 > sP[MB_P_COAX_RCLOW] * sP[MB_P_COAX_RCLOW] && s2 < sP[MB_P_COAX_RCHIGH] * sP[MB_P_COAX_RCHIGH];
        }
        if (threadIdx.x == 0) ctr[0] = n_sr - cnt;
        q_push4(qBP, &ctr[1], (to_bp & 1u) != 0u, qCR, &ctr[4], (to_bp & 2u) != 0u, qCX, &ctr[2], to_cx, qEV, &ctr[5], to_ev, wcnt, pk);
        MB_TICK(4)
        continue;
      }
      if (n_nl >= kU1 * kFB || (flush && n_nl > 0)) {
        // ---------------- phase 1: Debye-Hueckel on up to kU1 in-range pairs per thread (independent chains: the CTA
        // has only 16 warps, so latency is hidden inside the thread)
        const int cnt = n_nl >= kU1 * kFB ? kU1 * kFB : n_nl;
        const int first = n_nl - cnt;
        RegAcc<T, MB_P_DEBYE_KAPPA, 5> dacc;  // the batch's kU1 pairs of this thread; reduced once per batch below
        dacc.zero();
        T ev = T(0);
#pragma unroll
        for (int u = 0; u < kU1; ++u) {
          const int t = threadIdx.x + u * kFB;
          const bool valid = t < cnt;
          const uint32_t pk = valid ? qNL[first + t] : 0u;
          const int i = pk & 0xffff, j = pk >> 16;
          V3<T> db;
          if (CACHE_BACK) {
            db = disp(v3<T>(sB[3 * j], sB[3 * j + 1], sB[3 * j + 2]), v3<T>(sB[3 * i], sB[3 * i + 1], sB[3 * i + 2]), M.box);
          } else {
            const Nuc<T> ni = smem_nuc(sC, sQ, i), nj = smem_nuc(sC, sQ, j);
            db = disp(site(nj, g.back[0], g.back[1], g.back[2]), site(ni, g.back[0], g.back[1], g.back[2]), M.box);
          }
          T m = T(1);
          if (M.half_charged_ends) m = ((sF[i] & 4) ? T(0.5) : T(1)) * ((sF[j] & 4) ? T(0.5) : T(1));
          V3<T> gd;
          if (WP)
            ev += debye_term<T, false, true>(sP, 0, valid, db, m, cot[MB_TERM_DEBYE], gd, dacc);
          else
            ev += debye_term<T, false, false>(sP, 0, valid, db, m, cot[MB_TERM_DEBYE], gd, nacc);
        }
        warp_energy_to(ev, &eW[MB_TERM_DEBYE]);
        if (WP) {
          const int idx[4] = {MB_P_DEBYE_KAPPA, MB_P_DEBYE_PREF, MB_P_DEBYE_SMOOTH, MB_P_DEBYE_RCUT};
          const T val[4] = {dacc.r[MB_P_DEBYE_KAPPA - MB_P_DEBYE_KAPPA], dacc.r[MB_P_DEBYE_PREF - MB_P_DEBYE_KAPPA],
                            dacc.r[MB_P_DEBYE_SMOOTH - MB_P_DEBYE_KAPPA], dacc.r[MB_P_DEBYE_RCUT - MB_P_DEBYE_KAPPA]};
          acc_add_group(pacc, 0, idx, val);
        }
        __syncthreads();  // the batch has been read before the counter moves
        if (threadIdx.x == 0) ctr[3] = first;
        __syncthreads();
        MB_TICK(3)
        continue;
      }
      if (flush) break;
      // ---------------- producer: kSlice candidates per thread; a candidate inside the caller's centre cutoff goes to
      // queue DB if its backbone sites are inside the Debye-Hueckel cutoff and to queue SR if its centres are inside the
      // short-range cutoff (the exact supports of the terms: everything else contributes exactly zero)
      unsigned acc_db = 0, acc_sr = 0;
      uint32_t found[kSlice];
#pragma unroll
      for (int u = 0; u < kSlice; ++u) found[u] = 0u;
      if (!cells) {
        if (base >= count) {
          flush = true;
          continue;
        }
        base += (long long)kFB * kSlice;
        uint32_t cur_[kSlice];
#pragma unroll
        for (int u = 0; u < kSlice; ++u) cur_[u] = decode(pfi[u], pfj[u]);  // loaded one step ago
        issue_loads(base);                                                  // the next step's entries, used one step from now
#pragma unroll
        for (int u = 0; u < kSlice; ++u) {
          const uint32_t cur = cur_[u];
          if (cur != 0xffffffffu && tagged) {
            found[u] = (cur & 0x3fffu) | (((cur >> 14) & 0x3fffu) << 16);
            if (want_debye && (cur & (1u << 28))) acc_db |= 1u << u;
            if (want_sr && (cur & (1u << 29))) acc_sr |= 1u << u;
          } else if (cur != 0xffffffffu) {
            const int i = int(cur & 0xffffu), j = int(cur >> 16);
            {
              found[u] = cur;
              const V3<T> dc = disp(v3<T>(sC[3 * j], sC[3 * j + 1], sC[3 * j + 2]), v3<T>(sC[3 * i], sC[3 * i + 1], sC[3 * i + 2]), M.box);
              if (want_sr && dot(dc, dc) < sK[0]) acc_sr |= 1u << u;
              if (want_debye) {
                V3<T> db;
                if (CACHE_BACK) {
                  db = disp(v3<T>(sB[3 * j], sB[3 * j + 1], sB[3 * j + 2]), v3<T>(sB[3 * i], sB[3 * i + 1], sB[3 * i + 2]), M.box);
                } else {
                  const Nuc<T> ni = smem_nuc(sC, sQ, i), nj = smem_nuc(sC, sQ, j);
                  db = disp(site(nj, g.back[0], g.back[1], g.back[2]), site(ni, g.back[0], g.back[1], g.back[2]), M.box);
                }
                if (dot(db, db) < sK[3]) acc_db |= 1u << u;
              }
            }
          }
        }
      } else {
        // CTA-uniform exit test: every thread has exhausted its stream
        const int busy = __syncthreads_or(int(st_a & 0xffffu) < n);
        if (!busy) {
          flush = true;
          continue;
        }
        int st_pos = int(st_a & 0xffffu), st_row = int((st_a >> 16) & 0xffu), st_k = int(st_kl & 0xffffu), st_len = int(st_kl >> 16);
        bool st_fresh = (st_a >> 24) != 0u;
        RowRun st_run{int(st_r0 & 0xffffu), int(st_r0 >> 16), int(st_r1 & 0xffffu), int(st_r1 >> 16)};
        int st_ax = int(st_c & 1023u), st_ay = int((st_c >> 10) & 1023u), st_az = int(st_c >> 20);
        int st_p = 0, st_e0 = 0xffff, st_e1 = 0xffff;
        T st_x = 0, st_y = 0, st_z = 0;
        if (st_pos < n && !st_fresh) {  // reload the current nucleotide
          st_p = sCorder[st_pos];
          st_x = sC[3 * st_p];
          st_y = sC[3 * st_p + 1];
          st_z = sC[3 * st_p + 2];
          const unsigned ex = *reinterpret_cast<const unsigned*>(sExcl + st_p * kExcl);
          st_e0 = int(ex & 0xffffu);
          st_e1 = int(ex >> 16);
        }
        {
          // advance to a row that still has candidates, once per step with the warp converged (a row transition inside
          // the slice would make every lane wait for every other lane's row decode); the slice ends with the row
          while (st_pos < n && st_k >= st_len) {
            int skip = 0;
            if (st_fresh) {
              st_fresh = false;
              st_p = sCorder[st_pos];
              st_x = sC[3 * st_p];
              st_y = sC[3 * st_p + 1];
              st_z = sC[3 * st_p + 2];
              const unsigned ex = *reinterpret_cast<const unsigned*>(sExcl + st_p * kExcl);  // both partners in one word
              st_e0 = int(ex & 0xffffu);
              st_e1 = int(ex >> 16);
              int cc[3];  // the build-time cell table aliases the queues, so the coordinates are recomputed
              cell_of(grid, M.box, periodic, st_x, st_y, st_z, cc);
              st_ax = cc[0];
              st_ay = cc[1];
              st_az = cc[2];
              skip = st_pos - sCstart[st_ax + grid->n[0] * (st_ay + grid->n[1] * st_az)] + 1;
              st_row = 0;
            } else {
              ++st_row;
            }
            if (st_row >= n_rows) {  // next nucleotide of this thread
              st_pos += kFB;
              st_fresh = true;
              st_k = st_len = 0;
              continue;
            }
            st_run = row_run(grid, sCstart, periodic, st_ax, st_ay, st_az, st_row, skip);
            st_k = 0;
            st_len = st_run.len0 + st_run.len1;
          }
        }
        const int st_cnt = (st_pos < n) ? ((st_len - st_k) < kSlice ? (st_len - st_k) : kSlice) : 0;
#pragma unroll
        for (int u = 0; u < kSlice; ++u) {
          if (u < st_cnt) {
            const int kk = st_k + u;
            const int r = sCorder[kk < st_run.len0 ? st_run.s0 + kk : st_run.s1 + (kk - st_run.len0)];
            if (r != st_e0 && r != st_e1) {
              T dx = sC[3 * r] - st_x;
              if (periodic) dx = wrap1(dx, M.box[0]);
              if (dx * dx < sK[4]) {
                T dy = sC[3 * r + 1] - st_y, dz = sC[3 * r + 2] - st_z;
                if (periodic) {
                  dy = wrap1(dy, M.box[1]);
                  dz = wrap1(dz, M.box[2]);
                }
                const T d2 = dx * dx + dy * dy + dz * dz;
                if (d2 < sK[4]) {
                  const int i = st_p < r ? st_p : r, j = st_p < r ? r : st_p;
                  found[u] = uint32_t(i) | (uint32_t(j) << 16);
                  if (want_sr && d2 < sK[0]) acc_sr |= 1u << u;
                  if (want_debye) {
                    V3<T> db;
                    if (CACHE_BACK) {
                      db = disp(v3<T>(sB[3 * j], sB[3 * j + 1], sB[3 * j + 2]), v3<T>(sB[3 * i], sB[3 * i + 1], sB[3 * i + 2]), M.box);
                    } else {
                      const Nuc<T> ni = smem_nuc(sC, sQ, i), nj = smem_nuc(sC, sQ, j);
                      db = disp(site(nj, g.back[0], g.back[1], g.back[2]), site(ni, g.back[0], g.back[1], g.back[2]), M.box);
                    }
                    if (dot(db, db) < sK[3]) acc_db |= 1u << u;
                  }
                }
              }
            }
          }
        }
        st_k += st_cnt;
        st_a = unsigned(st_pos) | (unsigned(st_row) << 16) | (st_fresh ? (1u << 24) : 0u);
        st_kl = unsigned(st_k) | (unsigned(st_len) << 16);
        st_r0 = unsigned(st_run.s0) | (unsigned(st_run.len0) << 16);
        st_r1 = unsigned(st_run.s1) | (unsigned(st_run.len1) << 16);
        st_c = unsigned(st_ax) | (unsigned(st_ay) << 10) | (unsigned(st_az) << 20);
      }
      q_push2_multi<kSlice>(qNL, &ctr[3], acc_db, qSR, &ctr[0], acc_sr, wcnt, found);
      MB_TICK(2)
    }
  }

#ifdef MB_FRAME_PROFILE
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("frame-kernel cycles: stage+bonded %lld | cells %lld | producer %lld (%d steps) | phase1 %lld (%d) | phase2 %lld (%d) | "
           "hb %lld (%d) | cross %lld (%d) | coax %lld (%d)\n", prof_t[0], prof_t[1], prof_t[2], prof_n[2], prof_t[3], prof_n[3], prof_t[4],
           prof_n[4], prof_t[5], prof_n[5], prof_t[7], prof_n[7], prof_t[6], prof_n[6]);
#endif
  // ---------------------------------------------------------------- flush
#if MB_FRAME_GLOBACC
  if (WP) __threadfence();  // this thread's RED.ADDs are performed before the barrier releases the readers below
#endif
  __syncthreads();
  const bool poisoned = cells && ctr[6] != 0;
  if (threadIdx.x < MB_N_TERMS && a.terms) {
    T v = 0;
    for (int w = 0; w < kFWarps; ++w) v += sE[w * MB_N_TERMS + threadIdx.x];
    if (poisoned) v = T(NAN);
    if (v != T(0)) atomicAdd(&a.terms[(long long)frame * MB_N_TERMS + threadIdx.x], v);
  }
  if (WP) {
    T* out = a.d_params + (long long)frame * a.d_params_frame_stride;
#if MB_FRAME_GLOBACC
    // one warp per parameter: sum its 32 lane slots (fixed tree), add to the frame's row, leave the image zeroed
    for (int p = warp; p < MB_P_COUNT; p += kFWarps) {
      T v = __ldcg(gimg + p * 32 + lane);
      const bool any = __any_sync(kFull, v != T(0));
      if (any) gimg[p * 32 + lane] = T(0);
      if (lane < L.acc_rows) v += sAcc[lane * MB_P_COUNT + p];  // the shared-memory image (bonded phase, register accumulators)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
      if (lane == 0 && v != T(0)) atomicAdd(&out[p], v);
    }
#else
    for (int p = threadIdx.x; p < MB_P_COUNT; p += kFB) {
      T v = T(0);
      for (int r = 0; r < L.acc_rows; ++r) v += sAcc[r * MB_P_COUNT + p];
      if (v != T(0)) atomicAdd(&out[p], v);
    }
#endif
  }
  // ---------------------------------------------------------------- fused observables
  // propeller twist / rise / pitch angle / diameter of this frame from the nucleotides already staged in shared memory
  // (the reference re-derives every site of every frame in a second vmap pass inside the loss function)
  if (a.obs_out) {
    __syncthreads();  // sE (the per-warp energy rows) becomes the reduction scratch
    frame_observables_smem<T>(a.obs.base_pairs, a.obs.quartets, a.obs.n_base_pairs, a.obs.n_quartets, a.obs.sigma_backbone, M.box[0],
                              M.box[1], M.box[2], g.back[0], g.back[1], g.back[2], g.base, sC, sQ, sE,
                              a.obs_out + (long long)frame * MB_N_OBS);
  }
  }  // persistent loop over frames
}

template <class T>
static bool pick_layout(const EnergyDev<T>& a, bool wp, bool* cache_back, FrameSmem* L) {
  const bool cells = a.all_pairs_cutoff > T(0);
  if ((long long)a.n * 11 * (long long)sizeof(T) > 227 * 1024) return false;
  if (cells && (size_t)a.n * 4 > sizeof(uint32_t) * (kSrCap + 4 * kQCap)) return false;  // sCell aliases the SR/HB/CR/CX/EV queues
  for (int cb = 1; cb >= 0; --cb) {
    for (int rows = wp ? kFWarps : 1; rows >= 1; rows >>= 1) {  // as many image copies as fit (16 = one per warp)
      *L = frame_smem_layout<T>(a.n, wp, cb != 0, cells, rows);
      if (L->total <= 227u * 1024u) {
        *cache_back = cb != 0;
        return true;
      }
    }
  }
  return false;
}

template <class T>
bool frame_kernel_eligible(const EnergyDev<T>& a) {
  if (a.M.n_banks != 1 || a.n > 60000 || (a.tagged && a.n >= 16384)) return false;  // 16-bit indices, with headroom for the producer stream position
  bool cb;
  FrameSmem L;
  return pick_layout(a, true, &cb, &L);
}

size_t frame_scratch_bytes(int real_bytes) {
  return MB_FRAME_GLOBACC ? (size_t)kMaxResidentCtas * MB_P_COUNT * 32 * (size_t)real_bytes : 0;
}

template <class T, bool WP, bool CB, bool CELLS>
static int launch_two(cudaStream_t s, const EnergyDev<T>& a, const FrameSmem& L) {
  MB_CUDA_CHECK(cudaFuncSetAttribute(k_frame_energy<T, WP, CB, CELLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  // persistent CTAs: as many as are resident at once (one per SM at the DiffTRe size), each walks frames b, b + grid, ...
  int dev = 0, n_sm = 0, per_sm = 0;
  MB_CUDA_CHECK(cudaGetDevice(&dev));
  MB_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  MB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_frame_energy<T, WP, CB, CELLS>, kFB, L.total));
  long long grid = (long long)n_sm * (per_sm > 0 ? per_sm : 1);
  if (grid > kMaxResidentCtas) grid = kMaxResidentCtas;
  if (grid > a.n_frames) grid = a.n_frames;
  MB_REQUIRE(!(WP && MB_FRAME_GLOBACC) || a.acc_scratch, MB_ECAPACITY,
             "frame kernel: parameter gradients need the workspace of mythos_b200_energy_workspace_bytes()");
  k_frame_energy<T, WP, CB, CELLS><<<(unsigned)grid, kFB, L.total, s>>>(a, L);
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}
template <class T, bool WP, bool CB>
static int launch_one(cudaStream_t s, const EnergyDev<T>& a, const FrameSmem& L) {
  return a.all_pairs_cutoff > T(0) ? launch_two<T, WP, CB, true>(s, a, L) : launch_two<T, WP, CB, false>(s, a, L);
}

template <class T>
int launch_frame_kernel(cudaStream_t s, const EnergyDev<T>& a, bool wp) {
  bool cb;
  FrameSmem L;
  MB_REQUIRE(pick_layout(a, wp, &cb, &L), MB_ECAPACITY, "frame kernel: frame does not fit in shared memory");
  if (wp) return cb ? launch_one<T, true, true>(s, a, L) : launch_one<T, true, false>(s, a, L);
  return cb ? launch_one<T, false, true>(s, a, L) : launch_one<T, false, false>(s, a, L);
}

template <class T>
bool frame_kernel_fits(int n, bool wp) {
  EnergyDev<T> a{};
  a.n = n;
  a.M.n_banks = 1;
  a.all_pairs_cutoff = T(0);
  bool cb;
  FrameSmem L;
  return n <= 60000 && pick_layout(a, wp, &cb, &L);
}
template bool frame_kernel_fits<float>(int, bool);
template bool frame_kernel_fits<double>(int, bool);
template bool frame_kernel_eligible<float>(const EnergyDev<float>&);
template bool frame_kernel_eligible<double>(const EnergyDev<double>&);
template int launch_frame_kernel<float>(cudaStream_t, const EnergyDev<float>&, bool);
template int launch_frame_kernel<double>(cudaStream_t, const EnergyDev<double>&, bool);

}  // namespace mb
