// neighbors.cu -- batched cell-list neighbour build: integer binning, counting sort into cell order, count / scan / fill.
//
// Replaces mythos/utils/neighbors.py:12-59, i.e. jax_md.partition.neighbor_list(format=OrderedSparse,
// disable_cell_list=True, custom_mask_function=bonded mask), whose per-rebuild cost is an O(N^2) distance
// matrix.  Cells are only a superset generator: the accept test is the reference's arithmetic
//   dR = Ri - Rj (periodic: floor-mod wrap to [-L/2, L/2));  d2 = dx*dx + dy*dy + dz*dz;
//   keep iff d2 < (r_cutoff + dr_threshold)^2 (strict), i < j, (i,j) not bonded
// evaluated in the positions' dtype with explicit round-to-nearest mul/add (no FMA contraction), so the pair
// SET is bit-exact against an all-pairs evaluation of the same expression.
//
// All frames of a call are processed together (every kernel is one launch over F x N nucleotides or F x C cells):
//   1. bounds    per-frame bounding box (free space) by block reduction + ordered-integer atomics
//   2. grid      per frame: cell edge = cutoff / S (stencil half-width S = 2 when the cell table allows, else 1,
//                else coarser), grid dimensions, origin
//   3. bin       integer cell id per nucleotide, cell histogram
//   4. scan      one exclusive scan over all F x C cell counts: a cell's start in the frame-major sorted order
//   5. scatter   nucleotide ids into their cell (atomic cursor), then rank inside the cell by id (deterministic) and
//                write the cell-ordered records (coordinates + id, one 128/256-bit row each)
//   6. count     one thread per cell-ordered nucleotide walks its forward half shell: (1 + S + S(2S+1)) rows, each one
//                contiguous run of the sorted arrays (x-neighbouring cells have consecutive ids); neighbouring lanes sit
//                in the same or adjacent cells, so candidate loads are broadcasts out of L1
//   7. scan      exclusive scan of the per-nucleotide counts -> offsets; per-frame totals -> count[frame]
//   8. fill      same walk replaying the accept bits the count pass recorded (no coordinates, no arithmetic), writing
//                (min(i,j), max(i,j)) at the nucleotide's offset four pairs per 128-bit store; tail padded with N
// Support tags (MB_NL_TAG_SUPPORTS, an internal contract with this library's frame-resident energy kernel): the build ORs
// constant tag bits into the second index of every pair it writes and can append after the entries a previous build left
// in the same list.  The host runs it twice -- on the centres with the short-range cutoff (tag bit 30) and on the backbone
// sites with the Debye-Hueckel cutoff (tag bit 29) -- so that only pairs inside the support of some term are ever
// written, each already labelled with the phase queue it belongs to; every other pair contributes exactly zero.
// Warp-slot mode (MB_NL_WARP_SLOTS) replaces 6-8 by ONE walk: the partners a lane accepts are staged in shared memory and
// the warp writes its pairs, coalesced and in a fixed order, into its own fixed-width slot of the list, padding the rest of
// the slot with N.  No count pass, no scan, no replay -- at the price of ~20 % padding entries the consumers skip.
// Rows mode (MB_NL_ROWS) replaces 6-8 by ONE walk that writes a fixed-width row per nucleotide with the unused slots set
// to the padding value N: still a valid OrderedSparse list (consumers mask entries >= N wherever they are), at half the
// cost -- the shape the energy kernels of this library are fed with.
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace mb {

constexpr int kNlBlock = 128;
constexpr int kMaxExcl = 4;
constexpr int kNlUnroll = 4;      // candidates in flight per thread in the walk
constexpr int kNlBitWords = 8;    // accept bits the count pass records per nucleotide for the fill pass (32 each)
constexpr int kNlBitCap = 32 * kNlBitWords;
constexpr int kScanChunk = 2048;  // elements per block of the multi-block scan (256 threads x 8)

template <class T>
struct NlGrid {
  T origin[3], inv[3], width[3];
  int n[3], S, ncell, _pad;
};

// cell-ordered record of one nucleotide: coordinates + id in one 16-byte-aligned row (128-bit loads)
template <class T>
struct NlRec;
template <>
struct alignas(16) NlRec<double> {
  double x, y, z;
  int32_t id, pad;
};
template <>
struct alignas(16) NlRec<float> {
  float x, y, z;
  int32_t id;
};

template <class T>
struct NlDev {
  int n, n_frames, n_bonded, cmax;  // cmax = cell-table entries per frame
  const T* center;
  const int32_t* bonded;
  T box[3];
  int periodic;
  T cut;    // r_cutoff + dr_threshold
  T cut2;   // its square, in T
  int32_t* pairs;
  long long capacity;
  int32_t* count;
  int32_t* overflow;
  int32_t* max_row;  // rows mode: (F) longest row found, or nullptr
  int lane_slots;                // warp-slot mode: partner ids staged per lane in shared memory
  long long slot_base, slot_width;  // warp-slot mode: this build's slots start at slot_base of each frame's list
  uint32_t tag_bits;             // OR-ed into the second index of every pair written (MB_NL_TAG_SUPPORTS)
  const int32_t* append_count;   // (F) entries already in each frame's list: this build appends after them, or nullptr
  T* reference;                  // (F,N,3) centres at the last build: conditional rebuild (k_nl_frame), or nullptr
  T move2;                       // squared displacement beyond which a frame is rebuilt
  int32_t* rebuilds;             // (F) rebuild counters, or nullptr
  int packed;                    // k_nl_frame: warp slots packed back to back (MB_NL_PACKED_SLOTS)
  // workspace
  int32_t* excl;       // (N, kMaxExcl)
  unsigned long long* bounds;  // (F, 6) ordered-integer min / max corners
  NlGrid<T>* grid;     // (F)
  int32_t* cell;       // (F*N) cell id of each nucleotide
  int32_t* cstart;     // (F*cmax + 1) histogram, then exclusive scan (global offsets into the sorted arrays)
  int32_t* cursor;     // (F*cmax)
  int32_t* tmp_order;  // (F*N) ids in cell order, unsorted inside a cell
  NlRec<T>* srec;      // (F*N) cell-ordered (coordinates, id), ids ascending inside a cell
  int32_t* nbcount;    // (F*N + 1) per-nucleotide pair counts (cell order), then exclusive scan
  uint32_t* bitbuf;    // (kNlBitWords, F*N) accept bits of the count pass, word-major
  int32_t* scan_tmp;   // block sums of the multi-block scan
};

// ------------------------------------------------------------------------------------------------ exclusions
__global__ void k_nl_excl_init(int32_t* excl, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n * kMaxExcl) excl[k] = -1;
}
__global__ void k_nl_excl_fill(int32_t* excl, const int32_t* bonded, int nb, int n, int32_t* overflow) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nb) return;
  const int a = bonded[2 * k], b = bonded[2 * k + 1];
  if (a < 0 || b < 0 || a >= n || b >= n) return;
  for (int side = 0; side < 2; ++side) {
    const int me = side ? b : a, other = side ? a : b;
    bool placed = false;
    for (int s = 0; s < kMaxExcl && !placed; ++s) placed = (atomicCAS(&excl[me * kMaxExcl + s], -1, other) == -1);
    if (!placed) atomicOr(overflow, 2);
  }
}

// ------------------------------------------------------------------------------------------------ bounds
// order-preserving map of a real onto an unsigned integer, so that min / max can be taken with integer atomics
__device__ __forceinline__ unsigned long long ord_encode(double v) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_decode(unsigned long long e) {
  const unsigned long long u = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
  return __longlong_as_double((long long)u);
}

__global__ void k_nl_bounds_init(unsigned long long* bounds, int n_frames) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n_frames * 6) bounds[k] = (k % 6 < 3) ? ~0ull : 0ull;
}

template <class T>
__global__ void k_nl_bounds(NlDev<T> a) {
  const int f = blockIdx.y;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
    const T* c = a.center + 3ll * ((long long)f * a.n + i);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double x = double(c[d]);
      lo[d] = x < lo[d] ? x : lo[d];
      hi[d] = x > hi[d] ? x : hi[d];
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double l = __shfl_xor_sync(0xffffffffu, lo[d], o), h = __shfl_xor_sync(0xffffffffu, hi[d], o);
      lo[d] = l < lo[d] ? l : lo[d];
      hi[d] = h > hi[d] ? h : hi[d];
    }
    if ((threadIdx.x & 31) == 0) {
      if (lo[d] < 1e299) atomicMin(&a.bounds[6 * f + d], ord_encode(lo[d]));
      if (hi[d] > -1e299) atomicMax(&a.bounds[6 * f + 3 + d], ord_encode(hi[d]));
    }
  }
}

// grid dimensions of one frame from its bounding box (free space) or the box (periodic)
template <class T>
__device__ __forceinline__ NlGrid<T> make_grid(const NlDev<T>& a, const double lo[3], const double hi[3]) {
  NlGrid<T> g;
  T ext[3];
  for (int d = 0; d < 3; ++d) {
    if (a.periodic) {
      g.origin[d] = T(0);
      ext[d] = a.box[d];
    } else {
      const double l = lo[d], h = hi[d];
      g.origin[d] = T(l);
      ext[d] = T(h - l);
      if (!(ext[d] >= T(0))) ext[d] = T(0);
    }
  }
  int S = 2;
  T cs = a.cut * T(1.0001) / T(S);  // a hair wider than needed so rounding in the binning can never hide a pair
  while (true) {
    long long tot = 1;
    for (int d = 0; d < 3; ++d) {
      long long nd = a.periodic ? (long long)(ext[d] / cs) : (long long)(ext[d] / cs) + 1;
      if (nd < 1) nd = 1;
      if (a.periodic && nd < 2 * S + 1) nd = 1;  // too few cells along a periodic axis for the stencil: one cell spans it
      if (nd > 1 << 20) nd = 1 << 20;
      g.n[d] = int(nd);
      tot *= nd;
      if (tot > (1ll << 40)) tot = 1ll << 40;
    }
    if (tot <= a.cmax) {
      g.ncell = int(tot);
      break;
    }
    if (S > 1) {
      S = 1;
      cs = a.cut * T(1.0001);
    } else {
      cs *= T(1.26);
    }
  }
  g.S = S;
  for (int d = 0; d < 3; ++d) {
    g.width[d] = a.periodic ? ext[d] / T(g.n[d]) : cs;
    g.inv[d] = T(1) / g.width[d];
  }
  g._pad = 0;
  return g;
}

// one thread per frame: grid dimensions
template <class T>
__global__ void k_nl_grid(NlDev<T> a) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= a.n_frames) return;
  double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  if (!a.periodic)
    for (int d = 0; d < 3; ++d) {
      lo[d] = ord_decode(a.bounds[6 * f + d]);
      hi[d] = ord_decode(a.bounds[6 * f + 3 + d]);
    }
  a.grid[f] = make_grid(a, lo, hi);
}

template <class T>
__device__ __forceinline__ int cell_id(const NlDev<T>& a, const NlGrid<T>& g, const T* c, int cc[3]) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    T x = c[d] - g.origin[d];
    if (a.periodic) {  // positions may lie outside the primary image
      x = fmod(x, a.box[d]);
      if (x < T(0)) x += a.box[d];
    }
    int ci = int(x * g.inv[d]);
    ci = ci < 0 ? 0 : (ci >= g.n[d] ? g.n[d] - 1 : ci);
    cc[d] = ci;
  }
  return cc[0] + g.n[0] * (cc[1] + g.n[1] * cc[2]);
}

template <class T>
__global__ void k_nl_bin(NlDev<T> a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.n * a.n_frames) return;
  const int f = int(idx / a.n);
  int cc[3];
  const int cid = cell_id(a, a.grid[f], a.center + 3 * idx, cc);
  a.cell[idx] = cid;
  atomicAdd(&a.cstart[(long long)f * a.cmax + cid], 1);
}

// ------------------------------------------------------------------------------------------------ multi-block scan
// exclusive scan of data[0..n) in place; data[n] receives the total.  Three launches: local scans + block sums,
// scan of the block sums (one block), add.
__global__ void k_scan_local(int32_t* data, long long n, int32_t* sums) {
  __shared__ int32_t swarp[8];
  const long long base = (long long)blockIdx.x * kScanChunk + threadIdx.x * 8;
  int32_t v[8];
  int32_t t = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v[k] = (base + k < n) ? data[base + k] : 0;
    t += v[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t x = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) swarp[warp] = x;
  __syncthreads();
  int32_t before = 0;
  for (int w = 0; w < warp; ++w) before += swarp[w];
  int32_t run = before + x - t;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (base + k < n) data[base + k] = run;
    run += v[k];
  }
  if (threadIdx.x == 255) sums[blockIdx.x] = run;
}
__global__ void k_scan_sums(int32_t* sums, int nblocks) {
  __shared__ int32_t swarp[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int base = 0; base < nblocks; base += blockDim.x) {
    const int k = base + threadIdx.x;
    const int32_t v = (k < nblocks) ? sums[k] : 0;
    int32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) swarp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int32_t w = (lane < nw) ? swarp[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      swarp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int32_t before = carry + (warp ? swarp[warp - 1] : 0);
    if (k < nblocks) sums[k] = before + x - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += swarp[nw - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[nblocks] = carry;
}
__global__ void k_scan_add(int32_t* data, long long n, const int32_t* sums, int nblocks) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) data[k] += sums[k / kScanChunk];
  if (k == 0) data[n] = sums[nblocks];
}
static void scan_exclusive(cudaStream_t s, int32_t* data, long long n, int32_t* tmp) {
  const int nblocks = int((n + kScanChunk - 1) / kScanChunk);
  k_scan_local<<<nblocks, 256, 0, s>>>(data, n, tmp);
  k_scan_sums<<<1, 1024, 0, s>>>(tmp, nblocks);
  k_scan_add<<<ceil_div(n, 256), 256, 0, s>>>(data, n, tmp, nblocks);
}

// ------------------------------------------------------------------------------------------------ cell order
template <class T>
__global__ void k_nl_scatter(NlDev<T> a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.n * a.n_frames) return;
  const int f = int(idx / a.n);
  const long long c = (long long)f * a.cmax + a.cell[idx];
  const int pos = a.cstart[c] + atomicAdd(&a.cursor[c], 1);
  a.tmp_order[pos] = int(idx - (long long)f * a.n);
}

// rank inside the cell by id (members of a cell: a few to a few dozen), write the sorted copies
template <class T>
__global__ void k_nl_rank(NlDev<T> a) {
  const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= (long long)a.n * a.n_frames) return;
  const int f = int(pos / a.n);
  const int id = a.tmp_order[pos];
  const long long c = (long long)f * a.cmax + a.cell[(long long)f * a.n + id];
  const int lo = a.cstart[c], hi = a.cstart[c + 1];
  int rank = 0;
  for (int q = lo; q < hi; ++q) rank += a.tmp_order[q] < id;
  const T* ctr = a.center + 3 * ((long long)f * a.n + id);
  NlRec<T> r{};
  r.x = ctr[0];
  r.y = ctr[1];
  r.z = ctr[2];
  r.id = id;
  a.srec[lo + rank] = r;
}

// ------------------------------------------------------------------------------------------------ walk
template <class T>
__device__ __forceinline__ T wrap_nl(T d, T L) {
  T s = fmod(d + T(0.5) * L, L);
  if (s != T(0) && s < T(0)) s += L;
  return s - T(0.5) * L;
}
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

// MODE 0: count, 1: fill (compact list), 2: rows (one pass: fixed-width row per nucleotide, slot-major)
template <class T, int MODE, bool PERIODIC>
__global__ void __launch_bounds__(kNlBlock) k_nl_walk(NlDev<T> a) {
  constexpr bool FILL = MODE == 1, ROWS = MODE == 2, SLOTS = MODE == 3;
  extern __shared__ int32_t stage_raw[];  // SLOTS: per warp, 32 lanes x lane_slots partner ids
  long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int f;
  bool live = true;
  long long gw = 0;  // SLOTS: warp index within the frame
  if (SLOTS) {
    // warps never straddle two frames: warp w of frame f covers cell-order positions [32 w, 32 w + 32) of that frame
    const long long wpf = (a.n + 31) / 32, warp_id = pos >> 5;
    f = int(warp_id / wpf);
    gw = warp_id - (long long)f * wpf;
    if (f >= a.n_frames) return;  // whole warp
    const long long in_frame = gw * 32 + (threadIdx.x & 31);
    live = in_frame < a.n;
    pos = (long long)f * a.n + (live ? in_frame : 0);
  } else {
    if (pos >= (long long)a.n * a.n_frames) return;
    f = int(pos / a.n);
  }
  const NlGrid<T> g = a.grid[f];
  const NlRec<T> me = a.srec[pos];
  const int i = me.id;
  const T xi = me.x, yi = me.y, zi = me.z;
  int cc[3];
  {
    const T c[3] = {xi, yi, zi};
    cell_id(a, g, c, cc);
  }
  const int32_t* cs = a.cstart + (long long)f * a.cmax;
  int ex[kMaxExcl];
#pragma unroll
  for (int s = 0; s < kMaxExcl; ++s) ex[s] = a.excl[i * kMaxExcl + s];
  int found = 0;
  long long wpos = 0;
  int32_t* out0 = nullptr;
  int32_t* out1 = nullptr;
  if (FILL) {
    wpos = a.nbcount[pos] - a.nbcount[(long long)f * a.n] + (a.append_count ? a.append_count[f] : 0);
    out0 = a.pairs + (long long)f * 2 * a.capacity;
    out1 = out0 + a.capacity;
  }
  // rows mode: slot k of the nucleotide at cell-order position p (within its frame) is entry k * N + p, so that the
  // lanes of a warp, which advance through their slots at about the same pace, write neighbouring addresses
  const int row_width = ROWS ? int(a.capacity / a.n) : 0;
  if (ROWS) {
    out0 = a.pairs + (long long)f * 2 * a.capacity + (pos - (long long)f * a.n);
    out1 = out0 + a.capacity;
  }
  const long long R = (long long)a.n * a.n_frames;
  int ci = 0;        // running candidate index of this thread (same enumeration in the count and fill passes)
  unsigned bw = 0u;  // current word of accept bits
  // accept test of one candidate record; receiver = lower index (OrderedSparse keeps i < j): dR = R_low - R_high, as the
  // reference evaluates it (in free space the two orders give exact negatives, whose squares are identical)
  const int tag_bits = int(a.tag_bits);
  auto accept = [&](const NlRec<T>& c, int& j, bool& me_low) -> bool {
    j = c.id;
    me_low = i < j;
    T ddx = xi - c.x, ddy = yi - c.y, ddz = zi - c.z;
    if (PERIODIC) {
      if (!me_low) {
        ddx = c.x - xi;
        ddy = c.y - yi;
        ddz = c.z - zi;
      }
      ddx = wrap_nl(ddx, a.box[0]);
      ddy = wrap_nl(ddy, a.box[1]);
      ddz = wrap_nl(ddz, a.box[2]);
    }
    const T d2 = add_rn(add_rn(mul_rn(ddx, ddx), mul_rn(ddy, ddy)), mul_rn(ddz, ddz));
    return d2 < a.cut2 && j != ex[0] && j != ex[1] && j != ex[2] && j != ex[3];
  };
  // fill pass: accepted pairs go out four at a time as 128-bit stores once the write position is 16-byte aligned
  int l0 = 0, l1 = 0, l2 = 0, l3 = 0, h0 = 0, h1 = 0, h2 = 0, h3 = 0, nbuf = 0;
  const bool vec_ok = FILL && (a.capacity & 3) == 0 && (reinterpret_cast<unsigned long long>(a.pairs) & 15ull) == 0;
  auto emit = [&](int lo, int hi) {
    l0 = l1; l1 = l2; l2 = l3; l3 = lo;
    h0 = h1; h1 = h2; h2 = h3; h3 = hi;
    ++nbuf;
    ++found;
    if (!vec_ok || (wpos & 3) != 0 || wpos + 4 > a.capacity) {  // scalar: unaligned head, odd capacity, or the truncated end
      if (wpos < a.capacity) {
        out0[wpos] = lo;
        out1[wpos] = hi;
      }
      ++wpos;
      nbuf = 0;
    } else if (nbuf == 4) {
      *reinterpret_cast<int4*>(out0 + wpos) = make_int4(l0, l1, l2, l3);
      *reinterpret_cast<int4*>(out1 + wpos) = make_int4(h0, h1, h2, h3);
      wpos += 4;
      nbuf = 0;
    }
  };
  const int S = g.S, n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const int n_rows = live ? 1 + S + S * (2 * S + 1) : 0;
  const int lane_slots = SLOTS ? a.lane_slots : 0;
  int32_t* stage = SLOTS ? stage_raw + ((threadIdx.x >> 5) * 32 + (threadIdx.x & 31)) * lane_slots : nullptr;
  for (int row = 0; row < n_rows; ++row) {
    int dy = row, dz = 0;
    if (row > S) {
      const int k = row - S - 1;
      dz = 1 + k / (2 * S + 1);
      dy = k % (2 * S + 1) - S;
    }
    int by = cc[1] + dy, bz = cc[2] + dz;
    if (PERIODIC) {
      if ((n1 == 1 && dy) || (n2 == 1 && dz)) continue;
      by = by < 0 ? by + n1 : (by >= n1 ? by - n1 : by);
      bz = bz < 0 ? bz + n2 : (bz >= n2 ? bz - n2 : bz);
    } else if (by < 0 || bz < 0 || by >= n1 || bz >= n2) {
      continue;
    }
    const int rowbase = n0 * (by + n1 * bz);
    int xlo = (row == 0) ? cc[0] : cc[0] - S, xhi = cc[0] + S;
    // up to two runs of the sorted arrays (a periodic x wraps)
    int rs[2] = {0, 0}, re[2] = {0, 0};
    if (PERIODIC && n0 > 1) {
      if (xlo < 0) {
        rs[1] = cs[rowbase + n0 + xlo];
        re[1] = cs[rowbase + n0];
        xlo = 0;
      } else if (xhi >= n0) {
        rs[1] = cs[rowbase];
        re[1] = cs[rowbase + xhi - n0 + 1];
        xhi = n0 - 1;
      }
    } else {
      xlo = xlo < 0 ? 0 : xlo;
      xhi = xhi >= n0 ? n0 - 1 : xhi;
      if (PERIODIC) xlo = xhi = cc[0];
    }
    rs[0] = cs[rowbase + xlo];
    re[0] = cs[rowbase + xhi + 1];
    if (row == 0) rs[0] = int(pos) + 1 > rs[0] ? int(pos) + 1 : rs[0];  // own cell: ids ascend, partners come after
#pragma unroll
    for (int run = 0; run < 2; ++run) {
      if (FILL) {
        // second pass: replay the accept bits of the count pass (no coordinates, no arithmetic); candidates beyond the
        // recorded kNlBitCap are re-tested
        int q = rs[run];
        while (q < re[run]) {
          if (ci >= kNlBitCap) {
            for (; q < re[run]; ++q) {
              int j;
              bool me_low;
              if (accept(a.srec[q], j, me_low)) emit(me_low ? i : j, (me_low ? j : i) | tag_bits);
            }
            break;
          }
          if ((ci & 31) == 0) bw = a.bitbuf[(long long)(ci >> 5) * R + pos];
          int avail = 32 - (ci & 31);
          avail = avail < re[run] - q ? avail : re[run] - q;
          unsigned m = (bw >> (ci & 31)) & (avail == 32 ? 0xffffffffu : ((1u << avail) - 1u));
          while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1u;
            const int j = a.srec[q + bit].id;
            emit(i < j ? i : j, (i < j ? j : i) | tag_bits);
          }
          q += avail;
          ci += avail;
        }
      } else {
        for (int q0 = rs[run]; q0 < re[run]; q0 += kNlUnroll) {
          NlRec<T> c[kNlUnroll];
#pragma unroll
          for (int u = 0; u < kNlUnroll; ++u) {  // independent 128-bit loads first, tests after
            const int q = q0 + u < re[run] ? q0 + u : re[run] - 1;
            c[u] = a.srec[q];
          }
#pragma unroll
          for (int u = 0; u < kNlUnroll; ++u) {
            int j;
            bool me_low;
            const bool in = q0 + u < re[run];
            const bool ok = accept(c[u], j, me_low) && in;
            if (MODE == 0 && in) {  // record the decision for the fill pass
              if (ci < kNlBitCap) {
                bw |= (ok ? 1u : 0u) << (ci & 31);
                if ((ci & 31) == 31) {
                  a.bitbuf[(long long)(ci >> 5) * R + pos] = bw;
                  bw = 0u;
                }
              }
              ++ci;
            }
            if (ok) {
              if (ROWS && found < row_width) {
                out0[(long long)found * a.n] = me_low ? i : j;
                out1[(long long)found * a.n] = me_low ? j : i;
              }
              if (SLOTS && found < lane_slots) stage[found] = j;
              ++found;
            }
          }
        }
      }
    }
  }
  if (MODE == 0 && (ci & 31) != 0 && ci < kNlBitCap) a.bitbuf[(long long)(ci >> 5) * R + pos] = bw;
  if (FILL) {  // the last 1..3 buffered pairs
    const int lo_[4] = {l0, l1, l2, l3}, hi_[4] = {h0, h1, h2, h3};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k >= 4 - nbuf) {
        if (wpos < a.capacity) {
          out0[wpos] = lo_[k];
          out1[wpos] = hi_[k];
        }
        ++wpos;
      }
  }
  if (MODE == 0) a.nbcount[pos] = found;
  if (SLOTS) {
    // One pass, no scan, no second walk: the warp's pairs go into the warp's own fixed-width slot of the frame's list
    // (slot w = entries [slot_base + w * slot_width, + slot_width)); what is left of the slot is padded with N, which every
    // consumer skips.  Deterministic (lane-major inside the slot).  A lane with more partners than lane_slots or a warp with
    // more pairs than slot_width sets overflow bits 2 / 0; the longest lane row / warp total go to max_row for resizing.
    const int lane = threadIdx.x & 31;
    const int mine = found < lane_slots ? found : lane_slots;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int longest = found;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int y = __shfl_xor_sync(0xffffffffu, longest, o);
      longest = y > longest ? y : longest;
    }
    int32_t* o0 = a.pairs + (long long)f * 2 * a.capacity + a.slot_base + gw * a.slot_width;
    int32_t* o1 = o0 + a.capacity;
    __syncwarp();
    const int32_t* wstage = stage_raw + (threadIdx.x >> 5) * 32 * lane_slots;
    for (int L = 0; L < 32; ++L) {  // owner by owner: the owner's partners are one coalesced run
      const int cnt = __shfl_sync(0xffffffffu, mine, L), off = __shfl_sync(0xffffffffu, incl - mine, L);
      const int iL = __shfl_sync(0xffffffffu, i, L);
      for (int k = lane; k < cnt; k += 32) {
        if (off + k < a.slot_width) {
          const int j = wstage[L * lane_slots + k];
          o0[off + k] = iL < j ? iL : j;
          o1[off + k] = (iL < j ? j : iL) | tag_bits;
        }
      }
    }
    for (int k = total + lane; k < a.slot_width; k += 32) {
      o0[k] = a.n;
      o1[k] = a.n;
    }
    if (lane == 0) {
      if (total > a.slot_width) atomicOr(a.overflow, 1);
      if (longest > lane_slots) atomicOr(a.overflow, 4);
      atomicAdd(&a.count[f], total < a.slot_width ? total : a.slot_width);
      if (a.max_row) {
        atomicMax(&a.max_row[2 * f], longest);
        atomicMax(&a.max_row[2 * f + 1], total);
      }
    }
  }
  if (ROWS) {
    for (int k = found; k < row_width; ++k) {  // unused slots carry the padding value N, as the tail of a compact list
      out0[(long long)k * a.n] = a.n;
      out1[(long long)k * a.n] = a.n;
    }
    if (found > row_width) atomicOr(a.overflow, 1);
    // per-frame totals and the longest row (so that the caller can size the rows), one atomic per warp where possible
    int tot = found, mx = found;
    const unsigned am = __activemask();
    const int f0 = __shfl_sync(am, f, __ffs(am) - 1);
    const bool uniform = am == 0xffffffffu && __all_sync(am, f == f0);
    if (uniform) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        tot += __shfl_xor_sync(0xffffffffu, tot, o);
        const int m = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = m > mx ? m : mx;
      }
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&a.count[f], tot);
        if (a.max_row) atomicMax(&a.max_row[f], mx);
      }
    } else {
      atomicAdd(&a.count[f], found);
      if (a.max_row) atomicMax(&a.max_row[f], found);
    }
  }
}

// ------------------------------------------------------------------------------------------------ frame-resident build
// Warp-slot builds of systems whose cell table and cell-ordered records fit in shared memory (free space): ONE launch does
// what bounds / grid / bin / scan / scatter / rank / walk do above in eleven, a CTA per frame (persistent, two CTAs per SM):
//   bounds (block reduction) -> grid -> cell histogram (shared atomics) -> in-place exclusive scan -> scatter (the same
//   array is the cursor: after the scatter entry c+1 holds the end of cell c) -> rank by id inside the cell, records into
//   shared memory -> walk out of shared memory.
// The walk first collects a lane's non-empty candidate runs (rows of the forward half shell, unrolled for the stencil half
// width S: no division, one row base) into shared memory, then runs ONE flat loop over its candidates: the lanes of a warp
// stay together for max(total candidates) iterations instead of sum over rows of max(row candidates).
// Same grid, same cell order, same warp partition, same order inside a slot as k_nl_walk<T,3,false>: identical lists.
constexpr int kFrRowsMax = 14;  // S = 2: 1 + 2 + 2 * 5 rows, + the sentinel

template <class T, int kFrBlock>
struct NlFrameLayout {
  size_t off_rec, off_runs, off_stage, bytes;
  __host__ __device__ NlFrameLayout(int n, int cmax, int lane_slots) {
    size_t o = (sizeof(int32_t) * (size_t)(cmax + 2) + 15) & ~size_t(15);
    off_rec = o;
    o += sizeof(NlRec<T>) * (size_t)n;
    off_runs = o;
    const size_t runs = sizeof(uint32_t) * kFrRowsMax * kFrBlock, ids = sizeof(uint16_t) * 2 * (size_t)n;
    o += ((runs > ids ? runs : ids) + 15) & ~size_t(15);  // cell / order ids are dead once the records are written
    off_stage = o;
    o += (sizeof(uint16_t) * (size_t)kFrBlock * lane_slots + 15) & ~size_t(15);
    bytes = o;
  }
};

template <class T, int S, int kFrBlock>
__device__ __forceinline__ int frame_runs(const NlGrid<T>& g, const int32_t* arr, const int cc[3], int p, bool live, uint32_t* runs) {
  // the lane's candidate runs [rs, re) in cell order, row by row of the forward half shell; non-empty ones are kept
  const int n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const int xlo = cc[0] - S < 0 ? 0 : cc[0] - S, xhi1 = (cc[0] + S >= n0 ? n0 - 1 : cc[0] + S) + 1;
  const int base0 = n0 * (cc[1] + n1 * cc[2]), zstride = n0 * n1;
  int nruns = 0;
#pragma unroll
  for (int row = 0; row < 1 + S + S * (2 * S + 1); ++row) {
    const int dz = row <= S ? 0 : 1 + (row - S - 1) / (2 * S + 1);
    const int dy = row <= S ? row : (row - S - 1) % (2 * S + 1) - S;
    const int by = cc[1] + dy, bz = cc[2] + dz;
    if (live && by >= 0 && by < n1 && bz < n2) {
      const int rb = base0 + dy * n0 + dz * zstride;
      int rs = arr[rb + (row == 0 ? cc[0] : xlo)];
      const int re = arr[rb + xhi1];
      if (row == 0) rs = p + 1 > rs ? p + 1 : rs;  // own cell: ids ascend, partners come after
      if (rs < re) {
        runs[nruns * kFrBlock] = uint32_t(rs) | (uint32_t(re) << 16);
        ++nruns;
      }
    }
  }
  runs[nruns * kFrBlock] = 0u;  // sentinel: an empty run ends the lane's walk
  return nruns;
}

template <class T, int kFrBlock>
__global__ void __launch_bounds__(kFrBlock, 2) k_nl_frame(NlDev<T> a) {
  extern __shared__ __align__(16) unsigned char fsm[];
  __shared__ double s_lo[kFrBlock / 32][3], s_hi[kFrBlock / 32][3];
  __shared__ int32_t s_warp[kFrBlock / 32];
  __shared__ int32_t s_stat[4];  // pairs written, longest lane row, largest warp total, overflow bits
  __shared__ int32_t s_tot[2][kFrBlock / 32];  // packed slots: the warps' totals of a round (double buffered)
  const NlFrameLayout<T, kFrBlock> lay(a.n, a.cmax, a.lane_slots);
  int32_t* arr = reinterpret_cast<int32_t*>(fsm);  // arr[c] = start of cell c in cell order, arr[ncell] = n
  int32_t* hist = arr + 1;
  NlRec<T>* rec = reinterpret_cast<NlRec<T>*>(fsm + lay.off_rec);
  uint32_t* runs = reinterpret_cast<uint32_t*>(fsm + lay.off_runs);
  uint16_t* cell16 = reinterpret_cast<uint16_t*>(fsm + lay.off_runs);
  uint16_t* order16 = cell16 + a.n;
  uint16_t* stage_all = reinterpret_cast<uint16_t*>(fsm + lay.off_stage);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = a.n;
  const int lane_slots = a.lane_slots, wpf = (n + 31) / 32;
  const int tag_bits = int(a.tag_bits);

  for (int f = blockIdx.x; f < a.n_frames; f += gridDim.x) {
    const T* ctr = a.center + 3ll * f * n;
    if (a.reference) {
      // conditional rebuild: has any nucleotide moved farther than the threshold since the frame's last build?
      T* ref = a.reference + 3ll * f * n;
      T m = T(0);
      for (int i = tid; i < n; i += kFrBlock) {
        const T dx = ctr[3 * i] - ref[3 * i], dy = ctr[3 * i + 1] - ref[3 * i + 1], dz = ctr[3 * i + 2] - ref[3 * i + 2];
        const T d2 = dx * dx + dy * dy + dz * dz;
        m = d2 > m ? d2 : m;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const T y = __shfl_xor_sync(0xffffffffu, m, o);
        m = y > m ? y : m;
      }
      if (lane == 0) s_lo[warp][0] = double(m);
      __syncthreads();
      double mm = s_lo[0][0];
#pragma unroll
      for (int w = 1; w < kFrBlock / 32; ++w) mm = s_lo[w][0] > mm ? s_lo[w][0] : mm;
      __syncthreads();  // s_lo is reused by the bounds below
      if (!(mm > double(a.move2))) continue;  // (the same decision in every thread of the CTA)
      for (int i = tid; i < 3 * n; i += kFrBlock) ref[i] = ctr[i];
      if (tid == 0 && a.rebuilds) a.rebuilds[f] += 1;
    }
    // ---- bounds
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = tid; i < n; i += kFrBlock) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double x = double(ctr[3 * i + d]);
        lo[d] = x < lo[d] ? x : lo[d];
        hi[d] = x > hi[d] ? x : hi[d];
      }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double l = __shfl_xor_sync(0xffffffffu, lo[d], o), h = __shfl_xor_sync(0xffffffffu, hi[d], o);
        lo[d] = l < lo[d] ? l : lo[d];
        hi[d] = h > hi[d] ? h : hi[d];
      }
      if (lane == 0) {
        s_lo[warp][d] = lo[d];
        s_hi[warp][d] = hi[d];
      }
    }
    if (tid < 4) s_stat[tid] = 0;
    __syncthreads();
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lo[d] = s_lo[0][d];
      hi[d] = s_hi[0][d];
#pragma unroll
      for (int w = 1; w < kFrBlock / 32; ++w) {
        lo[d] = s_lo[w][d] < lo[d] ? s_lo[w][d] : lo[d];
        hi[d] = s_hi[w][d] > hi[d] ? s_hi[w][d] : hi[d];
      }
    }
    const NlGrid<T> g = make_grid(a, lo, hi);
    const int ncell = g.ncell;
    // ---- histogram
    for (int c = tid; c <= ncell; c += kFrBlock) arr[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kFrBlock) {
      int cc[3];
      const int cid = cell_id(a, g, ctr + 3 * i, cc);
      cell16[i] = uint16_t(cid);
      atomicAdd(&hist[cid], 1);
    }
    __syncthreads();
    // ---- exclusive scan of hist[0 .. ncell) in place (a thread owns a contiguous piece)
    {
      const int per = (ncell + kFrBlock - 1) / kFrBlock, c0 = tid * per, c1 = c0 + per < ncell ? c0 + per : ncell;
      int t = 0;
      for (int c = c0; c < c1; ++c) t += hist[c];
      int x = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (lane == 31) s_warp[warp] = x;
      __syncthreads();
      int run = x - t;
      for (int w = 0; w < warp; ++w) run += s_warp[w];
      for (int c = c0; c < c1; ++c) {
        const int v = hist[c];
        hist[c] = run;
        run += v;
      }
    }
    __syncthreads();
    // ---- scatter: hist[c] walks from the start of cell c to its end, i.e. arr[c+1] becomes the start of cell c+1
    for (int i = tid; i < n; i += kFrBlock) order16[atomicAdd(&hist[cell16[i]], 1)] = uint16_t(i);
    __syncthreads();
    // ---- rank inside the cell by id, cell-ordered records
    for (int pos = tid; pos < n; pos += kFrBlock) {
      const int id = order16[pos], c = cell16[id];
      const int b = arr[c], e = arr[c + 1];
      int rank = 0;
      for (int q = b; q < e; ++q) rank += order16[q] < id;
      NlRec<T> r{};
      r.x = ctr[3 * id];
      r.y = ctr[3 * id + 1];
      r.z = ctr[3 * id + 2];
      r.id = id;
      rec[b + rank] = r;
    }
    __syncthreads();
    // ---- walk: warp gw covers cell-order positions [32 gw, 32 gw + 32)
    int w_pairs = 0, w_lane = 0, w_total = 0, w_flags = 0;
    uint16_t* stage = stage_all + (size_t)tid * lane_slots;
    const uint16_t* wstage = stage_all + (size_t)(warp * 32) * lane_slots;
    // rounds of one task per warp (task gw = 32 consecutive cell-order positions).  Packed slots need the totals of all
    // earlier tasks: one barrier per round, every warp takes part in every round (a warp without a task has no live lane)
    const bool packed = a.packed != 0;
    int run_base = packed ? (a.slot_base == 0 ? 0 : a.count[f]) : 0;  // packed: where this build's first entry goes
    for (int r0 = 0, rr = 0; r0 < wpf; r0 += kFrBlock / 32, ++rr) {
      const int gw = r0 + warp;
      if (!packed && gw >= wpf) break;
      const int p = gw * 32 + lane;
      const bool live = p < n;
      const NlRec<T> me = rec[live ? p : 0];
      const int i = me.id;
      const T xi = me.x, yi = me.y, zi = me.z;
      int cc[3];
      {
        const T c[3] = {xi, yi, zi};
        cell_id(a, g, c, cc);
      }
      const int4 ex = *reinterpret_cast<const int4*>(a.excl + i * kMaxExcl);
      const int nruns = g.S == 2 ? frame_runs<T, 2, kFrBlock>(g, arr, cc, p, live, runs + tid)
                                 : frame_runs<T, 1, kFrBlock>(g, arr, cc, p, live, runs + tid);
      int found = 0;
      {
        const T cut2 = a.cut2;
        const uint32_t* rp = runs + tid;
        uint32_t v = *rp;
        int q = int(v & 0xffffu), qe = int(v >> 16);
        while (q < qe) {  // two candidates of the current run per turn (independent loads and distance chains)
          const bool two = q + 1 < qe;
          const NlRec<T> c0 = rec[q], c1 = rec[two ? q + 1 : q];
          q += two ? 2 : 1;
          if (q == qe) {  // next run (the list ends with an empty one)
            rp += kFrBlock;
            v = *rp;
            q = int(v & 0xffffu);
            qe = int(v >> 16);
          }
          const T ax = xi - c0.x, ay = yi - c0.y, az = zi - c0.z;
          const T bx = xi - c1.x, by = yi - c1.y, bz = zi - c1.z;
          const T d2a = add_rn(add_rn(mul_rn(ax, ax), mul_rn(ay, ay)), mul_rn(az, az));
          const T d2b = add_rn(add_rn(mul_rn(bx, bx), mul_rn(by, by)), mul_rn(bz, bz));
          const int ja = c0.id, jb = c1.id;
          if (d2a < cut2 && ja != ex.x && ja != ex.y && ja != ex.z && ja != ex.w) {
            if (found < lane_slots) stage[found] = uint16_t(ja);
            ++found;
          }
          if (two && d2b < cut2 && jb != ex.x && jb != ex.y && jb != ex.z && jb != ex.w) {
            if (found < lane_slots) stage[found] = uint16_t(jb);
            ++found;
          }
        }
      }
      // the warp's pairs into its slot in lane-major order (owner by owner), the rest of the slot padded with n
      const int mine = found < lane_slots ? found : lane_slots;
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      int longest = found;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const int y = __shfl_xor_sync(0xffffffffu, longest, o);
        longest = y > longest ? y : longest;
      }
      int32_t* o0 = a.pairs + (long long)f * 2 * a.capacity + (packed ? 0 : a.slot_base + (long long)gw * a.slot_width);
      int32_t* o1 = o0 + a.capacity;
      int sw = int(a.slot_width);
      const int excl_before = incl - mine;
      if (packed) {
        if (lane == 0) s_tot[rr & 1][warp] = total;
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < kFrBlock / 32; ++w) {
          const int t = s_tot[rr & 1][w];
          before += w < warp ? t : 0;
          all += t;
        }
        const long long room = a.capacity - (long long)(run_base + before);
        o0 += run_base + before;
        o1 += run_base + before;
        sw = room < 0 ? 0 : (room < total ? int(room) : total);  // entries of this task that fit; no padding behind them
        run_base += all;
      }
      __syncwarp();
      // four owners at a time, eight lanes each: an owner's partners are one contiguous run of the slot
      {
        const int grp = lane >> 3, gl = lane & 7;
#pragma unroll 2
        for (int m = 0; m < 8; ++m) {
          const int L = 4 * m + grp;
          const int cnt = __shfl_sync(0xffffffffu, mine, L), off = __shfl_sync(0xffffffffu, excl_before, L);
          const int iL = __shfl_sync(0xffffffffu, i, L);
          const uint16_t* src = wstage + L * lane_slots;
          for (int k = gl; k < cnt; k += 8) {
            if (off + k < sw) {
              const int j = src[k];
              o0[off + k] = iL < j ? iL : j;
              o1[off + k] = (iL < j ? j : iL) | tag_bits;
            }
          }
        }
        for (int k = total + lane; k < sw; k += 32) {
          o0[k] = n;
          o1[k] = n;
        }
      }
      __syncwarp();
      w_pairs += total < sw ? total : sw;
      w_lane = longest > w_lane ? longest : w_lane;
      w_total = total > w_total ? total : w_total;
      w_flags |= (total > sw ? 1 : 0) | (longest > lane_slots ? 4 : 0);
    }
    if (lane == 0) {
      atomicAdd(&s_stat[0], w_pairs);
      atomicMax(&s_stat[1], w_lane);
      atomicMax(&s_stat[2], w_total);
      if (w_flags) atomicOr(&s_stat[3], w_flags);
    }
    __syncthreads();
    if (tid == 0) {
      a.count[f] = (a.slot_base == 0 ? 0 : a.count[f]) + s_stat[0];
      if (a.max_row) {
        a.max_row[2 * f] = s_stat[1];
        a.max_row[2 * f + 1] = s_stat[2];
      }
      if (s_stat[3]) atomicOr(a.overflow, s_stat[3]);
    }
    __syncthreads();  // s_stat, the cell table and the records are reused by the next frame
  }
}

// rows mode: entries beyond N * row_width (capacity not a multiple of N) are padding
template <class T>
__global__ void k_nl_rows_tail(NlDev<T> a) {
  const int f = blockIdx.y;
  const long long first = (long long)(a.capacity / a.n) * a.n;
  int32_t* out0 = a.pairs + (long long)f * 2 * a.capacity;
  for (long long k = first + blockIdx.x * blockDim.x + threadIdx.x; k < a.capacity; k += (long long)gridDim.x * blockDim.x) {
    out0[k] = a.n;
    out0[a.capacity + k] = a.n;
  }
}
__global__ void k_nl_zero_counts(int32_t* count, int32_t* max_row, int n_frames, int per_frame) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n_frames) {
    if (count) count[k] = 0;
    if (max_row)
      for (int m = 0; m < per_frame; ++m) max_row[per_frame * k + m] = 0;
  }
}

template <class T>
__global__ void k_nl_finish(NlDev<T> a) {
  // pad the tail with N and publish count / overflow
  const int f = blockIdx.y;
  const int total = a.nbcount[(long long)(f + 1) * a.n] - a.nbcount[(long long)f * a.n] + (a.append_count ? a.append_count[f] : 0);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.count[f] = total;
    if (total > a.capacity) atomicOr(a.overflow, 1);
  }
  int32_t* out0 = a.pairs + (long long)f * 2 * a.capacity;
  for (long long k = (long long)(total < 0 ? 0 : total) + blockIdx.x * blockDim.x + threadIdx.x; k < a.capacity;
       k += (long long)gridDim.x * blockDim.x) {
    out0[k] = a.n;
    out0[a.capacity + k] = a.n;
  }
}

static int cells_per_frame(int n) {
  long long c = 2ll * n;
  if (c < 4096) c = 4096;
  if (c > (1 << 22)) c = 1 << 22;
  return int(c);
}
static size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

template <class T>
static size_t carve(NlDev<T>* a, void* ws, int n, int F) {
  const size_t C = (size_t)cells_per_frame(n), R = (size_t)F * n;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = ws ? static_cast<char*>(ws) + off : nullptr;
    off += align_up(bytes);
    return p;
  };
  void* p;
  p = take(sizeof(int32_t) * n * kMaxExcl);
  if (a) a->excl = static_cast<int32_t*>(p);
  p = take(sizeof(unsigned long long) * 6 * F);
  if (a) a->bounds = static_cast<unsigned long long*>(p);
  p = take(sizeof(NlGrid<double>) * F);
  if (a) a->grid = static_cast<NlGrid<T>*>(p);
  p = take(sizeof(int32_t) * R);
  if (a) a->cell = static_cast<int32_t*>(p);
  p = take(sizeof(int32_t) * (F * C + 1));
  if (a) a->cstart = static_cast<int32_t*>(p);
  p = take(sizeof(int32_t) * F * C);
  if (a) a->cursor = static_cast<int32_t*>(p);
  p = take(sizeof(int32_t) * R);
  if (a) a->tmp_order = static_cast<int32_t*>(p);
  p = take(sizeof(NlRec<double>) * R);
  if (a) a->srec = static_cast<NlRec<T>*>(p);
  p = take(sizeof(int32_t) * (R + 1));
  if (a) a->nbcount = static_cast<int32_t*>(p);
  p = take(sizeof(uint32_t) * kNlBitWords * R);
  if (a) a->bitbuf = static_cast<uint32_t*>(p);
  const size_t longest = (F * C + 1 > R + 1) ? F * C + 1 : R + 1;
  p = take(sizeof(int32_t) * (longest / kScanChunk + 2));
  if (a) a->scan_tmp = static_cast<int32_t*>(p);
  return off;
}

// MYTHOS_B200_NL_FRAME=0 keeps the multi-launch route for warp-slot builds (A/B tests compare the two routes' lists)
static bool frame_build_enabled() {
  const char* e = getenv("MYTHOS_B200_NL_FRAME");
  return !(e && e[0] == '0');
}

static int frame_block_override() {  // MYTHOS_B200_NL_FRAME_BLOCK=256|384|512: development switch
  const char* e = getenv("MYTHOS_B200_NL_FRAME_BLOCK");
  const int v = e ? atoi(e) : 0;
  return (v == 256 || v == 384 || v == 512) ? v : 0;
}

template <class T>
static int nl_impl(cudaStream_t s, const mb_nl_args* x) {
  MB_REQUIRE(x, MB_EINVAL_SHAPE, "nl_build: null args");
  MB_REQUIRE(x->n > 0 && x->n_frames > 0 && x->n_frames <= 65535, MB_EINVAL_SHAPE, "nl_build: bad n / n_frames");
  MB_REQUIRE((long long)x->n * x->n_frames < (1ll << 31) - 4096, MB_EINVAL_SHAPE, "nl_build: n * n_frames must fit in 31 bits");
  MB_REQUIRE(x->center && x->pairs && x->count && x->overflow, MB_EINVAL_SHAPE, "nl_build: missing buffers");
  MB_REQUIRE(x->n_bonded == 0 || x->bonded, MB_EINVAL_SHAPE, "nl_build: bonded list missing");
  MB_REQUIRE(x->capacity > 0, MB_EINVAL_SHAPE, "nl_build: capacity must be positive");
  MB_REQUIRE(x->r_cutoff + x->dr_threshold > 0, MB_EINVAL_SHAPE, "nl_build: cutoff must be positive");
  const size_t need = carve<T>(nullptr, nullptr, x->n, x->n_frames);
  MB_REQUIRE(x->workspace && x->workspace_bytes >= need, MB_ECAPACITY, "nl_build: workspace too small");
  const bool periodic = x->box[0] > 0 || x->box[1] > 0 || x->box[2] > 0;
  MB_REQUIRE(!periodic || (x->box[0] > 0 && x->box[1] > 0 && x->box[2] > 0), MB_EINVAL_SHAPE,
             "nl_build: box must be all zero (free) or all positive");
  NlDev<T> a;
  a.n = x->n;
  a.n_frames = x->n_frames;
  a.n_bonded = x->n_bonded;
  a.cmax = cells_per_frame(x->n);
  a.center = static_cast<const T*>(x->center);
  a.bonded = x->bonded;
  for (int d = 0; d < 3; ++d) a.box[d] = T(x->box[d]);
  a.periodic = periodic;
  a.cut = T(x->r_cutoff) + T(x->dr_threshold);
  a.cut2 = a.cut * a.cut;
  a.pairs = x->pairs;
  a.capacity = x->capacity;
  a.count = x->count;
  a.overflow = x->overflow;
  a.max_row = (x->flags & MB_NL_ROWS) ? x->max_row : nullptr;
  const bool tag = (x->flags & MB_NL_TAG_SUPPORTS) != 0;
  MB_REQUIRE(!tag || !(x->flags & MB_NL_ROWS), MB_EINVAL_SHAPE, "nl_build: support tags need the compact or the warp-slot layout");
  a.lane_slots = 0;
  a.slot_base = a.slot_width = 0;
  MB_REQUIRE(!tag || (x->n < (1 << 29) && !(x->tag_bits & 0x1fffffffu)), MB_EINVAL_SHAPE, "nl_build: tag bits must be in the top 3 bits, n < 2^29");
  a.tag_bits = tag ? x->tag_bits : 0u;
  a.append_count = tag ? x->append_count : nullptr;
  a.reference = static_cast<T*>(x->reference);
  a.move2 = T(x->move_threshold) * T(x->move_threshold);
  a.rebuilds = x->rebuilds;
  carve<T>(&a, x->workspace, x->n, x->n_frames);

  const int F = x->n_frames, n = x->n;
  const long long R = (long long)F * n, FC = (long long)F * a.cmax;
  if ((x->flags & MB_NL_WARP_SLOTS) && !periodic && n < 32768 && frame_build_enabled()) {
    // frame-resident route: cell table + records of a frame in shared memory, one launch (see k_nl_frame)
    const long long wpf = (n + 31) / 32;
    const bool packed = (x->flags & MB_NL_PACKED_SLOTS) != 0;
    MB_REQUIRE(x->lane_slots > 0 && x->lane_slots <= 256 && x->slot_base >= 0 &&
                   (packed || (x->slot_width > 0 && x->slot_base + wpf * x->slot_width <= x->capacity)),
               MB_EINVAL_SHAPE, "nl_build: warp-slot mode needs 0 < lane_slots <= 256 and slot_base + ceil(n/32) * slot_width <= capacity");
    MB_REQUIRE(!packed || !x->reference, MB_EINVAL_SHAPE, "nl_build: packed slots cannot be rebuilt conditionally (a second build appends behind the first)");
    a.packed = packed ? 1 : 0;
    int dev = 0, smem_max = 0, smem_sm = 0, sms = 0;
    MB_CUDA_CHECK(cudaGetDevice(&dev));
    MB_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    MB_CUDA_CHECK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    MB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // widest CTA of which two fit on an SM (the walk is latency bound: warps in flight are what counts)
    const size_t b512 = NlFrameLayout<T, 512>(n, a.cmax, x->lane_slots).bytes, b384 = NlFrameLayout<T, 384>(n, a.cmax, x->lane_slots).bytes,
                 b256 = NlFrameLayout<T, 256>(n, a.cmax, x->lane_slots).bytes;
    const size_t two = ((size_t)smem_sm - 2 * 1024) / 2;
    const int blk = frame_block_override() ? frame_block_override() : (b512 <= two ? 512 : (b384 <= two ? 384 : 256));
    const size_t bytes = blk == 512 ? b512 : (blk == 384 ? b384 : b256);
    if (bytes <= (size_t)smem_max) {
      a.lane_slots = x->lane_slots;
      a.slot_base = x->slot_base;
      a.slot_width = x->slot_width;
      a.max_row = x->max_row;
      a.tag_bits = x->tag_bits & 0xe0000000u;
      if (!(x->flags & MB_NL_REUSE_EXCLUSIONS)) {
        k_nl_excl_init<<<ceil_div((long long)n * kMaxExcl, 256), 256, 0, s>>>(a.excl, n);
        if (x->n_bonded > 0)
          k_nl_excl_fill<<<ceil_div(x->n_bonded, 256), 256, 0, s>>>(a.excl, a.bonded, x->n_bonded, n, a.overflow);
      }
      auto launch = [&](auto kern, int threads) -> int {
        // (the attribute only ever grows, the occupancy is remembered per size: nothing but the launch while a graph is captured)
        static std::mutex mu;
        // (the three block widths share one function-pointer type, hence one set of statics: the kernel is part of the key)
        static std::map<std::pair<const void*, int>, size_t> granted;
        static std::map<std::pair<std::pair<const void*, int>, size_t>, int> known;
        const std::pair<const void*, int> kd{reinterpret_cast<const void*>(kern), dev};
        int per_sm = 0;
        bool raise = false;
        {
          std::lock_guard<std::mutex> lock(mu);
          auto it = known.find({kd, bytes});
          if (it != known.end()) per_sm = it->second;
          raise = granted[kd] < bytes;
        }
        if (raise) {
          MB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
          std::lock_guard<std::mutex> lock(mu);
          if (granted[kd] < bytes) granted[kd] = bytes;
        }
        if (!per_sm) {
          MB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, bytes));
          if (per_sm < 1) per_sm = 1;
          std::lock_guard<std::mutex> lock(mu);
          known[{kd, bytes}] = per_sm;
        }
        const int blocks = F < per_sm * sms ? F : per_sm * sms;
        kern<<<blocks, threads, bytes, s>>>(a);
        return MB_OK;
      };
      const int rc = blk == 512 ? launch(k_nl_frame<T, 512>, 512) : (blk == 384 ? launch(k_nl_frame<T, 384>, 384) : launch(k_nl_frame<T, 256>, 256));
      if (rc != MB_OK) return rc;
      MB_CUDA_CHECK(cudaGetLastError());
      return MB_OK;
    }
  }
  MB_REQUIRE(!x->reference, MB_EINVAL_SHAPE, "nl_build: the conditional rebuild (reference) needs a free-space warp-slot build on the frame-resident route");
  MB_REQUIRE(!(x->flags & MB_NL_PACKED_SLOTS), MB_EINVAL_SHAPE, "nl_build: packed slots need the frame-resident route (mythos_b200_nl_conditional_supported)");
  a.packed = 0;
  MB_CUDA_CHECK(cudaMemsetAsync(a.cstart, 0, sizeof(int32_t) * (size_t)(FC + 1), s));
  MB_CUDA_CHECK(cudaMemsetAsync(a.cursor, 0, sizeof(int32_t) * (size_t)FC, s));
  k_nl_excl_init<<<ceil_div((long long)n * kMaxExcl, 256), 256, 0, s>>>(a.excl, n);
  if (x->n_bonded > 0)
    k_nl_excl_fill<<<ceil_div(x->n_bonded, 256), 256, 0, s>>>(a.excl, a.bonded, x->n_bonded, n, a.overflow);
  if (!periodic) {
    k_nl_bounds_init<<<ceil_div(6ll * F, 256), 256, 0, s>>>(a.bounds, F);
    int bx = ceil_div(n, 256 * 4);
    if (bx > 64) bx = 64;
    k_nl_bounds<T><<<dim3(bx, F), 256, 0, s>>>(a);
  }
  k_nl_grid<T><<<ceil_div(F, 128), 128, 0, s>>>(a);
  const int gr = ceil_div(R, kNlBlock);
  k_nl_bin<T><<<gr, kNlBlock, 0, s>>>(a);
  scan_exclusive(s, a.cstart, FC, a.scan_tmp);
  k_nl_scatter<T><<<gr, kNlBlock, 0, s>>>(a);
  k_nl_rank<T><<<gr, kNlBlock, 0, s>>>(a);
  if (x->flags & MB_NL_WARP_SLOTS) {
    const long long wpf = (n + 31) / 32;
    MB_REQUIRE(x->lane_slots > 0 && x->lane_slots <= 256 && x->slot_width > 0 && x->slot_base >= 0 &&
                   x->slot_base + wpf * x->slot_width <= x->capacity,
               MB_EINVAL_SHAPE, "nl_build: warp-slot mode needs 0 < lane_slots <= 256 and slot_base + ceil(n/32) * slot_width <= capacity");
    a.lane_slots = x->lane_slots;
    a.slot_base = x->slot_base;
    a.slot_width = x->slot_width;
    a.max_row = x->max_row;
    a.tag_bits = x->tag_bits & 0xe0000000u;
    // counts accumulate over the builds that share a list: only the first one (slot_base == 0) zeroes them
    k_nl_zero_counts<<<ceil_div(F, 256), 256, 0, s>>>(x->slot_base == 0 ? a.count : nullptr, a.max_row, F, 2);
    const long long warps = wpf * F;
    const size_t smem = sizeof(int32_t) * (kNlBlock / 32) * 32 * (size_t)x->lane_slots;
    const int blocks = ceil_div(warps * 32, kNlBlock);
    if (periodic) {
      MB_CUDA_CHECK(cudaFuncSetAttribute(k_nl_walk<T, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_nl_walk<T, 3, true><<<blocks, kNlBlock, smem, s>>>(a);
    } else {
      MB_CUDA_CHECK(cudaFuncSetAttribute(k_nl_walk<T, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_nl_walk<T, 3, false><<<blocks, kNlBlock, smem, s>>>(a);
    }
  } else if (x->flags & MB_NL_ROWS) {
    k_nl_zero_counts<<<ceil_div(F, 256), 256, 0, s>>>(a.count, a.max_row, F, 1);
    if (periodic) k_nl_walk<T, 2, true><<<gr, kNlBlock, 0, s>>>(a);
    else k_nl_walk<T, 2, false><<<gr, kNlBlock, 0, s>>>(a);
    if (x->capacity % n) k_nl_rows_tail<T><<<dim3(1, F), 256, 0, s>>>(a);
  } else {
    if (periodic) k_nl_walk<T, 0, true><<<gr, kNlBlock, 0, s>>>(a);
    else k_nl_walk<T, 0, false><<<gr, kNlBlock, 0, s>>>(a);
    scan_exclusive(s, a.nbcount, R, a.scan_tmp);
    if (periodic) k_nl_walk<T, 1, true><<<gr, kNlBlock, 0, s>>>(a);
    else k_nl_walk<T, 1, false><<<gr, kNlBlock, 0, s>>>(a);
    dim3 gf((F > 64 || a.append_count) ? 1 : 32, F);  // only the tail [count, capacity) is touched; one block when count aliases append_count
    k_nl_finish<T><<<gf, 256, 0, s>>>(a);
  }
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}

}  // namespace mb

extern "C" size_t mythos_b200_nl_workspace_bytes(int32_t n, int32_t n_frames) {
  if (n <= 0 || n_frames <= 0) return 0;
  return mb::carve<double>(nullptr, nullptr, n, n_frames);
}
extern "C" int mythos_b200_nl_conditional_supported(int32_t n, int32_t lane_slots, int32_t real_bytes) {
  if (n <= 0 || n >= 32768 || lane_slots <= 0 || lane_slots > 256 || !mb::frame_build_enabled()) return 0;
  int dev = 0, smem_max = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
  const int cmax = mb::cells_per_frame(n);
  const size_t b = real_bytes == 4 ? mb::NlFrameLayout<float, 256>(n, cmax, lane_slots).bytes : mb::NlFrameLayout<double, 256>(n, cmax, lane_slots).bytes;
  return b <= (size_t)smem_max ? 1 : 0;
}
extern "C" int mythos_b200_nl_build_f64(void* stream, const mb_nl_args* a) {
  return mb::nl_impl<double>(static_cast<cudaStream_t>(stream), a);
}
extern "C" int mythos_b200_nl_build_f32(void* stream, const mb_nl_args* a) {
  return mb::nl_impl<float>(static_cast<cudaStream_t>(stream), a);
}
