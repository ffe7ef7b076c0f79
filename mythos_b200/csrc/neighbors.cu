// neighbors.cu -- batched cell-list neighbour build: integer binning, counting sort, count / scan / fill.
//
// Replaces mythos/utils/neighbors.py:12-59, i.e. jax_md.partition.neighbor_list(format=OrderedSparse,
// disable_cell_list=True, custom_mask_function=bonded mask), whose per-rebuild cost is an O(N^2) distance
// matrix.  Cells are only a superset generator: the accept test is the reference's arithmetic
//   dR = Ri - Rj (periodic: floor-mod wrap to [-L/2, L/2));  d2 = dx*dx + dy*dy + dz*dz;
//   keep iff d2 < (r_cutoff + dr_threshold)^2 (strict), i < j, (i,j) not bonded
// evaluated in the positions' dtype with explicit round-to-nearest mul/add (no FMA contraction), so the pair
// SET is bit-exact against an all-pairs evaluation of the same expression.
//
// Per frame (frames are independent; grid.y = frame):
//   1. bounds    min corner of the frame (free space) -> cell origin
//   2. bin       integer cell coordinates (10 bits per axis) -> key; hashed bucket; bucket histogram
//   3. scan      exclusive scan of the histogram
//   4. scatter   particle ids into bucket order, then each bucket segment sorted by id (deterministic)
//   5. count     per particle: half-shell walk (own cell with j>i + 13 forward cells), accept test
//   6. scan      exclusive scan of the per-particle counts -> offsets, total -> count[frame]
//   7. fill      same walk, writing (min(i,j), max(i,j)) at the particle's offset; tail padded with N
#include "common.cuh"

namespace mb {

constexpr int kNlBlock = 128;
constexpr int kMaxExcl = 4;

template <class T>
struct NlDev {
  int n, n_frames, n_bonded, hbits;  // buckets per frame = 1 << hbits
  const T* center;
  const int32_t* bonded;
  T box[3];
  int periodic;
  T cell;   // cell edge >= cutoff
  T cut2;   // (r_cutoff + dr_threshold)^2 in T
  int32_t* pairs;
  long long capacity;
  int32_t* count;
  int32_t* overflow;
  // workspace
  int32_t* excl;     // (N, kMaxExcl)
  T* origin;         // (F, 3)
  int32_t* dims;     // (F, 3)
  uint32_t* key;     // (F, N)
  int32_t* bstart;   // (F, H + 1)  histogram, then exclusive scan
  int32_t* cursor;   // (F, H)
  int32_t* order;    // (F, N) particle ids in bucket order
  int32_t* nbcount;  // (F, N + 1) per-particle pair counts, then exclusive scan
};

__device__ __forceinline__ uint32_t bucket_of(uint32_t key, int hbits) { return (key * 2654435761u) >> (32 - hbits); }

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

template <class T>
__global__ void k_nl_bounds(NlDev<T> a) {
  const int f = blockIdx.x;
  __shared__ T smin[3][kNlBlock];
  T m[3] = {T(1e30), T(1e30), T(1e30)};
  if (!a.periodic) {
    for (int i = threadIdx.x; i < a.n; i += kNlBlock) {
      const T* c = a.center + 3ll * ((long long)f * a.n + i);
      for (int d = 0; d < 3; ++d) m[d] = c[d] < m[d] ? c[d] : m[d];
    }
  }
  for (int d = 0; d < 3; ++d) smin[d][threadIdx.x] = m[d];
  __syncthreads();
  for (int s = kNlBlock / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int d = 0; d < 3; ++d)
        smin[d][threadIdx.x] = smin[d][threadIdx.x + s] < smin[d][threadIdx.x] ? smin[d][threadIdx.x + s] : smin[d][threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x < 3) {
    const int d = threadIdx.x;
    if (a.periodic) {
      a.origin[3 * f + d] = T(0);
      int nd = int(a.box[d] / a.cell);
      if (nd < 3) nd = 1;  // fewer than 3 cells along a periodic axis: a single cell spanning it
      if (nd > 1024) nd = 1024;
      a.dims[3 * f + d] = nd;
    } else {
      a.origin[3 * f + d] = smin[d][0];
      a.dims[3 * f + d] = 1024;  // open grid; coordinates beyond are clipped into the last cell
    }
  }
}

template <class T>
__device__ __forceinline__ void cell_coords(const NlDev<T>& a, int f, const T* c, int cc[3]) {
  for (int d = 0; d < 3; ++d) {
    const int nd = a.dims[3 * f + d];
    T x = c[d] - a.origin[3 * f + d];
    int ci;
    if (a.periodic) {
      // positions may lie outside the primary image; cell width along this axis is box/nd >= cell
      T s = fmod(x, a.box[d]);
      if (s < T(0)) s += a.box[d];
      ci = int(s / (a.box[d] / T(nd)));
      if (ci >= nd) ci = nd - 1;
      if (ci < 0) ci = 0;
    } else {
      ci = int(x / a.cell);
      if (ci < 0) ci = 0;
      if (ci >= nd) ci = nd - 1;
    }
    cc[d] = ci;
  }
}

template <class T>
__global__ void k_nl_bin(NlDev<T> a) {
  const int f = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  int cc[3];
  cell_coords(a, f, a.center + 3ll * ((long long)f * a.n + i), cc);
  const uint32_t key = uint32_t(cc[0]) | (uint32_t(cc[1]) << 10) | (uint32_t(cc[2]) << 20);
  a.key[(long long)f * a.n + i] = key;
  const long long H = 1ll << a.hbits;
  atomicAdd(&a.bstart[(long long)f * (H + 1) + bucket_of(key, a.hbits)], 1);
}

// exclusive scan of `len` int32 per segment, one block per segment; writes the total at [len]
__global__ void k_seg_scan(int32_t* data, long long seg_stride, int len) {
  __shared__ int32_t swarp[32];
  __shared__ int32_t carry;
  int32_t* seg = data + (long long)blockIdx.x * seg_stride;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int base = 0; base < len; base += blockDim.x) {
    const int k = base + threadIdx.x;
    const int32_t v = (k < len) ? seg[k] : 0;
    int32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) swarp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int32_t w = (lane < nw) ? swarp[lane] : 0;
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      swarp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int32_t before = carry + (warp ? swarp[warp - 1] : 0);
    if (k < len) seg[k] = before + x - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += swarp[nw - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) seg[len] = carry;
}

template <class T>
__global__ void k_nl_scatter(NlDev<T> a) {
  const int f = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const long long H = 1ll << a.hbits;
  const uint32_t b = bucket_of(a.key[(long long)f * a.n + i], a.hbits);
  const int pos = atomicAdd(&a.cursor[(long long)f * H + b], 1);
  a.order[(long long)f * a.n + a.bstart[(long long)f * (H + 1) + b] + pos] = i;
}

template <class T>
__global__ void k_nl_sort_buckets(NlDev<T> a) {
  const int f = blockIdx.y;
  const long long H = 1ll << a.hbits;
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= H) return;
  const int32_t* bs = a.bstart + (long long)f * (H + 1);
  int32_t* o = a.order + (long long)f * a.n;
  const int lo = bs[b], hi = bs[b + 1];
  for (int p = lo + 1; p < hi; ++p) {
    const int v = o[p];
    int q = p - 1;
    while (q >= lo && o[q] > v) {
      o[q + 1] = o[q];
      --q;
    }
    o[q + 1] = v;
  }
}

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

template <class T, bool FILL>
__global__ void k_nl_walk(NlDev<T> a) {
  const int f = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const long long fb = (long long)f * a.n;
  const long long H = 1ll << a.hbits;
  const T* ci = a.center + 3 * (fb + i);
  const T xi = ci[0], yi = ci[1], zi = ci[2];
  const uint32_t mykey = a.key[fb + i];
  const int cx = mykey & 1023, cy = (mykey >> 10) & 1023, cz = (mykey >> 20) & 1023;
  const int nx = a.dims[3 * f], ny = a.dims[3 * f + 1], nz = a.dims[3 * f + 2];
  int ex[kMaxExcl];
  for (int s = 0; s < kMaxExcl; ++s) ex[s] = a.excl[i * kMaxExcl + s];
  const int32_t* bs = a.bstart + f * (H + 1);
  const int32_t* ord = a.order + fb;
  int found = 0;
  long long wpos = 0;
  int32_t* out0 = nullptr;
  int32_t* out1 = nullptr;
  if (FILL) {
    wpos = a.nbcount[(long long)f * (a.n + 1) + i];
    out0 = a.pairs + (long long)f * 2 * a.capacity;
    out1 = out0 + a.capacity;
  }
  // half shell: offsets (dx,dy,dz) that are lexicographically >= 0 in (dz,dy,dx) order
  for (int dz = 0; dz <= 1; ++dz) {
    for (int dy = (dz ? -1 : 0); dy <= 1; ++dy) {
      for (int dx = ((dz || dy) ? -1 : 0); dx <= 1; ++dx) {
        int ox = cx + dx, oy = cy + dy, oz = cz + dz;
        if (a.periodic) {
          if ((nx == 1 && dx) || (ny == 1 && dy) || (nz == 1 && dz)) continue;
          ox = (ox + nx) % nx;
          oy = (oy + ny) % ny;
          oz = (oz + nz) % nz;
        } else if (ox < 0 || oy < 0 || oz < 0 || ox >= nx || oy >= ny || oz >= nz) {
          continue;
        }
        const bool own = !(dx || dy || dz);
        const uint32_t nkey = uint32_t(ox) | (uint32_t(oy) << 10) | (uint32_t(oz) << 20);
        const uint32_t b = bucket_of(nkey, a.hbits);
        const int lo = bs[b], hi = bs[b + 1];
        for (int p = lo; p < hi; ++p) {
          const int j = ord[p];
          if (a.key[fb + j] != nkey) continue;
          if (own && j <= i) continue;
          if (j == ex[0] || j == ex[1] || j == ex[2] || j == ex[3]) continue;
          const int lo_i = i < j ? i : j, hi_j = i < j ? j : i;
          // receiver = lower index, as the OrderedSparse format keeps i < j
          const T* cl = (lo_i == i) ? ci : a.center + 3 * (fb + j);
          const T* ch = (lo_i == i) ? a.center + 3 * (fb + j) : ci;
          T ddx = cl[0] - ch[0], ddy = cl[1] - ch[1], ddz = cl[2] - ch[2];
          if (a.periodic) {
            ddx = wrap_nl(ddx, a.box[0]);
            ddy = wrap_nl(ddy, a.box[1]);
            ddz = wrap_nl(ddz, a.box[2]);
          }
          const T d2 = add_rn(add_rn(mul_rn(ddx, ddx), mul_rn(ddy, ddy)), mul_rn(ddz, ddz));
          if (d2 < a.cut2) {
            if (FILL) {
              if (wpos < a.capacity) {
                out0[wpos] = lo_i;
                out1[wpos] = hi_j;
              }
              ++wpos;
            }
            ++found;
          }
        }
      }
    }
  }
  (void)xi;
  (void)yi;
  (void)zi;
  if (!FILL) a.nbcount[(long long)f * (a.n + 1) + i] = found;
}

template <class T>
__global__ void k_nl_finish(NlDev<T> a) {
  // pad the tail with N and publish count / overflow
  const int f = blockIdx.y;
  const int total = a.nbcount[(long long)f * (a.n + 1) + a.n];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.count[f] = total;
    if (total > a.capacity) atomicOr(a.overflow, 1);
  }
  int32_t* out0 = a.pairs + (long long)f * 2 * a.capacity;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < a.capacity; k += (long long)gridDim.x * blockDim.x) {
    if (k >= total) {
      out0[k] = a.n;
      out0[a.capacity + k] = a.n;
    }
  }
}

static int hash_bits(int n) {
  int b = 5;
  while ((1ll << b) < 2ll * n && b < 24) ++b;
  return b;
}
static size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

template <class T>
static size_t carve(NlDev<T>* a, void* ws, int n, int F) {
  const size_t H = size_t(1) << hash_bits(n);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = ws ? static_cast<char*>(ws) + off : nullptr;
    off += align_up(bytes);
    return p;
  };
  void* p;
  p = take(sizeof(int32_t) * n * kMaxExcl);
  if (a) a->excl = static_cast<int32_t*>(p);
  p = take(sizeof(double) * 3 * F);
  if (a) a->origin = static_cast<T*>(p);
  p = take(sizeof(int32_t) * 3 * F);
  if (a) a->dims = static_cast<int32_t*>(p);
  p = take(sizeof(uint32_t) * (size_t)F * n);
  if (a) a->key = static_cast<uint32_t*>(p);
  p = take(sizeof(int32_t) * (size_t)F * (H + 1));
  if (a) a->bstart = static_cast<int32_t*>(p);
  p = take(sizeof(int32_t) * (size_t)F * H);
  if (a) a->cursor = static_cast<int32_t*>(p);
  p = take(sizeof(int32_t) * (size_t)F * n);
  if (a) a->order = static_cast<int32_t*>(p);
  p = take(sizeof(int32_t) * (size_t)F * (n + 1));
  if (a) a->nbcount = static_cast<int32_t*>(p);
  return off;
}

template <class T>
static int nl_impl(cudaStream_t s, const mb_nl_args* x) {
  MB_REQUIRE(x, MB_EINVAL_SHAPE, "nl_build: null args");
  MB_REQUIRE(x->n > 0 && x->n_frames > 0 && x->n_frames <= 65535, MB_EINVAL_SHAPE, "nl_build: bad n / n_frames");
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
  a.hbits = hash_bits(x->n);
  a.center = static_cast<const T*>(x->center);
  a.bonded = x->bonded;
  for (int d = 0; d < 3; ++d) a.box[d] = T(x->box[d]);
  a.periodic = periodic;
  const T cut = T(x->r_cutoff) + T(x->dr_threshold);
  a.cut2 = cut * cut;
  a.cell = cut * T(1.0001);  // a hair wider than the cutoff so rounding in the binning can never hide a pair
  a.pairs = x->pairs;
  a.capacity = x->capacity;
  a.count = x->count;
  a.overflow = x->overflow;
  carve<T>(&a, x->workspace, x->n, x->n_frames);

  const long long H = 1ll << a.hbits;
  const int F = x->n_frames, n = x->n;
  MB_CUDA_CHECK(cudaMemsetAsync(a.bstart, 0, sizeof(int32_t) * (size_t)F * (H + 1), s));
  MB_CUDA_CHECK(cudaMemsetAsync(a.cursor, 0, sizeof(int32_t) * (size_t)F * H, s));
  k_nl_excl_init<<<ceil_div((long long)n * kMaxExcl, 256), 256, 0, s>>>(a.excl, n);
  if (x->n_bonded > 0)
    k_nl_excl_fill<<<ceil_div(x->n_bonded, 256), 256, 0, s>>>(a.excl, a.bonded, x->n_bonded, n, a.overflow);
  k_nl_bounds<T><<<F, kNlBlock, 0, s>>>(a);
  dim3 gp(ceil_div(n, kNlBlock), F);
  k_nl_bin<T><<<gp, kNlBlock, 0, s>>>(a);
  k_seg_scan<<<F, 1024, 0, s>>>(a.bstart, H + 1, (int)H);
  k_nl_scatter<T><<<gp, kNlBlock, 0, s>>>(a);
  dim3 gb(ceil_div(H, kNlBlock), F);
  k_nl_sort_buckets<T><<<gb, kNlBlock, 0, s>>>(a);
  k_nl_walk<T, false><<<gp, kNlBlock, 0, s>>>(a);
  k_seg_scan<<<F, 1024, 0, s>>>(a.nbcount, n + 1, n);
  k_nl_walk<T, true><<<gp, kNlBlock, 0, s>>>(a);
  dim3 gf(min(ceil_div(x->capacity, 256), 1024), F);
  k_nl_finish<T><<<gf, 256, 0, s>>>(a);
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}

}  // namespace mb

extern "C" size_t mythos_b200_nl_workspace_bytes(int32_t n, int32_t n_frames) {
  if (n <= 0 || n_frames <= 0) return 0;
  return mb::carve<double>(nullptr, nullptr, n, n_frames);
}
extern "C" int mythos_b200_nl_build_f64(void* stream, const mb_nl_args* a) {
  return mb::nl_impl<double>(static_cast<cudaStream_t>(stream), a);
}
extern "C" int mythos_b200_nl_build_f32(void* stream, const mb_nl_args* a) {
  return mb::nl_impl<float>(static_cast<cudaStream_t>(stream), a);
}
