// trajectory.cu -- oxDNA trajectory text -> (F,N,3) centres + (F,N,4) quaternions on the device.
//
// Stands in for mythos/input/trajectory.py:192-320 (`from_file` / `_read_file`: "t = ..", "b = ..", "E = .." and N lines of
// 15 numbers per state, per-strand reversal when the file is 5'->3') and NucleotideState.quaternions
// (trajectory.py:163-175 -> mythos/utils/math.py:9-65: a1, a3 -> Tait-Bryan ZYX angles -> quaternion).  The reference
// parses line by line in Python; here the file's bytes are copied to the device once and
//   k_traj_count   counts the newlines of each 64 KiB chunk                         (HBM: reads the text once)
//   k_traj_scan    exclusive scan of the chunk counts (one block)
//   k_traj_lines   writes the byte offset of every line start, in file order       (reads the text a second time)
//   k_traj_parse   one thread per line: header lines -> times / box / energies; nucleotide lines -> the first nine numbers,
//                  correctly rounded (parse_decimal.cuh), quaternion, scattered to the internal nucleotide order
// Velocities / angular momenta (columns 10-15) are not part of the energy path and are skipped.
#include "common.cuh"
#include "parse_decimal.cuh"

namespace mb {

constexpr int kChunk = 65536, kTrajThreads = 256, kPerThread = kChunk / kTrajThreads;

__global__ void __launch_bounds__(kTrajThreads) k_traj_count(const unsigned char* __restrict__ text, long long n_bytes,
                                                             int* __restrict__ chunk_count) {
  const long long base = (long long)blockIdx.x * kChunk;
  int c = 0;
  // 16-byte loads, consecutive threads on consecutive vectors
  for (int v = threadIdx.x; v < kChunk / 16; v += kTrajThreads) {
    const long long off = base + 16ll * v;
    if (off + 16 <= n_bytes) {
      const uint4 w = *reinterpret_cast<const uint4*>(text + off);
      const unsigned words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned x = words[k] ^ 0x0a0a0a0au;  // zero byte where a newline was
        // exact zero-byte mask (no borrow between bytes): high bit set iff the byte is zero
        const unsigned t = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);
        c += __popc(t);
      }
    } else {
      for (long long b = off; b < n_bytes && b < off + 16; ++b) c += text[b] == '\n';
    }
  }
  __shared__ int red[kTrajThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < kTrajThreads / 32; ++w) s += red[w];
    chunk_count[blockIdx.x] = s;
  }
}

// one block: chunk_count -> exclusive offsets (in place), total -> *n_lines (lines = newlines, + 1 if the file does not end
// with one)
__global__ void __launch_bounds__(1024) k_traj_scan(int* chunk_count, int n_chunks, const unsigned char* text, long long n_bytes,
                                                    long long* n_lines) {
  __shared__ long long carry;
  __shared__ int warp_sum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_chunks; base += 1024) {
    const int k = base + threadIdx.x;
    const int v = k < n_chunks ? chunk_count[k] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_sum[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      warp_sum[threadIdx.x] = w;
    }
    __syncthreads();
    const long long before = carry + (threadIdx.x >= 32 ? warp_sum[(threadIdx.x >> 5) - 1] : 0) + (x - v);
    if (k < n_chunks) chunk_count[k] = int(before);  // (a trajectory has < 2^31 lines)
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_lines = carry + ((n_bytes > 0 && text[n_bytes - 1] != '\n') ? 1 : 0);
}

// line_start[0] = 0; line_start[k] = byte after the k-th newline.  Each thread owns kPerThread consecutive bytes.
__global__ void __launch_bounds__(kTrajThreads) k_traj_lines(const unsigned char* __restrict__ text, long long n_bytes,
                                                             const int* __restrict__ chunk_offset, long long* __restrict__ line_start,
                                                             long long capacity) {
  const long long base = (long long)blockIdx.x * kChunk + (long long)threadIdx.x * kPerThread;
  int c = 0;
  for (int b = 0; b < kPerThread; ++b) c += (base + b < n_bytes) && text[base + b] == '\n';
  __shared__ int warp_sum[kTrajThreads / 32];
  int x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
  __syncthreads();
  int before = x - c;
  for (int w = 0; w < (threadIdx.x >> 5); ++w) before += warp_sum[w];
  long long k = (long long)chunk_offset[blockIdx.x] + before + 1;  // index of the line that starts after my first newline
  if (blockIdx.x == 0 && threadIdx.x == 0) line_start[0] = 0;
  for (int b = 0; b < kPerThread; ++b)
    if (base + b < n_bytes && text[base + b] == '\n') {
      if (k < capacity) line_start[k] = base + b + 1;
      ++k;
    }
}

__device__ __forceinline__ void skip_blank(const unsigned char*& p, const unsigned char* end) {
  while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
}

template <class T>
__global__ void __launch_bounds__(128) k_traj_parse(const unsigned char* __restrict__ text, long long n_bytes,
                                                    const long long* __restrict__ line_start, long long n_lines, int n, int n_frames,
                                                    const int32_t* __restrict__ dest, const uint64_t* __restrict__ pow5,
                                                    T* __restrict__ center, T* __restrict__ quat, double* __restrict__ times,
                                                    double* __restrict__ box, double* __restrict__ energies, int* __restrict__ status) {
  const long long L = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = n + 3;
  if (L >= (long long)n_frames * per) return;
  const int frame = int(L / per), row = int(L % per);
  const unsigned char* p = text + line_start[L];
  const unsigned char* end = text + ((L + 1 < n_lines) ? line_start[L + 1] - 1 : n_bytes);  // (the newline itself excluded)
  if (end > p && end[-1] == '\n') --end;
  bool ok = true;
  if (row < 3) {  // "t = x" | "b = x y z" | "E = x y z"
    skip_blank(p, end);
    const char want = row == 0 ? 't' : (row == 1 ? 'b' : 'E');
    if (!(p < end && *p == want)) {
      atomicAdd(&status[1], 1);  // the file is not N nucleotide lines per state: wrong strand_lengths or a truncated state
      return;
    }
    while (p < end && *p != '=') ++p;
    if (p < end) ++p;
    double v[3] = {0, 0, 0};
    const int cnt = row == 0 ? 1 : 3;
    for (int k = 0; k < cnt; ++k) {
      skip_blank(p, end);
      v[k] = parse_number(p, end, pow5, ok);
    }
    if (row == 0) times[frame] = v[0];
    if (row == 1)
      for (int k = 0; k < 3; ++k) box[3 * frame + k] = v[k];
    if (row == 2)
      for (int k = 0; k < 3; ++k) energies[3 * frame + k] = v[k];
    if (!ok) atomicAdd(&status[0], 1);
    return;
  }
  double v[9];
#pragma unroll 1
  for (int k = 0; k < 9; ++k) {
    skip_blank(p, end);
    if (p >= end) ok = false;
    v[k] = parse_number(p, end, pow5, ok);
  }
  if (!ok) atomicAdd(&status[0], 1);
  // x = a1, z = a3, y = a3 x a1 -> Tait-Bryan ZYX angles (utils/math.py:35-37) -> quaternion (math.py:56-63)
  const double x0 = v[3], x1 = v[4], x2 = v[5], z0 = v[6], z1 = v[7], z2 = v[8];
  const double y2 = z0 * x1 - z1 * x0;
  const double psi = atan2(x1, x0), theta = asin(-fmin(fmax(x2, -1.0), 1.0)), phi = atan2(y2, z2);
  double s_psi, c_psi, s_th, c_th, s_ph, c_ph;
  sincos(0.5 * psi, &s_psi, &c_psi);
  sincos(0.5 * theta, &s_th, &c_th);
  sincos(0.5 * phi, &s_ph, &c_ph);
  const int r = row - 3;
  const long long o = (long long)frame * n + (dest ? dest[r] : r);
  center[3 * o] = T(v[0]);
  center[3 * o + 1] = T(v[1]);
  center[3 * o + 2] = T(v[2]);
  quat[4 * o] = T(s_psi * s_th * s_ph + c_psi * c_th * c_ph);
  quat[4 * o + 1] = T(-s_psi * s_th * c_ph + s_ph * c_psi * c_th);
  quat[4 * o + 2] = T(s_psi * c_th * s_ph + c_psi * s_th * c_ph);
  quat[4 * o + 3] = T(s_psi * c_th * c_ph - c_psi * s_th * s_ph);
}

}  // namespace mb

extern "C" {
size_t mythos_b200_traj_workspace_bytes(int64_t n_bytes) {
  const size_t chunks = size_t((n_bytes + mb::kChunk - 1) / mb::kChunk);
  return (chunks + 1) * sizeof(int) + 16;
}

int mythos_b200_traj_index(void* stream, const void* text, int64_t n_bytes, void* workspace, size_t workspace_bytes,
                           int64_t* line_start, int64_t line_capacity, int64_t* n_lines) {
  MB_REQUIRE(text && workspace && n_lines && n_bytes > 0, MB_EINVAL_SHAPE, "traj_index: text / workspace / n_lines required");
  MB_REQUIRE(workspace_bytes >= mythos_b200_traj_workspace_bytes(n_bytes), MB_ECAPACITY, "traj_index: workspace too small");
  MB_REQUIRE((reinterpret_cast<uintptr_t>(text) & 15) == 0, MB_EINVAL_SHAPE, "traj_index: text must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int chunks = int((n_bytes + mb::kChunk - 1) / mb::kChunk);
  int* counts = static_cast<int*>(workspace);
  const unsigned char* t = static_cast<const unsigned char*>(text);
  if (!line_start) {  // pass 1: count
    mb::k_traj_count<<<chunks, mb::kTrajThreads, 0, s>>>(t, n_bytes, counts);
    MB_CUDA_CHECK(cudaGetLastError());
    mb::k_traj_scan<<<1, 1024, 0, s>>>(counts, chunks, t, n_bytes, reinterpret_cast<long long*>(n_lines));
    MB_CUDA_CHECK(cudaGetLastError());
    return MB_OK;
  }
  // pass 2 (after pass 1 on the same workspace): the line starts
  mb::k_traj_lines<<<chunks, mb::kTrajThreads, 0, s>>>(t, n_bytes, counts, reinterpret_cast<long long*>(line_start), line_capacity);
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}

static int traj_parse_impl(void* stream, const mb_traj_args* a, bool f64) {
  MB_REQUIRE(a && a->text && a->line_start && a->pow5 && a->center && a->quat && a->times && a->box && a->energies && a->status,
             MB_EINVAL_SHAPE, "traj_parse: missing buffers");
  MB_REQUIRE(a->n > 0 && a->n_frames > 0, MB_EINVAL_SHAPE, "traj_parse: n and n_frames must be positive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long lines = (long long)a->n_frames * (a->n + 3);
  const unsigned blocks = unsigned((lines + 127) / 128);
  const unsigned char* t = static_cast<const unsigned char*>(a->text);
  const long long* ls = reinterpret_cast<const long long*>(a->line_start);
  MB_CUDA_CHECK(cudaMemsetAsync(a->status, 0, 4 * sizeof(int), s));
  if (f64)
    mb::k_traj_parse<double><<<blocks, 128, 0, s>>>(t, a->n_bytes, ls, a->n_lines, a->n, a->n_frames, a->dest, a->pow5,
                                                    static_cast<double*>(a->center), static_cast<double*>(a->quat), a->times, a->box,
                                                    a->energies, a->status);
  else
    mb::k_traj_parse<float><<<blocks, 128, 0, s>>>(t, a->n_bytes, ls, a->n_lines, a->n, a->n_frames, a->dest, a->pow5,
                                                   static_cast<float*>(a->center), static_cast<float*>(a->quat), a->times, a->box,
                                                   a->energies, a->status);
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}
int mythos_b200_traj_parse_f64(void* stream, const mb_traj_args* a) { return traj_parse_impl(stream, a, true); }
int mythos_b200_traj_parse_f32(void* stream, const mb_traj_args* a) { return traj_parse_impl(stream, a, false); }
}
