// parse_decimal.cuh -- correctly rounded decimal -> binary64 conversion for the trajectory parser.
//
// The reference parses every number of an oxDNA trajectory with np.fromstring (strtod: correctly rounded,
// mythos/input/trajectory.py:262).  To reproduce it bit for bit on the device:
//   * Clinger's fast path when the decimal significand fits 53 bits and |exponent| <= 22 (both operands exact, one IEEE
//     operation) -- every number oxDNA itself writes (15 significant digits);
//   * otherwise the Eisel-Lemire algorithm (D. Lemire, "Number parsing at a gigabyte per second", 2021) on a 64-bit
//     significand with a table of 128-bit truncated powers of five for exponents in [kPow5Min, kPow5Max] -- the 16/17-digit
//     numbers that mythos' own writer produces (str(float)).  Inside that exponent window the algorithm always decides;
//   * anything else (more than 19 significant digits, exponents outside the table) reports failure: the caller raises.
// __host__ __device__ so the CPU test-suite can check it against strtod on millions of inputs (tests/host_check).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MB_PD __host__ __device__ __forceinline__
#else
#define MB_PD inline
#endif

namespace mb {

constexpr int kPow5Min = -64, kPow5Max = 64;  // table rows: 5^q for q in [kPow5Min, kPow5Max], 128-bit truncated, normalised

MB_PD void mul64(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umul64hi(a, b);
#else
  const unsigned __int128 p = (unsigned __int128)a * b;
  lo = (uint64_t)p;
  hi = (uint64_t)(p >> 64);
#endif
}
MB_PD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return __builtin_clzll(x);
#endif
}
MB_PD double bits_to_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double d;
  __builtin_memcpy(&d, &b, 8);
  return d;
#endif
}

// w * 10^q, w != 0 a 64-bit decimal significand -> nearest double (ties to even).  pow5: (kPow5Max-kPow5Min+1, 2) uint64
// rows {hi, lo}.  Returns false if it cannot decide (q outside the table).
MB_PD bool eisel_lemire(uint64_t w, int q, const uint64_t* pow5, double& out) {
  if (q < kPow5Min || q > kPow5Max) return false;
  int lz = clz64(w);
  w <<= lz;
  const uint64_t p_hi = pow5[2 * (q - kPow5Min)], p_lo = pow5[2 * (q - kPow5Min) + 1];
  uint64_t hi, lo;
  mul64(w, p_hi, hi, lo);
  if ((hi & 0x1FF) == 0x1FF) {  // the truncated product may not determine the rounding: add the low word's contribution
    uint64_t hi2, lo2;
    mul64(w, p_lo, hi2, lo2);
    const uint64_t before = lo;
    lo += hi2;
    if (lo < before) ++hi;
  }
  const int upper = int(hi >> 63);
  uint64_t m = hi >> (upper + 64 - 52 - 3);  // 54 bits: 53 + one rounding bit
  lz += 1 ^ upper;
  // binary exponent of 10^q's leading bit: floor(log2(5^q)) + q = ((152170 + 65536) * q) >> 16
  int e2 = int((int64_t(217706) * q) >> 16) + 63 - lz + 1023 + 1;
  if (e2 <= 0 || e2 >= 2047) return false;  // subnormal / overflow: not a coordinate this parser accepts
  // exactly halfway between two doubles? only possible for small |q| (5^q must divide out); then round to even
  if (lo <= 1 && q >= -4 && q <= 23 && (m & 3) == 1 && (m << (upper + 64 - 52 - 3)) == hi) m &= ~uint64_t(1);
  m += (m & 1);
  m >>= 1;
  if (m >= (uint64_t(2) << 52)) {
    m = uint64_t(1) << 52;
    ++e2;
  }
  m &= ~(uint64_t(1) << 52);
  if (e2 >= 2047) return false;
  out = bits_to_double(m | (uint64_t(e2) << 52));
  return true;
}

// digits (as a 64-bit integer, at most 19 of them) x 10^q with sign -> double; false = cannot be converted exactly here
MB_PD bool decimal_to_double(uint64_t w, int q, bool negative, const uint64_t* pow5, double& out) {
  if (w == 0) {
    out = negative ? -0.0 : 0.0;
    return true;
  }
  double v;
  if (w < (uint64_t(1) << 53) && q >= -22 && q <= 22) {  // Clinger: exact operands, one correctly rounded operation
    const double p10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                            1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    v = q < 0 ? double(w) / p10[-q] : double(w) * p10[q];
  } else if (!eisel_lemire(w, q, pow5, v)) {
    return false;
  }
  out = negative ? -v : v;
  return true;
}

// One number starting at p (p < end): [+-]digits[.digits][(e|E)[+-]digits].  Advances p past it.  ok = false on a malformed
// token, more than 19 significant digits, or an exponent outside the table.
MB_PD double parse_number(const unsigned char*& p, const unsigned char* end, const uint64_t* pow5, bool& ok) {
  bool neg = false;
  if (p < end && (*p == '-' || *p == '+')) neg = (*p++ == '-');
  uint64_t w = 0;
  int digits = 0, q = 0, seen = 0;
  while (p < end && *p >= '0' && *p <= '9') {
    if (w != 0 || *p != '0') {
      if (digits < 19) {
        w = w * 10 + (*p - '0');
        ++digits;
      } else {
        ++q;  // (a 20th+ digit: keep the magnitude, flag below)
        ok = false;
      }
    }
    ++p;
    ++seen;
  }
  if (p < end && *p == '.') {
    ++p;
    while (p < end && *p >= '0' && *p <= '9') {
      if (w != 0 || *p != '0') {
        if (digits < 19) {
          w = w * 10 + (*p - '0');
          ++digits;
          --q;
        } else if (*p != '0') {
          ok = false;
        }
      } else {
        --q;  // leading zeros after the point
      }
      ++p;
      ++seen;
    }
  }
  if (!seen) ok = false;
  if (p < end && (*p == 'e' || *p == 'E')) {
    ++p;
    bool eneg = false;
    if (p < end && (*p == '-' || *p == '+')) eneg = (*p++ == '-');
    int ex = 0, eseen = 0;
    while (p < end && *p >= '0' && *p <= '9') {
      if (ex < 10000) ex = ex * 10 + (*p - '0');
      ++p;
      ++eseen;
    }
    if (!eseen) ok = false;
    q += eneg ? -ex : ex;
  }
  double v = 0.0;
  if (!decimal_to_double(w, q, neg, pow5, v)) ok = false;
  return v;
}

}  // namespace mb
