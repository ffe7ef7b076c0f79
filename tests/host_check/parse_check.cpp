// parse_check.cpp -- TEST INFRASTRUCTURE ONLY: the trajectory parser's decimal conversion (mythos_b200/csrc/parse_decimal.cuh)
// compiled for the host so the CPU suite can compare it with strtod on millions of inputs.  Not linked by the library.
#include <cstdlib>
#include <cstring>

#include "../../mythos_b200/csrc/parse_decimal.cuh"

extern "C" {
// parse `n` NUL-separated tokens laid end to end in `buf` (total `len` bytes); out[k] = value, okf[k] = 1 if converted
void parse_check_tokens(const unsigned char* buf, long len, int n, const uint64_t* pow5, double* out, int* okf) {
  const unsigned char* p = buf;
  const unsigned char* end = buf + len;
  for (int k = 0; k < n; ++k) {
    const unsigned char* tok_end = p;
    while (tok_end < end && *tok_end) ++tok_end;
    bool ok = true;
    const unsigned char* q = p;
    out[k] = mb::parse_number(q, tok_end, pow5, ok);
    okf[k] = (ok && q == tok_end) ? 1 : 0;
    p = tok_end + 1;
  }
}
}
