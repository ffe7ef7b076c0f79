// theta_tape.cu -- host-side evaluator of the theta -> kernel-parameter-bank chain and of its vector-Jacobian product.
//
// The reference runs `init_params` (mythos/energy/configuration.py:110-113; e.g. dna1/stacking.py:120-183) and the
// smoothing solvers (mythos/energy/dna1/base_smoothing_functions.py:48-142) inside every jitted step, so XLA compiles
// them once into a few fused kernels.  Here the same chain is recorded ONCE per energy function as a straight-line
// scalar tape (mythos_b200/energy/theta_tape.py lowers the traced graph of the Python chain) and replayed by this file:
// a few thousand double-precision operations forward, the same number in the reverse sweep -- microseconds instead of the
// milliseconds a 700-node eager autograd graph costs per DiffTRe step.  Plain host code: no device, no stream.
#include <cmath>

#include "common.cuh"

namespace mb {
static inline int tape_forward(const mb_theta_tape* t, const double* in, double* v) {
  for (int k = 0; k < t->n_nodes; ++k) {
    const int a = t->arg0[k], b = t->arg1[k];
    double r;
    switch (t->op[k]) {
      case MB_TAPE_CONST: r = t->imm[k]; break;
      case MB_TAPE_INPUT: r = in[a]; break;
      case MB_TAPE_ADD: r = v[a] + v[b]; break;
      case MB_TAPE_SUB: r = v[a] - v[b]; break;
      case MB_TAPE_MUL: r = v[a] * v[b]; break;
      case MB_TAPE_DIV: r = v[a] / v[b]; break;
      case MB_TAPE_NEG: r = -v[a]; break;
      case MB_TAPE_RECIP: r = 1.0 / v[a]; break;
      case MB_TAPE_EXP: r = std::exp(v[a]); break;
      case MB_TAPE_LOG: r = std::log(v[a]); break;
      case MB_TAPE_SQRT: r = std::sqrt(v[a]); break;
      case MB_TAPE_POW: {
        const double p = t->imm[k];
        r = (p == 2.0) ? v[a] * v[a] : std::pow(v[a], p);
        break;
      }
      default: return -1;
    }
    v[k] = r;
  }
  return 0;
}
}  // namespace mb

extern "C" {
int mythos_b200_theta_tape_forward(const mb_theta_tape* t, const double* inputs, double* values, double* outputs) {
  MB_REQUIRE(t && inputs && values && outputs && t->n_nodes >= 0, MB_EINVAL_SHAPE, "theta tape: null argument");
  for (int k = 0; k < t->n_nodes; ++k) {
    const int op = t->op[k], a = t->arg0[k], b = t->arg1[k];
    const bool unary = op >= MB_TAPE_NEG, binary = op >= MB_TAPE_ADD && op <= MB_TAPE_DIV;
    MB_REQUIRE(op >= 0 && op <= MB_TAPE_POW, MB_EINVAL_SHAPE, "theta tape: unknown opcode");
    MB_REQUIRE(op != MB_TAPE_INPUT || (a >= 0 && a < t->n_inputs), MB_EINVAL_SHAPE, "theta tape: input index out of range");
    MB_REQUIRE(!(unary || binary) || (a >= 0 && a < k), MB_EINVAL_SHAPE, "theta tape: operand does not precede its use");
    MB_REQUIRE(!binary || (b >= 0 && b < k), MB_EINVAL_SHAPE, "theta tape: operand does not precede its use");
  }
  MB_REQUIRE(mb::tape_forward(t, inputs, values) == 0, MB_EINVAL_SHAPE, "theta tape: unknown opcode");
  for (int o = 0; o < t->n_outputs; ++o) {
    const int k = t->out[o];
    MB_REQUIRE(k >= -1 && k < t->n_nodes, MB_EINVAL_SHAPE, "theta tape: output index out of range");
    outputs[o] = k < 0 ? 0.0 : values[k];  // -1: a structurally empty bank slot
  }
  return MB_OK;
}

int mythos_b200_theta_tape_vjp(const mb_theta_tape* t, const double* values, const double* out_cot, double* adjoint,
                               double* in_grad) {
  MB_REQUIRE(t && values && out_cot && adjoint && in_grad, MB_EINVAL_SHAPE, "theta tape: null argument");
  const double* v = values;
  double* g = adjoint;
  for (int k = 0; k < t->n_nodes; ++k) g[k] = 0.0;
  for (int i = 0; i < t->n_inputs; ++i) in_grad[i] = 0.0;
  for (int o = 0; o < t->n_outputs; ++o)
    if (t->out[o] >= 0) g[t->out[o]] += out_cot[o];
  for (int k = t->n_nodes - 1; k >= 0; --k) {
    const double gk = g[k];
    if (gk == 0.0) continue;
    const int a = t->arg0[k], b = t->arg1[k];
    switch (t->op[k]) {
      case MB_TAPE_CONST: break;
      case MB_TAPE_INPUT: in_grad[a] += gk; break;
      case MB_TAPE_ADD: g[a] += gk; g[b] += gk; break;
      case MB_TAPE_SUB: g[a] += gk; g[b] -= gk; break;
      case MB_TAPE_MUL: g[a] += gk * v[b]; g[b] += gk * v[a]; break;
      case MB_TAPE_DIV: g[a] += gk / v[b]; g[b] -= gk * v[k] / v[b]; break;
      case MB_TAPE_NEG: g[a] -= gk; break;
      case MB_TAPE_RECIP: g[a] -= gk * v[k] * v[k]; break;
      case MB_TAPE_EXP: g[a] += gk * v[k]; break;
      case MB_TAPE_LOG: g[a] += gk / v[a]; break;
      case MB_TAPE_SQRT: g[a] += gk * 0.5 / v[k]; break;
      case MB_TAPE_POW: {
        const double p = t->imm[k];
        g[a] += gk * ((p == 2.0) ? 2.0 * v[a] : p * std::pow(v[a], p - 1.0));
        break;
      }
      default: return MB_EINVAL_SHAPE;
    }
  }
  return MB_OK;
}
}
