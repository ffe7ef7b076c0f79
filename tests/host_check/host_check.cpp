// host_check.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the device math header (mythos_b200/csrc/oxdna_device.cuh) for the host with g++ so that the
// analytic gradients can be unit-checked against the oracle's autograd in the CPU test suite, without a GPU.
// Nothing in the shipped library links or calls this file; the product path is CUDA-only.
#include <cstring>
#include <vector>

#include "../../mythos_b200/csrc/oxdna_device.cuh"

namespace {
struct ArrayAcc {
  double* out;
  template <class T>
  void add(int bank, int idx, T v) {
    out[bank * MB_P_COUNT + idx] += double(v);
  }
  template <class T>
  void add_scatter(int bank, int idx, T v, bool pred) {
    if (pred) out[bank * MB_P_COUNT + idx] += double(v);
  }
};

template <class T>
void run(const mb_model* model, int n, const double* center, const double* quat, const int* seq, const int* nt_type,
         const int* nt_type_stack, const int* is_end, const int* bonded, int nb, const int* pairs, long cap,
         const double* params, const double* cot_in, unsigned mask, double* terms, double* d_center, double* d_quat,
         double* d_params) {
  using namespace mb;
  ModelT<T> M;
  M.load(*model);
  const int np = model->n_banks * MB_P_COUNT;
  std::vector<T> P(np);
  for (int k = 0; k < np; ++k) P[k] = T(params[k]);
  T cot[MB_N_TERMS];
  for (int t = 0; t < MB_N_TERMS; ++t) cot[t] = cot_in ? T(cot_in[t]) : T(1);
  std::vector<Nuc<T>> nuc(n);
  for (int i = 0; i < n; ++i) {
    nuc[i].c = v3<T>(T(center[3 * i]), T(center[3 * i + 1]), T(center[3 * i + 2]));
    axes_from_quat(T(quat[4 * i]), T(quat[4 * i + 1]), T(quat[4 * i + 2]), T(quat[4 * i + 3]), nuc[i].a1, nuc[i].a2,
                   nuc[i].a3);
  }
  std::vector<NucGrad<T>> G(n);
  for (auto& g : G) g.zero();
  T e[MB_N_TERMS] = {0, 0, 0, 0, 0, 0, 0, 0};
  ArrayAcc acc{d_params};
  auto nt = [&](const int* a, int i) { return a ? a[i] : 1; };
  const int* snt = nt_type_stack ? nt_type_stack : nt_type;
  for (int k = 0; k < nb; ++k) {
    const int i = bonded[2 * k], j = bonded[2 * k + 1];
    NucGrad<T> gi, gj;
    gi.zero();
    gj.zero();
    bonded_pair<T, true, true>(M, P.data(), true, nuc[i], nuc[j], seq[i], seq[j], nt(nt_type, i), nt(nt_type, j),
                               nt(snt, i), nt(snt, j), mask, cot, e, gi, gj, acc);
    G[i].c = G[i].c + gi.c; G[i].a1 = G[i].a1 + gi.a1; G[i].a2 = G[i].a2 + gi.a2; G[i].a3 = G[i].a3 + gi.a3;
    G[j].c = G[j].c + gj.c; G[j].a1 = G[j].a1 + gj.a1; G[j].a2 = G[j].a2 + gj.a2; G[j].a3 = G[j].a3 + gj.a3;
  }
  for (long k = 0; k < cap; ++k) {
    const int i = pairs[k], j = pairs[cap + k];
    if (i >= n || j >= n) continue;
    T m = T(1);
    if (model->half_charged_ends && is_end) m = (is_end[i] ? T(0.5) : T(1)) * (is_end[j] ? T(0.5) : T(1));
    NucGrad<T> gi, gj;
    gi.zero();
    gj.zero();
    unbonded_pair<T, true, true>(M, P.data(), true, nuc[i], nuc[j], seq[i], seq[j], nt(nt_type, i), nt(nt_type, j), m,
                                 mask, cot, e, gi, gj, acc);
    G[i].c = G[i].c + gi.c; G[i].a1 = G[i].a1 + gi.a1; G[i].a2 = G[i].a2 + gi.a2; G[i].a3 = G[i].a3 + gi.a3;
    G[j].c = G[j].c + gj.c; G[j].a1 = G[j].a1 + gj.a1; G[j].a2 = G[j].a2 + gj.a2; G[j].a3 = G[j].a3 + gj.a3;
  }
  for (int t = 0; t < MB_N_TERMS; ++t) terms[t] = double(e[t]);
  for (int i = 0; i < n; ++i) {
    d_center[3 * i] = double(G[i].c.x);
    d_center[3 * i + 1] = double(G[i].c.y);
    d_center[3 * i + 2] = double(G[i].c.z);
    T dq[4];
    quat_grad(G[i], T(quat[4 * i]), T(quat[4 * i + 1]), T(quat[4 * i + 2]), T(quat[4 * i + 3]), dq);
    for (int c = 0; c < 4; ++c) d_quat[4 * i + c] = double(dq[c]);
  }
}
}  // namespace

extern "C" void host_check_eval(int use_f32, const mb_model* model, int n, const double* center, const double* quat,
                                const int* seq, const int* nt_type, const int* nt_type_stack, const int* is_end,
                                const int* bonded, int nb, const int* pairs, long cap, const double* params,
                                const double* cot, unsigned mask, double* terms, double* d_center, double* d_quat,
                                double* d_params) {
  std::memset(d_params, 0, sizeof(double) * model->n_banks * MB_P_COUNT);
  if (use_f32)
    run<float>(model, n, center, quat, seq, nt_type, nt_type_stack, is_end, bonded, nb, pairs, cap, params, cot, mask,
               terms, d_center, d_quat, d_params);
  else
    run<double>(model, n, center, quat, seq, nt_type, nt_type_stack, is_end, bonded, nb, pairs, cap, params, cot, mask,
                terms, d_center, d_quat, d_params);
}

extern "C" const char* host_check_param_name(int i) {
  static const char* names[] = {
#define MB_X_NAME(id, name) name,
      MB_PARAM_LIST(MB_X_NAME)
#undef MB_X_NAME
  };
  return (i >= 0 && i < MB_P_COUNT_RAW) ? names[i] : nullptr;
}

// scalar primitives f1..f6 of the device header (value and d/dx), for the reference's scalar known-answer tests
// (tests/test_scalar_kats.py).  Block layouts as documented in oxdna_device.cuh; f1 has eps == 1 by construction.
extern "C" double host_check_scalar(int use_f32, int kind, double x, const double* p, double eps, double* df_out) {
  using namespace mb;
  double val = 0, df = 0;
  if (use_f32) {
    float pf[16], d = 0;
    for (int k = 0; k < 16; ++k) pf[k] = float(p[k]);
    switch (kind) {
      case 1: val = f1_val<float>(float(x), pf, d); break;
      case 2: val = f2_val<float>(float(x), pf, d); break;
      case 3: val = f3_val<float>(float(x), pf, float(eps), d); break;
      case 4: val = f4_val<float>(float(x), pf, d); break;
      case 5: val = f5_val<float>(float(x), pf, d); break;
      case 6: val = f6_val<float>(float(x), pf, d); break;
    }
    df = d;
  } else {
    switch (kind) {
      case 1: val = f1_val<double>(x, p, df); break;
      case 2: val = f2_val<double>(x, p, df); break;
      case 3: val = f3_val<double>(x, p, eps, df); break;
      case 4: val = f4_val<double>(x, p, df); break;
      case 5: val = f5_val<double>(x, p, df); break;
      case 6: val = f6_val<double>(x, p, df); break;
    }
  }
  if (df_out) *df_out = df;
  return val;
}
