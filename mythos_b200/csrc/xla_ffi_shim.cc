// xla_ffi_shim.cc -- XLA FFI adapters over the plain C-ABI (include/mythos_b200.h).
//
// NOT part of libmythos_b200.so: it needs the XLA FFI headers that ship inside jaxlib (jax.ffi.include_dir()), which
// do not exist in the build image (no jax / jaxlib).  Where JAX is installed, build it with
//   g++ -O2 -std=c++17 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())") -I include \
//       mythos_b200/csrc/xla_ffi_shim.cc -L mythos_b200 -lmythos_b200 -o mythos_b200/libmythos_b200_xla.so
// and register the handlers from Python as mythos_b200/jax_ffi.py does.  Each handler only ENQUEUES on the stream XLA
// hands it (no host sync, re-entrant), converts a non-zero status into ffi::Error, and forwards buffers untouched:
// the differentiation contract (forward saves inputs only; backward is a second custom call) is wired in Python with
// jax.custom_vjp (jax_ffi.py).  Untested here: it has never been compiled or run in this repository's CI.
#include <cstdint>
#include <cstring>

#include "../../include/mythos_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error status_to_error(int st, const char* what) {
  if (st == MB_OK) return ffi::Error::Success();
  const char* msg = mythos_b200_last_error();
  const auto code = (st == MB_ECUDA) ? ffi::ErrorCode::kInternal : ffi::ErrorCode::kInvalidArgument;
  return ffi::Error(code, std::string(what) + ": " + (msg ? msg : "unknown error"));
}

// The model description travels as one opaque byte attribute (the bytes of mb_model), packed on the Python side.
bool load_model(ffi::Span<const uint8_t> bytes, mb_model* m) {
  if (bytes.size() != sizeof(mb_model)) return false;
  std::memcpy(m, bytes.data(), sizeof(mb_model));
  return true;
}

template <bool F64>
ffi::Error EnergyImpl(cudaStream_t stream, ffi::AnyBuffer center, ffi::AnyBuffer quat, ffi::AnyBuffer params,
                      ffi::AnyBuffer cot, ffi::Buffer<ffi::S32> seq, ffi::Buffer<ffi::S32> nt_type,
                      ffi::Buffer<ffi::S32> is_end, ffi::Buffer<ffi::S32> bonded, ffi::Buffer<ffi::S32> pairs,
                      ffi::Result<ffi::AnyBuffer> terms, ffi::Result<ffi::AnyBuffer> d_center,
                      ffi::Result<ffi::AnyBuffer> d_quat, ffi::Result<ffi::AnyBuffer> d_params,
                      ffi::Span<const uint8_t> model_bytes, int32_t term_mask, int32_t want_grads) {
  mb_model model;
  if (!load_model(model_bytes, &model)) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "bad mb_model attribute");
  const auto dims = center.dimensions();  // (F, N, 3)
  if (dims.size() != 3 || dims[2] != 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "center must be (F,N,3)");
  mb_energy_args a{};
  a.model = &model;
  a.n_frames = static_cast<int32_t>(dims[0]);
  a.n = static_cast<int32_t>(dims[1]);
  a.center = center.untyped_data();
  a.quat = quat.untyped_data();
  a.seq = seq.typed_data();
  a.nt_type = nt_type.element_count() ? nt_type.typed_data() : nullptr;
  a.is_end = is_end.element_count() ? is_end.typed_data() : nullptr;
  a.bonded = bonded.typed_data();
  a.n_bonded = static_cast<int32_t>(bonded.element_count() / 2);
  const auto pd = pairs.dimensions();  // (2,U) or (F,2,U)
  a.pairs = pairs.element_count() ? pairs.typed_data() : nullptr;
  a.pair_capacity = pd.empty() ? 0 : pd.back();
  a.pair_frame_stride = (pd.size() == 3) ? 2 * pd.back() : 0;
  a.params = params.untyped_data();
  a.cot = cot.element_count() ? cot.untyped_data() : nullptr;
  a.term_mask = static_cast<uint32_t>(term_mask);
  a.terms = terms->untyped_data();
  if (want_grads) {
    a.d_center = d_center->untyped_data();
    a.d_quat = d_quat->untyped_data();
    a.d_params = d_params->untyped_data();
  }
  const int st = F64 ? mythos_b200_energy_f64(stream, &a) : mythos_b200_energy_f32(stream, &a);
  return status_to_error(st, "mythos_b200_energy");
}

// ---- neighbour list: jax_md.partition.neighbor_list(OrderedSparse) as called from mythos/utils/neighbors.py:51-59 -------
// center (F,N,3) -> idx (F,2,capacity) int32 padded with N, count (F), overflow (1) (= did_buffer_overflow).  The scratch
// the build needs arrives as one more result buffer (XLA owns every allocation; size from mythos_b200_nl_workspace_bytes).
template <bool F64>
ffi::Error NlBuildImpl(cudaStream_t stream, ffi::AnyBuffer center, ffi::Buffer<ffi::S32> bonded,
                       ffi::Result<ffi::Buffer<ffi::S32>> pairs, ffi::Result<ffi::Buffer<ffi::S32>> count,
                       ffi::Result<ffi::Buffer<ffi::S32>> overflow, ffi::Result<ffi::Buffer<ffi::U8>> workspace,
                       ffi::Span<const double> box, double r_cutoff, double dr_threshold) {
  const auto dims = center.dimensions();  // (F, N, 3)
  if (dims.size() != 3 || dims[2] != 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "center must be (F,N,3)");
  if (box.size() != 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "box must have 3 entries (zeros = free space)");
  const auto pd = pairs->dimensions();  // (F, 2, capacity)
  if (pd.size() != 3 || pd[0] != dims[0] || pd[1] != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "pairs must be (F,2,capacity)");
  mb_nl_args a{};
  a.n_frames = static_cast<int32_t>(dims[0]);
  a.n = static_cast<int32_t>(dims[1]);
  a.center = center.untyped_data();
  a.bonded = bonded.typed_data();
  a.n_bonded = static_cast<int32_t>(bonded.element_count() / 2);
  for (int d = 0; d < 3; ++d) a.box[d] = box[d];
  a.r_cutoff = r_cutoff;
  a.dr_threshold = dr_threshold;
  a.pairs = pairs->typed_data();
  a.capacity = pd[2];
  a.count = count->typed_data();
  a.overflow = overflow->typed_data();
  a.workspace = workspace->typed_data();
  a.workspace_bytes = workspace->element_count();
  const int st = F64 ? mythos_b200_nl_build_f64(stream, &a) : mythos_b200_nl_build_f32(stream, &a);
  return status_to_error(st, "mythos_b200_nl_build");
}

// ---- rigid-body Langevin step: step_fn of jax_md.simulate.nvt_langevin on RigidBody (simulators/jax_md/jaxmd.py:73,82-94)
// XLA buffers are immutable values, so the state goes in as operands and comes out as results that ALIAS them
// (input_output_aliases in the ffi_call): the kernel updates in place, exactly as the C-ABI expects.
template <bool F64>
ffi::Error LangevinImpl(cudaStream_t stream, ffi::AnyBuffer center, ffi::AnyBuffer quat, ffi::AnyBuffer p_center,
                        ffi::AnyBuffer p_quat, ffi::AnyBuffer d_center, ffi::AnyBuffer d_quat, ffi::AnyBuffer noise,
                        ffi::Result<ffi::AnyBuffer> center_out, ffi::Result<ffi::AnyBuffer> quat_out,
                        ffi::Result<ffi::AnyBuffer> p_center_out, ffi::Result<ffi::AnyBuffer> p_quat_out, double dt, double kT,
                        double gamma_center, double gamma_quat, double mass, ffi::Span<const double> inertia,
                        ffi::Span<const double> box, int64_t seed, int64_t step, int32_t phase) {
  const auto dims = center.dimensions();  // (N, 3)
  if (dims.size() != 2 || dims[1] != 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "center must be (N,3)");
  if (inertia.size() != 3 || box.size() != 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "inertia and box must have 3 entries");
  if (center_out->untyped_data() != center.untyped_data() || quat_out->untyped_data() != quat.untyped_data() ||
      p_center_out->untyped_data() != p_center.untyped_data() || p_quat_out->untyped_data() != p_quat.untyped_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "state results must alias the state operands (input_output_aliases)");
  mb_langevin_args a{};
  a.n = static_cast<int32_t>(dims[0]);
  a.center = center_out->untyped_data();
  a.quat = quat_out->untyped_data();
  a.p_center = p_center_out->untyped_data();
  a.p_quat = p_quat_out->untyped_data();
  a.d_center = d_center.untyped_data();
  a.d_quat = d_quat.untyped_data();
  a.dt = dt;
  a.kT = kT;
  a.gamma_center = gamma_center;
  a.gamma_quat = gamma_quat;
  a.mass = mass;
  for (int d = 0; d < 3; ++d) {
    a.inertia[d] = inertia[d];
    a.box[d] = box[d];
  }
  a.seed = static_cast<uint64_t>(seed);
  a.step = static_cast<uint64_t>(step);
  a.noise = noise.element_count() ? noise.untyped_data() : nullptr;  // (N,6) injected normals (tests) or Philox inside
  a.phase = phase;
  const int st = F64 ? mythos_b200_langevin_f64(stream, &a) : mythos_b200_langevin_f32(stream, &a);
  return status_to_error(st, "mythos_b200_langevin");
}

// ---- DiffTRe weights: compute_weights_and_neff (mythos/optimization/objective.py:139-163) --------------------------------
template <bool F64>
ffi::Error WeightsImpl(cudaStream_t stream, ffi::AnyBuffer beta, ffi::AnyBuffer e_new, ffi::AnyBuffer e_ref,
                       ffi::Result<ffi::AnyBuffer> weights, ffi::Result<ffi::AnyBuffer> sums) {
  const auto n = e_new.element_count();
  if (beta.element_count() != n || e_ref.element_count() != n || weights->element_count() != n || sums->element_count() != 4)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "beta, e_new, e_ref, weights must be (F) and sums (4)");
  mb_weights_args a{};
  a.n_frames = static_cast<int32_t>(n);
  a.beta = beta.untyped_data();
  a.e_new = e_new.untyped_data();
  a.e_ref = e_ref.untyped_data();
  a.weights = weights->untyped_data();
  a.sums = sums->untyped_data();
  const int st = F64 ? mythos_b200_weights_neff_f64(stream, &a) : mythos_b200_weights_neff_f32(stream, &a);
  return status_to_error(st, "mythos_b200_weights_neff");
}

}  // namespace

#define MB_BIND_NL()                                                                                                 \
  ffi::Ffi::Bind()                                                                                                   \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()                                                                      \
      .Arg<ffi::AnyBuffer>()            /* center */                                                                 \
      .Arg<ffi::Buffer<ffi::S32>>()     /* bonded */                                                                 \
      .Ret<ffi::Buffer<ffi::S32>>()     /* pairs (F,2,capacity) */                                                   \
      .Ret<ffi::Buffer<ffi::S32>>()     /* count (F) */                                                              \
      .Ret<ffi::Buffer<ffi::S32>>()     /* overflow (1) */                                                           \
      .Ret<ffi::Buffer<ffi::U8>>()      /* workspace */                                                              \
      .Attr<ffi::Span<const double>>("box")                                                                          \
      .Attr<double>("r_cutoff")                                                                                      \
      .Attr<double>("dr_threshold")

#define MB_BIND_LANGEVIN()                                                                                           \
  ffi::Ffi::Bind()                                                                                                   \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()                                                                      \
      .Arg<ffi::AnyBuffer>() /* center */                                                                            \
      .Arg<ffi::AnyBuffer>() /* quat */                                                                              \
      .Arg<ffi::AnyBuffer>() /* p_center */                                                                          \
      .Arg<ffi::AnyBuffer>() /* p_quat */                                                                            \
      .Arg<ffi::AnyBuffer>() /* d_center */                                                                          \
      .Arg<ffi::AnyBuffer>() /* d_quat */                                                                            \
      .Arg<ffi::AnyBuffer>() /* noise (N,6) or empty */                                                              \
      .Ret<ffi::AnyBuffer>() /* center (aliased) */                                                                  \
      .Ret<ffi::AnyBuffer>() /* quat (aliased) */                                                                    \
      .Ret<ffi::AnyBuffer>() /* p_center (aliased) */                                                                \
      .Ret<ffi::AnyBuffer>() /* p_quat (aliased) */                                                                  \
      .Attr<double>("dt")                                                                                            \
      .Attr<double>("kT")                                                                                            \
      .Attr<double>("gamma_center")                                                                                  \
      .Attr<double>("gamma_quat")                                                                                    \
      .Attr<double>("mass")                                                                                          \
      .Attr<ffi::Span<const double>>("inertia")                                                                      \
      .Attr<ffi::Span<const double>>("box")                                                                          \
      .Attr<int64_t>("seed")                                                                                         \
      .Attr<int64_t>("step")                                                                                         \
      .Attr<int32_t>("phase")

#define MB_BIND_WEIGHTS()                                                                                            \
  ffi::Ffi::Bind()                                                                                                   \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()                                                                      \
      .Arg<ffi::AnyBuffer>() /* beta */                                                                              \
      .Arg<ffi::AnyBuffer>() /* e_new */                                                                             \
      .Arg<ffi::AnyBuffer>() /* e_ref */                                                                             \
      .Ret<ffi::AnyBuffer>() /* weights (F) */                                                                       \
      .Ret<ffi::AnyBuffer>() /* sums (4) */

XLA_FFI_DEFINE_HANDLER_SYMBOL(mythos_b200_xla_nl_build_f64, NlBuildImpl<true>, MB_BIND_NL());
XLA_FFI_DEFINE_HANDLER_SYMBOL(mythos_b200_xla_nl_build_f32, NlBuildImpl<false>, MB_BIND_NL());
XLA_FFI_DEFINE_HANDLER_SYMBOL(mythos_b200_xla_langevin_f64, LangevinImpl<true>, MB_BIND_LANGEVIN());
XLA_FFI_DEFINE_HANDLER_SYMBOL(mythos_b200_xla_langevin_f32, LangevinImpl<false>, MB_BIND_LANGEVIN());
XLA_FFI_DEFINE_HANDLER_SYMBOL(mythos_b200_xla_weights_neff_f64, WeightsImpl<true>, MB_BIND_WEIGHTS());
XLA_FFI_DEFINE_HANDLER_SYMBOL(mythos_b200_xla_weights_neff_f32, WeightsImpl<false>, MB_BIND_WEIGHTS());

#define MB_BIND_ENERGY()                                                                                             \
  ffi::Ffi::Bind()                                                                                                   \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()                                                                      \
      .Arg<ffi::AnyBuffer>() /* center */                                                                            \
      .Arg<ffi::AnyBuffer>() /* quat   */                                                                            \
      .Arg<ffi::AnyBuffer>() /* params */                                                                            \
      .Arg<ffi::AnyBuffer>() /* cot    */                                                                            \
      .Arg<ffi::Buffer<ffi::S32>>() /* seq */                                                                        \
      .Arg<ffi::Buffer<ffi::S32>>() /* nt_type */                                                                    \
      .Arg<ffi::Buffer<ffi::S32>>() /* is_end */                                                                     \
      .Arg<ffi::Buffer<ffi::S32>>() /* bonded */                                                                     \
      .Arg<ffi::Buffer<ffi::S32>>() /* pairs */                                                                      \
      .Ret<ffi::AnyBuffer>() /* terms */                                                                             \
      .Ret<ffi::AnyBuffer>() /* d_center */                                                                          \
      .Ret<ffi::AnyBuffer>() /* d_quat */                                                                            \
      .Ret<ffi::AnyBuffer>() /* d_params */                                                                          \
      .Attr<ffi::Span<const uint8_t>>("model")                                                                       \
      .Attr<int32_t>("term_mask")                                                                                    \
      .Attr<int32_t>("want_grads")

XLA_FFI_DEFINE_HANDLER_SYMBOL(mythos_b200_xla_energy_f64, EnergyImpl<true>, MB_BIND_ENERGY());
XLA_FFI_DEFINE_HANDLER_SYMBOL(mythos_b200_xla_energy_f32, EnergyImpl<false>, MB_BIND_ENERGY());
