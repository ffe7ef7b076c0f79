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

}  // namespace

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
