// observables.cu -- standalone per-frame observables kernel + C-ABI entry points (see observables_dev.cuh).
#include "observables_dev.cuh"

namespace mb {

constexpr int kObsThreads = 128;

template <class T>
__global__ void __launch_bounds__(kObsThreads) k_observables(const ModelT<T> M, int n, const T* __restrict__ center,
                                                            const T* __restrict__ quat, const int32_t* __restrict__ nt_type,
                                                            const ObsDev o, T* __restrict__ out) {
  __shared__ T red[4 * (kObsThreads / 32)];
  const long long base = (long long)blockIdx.x * n;
  auto nuc = [&](int i) {
    T q[4];
    return load_nuc(center, quat, base + i, q);
  };
  auto geom = [&](int i) -> const Geom<T>& { return M.geom[(nt_type && nt_type[i] == 2) ? 1 : 0]; };
  frame_observables<T>(o, M.box, nuc, geom, red, out + (long long)blockIdx.x * MB_N_OBS);
}

template <class T>
int launch_observables(cudaStream_t s, const ModelT<T>& M, int n, int n_frames, const T* center, const T* quat,
                       const int32_t* nt_type, const ObsDev& o, T* out) {
  k_observables<T><<<n_frames, kObsThreads, 0, s>>>(M, n, center, quat, nt_type, o, out);
  MB_CUDA_CHECK(cudaGetLastError());
  return MB_OK;
}
template int launch_observables<float>(cudaStream_t, const ModelT<float>&, int, int, const float*, const float*, const int32_t*, const ObsDev&, float*);
template int launch_observables<double>(cudaStream_t, const ModelT<double>&, int, int, const double*, const double*, const int32_t*, const ObsDev&, double*);

int check_observable_spec(const mb_observable_spec* spec, ObsDev* o) {
  MB_REQUIRE(spec, MB_EINVAL_SHAPE, "observables: null spec");
  MB_REQUIRE(spec->n_base_pairs >= 0 && spec->n_quartets >= 0, MB_EINVAL_SHAPE, "observables: negative list length");
  MB_REQUIRE(spec->n_base_pairs == 0 || spec->base_pairs, MB_EINVAL_SHAPE, "observables: base_pairs missing");
  MB_REQUIRE(spec->n_quartets == 0 || spec->quartets, MB_EINVAL_SHAPE, "observables: quartets missing");
  o->base_pairs = spec->base_pairs;
  o->quartets = spec->quartets;
  o->n_base_pairs = spec->n_base_pairs;
  o->n_quartets = spec->n_quartets;
  o->sigma_backbone = spec->sigma_backbone;
  return MB_OK;
}

template <class T>
static int observables_impl(cudaStream_t s, const mb_model* model, int n, int n_frames, const void* center, const void* quat,
                            const int32_t* nt_type, const mb_observable_spec* spec, void* out) {
  MB_REQUIRE(model && center && quat && out, MB_EINVAL_SHAPE, "observables: model / center / quat / out required");
  MB_REQUIRE(n > 0 && n_frames > 0, MB_EINVAL_SHAPE, "observables: n and n_frames must be positive");
  MB_REQUIRE(model->n_banks == 1 || nt_type, MB_EINVAL_SHAPE, "observables: nt_type required for the 3-bank (NA1) model");
  ObsDev o;
  const int st = check_observable_spec(spec, &o);
  if (st != MB_OK) return st;
  ModelT<T> M;
  M.load(*model);
  return launch_observables<T>(s, M, n, n_frames, static_cast<const T*>(center), static_cast<const T*>(quat),
                               model->n_banks == 1 ? nullptr : nt_type, o, static_cast<T*>(out));
}

}  // namespace mb

extern "C" {
int mythos_b200_observables_f64(void* stream, const mb_model* model, int32_t n, int32_t n_frames, const void* center, const void* quat,
                                const int32_t* nt_type, const mb_observable_spec* spec, void* out) {
  return mb::observables_impl<double>(static_cast<cudaStream_t>(stream), model, n, n_frames, center, quat, nt_type, spec, out);
}
int mythos_b200_observables_f32(void* stream, const mb_model* model, int32_t n, int32_t n_frames, const void* center, const void* quat,
                                const int32_t* nt_type, const mb_observable_spec* spec, void* out) {
  return mb::observables_impl<float>(static_cast<cudaStream_t>(stream), model, n, n_frames, center, quat, nt_type, spec, out);
}
}
