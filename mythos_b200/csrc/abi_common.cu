// abi_common.cu -- introspection entry points and the thread-local error string of the C-ABI.
#include <cstdarg>

#include "common.cuh"

namespace mb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static const char* const kParamNames[] = {
#define MB_X_NAME(id, name) name,
    MB_PARAM_LIST(MB_X_NAME)
#undef MB_X_NAME
};
}  // namespace mb

extern "C" {
int mythos_b200_abi_version(void) { return MB_ABI_VERSION; }
int mythos_b200_param_count(void) { return MB_P_COUNT; }
const char* mythos_b200_param_name(int index) {
  return (index >= 0 && index < MB_P_COUNT_RAW) ? mb::kParamNames[index] : nullptr;
}
int mythos_b200_param_index(const char* name) {
  if (!name) return -1;
  for (int i = 0; i < MB_P_COUNT_RAW; ++i)
    if (std::strcmp(name, mb::kParamNames[i]) == 0) return i;
  return -1;
}
const char* mythos_b200_last_error(void) { return mb::g_err; }
size_t mythos_b200_sizeof_model(void) { return sizeof(mb_model); }
size_t mythos_b200_sizeof_energy_args(void) { return sizeof(mb_energy_args); }
size_t mythos_b200_sizeof_nl_args(void) { return sizeof(mb_nl_args); }
}
