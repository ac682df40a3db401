// folp_nccl.cu -- run-time binding of the handful of NCCL entry points the row-
// partitioned solver uses (see folp_nccl.h).
#include "folp_nccl.h"

#include <dlfcn.h>

#include <mutex>

namespace folp {

namespace {
NcclApi g_api;
bool g_ok = false;
std::string g_err;
std::once_flag g_once;

template <class F>
bool bind(void* lib, const char* name, F* out) {
  *out = reinterpret_cast<F>(dlsym(lib, name));
  if (!*out) g_err = std::string("libnccl.so.2 lacks ") + name;
  return *out != nullptr;
}

void load() {
  // RTLD_NOLOAD first: reuse the copy the host process (e.g. torch) already mapped
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!lib) {
    const char* e = dlerror();
    g_err = std::string("cannot load libnccl.so.2: ") + (e ? e : "unknown error");
    return;
  }
  g_ok = bind(lib, "ncclGetUniqueId", &g_api.GetUniqueId) &&
         bind(lib, "ncclCommInitRank", &g_api.CommInitRank) &&
         bind(lib, "ncclCommDestroy", &g_api.CommDestroy) &&
         bind(lib, "ncclAllGather", &g_api.AllGather) &&
         bind(lib, "ncclReduceScatter", &g_api.ReduceScatter) &&
         bind(lib, "ncclGetErrorString", &g_api.GetErrorString) &&
         bind(lib, "ncclGetVersion", &g_api.GetVersion);
}
}  // namespace

const NcclApi* nccl_api(std::string* err) {
  std::call_once(g_once, load);
  if (!g_ok) {
    if (err) *err = g_err;
    return nullptr;
  }
  return &g_api;
}

}  // namespace folp
