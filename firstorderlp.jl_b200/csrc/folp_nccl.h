// folp_nccl.h -- NCCL bound at run time (dlopen), so that libfolp_b200.so has no
// link-time dependency: a single-GPU host never loads NCCL, and inside a process
// that already carries an NCCL (torch's bundled one) the same copy is reused.
#pragma once
#include <nccl.h>

#include <string>

namespace folp {

struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclReduceScatter) ReduceScatter = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
};

// Returns nullptr and fills *err if no libnccl.so.2 can be loaded.
const NcclApi* nccl_api(std::string* err);

}  // namespace folp
