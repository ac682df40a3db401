// folp_vmm.h -- the exchange region of the partitioned mode as an NVSwitch MULTICAST object (NVLS):
// every rank's region is physical memory from cuMemCreate, mapped three ways on every rank --
// its own memory, every peer's memory (unicast, for flags and scalars) and ONE multicast address
// range on which a single store lands in all ranks' regions at the same offset (replicated inside
// the switch). take_step pushes xbar and y+ with one multimem.st per element instead of
// world-1 posted stores (SURVEY 8e "K1 broadcasts xbar with multimem stores"); measured on 8 B200s
// (tools/mc_probe.cu, profiles/r02_mc_probe_x8.txt): 10 MB per GPU to all, 113 us as unicast stores,
// 97 us through the multicast object; behind an SpMV-like producer 137 us against 113 us.
//
// The driver API is bound at run time (cudaGetDriverEntryPoint): libfolp_b200.so keeps no link-time
// dependency on libcuda and loads on hosts without a driver. Across processes the allocation handles
// travel as POSIX file descriptors over abstract unix-domain sockets (SCM_RIGHTS).
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>

namespace folp {

constexpr int kVmmMaxRanks = 8;
constexpr int kVmmWords = 4;

// How the ranks of one handle talk while the region is set up: every rank contributes kVmmWords
// words and receives world * kVmmWords (rank-major). Must be called by all ranks in the same order.
// Returns false when the exchange itself failed (then on every rank).
using VmmAllgather = std::function<bool(const uint64_t* mine, uint64_t* all)>;

struct VmmRegion {
  bool active = false;
  bool cross_process = false;
  int world = 0, rank = 0, device = 0;
  size_t size = 0;                          // mapped bytes (a multiple of the multicast granularity)
  void* own = nullptr;                      // this rank's memory
  void* peer[kVmmMaxRanks] = {};            // every rank's memory as this rank addresses it (peer[rank] == own)
  void* mc = nullptr;                       // the multicast range
  // driver objects (CUmemGenericAllocationHandle / CUdeviceptr as integers)
  unsigned long long h_own = 0, h_mc = 0, h_peer[kVmmMaxRanks] = {};
  bool bound = false;
};

// Collective over the ranks of a handle. On success (true on every rank) the region is mapped and
// zero-filled and no rank has written into a peer yet; on failure (false on every rank) nothing stays
// allocated and *why says what this rank saw. `devices` = the CUDA device of every rank when all
// ranks are threads of this process (handles are shared by value), nullptr across processes.
bool vmm_region_create(const VmmAllgather& allgather, int rank, int world, int device, const int* devices,
                       size_t bytes, VmmRegion* out, const char** why);
// Tear-down in two steps: ranks that are threads of one process share their handles by value, so every
// rank must have unmapped / unbound (vmm_region_unmap) before any rank releases (vmm_region_release).
// Separate processes hold their own references and may call vmm_region_destroy (= both) at any time.
void vmm_region_unmap(VmmRegion* r);
void vmm_region_release(VmmRegion* r);
void vmm_region_destroy(VmmRegion* r);

}  // namespace folp
