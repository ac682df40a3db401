// folp_vmm.cu -- see folp_vmm.h: the exchange region as cuMemCreate memory bound to an NVSwitch
// multicast object. Host-side setup only; the stores are in folp_kernels.cu (mc_store).
#include "folp_vmm.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace folp {
namespace {

struct DriverApi {
  decltype(&cuDeviceGet) DeviceGet = nullptr;
  decltype(&cuDeviceGetAttribute) DeviceGetAttribute = nullptr;
  decltype(&cuMemCreate) MemCreate = nullptr;
  decltype(&cuMemRelease) MemRelease = nullptr;
  decltype(&cuMemAddressReserve) MemAddressReserve = nullptr;
  decltype(&cuMemAddressFree) MemAddressFree = nullptr;
  decltype(&cuMemMap) MemMap = nullptr;
  decltype(&cuMemUnmap) MemUnmap = nullptr;
  decltype(&cuMemSetAccess) MemSetAccess = nullptr;
  decltype(&cuMemGetAllocationGranularity) MemGetAllocationGranularity = nullptr;
  decltype(&cuMemExportToShareableHandle) MemExportToShareableHandle = nullptr;
  decltype(&cuMemImportFromShareableHandle) MemImportFromShareableHandle = nullptr;
  decltype(&cuMulticastCreate) MulticastCreate = nullptr;
  decltype(&cuMulticastAddDevice) MulticastAddDevice = nullptr;
  decltype(&cuMulticastBindMem) MulticastBindMem = nullptr;
  decltype(&cuMulticastUnbind) MulticastUnbind = nullptr;
  decltype(&cuMulticastGetGranularity) MulticastGetGranularity = nullptr;
  bool ok = false;
};

template <class F>
bool bind_entry(const char* name, F* f) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  *f = reinterpret_cast<F>(p);
  return true;
}

const DriverApi* driver_api() {
  static const DriverApi api = [] {
    DriverApi a;
    a.ok = bind_entry("cuDeviceGet", &a.DeviceGet) && bind_entry("cuDeviceGetAttribute", &a.DeviceGetAttribute) &&
           bind_entry("cuMemCreate", &a.MemCreate) && bind_entry("cuMemRelease", &a.MemRelease) &&
           bind_entry("cuMemAddressReserve", &a.MemAddressReserve) && bind_entry("cuMemAddressFree", &a.MemAddressFree) &&
           bind_entry("cuMemMap", &a.MemMap) && bind_entry("cuMemUnmap", &a.MemUnmap) &&
           bind_entry("cuMemSetAccess", &a.MemSetAccess) &&
           bind_entry("cuMemGetAllocationGranularity", &a.MemGetAllocationGranularity) &&
           bind_entry("cuMemExportToShareableHandle", &a.MemExportToShareableHandle) &&
           bind_entry("cuMemImportFromShareableHandle", &a.MemImportFromShareableHandle) &&
           bind_entry("cuMulticastCreate", &a.MulticastCreate) && bind_entry("cuMulticastAddDevice", &a.MulticastAddDevice) &&
           bind_entry("cuMulticastBindMem", &a.MulticastBindMem) && bind_entry("cuMulticastUnbind", &a.MulticastUnbind) &&
           bind_entry("cuMulticastGetGranularity", &a.MulticastGetGranularity);
    return a;
  }();
  return api.ok ? &api : nullptr;
}

// ---- file descriptors between the ranks' processes: abstract unix datagram sockets ----
socklen_t make_addr(sockaddr_un* a, uint64_t nonce, int rank) {
  memset(a, 0, sizeof(*a));
  a->sun_family = AF_UNIX;  // sun_path[0] == 0: abstract namespace, nothing to unlink
  const int k = snprintf(a->sun_path + 1, sizeof(a->sun_path) - 1, "folp-b200-%016llx-%d",
                         static_cast<unsigned long long>(nonce), rank);
  return static_cast<socklen_t>(offsetof(sockaddr_un, sun_path) + 1 + k);
}
int open_socket(uint64_t nonce, int rank) {
  const int s = socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
  if (s < 0) return -1;
  sockaddr_un a;
  const socklen_t len = make_addr(&a, nonce, rank);
  if (bind(s, reinterpret_cast<sockaddr*>(&a), len) != 0) {
    close(s);
    return -1;
  }
  return s;
}
bool send_fd(int s, uint64_t nonce, int to, int from, int kind, int fd) {
  int payload[2] = {from, kind};
  iovec io{payload, sizeof(payload)};
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  sockaddr_un a;
  const socklen_t len = make_addr(&a, nonce, to);
  msghdr m;
  memset(&m, 0, sizeof(m));
  m.msg_name = &a;
  m.msg_namelen = len;
  m.msg_iov = &io;
  m.msg_iovlen = 1;
  m.msg_control = ctrl;
  m.msg_controllen = sizeof(ctrl);
  cmsghdr* c = CMSG_FIRSTHDR(&m);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(c), &fd, sizeof(int));
  for (int attempt = 0; attempt < 200; ++attempt) {
    if (sendmsg(s, &m, 0) == static_cast<ssize_t>(sizeof(payload))) return true;
    usleep(5000);  // receiver's queue full (or not scheduled yet): try again for a second
  }
  return false;
}
bool recv_fd(int s, int* from, int* kind, int* fd, int timeout_ms) {
  pollfd p{s, POLLIN, 0};
  if (poll(&p, 1, timeout_ms) <= 0) return false;
  int payload[2] = {-1, -1};
  iovec io{payload, sizeof(payload)};
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  msghdr m;
  memset(&m, 0, sizeof(m));
  m.msg_iov = &io;
  m.msg_iovlen = 1;
  m.msg_control = ctrl;
  m.msg_controllen = sizeof(ctrl);
  if (recvmsg(s, &m, MSG_CMSG_CLOEXEC) != static_cast<ssize_t>(sizeof(payload))) return false;
  cmsghdr* c = CMSG_FIRSTHDR(&m);
  if (!c || c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS) return false;
  memcpy(fd, CMSG_DATA(c), sizeof(int));
  *from = payload[0];
  *kind = payload[1];
  return true;
}

bool map_view(const DriverApi* api, unsigned long long handle, size_t size, size_t gran, int device, void** out) {
  CUdeviceptr va = 0;
  if (api->MemAddressReserve(&va, size, gran, 0, 0) != CUDA_SUCCESS) return false;
  if (api->MemMap(va, size, 0, static_cast<CUmemGenericAllocationHandle>(handle), 0) != CUDA_SUCCESS) {
    api->MemAddressFree(va, size);
    return false;
  }
  CUmemAccessDesc d;
  memset(&d, 0, sizeof(d));
  d.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  d.location.id = device;
  d.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (api->MemSetAccess(va, size, &d, 1) != CUDA_SUCCESS) {
    api->MemUnmap(va, size);
    api->MemAddressFree(va, size);
    return false;
  }
  *out = reinterpret_cast<void*>(va);
  return true;
}
void unmap_view(const DriverApi* api, void** p, size_t size) {
  if (!*p) return;
  const CUdeviceptr va = reinterpret_cast<CUdeviceptr>(*p);
  api->MemUnmap(va, size);
  api->MemAddressFree(va, size);
  *p = nullptr;
}

}  // namespace

void vmm_region_unmap(VmmRegion* r) {
  const DriverApi* api = driver_api();
  if (!api) return;
  unmap_view(api, &r->mc, r->size);
  if (r->bound) {
    CUdevice dev;
    if (api->DeviceGet(&dev, r->device) == CUDA_SUCCESS)
      api->MulticastUnbind(static_cast<CUmemGenericAllocationHandle>(r->h_mc), dev, 0, r->size);
    r->bound = false;
  }
  for (int k = 0; k < r->world && k < kVmmMaxRanks; ++k)
    if (k != r->rank) unmap_view(api, &r->peer[k], r->size);
  unmap_view(api, &r->own, r->size);
  r->peer[r->rank] = nullptr;
  r->active = false;
}

void vmm_region_release(VmmRegion* r) {
  const DriverApi* api = driver_api();
  if (!api) return;
  for (int k = 0; k < r->world && k < kVmmMaxRanks; ++k) {
    // handles of the same process are shared by value and released by their owner
    if (k != r->rank && r->cross_process && r->h_peer[k]) api->MemRelease(static_cast<CUmemGenericAllocationHandle>(r->h_peer[k]));
    r->h_peer[k] = 0;
  }
  if (r->h_mc && (r->cross_process || r->rank == 0)) api->MemRelease(static_cast<CUmemGenericAllocationHandle>(r->h_mc));
  if (r->h_own) api->MemRelease(static_cast<CUmemGenericAllocationHandle>(r->h_own));
  r->h_mc = r->h_own = 0;
}

void vmm_region_destroy(VmmRegion* r) {
  vmm_region_unmap(r);
  vmm_region_release(r);
}

bool vmm_region_create(const VmmAllgather& allgather, int rank, int world, int device, const int* devices,
                       size_t bytes, VmmRegion* out, const char** why) {
  *out = VmmRegion();
  out->world = world;
  out->rank = rank;
  out->device = device;
  out->cross_process = devices == nullptr;
  const bool cross = out->cross_process;
  *why = "";
  std::vector<uint64_t> all(static_cast<size_t>(world > 0 ? world : 1) * kVmmWords, 0);
  uint64_t mine[kVmmWords] = {0, 0, 0, 0};
  int sock = -1, fd_own = -1, fd_mc = -1;
  // one agreement round: every rank learns whether every rank is still fine
  auto agree = [&](bool ok) {
    mine[0] = ok ? 1 : 0;
    if (!allgather(mine, all.data())) return false;
    for (int r = 0; r < world; ++r)
      if (!all[static_cast<size_t>(r) * kVmmWords]) {
        if (ok) *why = "refused by another rank";
        return false;
      }
    return true;
  };
  auto fail = [&]() {
    if (sock >= 0) close(sock);
    if (fd_own >= 0) close(fd_own);
    if (fd_mc >= 0) close(fd_mc);
    vmm_region_unmap(out);
    if (!cross) allgather(mine, all.data());  // threads of one process share handles by value: all unmapped before any release
    vmm_region_release(out);
    *out = VmmRegion();
    return false;
  };

  // ---- round 0: can every rank do it? ----
  const DriverApi* api = nullptr;
  // Default: from 4 ranks up. A multicast store also comes back to its sender through the switch, so every
  // rank takes in world/(world-1) times the bytes of unicast pushes: twice as much on 2 GPUs (measured: xbar
  // phase 19.3 us against 13.5 us on the 1e6 x 1e6 workload), even on 4, ahead on 8.
  const char* force = getenv("FOLP_MULTICAST");
  const bool wanted = force ? atoi(force) != 0 : world >= 4;
  bool ok = world >= 2 && world <= kVmmMaxRanks && wanted && getenv("FOLP_NO_MULTICAST") == nullptr && getenv("FOLP_NO_P2P") == nullptr;
  if (!ok) *why = "switched off (FOLP_MULTICAST / FOLP_NO_MULTICAST / FOLP_NO_P2P, or fewer than 4 ranks)";
  if (ok && !(api = driver_api())) { ok = false; *why = "driver entry points not available"; }
  CUdevice cudev = 0;
  if (ok) {
    int sup = 0;
    if (api->DeviceGet(&cudev, device) != CUDA_SUCCESS ||
        api->DeviceGetAttribute(&sup, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cudev) != CUDA_SUCCESS || !sup) {
      ok = false;
      *why = "device without multicast support";
    }
  }
  if (ok && devices)
    for (int a = 0; a < world; ++a)
      for (int b = a + 1; b < world; ++b)
        if (devices[a] == devices[b]) { ok = false; *why = "two ranks on one device"; }
  static std::atomic<uint64_t> counter{0};
  mine[1] = (static_cast<uint64_t>(getpid()) << 32) ^
            static_cast<uint64_t>(std::chrono::steady_clock::now().time_since_epoch().count()) ^ (counter.fetch_add(1) << 20);
  if (!agree(ok)) return fail();
  const uint64_t nonce = all[1];  // rank 0's

  // ---- round 1: this rank's memory, the multicast object (rank 0), the socket ----
  const CUmemAllocationHandleType ht = cross ? CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR : CU_MEM_HANDLE_TYPE_NONE;
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof(mp));
  mp.numDevices = static_cast<unsigned>(world);
  mp.size = bytes;
  mp.handleTypes = ht;
  CUmemAllocationProp ap;
  memset(&ap, 0, sizeof(ap));
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = device;
  ap.requestedHandleTypes = ht;
  size_t g_min = 0, g_rec = 0, g_alloc = 0, gran = 0, size = 0;
  if (api->MulticastGetGranularity(&g_min, &mp, CU_MULTICAST_GRANULARITY_MINIMUM) != CUDA_SUCCESS ||
      api->MulticastGetGranularity(&g_rec, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS ||
      api->MemGetAllocationGranularity(&g_alloc, &ap, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || !g_min || !g_alloc) {
    ok = false;
    *why = "granularity query failed";
  } else {
    gran = g_min > g_alloc ? g_min : g_alloc;
    size = (bytes + gran - 1) / gran * gran;
    if (size >= (static_cast<size_t>(96) << 20) && g_rec > gran) {  // large regions: the recommended (huge-page) granularity
      gran = g_rec;
      size = (bytes + gran - 1) / gran * gran;
    }
    out->size = size;
  }
  if (ok) {
    CUmemGenericAllocationHandle hnd = 0;
    if (api->MemCreate(&hnd, size, &ap, 0) != CUDA_SUCCESS) { ok = false; *why = "cuMemCreate failed"; }
    else out->h_own = hnd;
  }
  if (ok && !map_view(api, out->h_own, size, gran, device, &out->own)) { ok = false; *why = "mapping the region failed"; }
  if (ok) {
    out->peer[rank] = out->own;
    if (cudaMemset(out->own, 0, size) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
      cudaGetLastError();
      ok = false;
      *why = "zero-filling the region failed";
    }
  }
  if (ok && cross) {
    if (api->MemExportToShareableHandle(&fd_own, static_cast<CUmemGenericAllocationHandle>(out->h_own), ht, 0) != CUDA_SUCCESS) {
      fd_own = -1;
      ok = false;
      *why = "exporting the allocation failed";
    } else if ((sock = open_socket(nonce, rank)) < 0) {
      ok = false;
      *why = "unix socket for the handle exchange failed";
    }
  }
  if (ok && rank == 0) {
    mp.size = size;
    CUmemGenericAllocationHandle hnd = 0;
    if (api->MulticastCreate(&hnd, &mp) != CUDA_SUCCESS) { ok = false; *why = "cuMulticastCreate failed"; }
    else {
      out->h_mc = hnd;
      if (cross && api->MemExportToShareableHandle(&fd_mc, hnd, ht, 0) != CUDA_SUCCESS) {
        fd_mc = -1;
        ok = false;
        *why = "exporting the multicast object failed";
      }
    }
  }
  mine[1] = size;
  mine[2] = cross ? 0 : out->h_own;
  mine[3] = (cross || rank != 0) ? 0 : out->h_mc;
  if (!agree(ok)) return fail();
  for (int r = 0; r < world; ++r)
    if (all[static_cast<size_t>(r) * kVmmWords + 1] != size) { ok = false; *why = "ranks disagree on the region size"; }

  // ---- round 2: every peer's memory mapped here, this device added to the multicast object ----
  if (ok && !cross) {
    for (int r = 0; r < world; ++r)
      if (r != rank) out->h_peer[r] = all[static_cast<size_t>(r) * kVmmWords + 2];
    out->h_mc = all[3];
  }
  if (ok && cross) {
    for (int r = 0; r < world && ok; ++r) {
      if (r == rank) continue;
      if (!send_fd(sock, nonce, r, rank, 0, fd_own)) { ok = false; *why = "sending the allocation handle failed"; }
      if (ok && rank == 0 && !send_fd(sock, nonce, r, rank, 1, fd_mc)) { ok = false; *why = "sending the multicast handle failed"; }
    }
    const int expect = (world - 1) + (rank != 0 ? 1 : 0);
    for (int k = 0; k < expect && ok; ++k) {
      int from = -1, kind = -1, fd = -1;
      if (!recv_fd(sock, &from, &kind, &fd, 30000)) { ok = false; *why = "no handle from a peer rank"; break; }
      CUmemGenericAllocationHandle hnd = 0;
      const bool good = from >= 0 && from < world && from != rank && (kind == 0 || (kind == 1 && from == 0 && rank != 0)) &&
                        api->MemImportFromShareableHandle(&hnd, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), ht) == CUDA_SUCCESS;
      close(fd);
      if (!good) { ok = false; *why = "importing a peer's handle failed"; break; }
      if (kind == 0) out->h_peer[from] = hnd;
      else out->h_mc = hnd;
    }
  }
  if (sock >= 0) { close(sock); sock = -1; }
  if (fd_own >= 0) { close(fd_own); fd_own = -1; }
  if (fd_mc >= 0) { close(fd_mc); fd_mc = -1; }
  for (int r = 0; r < world && ok; ++r) {
    if (r == rank) continue;
    if (!out->h_peer[r] || !map_view(api, out->h_peer[r], size, gran, device, &out->peer[r])) { ok = false; *why = "mapping a peer's region failed"; }
  }
  if (ok && (!out->h_mc || api->MulticastAddDevice(static_cast<CUmemGenericAllocationHandle>(out->h_mc), cudev) != CUDA_SUCCESS)) {
    ok = false;
    *why = "cuMulticastAddDevice failed";
  }
  if (!agree(ok)) return fail();

  // ---- round 3: all devices are in: bind this rank's memory, map the multicast range ----
  if (api->MulticastBindMem(static_cast<CUmemGenericAllocationHandle>(out->h_mc), 0,
                            static_cast<CUmemGenericAllocationHandle>(out->h_own), 0, size, 0) != CUDA_SUCCESS) {
    ok = false;
    *why = "cuMulticastBindMem failed";
  } else {
    out->bound = true;
  }
  if (ok && !map_view(api, out->h_mc, size, gran, device, &out->mc)) { ok = false; *why = "mapping the multicast range failed"; }
  if (!agree(ok)) return fail();
  out->active = true;
  return true;
}

}  // namespace folp
