// folp_api.cu -- host side of libfolp_b200.so: the C ABI of include/folp_b200.h.
//
// Replaces the while-loop of optimize(::PdhgParameters, qp)
// (src/primal_dual_hybrid_gradient.jl:862-1048). The host keeps only scalar
// control flow (evaluation cadence :892-895, termination term.jl:233-273, restart
// decisions sp.jl:688-846, primal weight sp.jl:862-891); every vector lives in
// HBM and every O(n), O(m), O(nnz) operation is a kernel in folp_kernels.cu.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <tuple>
#include <atomic>
#include <new>
#include <thread>

#include "folp_kernels.cuh"
#include "folp_nccl.h"
#include "folp_vmm.h"

using namespace folp;

namespace {
thread_local std::string g_create_error;

// NCCL communicators outlive handles: ncclCommInitRank costs 1.3 s (2 ranks) to 4.8 s (8 ranks), which
// every folp_create of a process used to pay again. One communicator per (world, rank, device) is kept
// for the lifetime of the process and reused by later handles (all ranks create their handles in the
// same order, so either every rank finds its entry or none does). FOLP_NO_COMM_CACHE=1 disables it.
std::mutex g_comm_mutex;
std::map<std::tuple<int, int, int>, ncclComm_t> g_comm_cache;

// Multicast exchange regions outlive handles for the same reason: creating the multicast object, adding the
// devices and binding memory programs the NVSwitch and costs 40-110 ms per folp_create (2 GPUs, measured), more
// than the rest of a small handle's set-up. One-process-per-GPU mode only; a later handle of the same
// (world, rank, device) whose region fits reuses the mapping after zero-filling it (the ranks agree on the
// entry by its creation number: all ranks create their handles in the same order). FOLP_NO_REGION_CACHE=1
// disables it. Entries live until the process exits.
struct VmmCacheEntry {
  VmmRegion r;
  uint64_t id = 0;
  bool busy = false;
};
std::mutex g_vmm_mutex;
std::vector<VmmCacheEntry> g_vmm_cache;
uint64_t g_vmm_next_id = 1;

double now_sec() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

constexpr unsigned kMaxHostThreadsU = 16;
// Splits [begin, end) over up to kMaxHostThreadsU host threads (folp_create's O(nnz) loops: index conversion,
// transposition, position-major packing). fn(lo, hi, thread_index).
template <class F>
void parallel_for(int64_t begin, int64_t end, int64_t min_chunk, F fn) {
  const int64_t len = end - begin;
  unsigned hw = std::thread::hardware_concurrency();
  int T = static_cast<int>(std::min<int64_t>(std::min<unsigned>(hw ? hw : 1u, kMaxHostThreadsU), std::max<int64_t>(1, len / min_chunk)));
  if (T <= 1) {
    if (len > 0) fn(begin, end, 0);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(static_cast<size_t>(T));
  for (int t = 0; t < T; ++t) {
    const int64_t lo = begin + len * t / T, hi = begin + len * (t + 1) / T;
    th.emplace_back([=, &fn] { fn(lo, hi, t); });
  }
  for (auto& x : th) x.join();
}

// std::vector without the zero fill of resize() / the sizing constructor: folp_create's
// O(nnz) scratch arrays are written exactly once, by several threads.
// Large blocks are 2 MB aligned and can be advised as transparent huge pages (FOLP_THP=1): the
// arrays are touched once, front to back, and 4 KB first-touch faults are a fifth of the work.
template <class T>
struct NoInit {
  using value_type = T;
  template <class U> struct rebind { using other = NoInit<U>; };
  NoInit() = default;
  template <class U> NoInit(const NoInit<U>&) {}
  T* allocate(size_t count) {
    const size_t bytes = count * sizeof(T);
    void* q = nullptr;
    constexpr size_t kHuge = size_t{2} << 20;
    if (bytes >= 2 * kHuge) {
      const size_t rounded = (bytes + kHuge - 1) / kHuge * kHuge;
      q = aligned_alloc(kHuge, rounded);
      // opt-in: 38 vs 46 ms of host preparation on the GPU box when huge pages are at hand, but a
      // first-touch compaction stall of several hundred ms when they are not (seen on a long-running host)
      static const bool thp = getenv("FOLP_THP") != nullptr;
      if (q && thp) madvise(q, rounded, MADV_HUGEPAGE);
    } else {
      q = malloc(bytes ? bytes : 1);
    }
    if (!q) throw std::bad_alloc();
    return static_cast<T*>(q);
  }
  void deallocate(T* q, size_t) noexcept { free(q); }
  template <class U> bool operator==(const NoInit<U>&) const { return true; }
  template <class U> bool operator!=(const NoInit<U>&) const { return false; }
  template <class U> void construct(U* p) noexcept { ::new (static_cast<void*>(p)) U; }
  template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
using IVec = std::vector<int, NoInit<int>>;
using DVec = std::vector<double, NoInit<double>>;

// ---------------------------------------------------------------------------
// Process-wide PINNED host scratch for folp_create on one GPU. A create of the 1e6 x 1e6 x 1e7 problem writes
// ~620 MB of host scratch (transposition buckets, the unpacked CSR of A, both packed matrices, the renumbered
// vectors): as fresh heap pages that is ~150 K page faults per call, and the uploads leave pageable memory
// through the driver's staging at 6-8 GB/s. The scratch is therefore kept between creates, pinned
// (cudaHostAlloc, portable): no faults, and every upload is an asynchronous DMA that overlaps the packing.
// The FIRST create of a process that wants it runs on ordinary vectors and, when it is over, starts a
// background thread that pins a pool of the wanted size (pinning 600 MB takes ~0.2 s; started during the
// create it made that create 0.2 s slower); later creates lease it if it is free and large enough. FOLP_NO_HOST_POOL=1 switches it off, FOLP_HOST_POOL_MB caps it (default 1536).
// ---------------------------------------------------------------------------
struct HostPool {
  std::mutex mu;
  std::condition_variable cv;
  char* base = nullptr;
  size_t cap = 0;
  bool busy = false, growing = false;
  std::thread grower;
  ~HostPool() {  // process exit: a running cudaHostAlloc finishes (or fails with "unloading") before the statics go
    if (grower.joinable()) grower.join();
  }
};
HostPool g_host_pool;

// Bump allocator over a lease of the pool (64-byte aligned blocks; nullptr when there is no lease or no room).
struct HostCarver {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  template <class T>
  T* take(size_t count) {
    const size_t bytes = (count * sizeof(T) + 63) / 64 * 64;
    if (!base || used + bytes > cap) return nullptr;
    T* q = reinterpret_cast<T*>(base + used);
    used += bytes;
    return q;
  }
};

// Leases the pool for one create (RAII). want = bytes this create would carve.
struct HostPoolLease {
  HostCarver carver;
  bool held = false;
  void acquire(size_t want, int device) {
    if (getenv("FOLP_NO_HOST_POOL") != nullptr || want < (static_cast<size_t>(32) << 20)) return;
    size_t limit = static_cast<size_t>(1536) << 20;
    if (const char* e = getenv("FOLP_HOST_POOL_MB")) limit = static_cast<size_t>(atoll(e)) << 20;
    HostPool& P = g_host_pool;
    std::lock_guard<std::mutex> lock(P.mu);
    if (!P.busy && P.base && P.cap >= want) {
      P.busy = true;
      held = true;
      carver.base = P.base;
      carver.cap = P.cap;
      return;
    }
    if (P.growing || P.busy || want > limit || P.cap >= want) return;
    P.growing = true;
    grow_bytes = std::min(limit, want + want / 8);  // pinned once this create is over (destructor)
    grow_device = device;
  }
  size_t grow_bytes = 0;
  int grow_device = 0;
  void start_growth() {
    HostPool& P = g_host_pool;
    std::lock_guard<std::mutex> lock(P.mu);
    if (P.grower.joinable()) P.grower.join();  // a finished earlier growth
    const size_t bytes = grow_bytes;
    const int device = grow_device;
    P.grower = std::thread([bytes, device] {
      HostPool& Q = g_host_pool;
      void* q = nullptr;
      const bool ok = cudaSetDevice(device) == cudaSuccess && cudaHostAlloc(&q, bytes, cudaHostAllocPortable) == cudaSuccess;
      void* old = nullptr;
      {
        std::unique_lock<std::mutex> lk(Q.mu);
        Q.cv.wait(lk, [&] { return !Q.busy; });
        if (ok) {
          old = Q.base;
          Q.base = static_cast<char*>(q);
          Q.cap = bytes;
        }
        Q.growing = false;
      }
      if (old) cudaFreeHost(old);
    });
  }
  ~HostPoolLease() {
    if (grow_bytes) start_growth();  // in the background, while the caller goes on to solve
    if (!held) return;
    HostPool& P = g_host_pool;
    {
      std::lock_guard<std::mutex> lock(P.mu);
      P.busy = false;
    }
    P.cv.notify_all();
  }
};
}  // namespace

// ---------------------------------------------------------------------------
// Single-process multi-GPU (folp_create_multi): the wrapper handle owns one sub-handle per device,
// each driven by its own host thread running exactly the per-rank code of the one-process-per-GPU
// mode; the threads meet only at create time (to swap the addresses of their exchange regions --
// direct peer access inside one process, no CUDA IPC, no NCCL) and otherwise synchronise through the
// same device-side flags as separate processes do.
// ---------------------------------------------------------------------------
struct MultiCtx {
  int world = 0;
  std::vector<folp_handle*> sub;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  std::function<int(int)> job;
  uint64_t job_id = 0;
  int pending = 0;
  std::vector<int> rcs;
  bool stop = false;
  // create-time rendezvous of the ranks
  std::vector<void*> regions;
  std::vector<int> peer_ok;
  std::vector<uint64_t> words;  // world * kVmmWords: the ranks' contributions to an allgather of the multicast setup
  int bar_count = 0;
  uint64_t bar_gen = 0;
  std::condition_variable cv_bar;
  void barrier() {
    std::unique_lock<std::mutex> lock(mu);
    const uint64_t g = bar_gen;
    if (++bar_count == world) {
      bar_count = 0;
      bar_gen += 1;
      cv_bar.notify_all();
    } else {
      cv_bar.wait(lock, [&] { return bar_gen != g; });
    }
  }
  // runs f(rank) on every rank's thread; returns the first non-zero status
  int call(std::function<int(int)> f) {
    std::unique_lock<std::mutex> lock(mu);
    job = std::move(f);
    job_id += 1;
    pending = world;
    cv_job.notify_all();
    cv_done.wait(lock, [&] { return pending == 0; });
    for (int r = 0; r < world; ++r)
      if (rcs[r]) return rcs[r];
    return 0;
  }
  void worker(int rank) {
    uint64_t seen = 0;
    for (;;) {
      std::function<int(int)> f;
      {
        std::unique_lock<std::mutex> lock(mu);
        cv_job.wait(lock, [&] { return stop || job_id != seen; });
        if (stop) return;
        seen = job_id;
        f = job;
      }
      const int rc = f(rank);
      {
        std::unique_lock<std::mutex> lock(mu);
        rcs[rank] = rc;
        if (--pending == 0) cv_done.notify_all();
      }
    }
  }
};

struct folp_handle {
  std::string err;
  MultiCtx* multi = nullptr;    // wrapper handle of folp_create_multi: everything below lives in multi->sub[r]
  MultiCtx* shared = nullptr;   // sub-handle of a single-process multi-GPU solve: its ranks' rendezvous
  bool shared_rendezvous_done = false;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t n = 0, m = 0, nnz = 0, neq = 0;
  folp_params prm{};
  double cache[4] = {0, 0, 0, 0};
  double objective_constant = 0.0;
  SpmvMat A, At;
  SpmvMat Q;  // CSR of the scaled objective matrix (QP only; B.has_q)
  Bufs B;
  std::vector<void*> allocs;
  char* arena = nullptr;  // see arena_reserve
  size_t arena_cap = 0, arena_used = 0;
  // pinned host mirrors
  DevState* hs = nullptr;
  double* h_red = nullptr;   // 4 * kMaxScalars
  TrState* h_trs = nullptr;
  TrState* d_trs = nullptr;
  // RestartInfo (sp.jl:158-197), scalars
  int has_last_gap = 0;
  double last_gap = 0.0;
  int64_t last_restart_length = 1;
  double pd_last = 0.0, dd_last = 0.0, gap_reduction_ratio_last_trial = 1.0;
  // loop bookkeeping (the reference's `iteration`, pdhg.jl:885-887)
  int64_t iteration = 0;
  int need_step = 0, terminated = 0;
  double start_time = 0.0, basic_time = 0.0;
  folp_eval last_eval{};
  int64_t launches = 0;
  int64_t tr_passes = 0, tr_solves = 0;
  int tr_grid = 0;  // > 0: trust-region solves run as one cooperative kernel on this many blocks
  // results of the evaluation block's device work when it was enqueued blind (evaluate_enqueue_blind):
  // bit k of pre_mask = trust-region slot k is in pre[k]; pre_dist / pre_ax_cur = the distance sums are in
  // h_red / A * x_cur is in B.ax_cur already
  TrState pre[kTrSlots];
  unsigned pre_mask = 0;
  bool pre_dist = false, pre_ax_cur = false;
  int trm_grid = 0;   // > 0: all trust-region solves of an evaluation run as one cooperative kernel (k_tr_multi)
  std::vector<double> h_tmp;  // staging of primal-indexed vectors at the boundary (renumbering)
  std::vector<int> new2old;   // device numbering of the variables -> the caller's (empty: identity); order_variables
  bool take_cluster = false;  // the take_grid CTAs form one thread-block cluster (tiny instances, single GPU)
  int take_grid = 0;  // > 0: batches of take_step attempts run as one cooperative kernel (k_take_steps) on this many blocks
  unsigned long long* d_timers = nullptr;  // phase timers of k_take_steps (folp_debug_profile_attempts)
  std::map<int, cudaGraphExec_t> step_graphs;
  bool use_graphs = true;
  // ---- row-partitioned mode (folp_dist.world_size > 1) ----
  int rank = 0, world = 1;
  const NcclApi* nccl = nullptr;
  ncclComm_t comm = nullptr;
  bool comm_cached = false;         // the communicator belongs to the process-wide cache
  int64_t n_glob = 0, m_glob = 0;  // global sizes; n / m above are the local slice / rows
  int64_t n_pad = 0, m_pad = 0;    // exchange strides: ceil(n/world) (even), max rows per rank
  int64_t col0 = 0, row0 = 0;      // first global column / row owned by this rank
  std::vector<int64_t> row_begin;  // world + 1
  double* d_rows = nullptr;        // staging for row-indexed gathers, world * m_pad
  double* d_cols = nullptr;        // staging for column-indexed gathers, world * n_pad
  double* h_sc = nullptr;          // pinned, world * kScBlock
  // peer-memory exchange region {xbar | y_full | sc_recv | scx | flags} and the peers' mappings
  void* region = nullptr;
  void* peer_region[kMaxWorld] = {};
  VmmRegion vmm;                    // active: the region is cuMemCreate memory bound to an NVSwitch multicast object (folp_vmm.h)
  uint64_t vmm_cache_id = 0;        // != 0: the region belongs to the process-wide cache (g_vmm_cache)
  unsigned long long xchg_seq = 0;  // evaluation-block scalar exchanges done so far (same on every rank)
};

#define TRY(expr) FOLP_CUDA_TRY(h, expr)
#define CHECK_LAUNCH() TRY(cudaGetLastError())
#define NCCL_TRY(expr)                                                             \
  do {                                                                             \
    ncclResult_t _r = (expr);                                                      \
    if (_r != ncclSuccess) {                                                       \
      h->err = std::string(#expr) + ": " + h->nccl->GetErrorString(_r);            \
      return FOLP_NCCL_ERROR;                                                      \
    }                                                                              \
  } while (0)

// ---------------------------------------------------------------------------
// setup helpers
// ---------------------------------------------------------------------------
// One cudaMalloc for (nearly) everything a handle owns: ~60 separate allocations cost tens of
// milliseconds of folp_create and as many cudaFree calls in folp_destroy. dev_alloc carves
// 256-byte aligned blocks out of the arena and falls back to cudaMalloc when it is absent or full.
// Single-process multi-GPU: once cudaDeviceEnablePeerAccess is on, every cudaMalloc / cudaFree also
// maps / unmaps the block on all peers (measured: folp_create_multi 1.65 s for ~60 allocations on 2
// GPUs against 0.24 s with one process per GPU). Only the exchange region has to be peer-visible:
// everything else comes from the device's stream-ordered memory pool, which peers do not map.
static cudaError_t private_malloc(folp_handle* h, void** q, size_t bytes) {
  if (!h->shared) return cudaMalloc(q, bytes);
  cudaError_t e = cudaMallocAsync(q, bytes, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  return e;
}
static int arena_reserve(folp_handle* h, size_t bytes) {
  if (h->arena || bytes == 0) return FOLP_OK;
  void* q = nullptr;
  TRY(private_malloc(h, &q, bytes));
  h->allocs.push_back(q);
  h->arena = static_cast<char*>(q);
  h->arena_cap = bytes;
  h->arena_used = 0;
  return FOLP_OK;
}
template <class T>
static int dev_alloc(folp_handle* h, T** p, size_t count) {
  void* q = nullptr;
  // 16 elements of slack: vectorised kernels round their extents up and may read (never use)
  // a few elements past the end
  const size_t bytes = ((count + 16) * sizeof(T) + 255) / 256 * 256;
  if (h->arena && h->arena_used + bytes <= h->arena_cap) {
    q = h->arena + h->arena_used;
    h->arena_used += bytes;
  } else {
    TRY(private_malloc(h, &q, bytes));
    h->allocs.push_back(q);
  }
  TRY(cudaMemsetAsync(static_cast<char*>(q) + count * sizeof(T), 0, 16 * sizeof(T), h->stream));
  *p = static_cast<T*>(q);
  return FOLP_OK;
}
static int dev_zeros(folp_handle* h, double** p, size_t count) {
  int rc = dev_alloc(h, p, count);
  if (rc) return rc;
  TRY(cudaMemsetAsync(*p, 0, std::max<size_t>(count, 1) * sizeof(double), h->stream));
  return FOLP_OK;
}
static int dev_upload(folp_handle* h, double** p, const double* src, size_t count, double fill) {
  int rc = dev_alloc(h, p, count);
  if (rc) return rc;
  if (src) {
    TRY(cudaMemcpyAsync(*p, src, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  } else {
    launch_fill(*p, fill, static_cast<int64_t>(count), h->stream);
  }
  return FOLP_OK;
}

// Host half of the matrix set-up: cuts the rows into warp-sized work items (see
// folp_internal.cuh) and rewrites colidx / vals so that narrow groups are position-major.
// Pure host code (no CUDA call), also reachable through folp_debug_host_spmv.
struct PackedMatrix {
  std::vector<Tile> tiles;
  std::vector<int> rowid;  // slot -> row; identity outside sorted windows
  bool any_sorted = false;
  int nlong = 0, nchunks_total = 0;
};
// warps_total = warps of the grid k_spmv runs on: work item i goes to warp i % warps_total in
// that warp's trip i / warps_total (static striding, see k_spmv).
static void plan_tiles(int rows, const IVec& rowptr, int warps_total, PackedMatrix* out) {
  std::vector<Tile>& tiles = out->tiles;
  std::vector<int>& rowid = out->rowid;
  bool& any_sorted = out->any_sorted;
  int& nlong = out->nlong;
  int& nchunks_total = out->nchunks_total;
  tiles.reserve(static_cast<size_t>(rows) / 32 + 16);
  rowid.resize(static_cast<size_t>(rows));
  for (int q = 0; q < rows; ++q) rowid[q] = q;
  const bool sort_rows = getenv("FOLP_NO_ROW_SORT") == nullptr;
  constexpr int kGatherUnrollHost = FOLP_GATHER_UNROLL;  // positions per round of k_spmv
  int r = 0;
  while (r < rows) {
    const int len = rowptr[r + 1] - rowptr[r];
    if (len > kChunkNnz) {  // long row: chunks
      const int nch = (len + kChunkNnz - 1) / kChunkNnz;
      for (int c = 0; c < nch; ++c) {
        Tile t{};
        t.row_begin = r;
        t.nnz_begin = rowptr[r] + c * kChunkNnz;
        t.nnz_end = std::min(rowptr[r + 1], t.nnz_begin + kChunkNnz);
        t.rows_kind = (kTileLongChunk << 16) | 1;
        t.long_id = nlong; t.chunk_first = nchunks_total; t.chunk_count = nch; t.chunk_index = c;
        tiles.push_back(t);
      }
      nlong += 1;
      nchunks_total += nch;
      r += 1;
      continue;
    }
    if (len > kNarrowMax) {  // wide row: one warp
      Tile t{};
      t.row_begin = r;
      t.nnz_begin = rowptr[r];
      t.nnz_end = rowptr[r + 1];
      t.rows_kind = (kTileWarpPerRow << 16) | 1;
      tiles.push_back(t);
      r += 1;
      continue;
    }
    // a run of consecutive narrow rows, cut into windows of kSortWindow rows
    const int run0 = r;
    while (r < rows && rowptr[r + 1] - rowptr[r] <= kNarrowMax) r += 1;
    for (int w0 = run0; w0 < r; w0 += kSortWindow) {
      const int w1 = std::min(r, w0 + kSortWindow);
      // rounds of kGatherUnroll positions the warps spend on this window, identity vs sorted order
      auto rounds = [&](const int* ids) {
        int total = 0;
        for (int g = w0; g < w1; g += 32) {
          int mx = 0;
          for (int q = g; q < std::min(w1, g + 32); ++q) {
            const int row = ids ? ids[q - w0] : q;
            mx = std::max(mx, rowptr[row + 1] - rowptr[row]);
          }
          total += (mx + kGatherUnrollHost - 1) / kGatherUnrollHost;
        }
        return total;
      };
      // stable counting sort of the window's rows by decreasing length (lengths <= kNarrowMax)
      int ids[kSortWindow], start[kNarrowMax + 2] = {0};
      for (int q = w0; q < w1; ++q) start[kNarrowMax - (rowptr[q + 1] - rowptr[q]) + 1] += 1;
      for (int b = 0; b <= kNarrowMax; ++b) start[b + 1] += start[b];
      for (int q = w0; q < w1; ++q) ids[start[kNarrowMax - (rowptr[q + 1] - rowptr[q])]++] = q;
      // Sorting scatters the epilogue's per-row accesses over the window, which costs about as
      // much as a third of the rounds: on Poisson(10) columns (1e6 x 1e6 x 1e7 workload, 32 % fewer
      // rounds) K3 measured 65.4 us sorted against 63.6 us in place. Only windows whose rounds at
      // least halve are sorted (a few long rows among many short ones).
      const bool sorted = sort_rows && 2 * rounds(ids) <= rounds(nullptr);
      if (sorted) {
        any_sorted = true;
        // The groups of a sorted window differ in length by design, and work item i always
        // lands on warp i % warps_total: without care the same warp of every CTA would get the
        // longest group of every window it visits. Rotate the order of the window's FULL groups
        // by the trip number, so that a warp meets every rank in turn.
        const int full = (w1 - w0) / 32;  // groups of exactly 32 rows; a shorter tail group stays last
        const int rot = full > 1 ? static_cast<int>((tiles.size() / static_cast<size_t>(warps_total)) % full) : 0;
        for (int g = 0; g < full; ++g) {
          const int src = ((g + rot) % full) * 32;
          for (int q = 0; q < 32; ++q) rowid[w0 + g * 32 + q] = ids[src + q];
        }
        for (int q = w0 + full * 32; q < w1; ++q) rowid[q] = ids[q - w0];
      }
      int k = rowptr[w0];
      for (int g = w0; g < w1; g += 32) {
        const int cnt = std::min(w1, g + 32) - g;
        Tile t{};
        t.row_begin = g;  // first row, or first slot of rowid
        t.nnz_begin = k;
        for (int q = g; q < g + cnt; ++q) k += rowptr[rowid[q] + 1] - rowptr[rowid[q]];
        t.nnz_end = k;
        t.rows_kind = ((sorted ? kTileThreadPerRowSorted : kTileThreadPerRow) << 16) | cnt;
        tiles.push_back(t);
      }
    }
  }
  // A cost-balanced STATIC order of the work items (FOLP_NO_BALANCE_TILES=1 keeps the matrix order).
  // k_spmv hands item i to warp i % warps_total. Items are dealt to the least loaded warp in order of
  // decreasing cost (rounds of a group, 128-entry trips of a warp-per-row item, + 1), then laid out
  // so that warp w's k-th item sits at index k * warps_total + w, short lists padded with empty
  // groups. Deterministic (a function of the matrix and the grid), no atomics; the per-thread
  // partial sums of the reductions are simply formed over other rows. Measured: nothing on rows in
  // matrix order (round 1: every group of a random matrix costs about the same), but with the
  // variables renumbered by column length (order_variables) groups differ by up to 5x and the
  // balance is what turns fewer rounds into time: A'*y 60.6 -> 57.8 us on the 1e6 x 1e6 x 1e7 workload,
  // 57.5 -> 48.3 us (and A*xbar 62.3 -> 56.2 us) on the PageRank LP.
  if (getenv("FOLP_NO_BALANCE_TILES") == nullptr && warps_total > 0 &&
      tiles.size() > static_cast<size_t>(warps_total)) {
    const size_t T = tiles.size();
    std::vector<int> cost(T);
    for (size_t i = 0; i < T; ++i) {
      const Tile& t = tiles[i];
      const int kind = t.rows_kind >> 16;
      if (kind == kTileThreadPerRow || kind == kTileThreadPerRowSorted) {
        int mx = 0;
        for (int q = t.row_begin; q < t.row_begin + (t.rows_kind & 0xffff); ++q)
          mx = std::max(mx, rowptr[rowid[q] + 1] - rowptr[rowid[q]]);
        cost[i] = (mx + kGatherUnrollHost - 1) / kGatherUnrollHost + 1;
      } else {
        cost[i] = (t.nnz_end - t.nnz_begin + 127) / 128 + 1;
      }
    }
    std::vector<int> order(T);
    for (size_t i = 0; i < T; ++i) order[i] = static_cast<int>(i);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    // least-loaded-first with a deterministic tie break (load, then warp index)
    std::vector<std::vector<int>> mine(static_cast<size_t>(warps_total));
    std::vector<std::pair<int64_t, int>> heap;  // (-load, -warp) as a max-heap == min load, min warp
    heap.reserve(static_cast<size_t>(warps_total));
    for (int w = 0; w < warps_total; ++w) heap.emplace_back(0, -w);
    std::make_heap(heap.begin(), heap.end());
    for (int idx : order) {
      std::pop_heap(heap.begin(), heap.end());
      auto& top = heap.back();
      mine[static_cast<size_t>(-top.second)].push_back(idx);
      top.first -= cost[idx];
      std::push_heap(heap.begin(), heap.end());
    }
    size_t depth = 0;
    for (const auto& v : mine) depth = std::max(depth, v.size());
    Tile empty{};
    empty.rows_kind = kTileThreadPerRow << 16;  // a group of zero rows: every lane idles through it
    std::vector<Tile> balanced(depth * static_cast<size_t>(warps_total), empty);
    for (int w = 0; w < warps_total; ++w) {
      // within a warp keep the matrix order: its streams then move forward through memory
      std::sort(mine[w].begin(), mine[w].end());
      for (size_t k = 0; k < mine[w].size(); ++k)
        balanced[k * static_cast<size_t>(warps_total) + w] = tiles[mine[w][k]];
    }
    tiles.swap(balanced);
  }
}

// Writes the packed arrays: position-major inside every narrow group (each row keeps its own
// ascending order), plain order for wide rows and long-row chunks. Out of place: entry k of the
// CSR source is read through col(k) / val(k) -- the caller's own Int64 arrays, or scratch.
template <class GetCol, class GetVal>
static void fill_packed(const PackedMatrix& pk, const IVec& rowptr, GetCol col, GetVal val,
                        int* dcol, double* dval, const int* src_start = nullptr) {
  // src_start (optional): entry `pos` of row `row` is read at col/val(src_start[row] + pos) instead of
  // (rowptr[row] + pos): the rows of the source are a permutation of the packed rows (order_variables)
  const std::vector<Tile>& tiles = pk.tiles;
  const std::vector<int>& rowid = pk.rowid;
  parallel_for(0, static_cast<int64_t>(tiles.size()), 1 << 11, [&](int64_t lo, int64_t hi, int) {
    for (int64_t ti = lo; ti < hi; ++ti) {
      const Tile& t = tiles[ti];
      const int kind = t.rows_kind >> 16;
      if (kind != kTileThreadPerRow && kind != kTileThreadPerRowSorted) {
        const int row = t.row_begin;
        const int shift = src_start ? src_start[row] - rowptr[row] : 0;
        for (int k = t.nnz_begin; k < t.nnz_end; ++k) {
          dcol[k] = col(k + shift);
          dval[k] = val(k + shift);
        }
        continue;
      }
      const int g = t.row_begin, g1 = g + (t.rows_kind & 0xffff);
      int out = t.nnz_begin;
      for (int pos = 0; out < t.nnz_end; ++pos)
        for (int q = g; q < g1; ++q) {
          const int row = rowid[q];
          if (rowptr[row + 1] - rowptr[row] > pos) {
            const int k = (src_start ? src_start[row] : rowptr[row]) + pos;
            dcol[out] = col(k);
            dval[out] = val(k);
            ++out;
          }
        }
    }
  });
}

// Uploads a packed matrix.
// sync = false: the host arrays outlive the copies (pinned scratch leased for the whole create, which ends
// with a synchronisation of the stream)
static int upload_matrix(folp_handle* h, SpmvMat* M, int rows, int cols, const IVec& rowptr,
                         const PackedMatrix& pk, const int* colidx, const double* vals, bool sync = true) {
  M->rows = rows;
  M->cols = cols;
  M->nnz = rowptr[rows];
  const std::vector<Tile>& tiles = pk.tiles;
  const std::vector<int>& rowid = pk.rowid;
  const bool any_sorted = pk.any_sorted;
  const int nlong = pk.nlong, nchunks_total = pk.nchunks_total;
  M->ntiles = static_cast<int>(tiles.size());
  M->nlong = nlong;
  const size_t pad = 16;
  int rc;
  if ((rc = dev_alloc(h, &M->rowptr, static_cast<size_t>(rows) + 1))) return rc;
  if ((rc = dev_alloc(h, &M->colidx, static_cast<size_t>(M->nnz) + pad))) return rc;
  if ((rc = dev_alloc(h, &M->vals, static_cast<size_t>(M->nnz) + pad))) return rc;
  if ((rc = dev_alloc(h, &M->tiles, tiles.size()))) return rc;
  std::vector<int2> slot_row_len;
  if (any_sorted) {  // (row, length) of every slot: one coalesced 8-byte load per lane of a sorted group
    slot_row_len.resize(static_cast<size_t>(rows));
    for (int q = 0; q < rows; ++q)
      slot_row_len[q] = make_int2(rowid[q], rowptr[rowid[q] + 1] - rowptr[rowid[q]]);
    if ((rc = dev_alloc(h, &M->rowid, static_cast<size_t>(rows)))) return rc;
    TRY(cudaMemcpyAsync(M->rowid, slot_row_len.data(), static_cast<size_t>(rows) * sizeof(int2),
                        cudaMemcpyHostToDevice, h->stream));
  }
  if ((rc = dev_alloc(h, &M->long_partials, static_cast<size_t>(nchunks_total)))) return rc;
  if ((rc = dev_alloc(h, &M->long_tickets, static_cast<size_t>(nlong)))) return rc;
  TRY(cudaMemsetAsync(M->colidx + M->nnz, 0, pad * sizeof(int), h->stream));
  TRY(cudaMemsetAsync(M->vals + M->nnz, 0, pad * sizeof(double), h->stream));
  TRY(cudaMemsetAsync(M->long_tickets, 0, std::max(nlong, 1) * sizeof(unsigned), h->stream));
  TRY(cudaMemcpyAsync(M->rowptr, rowptr.data(), (static_cast<size_t>(rows) + 1) * sizeof(int),
                      cudaMemcpyHostToDevice, h->stream));
  if (M->nnz) {
    TRY(cudaMemcpyAsync(M->colidx, colidx, static_cast<size_t>(M->nnz) * sizeof(int),
                        cudaMemcpyHostToDevice, h->stream));
    TRY(cudaMemcpyAsync(M->vals, vals, static_cast<size_t>(M->nnz) * sizeof(double),
                        cudaMemcpyHostToDevice, h->stream));
  }
  if (!tiles.empty())
    TRY(cudaMemcpyAsync(M->tiles, tiles.data(), tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice,
                        h->stream));
  // rowptr, tiles and slot_row_len are pageable: those copies have left the host when the calls return
  if (sync) TRY(cudaStreamSynchronize(h->stream));  // host vectors die with the caller's scope
  return FOLP_OK;
}

// plan + pack (from unpacked CSR scratch) + upload. carver (optional): the packed arrays are carved from
// the leased pinned scratch and uploaded asynchronously (they live until the create returns).
static int build_matrix(folp_handle* h, SpmvMat* M, int rows, int cols, const IVec& rowptr,
                        const int* colidx, const double* vals, HostCarver* carver = nullptr) {
  PackedMatrix pk;
  plan_tiles(rows, rowptr, h->sm_count * kSpmvCtasPerSm * kSpmvWarps, &pk);
  const size_t nz = static_cast<size_t>(rowptr[rows]);
  int* pc = carver ? carver->take<int>(nz) : nullptr;
  double* pv = carver ? carver->take<double>(nz) : nullptr;
  const bool pooled = pc && pv;
  IVec pc_own;
  DVec pv_own;
  if (!pooled) {
    pc_own.resize(nz);
    pv_own.resize(nz);
    pc = pc_own.data();
    pv = pv_own.data();
  }
  fill_packed(pk, rowptr, [=](int k) { return colidx[k]; }, [=](int k) { return vals[k]; }, pc, pv);
  return upload_matrix(h, M, rows, cols, rowptr, pk, pc, pv, !pooled);
}

// collective: called by every rank thread of a single-process multi-GPU solve at the same time
static void free_handle(folp_handle* h, bool collective = false) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (auto& kv : h->step_graphs) cudaGraphExecDestroy(kv.second);
  for (void* p : h->allocs) {
    if (h->shared && h->stream) cudaFreeAsync(p, h->stream);  // private_malloc
    else cudaFree(p);
  }
  if (h->shared && h->stream) cudaStreamSynchronize(h->stream);
  if (h->hs) cudaFreeHost(h->hs);  // one block: hs | h_red | h_trs | h_sc
  if (h->vmm_cache_id) {  // back to the cache, mapped as it is
    std::lock_guard<std::mutex> lock(g_vmm_mutex);
    for (VmmCacheEntry& e : g_vmm_cache)
      if (e.id == h->vmm_cache_id) e.busy = false;
  } else if (h->vmm.h_own) {
    if (h->vmm.active) vmm_region_unmap(&h->vmm);
    if (collective) h->shared->barrier();  // handles are shared by value between the ranks of a process
    vmm_region_release(&h->vmm);
  } else {
    for (int r = 0; r < kMaxWorld; ++r)
      if (h->peer_region[r]) cudaIpcCloseMemHandle(h->peer_region[r]);
    if (h->region) cudaFree(h->region);
  }
  if (h->comm && h->nccl && !h->comm_cached) h->nccl->CommDestroy(h->comm);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

static int push_state(folp_handle* h) {
  TRY(cudaMemcpyAsync(h->B.st, h->hs, sizeof(DevState), cudaMemcpyHostToDevice, h->stream));
  return FOLP_OK;
}
static int pull_state(folp_handle* h) {
  TRY(cudaMemcpyAsync(h->hs, h->B.st, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
  TRY(cudaStreamSynchronize(h->stream));
  return FOLP_OK;
}

// ---------------------------------------------------------------------------
// 1-D row partition: contiguous row blocks balanced by
// cost = nonzeros + 2 per row, and equal column slices of ceil(n / world) rounded
// up to an even count (K1 moves 16-byte pairs). Pure host arithmetic; exported so
// that CPU tests can check it without a device.
// ---------------------------------------------------------------------------
static void partition_rows(int64_t m, const std::vector<int64_t>& row_cost_prefix, int world,
                           int64_t* row_begin) {
  const int64_t total = row_cost_prefix[m];
  row_begin[0] = 0;
  for (int r = 1; r < world; ++r) {
    const int64_t target = static_cast<int64_t>((static_cast<__int128>(total) * r) / world);
    const int64_t i = std::lower_bound(row_cost_prefix.begin(), row_cost_prefix.end(), target) -
                      row_cost_prefix.begin();
    row_begin[r] = std::min<int64_t>(std::max<int64_t>(i, row_begin[r - 1]), m);
  }
  row_begin[world] = m;
}

extern "C" int folp_partition(int64_t m, int64_t n, int64_t nnz, const int64_t* rowval,
                              int32_t index_base, int32_t world_size, int64_t* row_begin_out,
                              int64_t* col_begin_out) {
  if (m < 0 || n < 0 || nnz < 0 || world_size < 1 || !row_begin_out || !col_begin_out ||
      (nnz > 0 && !rowval))
    return FOLP_INVALID_ARGUMENT;
  std::vector<int64_t> prefix(static_cast<size_t>(m) + 1, 0);
  for (int64_t k = 0; k < nnz; ++k) {
    const int64_t r = rowval[k] - index_base;
    if (r < 0 || r >= m) return FOLP_INVALID_ARGUMENT;
    prefix[r + 1] += 1;
  }
  for (int64_t i = 0; i < m; ++i) prefix[i + 1] += prefix[i] + 2;
  partition_rows(m, prefix, world_size, row_begin_out);
  int64_t n_pad = (n + world_size - 1) / world_size;
  n_pad += n_pad & 1;
  for (int r = 0; r <= world_size; ++r) col_begin_out[r] = std::min<int64_t>(n, r * n_pad);
  return FOLP_OK;
}

// ---------------------------------------------------------------------------
// Peer-memory exchange: every rank opens every other rank's region through CUDA IPC
// (NVLink / NVSwitch peer access), so that take_step needs no NCCL call: K1 pushes its
// slice of xbar into all copies, K2 pushes its rows of y+, four scalars and three flags per
// attempt travel as plain stores. Falls back to the NCCL exchanges
// (FOLP_NO_P2P=1, more than kMaxWorld ranks, or IPC refused on any rank).
// ---------------------------------------------------------------------------
// The pointers of one rank's exchange region as the kernels use them (own region or a peer's mapping).
static void set_peer_pointers(folp_handle* h, int r, void* region) {
  Bufs& B = h->B;
  const int P = h->world;
  const size_t fx = static_cast<size_t>(P) * h->n_pad, fy = static_cast<size_t>(P) * h->m_pad;
  double* base = static_cast<double*>(region);
  B.xbar_peer[r] = base;
  B.yfull_peer[r] = base + fx;
  B.sc_peer[r] = base + fx + fy;
  B.scx_peer[r] = base + fx + fy + static_cast<size_t>(P) * kScBlock;
  B.hx_peer[r] = base + fx + fy + 3 * static_cast<size_t>(P) * kScBlock;
  B.flag_peer[r] = reinterpret_cast<unsigned long long*>(base + fx + fy + 5 * static_cast<size_t>(P) * kScBlock);
}

// The exchange region has an address: wire the local pointers into it.
static void wire_region(folp_handle* h) {
  Bufs& B = h->B;
  const int P = h->world;
  const size_t fx = static_cast<size_t>(P) * h->n_pad, fy = static_cast<size_t>(P) * h->m_pad;
  double* base = static_cast<double*>(h->region);
  B.xbar = base;
  B.y_full = base + fx;
  B.m_pad = static_cast<int>(h->m_pad);
  B.sc_recv = base + fx + fy;
  B.scx = base + fx + fy + static_cast<size_t>(P) * kScBlock;
  B.hx = base + fx + fy + 3 * static_cast<size_t>(P) * kScBlock;
  B.flags = reinterpret_cast<unsigned long long*>(base + fx + fy + 5 * static_cast<size_t>(P) * kScBlock);
}

// The region is bound to an NVSwitch multicast object (folp_vmm.h): adopt its mappings.
static void adopt_multicast_region(folp_handle* h) {
  Bufs& B = h->B;
  const int P = h->world;
  h->region = h->vmm.own;
  wire_region(h);
  for (int r = 0; r < P; ++r) set_peer_pointers(h, r, h->vmm.peer[r]);
  double* mc = static_cast<double*>(h->vmm.mc);
  B.xbar_mc = mc;
  B.yfull_mc = mc + static_cast<size_t>(P) * h->n_pad;
}

static void read_exchange_switches(folp_handle* h) {
  Bufs& B = h->B;
  if (const char* d = getenv("FOLP_DEBUG_FLAGS")) B.dbg = atoi(d);
  if (getenv("FOLP_BULK_PUSH") != nullptr) B.dbg |= 8;
  if (B.dbg & 8) B.xbar_mc = B.yfull_mc = nullptr;  // the copy-engine pushes address the peers one by one
}

// Single-process multi-GPU: the ranks are threads of this process. The region becomes a multicast
// object when the devices sit behind an NVSwitch that offers it; otherwise every device enables direct
// peer access to every other one and the addresses of plain allocations are swapped through host memory.
static int setup_peer_exchange_in_process(folp_handle* h, size_t bytes) {
  MultiCtx* mc = h->shared;
  const int P = h->world;
  // has every rank got this far? (a rank that failed earlier passes these two barriers with peer_ok = 0)
  mc->peer_ok[h->rank] = P <= kMaxWorld ? 1 : 0;
  h->shared_rendezvous_done = true;
  mc->barrier();
  int ok = 1;
  for (int r = 0; r < P; ++r) ok = ok && mc->peer_ok[r];
  mc->barrier();
  if (!ok) {
    h->err = "single-process multi-GPU: a rank failed before the exchange was set up (or more than 8 devices)";
    return FOLP_UNSUPPORTED;
  }
  // from here on all ranks are alive and move in lockstep
  std::vector<int> devs(static_cast<size_t>(P));
  for (int r = 0; r < P; ++r) devs[r] = mc->sub[r]->device;
  VmmAllgather gather = [&](const uint64_t* mine, uint64_t* all) {
    memcpy(&mc->words[static_cast<size_t>(h->rank) * kVmmWords], mine, sizeof(uint64_t) * kVmmWords);
    mc->barrier();
    memcpy(all, mc->words.data(), sizeof(uint64_t) * kVmmWords * P);
    mc->barrier();
    return true;
  };
  const char* why = "";
  if (vmm_region_create(gather, h->rank, P, h->device, devs.data(), bytes, &h->vmm, &why)) {
    adopt_multicast_region(h);
    h->B.p2p = 1;
    read_exchange_switches(h);
    return FOLP_OK;
  }
  if (getenv("FOLP_TIMING") != nullptr && h->rank == 0) fprintf(stderr, "[folp_create] no multicast region: %s\n", why);
  for (int r = 0; r < P && ok; ++r) {
    const int peer = mc->sub[r]->device;
    if (r == h->rank || peer == h->device) continue;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, h->device, peer) != cudaSuccess || !can) { cudaGetLastError(); ok = 0; break; }
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
    cudaGetLastError();
  }
  // the region is zeroed before any peer can write into it
  if (ok && (cudaMalloc(&h->region, bytes) != cudaSuccess || cudaMemsetAsync(h->region, 0, bytes, h->stream) != cudaSuccess ||
             cudaStreamSynchronize(h->stream) != cudaSuccess)) {
    cudaGetLastError();
    ok = 0;
  }
  mc->regions[h->rank] = h->region;
  mc->peer_ok[h->rank] = ok;
  mc->barrier();
  for (int r = 0; r < P; ++r) ok = ok && mc->peer_ok[r];
  if (!ok) {
    h->err = "single-process multi-GPU needs direct peer access between all selected devices";
    mc->barrier();
    return FOLP_UNSUPPORTED;
  }
  wire_region(h);
  for (int r = 0; r < P; ++r) set_peer_pointers(h, r, mc->regions[r]);
  h->B.p2p = 1;
  read_exchange_switches(h);
  mc->barrier();
  return FOLP_OK;
}

// Allocates the exchange region (bytes) and makes it reachable from every rank: multicast object,
// else CUDA IPC of a plain allocation, else (FOLP_NO_P2P, IPC refused) private with NCCL exchanges.
static int setup_peer_exchange(folp_handle* h, size_t bytes) {
  Bufs& B = h->B;
  const int P = h->world;
  if (h->shared) return setup_peer_exchange_in_process(h, bytes);
  {  // NVSwitch multicast region; its create-time agreement rounds ride on the NCCL communicator
    uint64_t* boot = nullptr;  // device: [kVmmWords] send | [P * kVmmWords] receive
    TRY(cudaMalloc(&boot, sizeof(uint64_t) * kVmmWords * (static_cast<size_t>(P) + 1)));
    std::string gather_err;
    VmmAllgather gather = [&](const uint64_t* mine, uint64_t* all) {
      static_assert(sizeof(uint64_t) == sizeof(double), "words travel as doubles (copied, never computed on)");
      if (cudaMemcpyAsync(boot, mine, sizeof(uint64_t) * kVmmWords, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) return false;
      if (h->nccl->AllGather(boot, boot + kVmmWords, kVmmWords, ncclDouble, h->comm, h->stream) != ncclSuccess) return false;
      if (cudaMemcpyAsync(all, boot + kVmmWords, sizeof(uint64_t) * kVmmWords * P, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) return false;
      return cudaStreamSynchronize(h->stream) == cudaSuccess;
    };
    const char* why = "";
    bool up = false;
    const bool use_cache = getenv("FOLP_NO_REGION_CACHE") == nullptr;
    {  // a cached region of an earlier handle, if every rank holds the same one
      uint64_t mine[kVmmWords] = {0, 0, 0, 0};
      std::vector<uint64_t> all(static_cast<size_t>(P) * kVmmWords, 0);
      if (use_cache && getenv("FOLP_NO_MULTICAST") == nullptr && getenv("FOLP_NO_P2P") == nullptr) {
        const char* force = getenv("FOLP_MULTICAST");
        std::lock_guard<std::mutex> lock(g_vmm_mutex);
        if (force ? atoi(force) != 0 : P >= 4)  // the default of vmm_region_create
          for (VmmCacheEntry& e : g_vmm_cache)
            if (!e.busy && e.r.world == P && e.r.rank == h->rank && e.r.device == h->device && e.r.size >= bytes) {
              mine[0] = e.id;
              mine[1] = e.r.size;
              break;
            }
      }
      if (!gather(mine, all.data())) {
        cudaFree(boot);
        h->err = "exchange of the ranks at create time failed";
        return FOLP_NCCL_ERROR;
      }
      bool same = mine[0] != 0;
      for (int r = 0; r < P; ++r) same = same && all[static_cast<size_t>(r) * kVmmWords] == mine[0] && all[static_cast<size_t>(r) * kVmmWords + 1] == mine[1];
      if (same) {
        {
          std::lock_guard<std::mutex> lock(g_vmm_mutex);
          for (VmmCacheEntry& e : g_vmm_cache)
            if (e.id == mine[0]) { e.busy = true; h->vmm = e.r; }
        }
        h->vmm_cache_id = mine[0];
        // zero-filled before any rank of the new handle can push into it
        const bool zeroed = cudaMemsetAsync(h->vmm.own, 0, bytes, h->stream) == cudaSuccess && cudaStreamSynchronize(h->stream) == cudaSuccess;
        mine[0] = zeroed ? 1 : 0;
        bool ok = gather(mine, all.data());
        for (int r = 0; r < P; ++r) ok = ok && all[static_cast<size_t>(r) * kVmmWords] == 1;
        if (!ok) {
          cudaFree(boot);
          h->err = "zero-filling the cached exchange region failed";
          return FOLP_CUDA_ERROR;
        }
        up = true;
      }
    }
    if (!up) {
      up = vmm_region_create(gather, h->rank, P, h->device, nullptr, bytes, &h->vmm, &why);
      if (up && use_cache) {
        std::lock_guard<std::mutex> lock(g_vmm_mutex);
        VmmCacheEntry e;
        e.r = h->vmm;
        e.id = g_vmm_next_id++;
        e.busy = true;
        g_vmm_cache.push_back(e);
        h->vmm_cache_id = e.id;
      }
    }
    cudaFree(boot);
    if (up) {
      adopt_multicast_region(h);
      B.p2p = 1;
      read_exchange_switches(h);
      if (B.dbg & 4) {  // probe: gather from private copies of the exchanged vectors
        B.xbar_priv = h->d_cols;
        B.yfull_priv = h->d_rows;
      }
      return FOLP_OK;
    }
    cudaGetLastError();
    if (getenv("FOLP_TIMING") != nullptr && h->rank == 0) fprintf(stderr, "[folp_create] no multicast region: %s\n", why);
  }
  TRY(cudaMalloc(&h->region, bytes));
  TRY(cudaMemsetAsync(h->region, 0, bytes, h->stream));
  wire_region(h);
  static_assert(sizeof(cudaIpcMemHandle_t) + sizeof(int) <= kScBlock * sizeof(double), "handle fits a block");
  struct Msg {
    cudaIpcMemHandle_t handle;
    int ok;
  } mine;
  memset(&mine, 0, sizeof(mine));
  mine.ok = (P <= kMaxWorld && getenv("FOLP_NO_P2P") == nullptr) ? 1 : 0;
  if (mine.ok && cudaIpcGetMemHandle(&mine.handle, h->region) != cudaSuccess) {
    cudaGetLastError();
    mine.ok = 0;
  }
  TRY(cudaMemcpyAsync(B.sc_send, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
  NCCL_TRY(h->nccl->AllGather(B.sc_send, B.sc_recv, kScBlock, ncclDouble, h->comm, h->stream));
  TRY(cudaMemcpyAsync(h->h_sc, B.sc_recv, sizeof(double) * P * kScBlock, cudaMemcpyDeviceToHost,
                      h->stream));
  TRY(cudaStreamSynchronize(h->stream));
  bool all_ok = true;
  std::vector<Msg> msgs(static_cast<size_t>(P));
  for (int r = 0; r < P; ++r) {
    memcpy(&msgs[r], h->h_sc + static_cast<size_t>(r) * kScBlock, sizeof(Msg));
    all_ok = all_ok && msgs[r].ok;
  }
  int opened = all_ok ? 1 : 0;
  if (all_ok) {
    for (int r = 0; r < P && opened; ++r) {
      if (r == h->rank) continue;
      if (cudaIpcOpenMemHandle(&h->peer_region[r], msgs[r].handle, cudaIpcMemLazyEnablePeerAccess) !=
          cudaSuccess) {
        cudaGetLastError();
        h->peer_region[r] = nullptr;
        opened = 0;
      }
    }
  }
  // second round: peer mode only if every rank opened every handle
  memset(&mine, 0, sizeof(mine));
  mine.ok = opened;
  TRY(cudaMemsetAsync(B.sc_recv, 0, sizeof(double) * P * kScBlock, h->stream));
  TRY(cudaMemcpyAsync(B.sc_send, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
  NCCL_TRY(h->nccl->AllGather(B.sc_send, B.sc_recv, kScBlock, ncclDouble, h->comm, h->stream));
  TRY(cudaMemcpyAsync(h->h_sc, B.sc_recv, sizeof(double) * P * kScBlock, cudaMemcpyDeviceToHost,
                      h->stream));
  TRY(cudaMemsetAsync(B.sc_recv, 0, sizeof(double) * P * kScBlock, h->stream));
  TRY(cudaMemsetAsync(B.sc_send, 0, sizeof(double) * kScBlock, h->stream));
  TRY(cudaStreamSynchronize(h->stream));
  for (int r = 0; r < P; ++r) {
    Msg m2;
    memcpy(&m2, h->h_sc + static_cast<size_t>(r) * kScBlock, sizeof(Msg));
    opened = opened && m2.ok;
  }
  if (!opened) {
    for (int r = 0; r < P; ++r)
      if (h->peer_region[r]) {
        cudaIpcCloseMemHandle(h->peer_region[r]);
        h->peer_region[r] = nullptr;
      }
    B.p2p = 0;
    return FOLP_OK;
  }
  for (int r = 0; r < P; ++r) set_peer_pointers(h, r, r == h->rank ? h->region : h->peer_region[r]);
  B.p2p = 1;
  read_exchange_switches(h);
  if (B.dbg & 4) {  // probe: gather from private copies of the exchanged vectors
    B.xbar_priv = h->d_cols;
    B.yfull_priv = h->d_rows;
  }
  return FOLP_OK;
}

// ---------------------------------------------------------------------------
// folp_create
// ---------------------------------------------------------------------------
// CSC (n columns, row index row_of(k), value val_of(k)) -> CSR by a stable two-level counting
// sort that keeps every write stream cache-friendly:
//   1. column blocks (one per thread, balanced by nonzeros) are distributed into buckets by ROW
//      BLOCK (2^S rows, ~1-2 MB of output each); bucket (row block b, thread t) sits at a fixed
//      offset, so a row block's entries end up contiguous and in ascending column order;
//   2. row blocks are finished independently: a counting sort by row inside a range that fits
//      a core's cache.
// The result does not depend on the thread count. Returns false if a row index is out of range
// (nothing is written out of bounds).
// col_id (optional, n entries): the column index stored for column j (a renumbering of the columns;
// the entries of a row keep the order of the ORIGINAL column numbers, so that row sums are formed
// in the reference's order).
template <class RowOf, class ValOf>
static bool transpose_to_csr(int64_t n, int64_t m, int64_t nnz, const IVec& rp, RowOf row_of,
                             ValOf val_of, IVec* rp2_out, IVec* ci2_out, DVec* v2_out,
                             const int* col_id = nullptr, HostCarver* carver = nullptr, int** ci_ext = nullptr,
                             double** v_ext = nullptr) {
  // carver + ci_ext / v_ext: column indices and values go to carved (pinned, already faulted-in) memory and
  // are returned through the pointers; *ci2_out / *v2_out stay empty then
  IVec& rp2 = *rp2_out;
  rp2.resize(static_cast<size_t>(m) + 1);
  int* ci2 = carver && ci_ext ? carver->take<int>(static_cast<size_t>(nnz)) : nullptr;
  double* v2 = carver && v_ext ? carver->take<double>(static_cast<size_t>(nnz)) : nullptr;
  if (!ci2 || !v2) {
    ci2_out->resize(static_cast<size_t>(nnz));
    v2_out->resize(static_cast<size_t>(nnz));
    ci2 = ci2_out->data();
    v2 = v2_out->data();
  }
  if (ci_ext) *ci_ext = ci2;
  if (v_ext) *v_ext = v2;
  if (m == 0) { rp2[0] = 0; return nnz == 0; }
  unsigned hw = std::thread::hardware_concurrency();
  const int T = static_cast<int>(std::min<int64_t>(std::min<unsigned>(hw ? hw : 1u, kMaxHostThreadsU),
                                                   std::max<int64_t>(1, nnz / (1 << 18))));
  std::vector<int64_t> cut(static_cast<size_t>(T) + 1, n);
  for (int t = 0; t <= T; ++t) {  // balance the column blocks by nonzeros
    const int64_t target = nnz * t / T;
    cut[t] = std::lower_bound(rp.begin(), rp.end(), static_cast<int>(target)) - rp.begin();
  }
  cut[0] = 0;
  cut[T] = n;
  // rows per block: ~128 K nonzeros of output per block, at most 4096 blocks
  int S = 0;
  {
    const double avg = std::max(1.0, static_cast<double>(nnz) / static_cast<double>(m));
    while ((int64_t{1} << S) * avg < 131072.0 && (int64_t{1} << S) < m) ++S;
    while (((m - 1) >> S) + 1 > 4096) ++S;
  }
  const int64_t NB = ((m - 1) >> S) + 1;
  struct Entry { int row, col; double val; };
  std::vector<Entry, NoInit<Entry>> buf_own;
  const size_t carver_mark = carver ? carver->used : 0;  // buf is dead when this function returns: its room is handed back
  Entry* buf = carver ? carver->take<Entry>(static_cast<size_t>(nnz)) : nullptr;
  if (!buf) {
    buf_own.resize(static_cast<size_t>(nnz));
    buf = buf_own.data();
  }
  std::vector<int64_t> cnt(static_cast<size_t>(NB) * T, 0);  // [b * T + t]
  std::atomic<int> bad{0};
  std::vector<std::thread> th;
  auto run = [&](auto fn) {
    for (int t = 0; t < T; ++t) th.emplace_back(fn, t);
    for (auto& x : th) x.join();
    th.clear();
  };
  run([&](int t) {
    std::vector<int64_t> c(static_cast<size_t>(NB), 0);
    for (int64_t k = rp[cut[t]]; k < rp[cut[t + 1]]; ++k) {
      const int64_t r = row_of(k);
      if (r < 0 || r >= m) { bad.store(1); return; }
      c[r >> S] += 1;
    }
    for (int64_t b = 0; b < NB; ++b) cnt[b * T + t] = c[b];
  });
  if (bad.load()) return false;
  std::vector<int64_t> block_start(static_cast<size_t>(NB) + 1, 0);
  {
    int64_t run_sum = 0;
    for (int64_t b = 0; b < NB; ++b) {
      block_start[b] = run_sum;
      for (int t = 0; t < T; ++t) {
        const int64_t c = cnt[b * T + t];
        cnt[b * T + t] = run_sum;  // becomes the write cursor of bucket (b, t)
        run_sum += c;
      }
    }
    block_start[NB] = run_sum;
  }
  run([&](int t) {
    std::vector<int64_t> cur(static_cast<size_t>(NB));
    for (int64_t b = 0; b < NB; ++b) cur[b] = cnt[b * T + t];
    for (int64_t j = cut[t]; j < cut[t + 1]; ++j)
      for (int64_t k = rp[j]; k < rp[j + 1]; ++k) {
        const int r = static_cast<int>(row_of(k));
        buf[cur[r >> S]++] = Entry{r, col_id ? col_id[j] : static_cast<int>(j), val_of(k)};
      }
  });
  std::atomic<int64_t> next_block{0};
  run([&](int) {
    IVec hist(static_cast<size_t>(1) << S);
    for (;;) {
      const int64_t b = next_block.fetch_add(1);
      if (b >= NB) break;
      const int64_t r0 = b << S, r1 = std::min<int64_t>(m, r0 + (int64_t{1} << S));
      std::fill(hist.begin(), hist.begin() + (r1 - r0), 0);
      const int64_t e0 = block_start[b], e1 = block_start[b + 1];
      for (int64_t e = e0; e < e1; ++e) hist[buf[e].row - r0] += 1;
      int64_t pos = e0;
      for (int64_t i = r0; i < r1; ++i) {
        rp2[i] = static_cast<int>(pos);
        const int c = hist[i - r0];
        hist[i - r0] = static_cast<int>(pos);
        pos += c;
      }
      for (int64_t e = e0; e < e1; ++e) {
        const Entry& q = buf[e];
        const int at = hist[q.row - r0]++;
        ci2[at] = q.col;
        v2[at] = q.val;
      }
    }
  });
  rp2[m] = static_cast<int>(nnz);
  if (carver) carver->used = carver_mark;
  return true;
}

namespace folp {
void set_create_error(const std::string& msg) { g_create_error = msg; }

// The CSR view of a CSC matrix as a permutation (used by folp_rescale.cu): rowptr (m+1) and, for
// the k-th entry of the CSR order (rows ascending, columns ascending inside a row), its CSC
// position perm[k]. False if a row index is out of range.
bool csr_permutation(int64_t n, int64_t m, int64_t nnz, const int64_t* colptr, const int64_t* rowval,
                     int base, std::vector<int>* rowptr, std::vector<int>* perm) {
  IVec rp(static_cast<size_t>(n) + 1);
  for (int64_t j = 0; j <= n; ++j) rp[j] = static_cast<int>(colptr[j] - base);
  for (int64_t j = 0; j < n; ++j)
    if (rp[j] > rp[j + 1] || rp[j] < 0 || rp[j + 1] > nnz) return false;
  IVec rp2, ci2;
  DVec pos;
  auto row_of = [=](int64_t k) { return rowval[k] - base; };
  auto val_of = [](int64_t k) { return static_cast<double>(k); };  // exact: k < 2^31
  if (!transpose_to_csr(n, m, nnz, rp, row_of, val_of, &rp2, &ci2, &pos)) return false;
  rowptr->assign(rp2.begin(), rp2.end());
  perm->resize(static_cast<size_t>(nnz));
  parallel_for(0, nnz, 1 << 18, [&](int64_t lo, int64_t hi, int) {
    for (int64_t k = lo; k < hi; ++k) (*perm)[k] = static_cast<int>(pos[k]);
  });
  return true;
}
}  // namespace folp

// ---------------------------------------------------------------------------
// Physical reordering of the VARIABLES (= the rows of A') by column length.
//
// k_spmv gives one lane one row, and the 32 lanes of a group wait for the longest row: on the
// columns of a random matrix (Poisson(10) lengths, the bench workload) a group runs 6 rounds where
// 3.5 would do, and A'*y costs 64 us against 55 us for the uniform rows of A. Sorting rows by length
// through a slot table (the length-sorted windows of plan_tiles) removes the rounds but scatters the
// epilogue's per-row accesses and measured no gain. Here the variables themselves are renumbered --
// inside windows of kVarSortWindow consecutive variables, by decreasing column length, stably -- and
// EVERYTHING primal-indexed lives in the new numbering on the device (c, l, u, x, A'y, averages, D,
// the column indices of A, rows and columns of Q): no indirection anywhere in a kernel. The caller's
// numbering exists only at the boundary: inputs are gathered through new2old when they are uploaded,
// outputs scattered back when they are fetched. Row sums keep the reference's order: the entries of
// a row of A are stored in the order of their ORIGINAL column numbers, a row of A' (one variable) is
// untouched. Windows keep the renumbering local (structured LPs keep their column locality).
// FOLP_NO_VAR_SORT=1 disables it.
// ---------------------------------------------------------------------------
constexpr int kVarSortWindow = 2048;
struct VarOrder {
  bool identity = true;
  std::vector<int> new2old, old2new;
  IVec rp_new;      // column pointers in the new numbering
  IVec src_start;   // first entry of new column j' in the caller's arrays
};
static void order_variables(int64_t n, const IVec& rp, VarOrder* vo) {
  vo->identity = true;
  if (getenv("FOLP_NO_VAR_SORT") != nullptr || n < 64) return;
  int64_t window = kVarSortWindow;
  if (const char* w = getenv("FOLP_VAR_SORT_WINDOW")) window = std::max<int64_t>(64, atoll(w) / 32 * 32);
  vo->new2old.resize(static_cast<size_t>(n));
  constexpr int kCap = 2 * kNarrowMax;  // longer columns are warp-per-row items anyway: one bucket
  std::atomic<bool> changed{false};
  const int64_t nwin = (n + window - 1) / window;
  // windows are independent: spread over the host threads (the result does not depend on their number)
  parallel_for(0, nwin, 64, [&](int64_t wlo, int64_t whi, int) {
    bool any = false;
    for (int64_t wi = wlo; wi < whi; ++wi) {
      const int64_t w0 = wi * window;
      const int64_t w1 = std::min<int64_t>(n, w0 + window);
      int start[kCap + 2] = {0};
      auto bucket = [&](int64_t j) { return kCap - std::min<int>(kCap, rp[j + 1] - rp[j]); };  // decreasing length
      for (int64_t j = w0; j < w1; ++j) start[bucket(j) + 1] += 1;
      for (int b = 0; b <= kCap; ++b) start[b + 1] += start[b];
      // The groups of a sorted window run from the longest rows to the shortest, and k_spmv hands work
      // item i to warp i % warps_total: with a window of 64 groups and 4736 warps a warp would meet the
      // SAME rank of every window it visits -- always the longest group, or always the shortest
      // (measured: A'*y 79 us instead of 64). The order of the window's full groups is therefore rotated
      // by a per-window pseudo-random amount, so that every warp meets all ranks.
      const int64_t full = (w1 - w0) / 32;  // groups of exactly 32 variables; a shorter tail stays last
      const int64_t rot = full > 1 ? static_cast<int64_t>((static_cast<uint64_t>(w0 / window) * 2654435761ull >> 7) % full) : 0;
      for (int64_t j = w0; j < w1; ++j) {
        int64_t at = start[bucket(j)]++;  // rank inside the window
        if (at < full * 32) at = ((at / 32 + rot) % full) * 32 + at % 32;
        at += w0;
        vo->new2old[at] = static_cast<int>(j);
        any = any || at != j;
      }
    }
    if (any) changed.store(true, std::memory_order_relaxed);
  });
  if (!changed.load()) {
    vo->new2old.clear();
    return;
  }
  vo->identity = false;
  vo->old2new.resize(static_cast<size_t>(n));
  vo->rp_new.resize(static_cast<size_t>(n) + 1);
  vo->src_start.resize(static_cast<size_t>(n) + 1);
  vo->rp_new[0] = 0;
  parallel_for(0, n, 1 << 15, [&](int64_t lo, int64_t hi, int) {
    for (int64_t j = lo; j < hi; ++j) {
      const int o = vo->new2old[j];
      vo->old2new[o] = static_cast<int>(j);
      vo->rp_new[j + 1] = rp[o + 1] - rp[o];  // lengths; summed below
      vo->src_start[j] = rp[o];
    }
  });
  for (int64_t j = 0; j < n; ++j) vo->rp_new[j + 1] += vo->rp_new[j];
  vo->src_start[n] = 0;
}

// Host half of folp_create on one GPU: A' (= the caller's CSC, read as CSR) is planned and packed
// straight from the caller's Int64 arrays on a second thread while this one transposes into the
// CSR of A and packs that. No CUDA call. False if a row index is out of range.
struct HostMatrices {
  PackedMatrix pk_t, pk_a;
  IVec rp2;
  // packed column indices / values of A' and A: carved from the pinned scratch (pooled) or owned
  int *atc = nullptr, *ac = nullptr;
  double *atv = nullptr, *av = nullptr;
  bool pooled = false;
  IVec own_atc, own_ac;
  DVec own_atv, own_av;
};
// at_ready (optional) runs on the second thread as soon as A' is packed (folp_create uploads it from
// there while this thread is still transposing).
static bool prepare_host_matrices(const folp_problem* p, const IVec& rp, const VarOrder& vo, int warps_total,
                                  HostMatrices* out, const std::function<void()>& at_ready = nullptr,
                                  HostCarver* carver = nullptr) {
  const int64_t n = p->num_variables, m = p->num_constraints, nnz = p->num_nonzeros;
  const int base = p->index_base;
  const int64_t* src_row = p->rowval;
  const double* src_val = p->nzval;
  auto row_of = [=](int64_t k) { return src_row[k] - base; };  // may be out of range until checked
  auto val_of = [=](int64_t k) { return src_val[k]; };
  if (carver) {
    out->atc = carver->take<int>(static_cast<size_t>(nnz));
    out->atv = carver->take<double>(static_cast<size_t>(nnz));
    out->ac = carver->take<int>(static_cast<size_t>(nnz));
    out->av = carver->take<double>(static_cast<size_t>(nnz));
  }
  out->pooled = out->atc && out->atv && out->ac && out->av;
  if (!out->pooled) {
    out->own_atc.resize(static_cast<size_t>(nnz));
    out->own_atv.resize(static_cast<size_t>(nnz));
    out->own_ac.resize(static_cast<size_t>(nnz));
    out->own_av.resize(static_cast<size_t>(nnz));
    out->atc = out->own_atc.data();
    out->atv = out->own_atv.data();
    out->ac = out->own_ac.data();
    out->av = out->own_av.data();
  }
  const bool timing = getenv("FOLP_TIMING") != nullptr;
  const double t0 = now_sec();
  double t_side_plan = 0, t_side = 0;
  const IVec& rp_t = vo.identity ? rp : vo.rp_new;  // rows of A' in the device numbering of the variables
  std::thread side([&] {
    plan_tiles(static_cast<int>(n), rp_t, warps_total, &out->pk_t);
    t_side_plan = now_sec();
    fill_packed(out->pk_t, rp_t, [=](int k) { return static_cast<int>(row_of(k)); }, val_of,
                out->atc, out->atv, vo.identity ? nullptr : vo.src_start.data());
    t_side = now_sec();
    if (at_ready) at_ready();
  });
  IVec ci2_own;
  DVec v2_own;
  int* ci2 = nullptr;
  double* v2 = nullptr;
  const bool ok = transpose_to_csr(n, m, nnz, rp, row_of, val_of, &out->rp2, &ci2_own, &v2_own,
                                   vo.identity ? nullptr : vo.old2new.data(), out->pooled ? carver : nullptr, &ci2, &v2);
  const double t1 = now_sec();
  double t2 = t1;
  if (ok) {
    plan_tiles(static_cast<int>(m), out->rp2, warps_total, &out->pk_a);
    t2 = now_sec();
    fill_packed(out->pk_a, out->rp2, [&](int k) { return ci2[k]; }, [&](int k) { return v2[k]; },
                out->ac, out->av);
  }
  const double t3 = now_sec();
  side.join();
  if (timing)
    fprintf(stderr, "[folp host] main: transpose %.1f, plan A %.1f, pack A %.1f | side: plan A' %.1f, pack A' %.1f ms\n",
            (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t_side_plan - t0) * 1e3,
            (t_side - t_side_plan) * 1e3);
  return ok;
}

// Test / measurement hook without any CUDA call: the host half of folp_create for one GPU
// (index conversion, transposition, work-item planning and packing of both matrices);
// *milliseconds receives its wall-clock time.
extern "C" int folp_debug_host_prepare(const folp_problem* p, double* milliseconds) {
  if (!p || !milliseconds) return FOLP_INVALID_ARGUMENT;
  const int64_t n = p->num_variables, nnz = p->num_nonzeros;
  const double t0 = now_sec();
  IVec rp(static_cast<size_t>(n) + 1);
  for (int64_t j = 0; j <= n; ++j) rp[j] = nnz ? static_cast<int>(p->colptr[j] - p->index_base) : 0;
  HostMatrices hm;
  VarOrder vo;
  order_variables(n, rp, &vo);
  const bool ok = prepare_host_matrices(p, rp, vo, 148 * kSpmvCtasPerSm * kSpmvWarps, &hm);
  *milliseconds = (now_sec() - t0) * 1e3;
  return ok ? FOLP_OK : FOLP_INVALID_ARGUMENT;
}

// FOLP_TIMING=1: wall-clock phases of folp_create on stderr (development aid)
struct PhaseTimer {
  bool on = getenv("FOLP_TIMING") != nullptr;
  double t0 = now_sec(), last = t0;
  void mark(const char* what) {
    if (!on) return;
    const double t = now_sec();
    fprintf(stderr, "[folp_create] %-28s %8.2f ms\n", what, (t - last) * 1e3);
    last = t;
  }
  void total() {
    if (on) fprintf(stderr, "[folp_create] %-28s %8.2f ms\n", "TOTAL", (now_sec() - t0) * 1e3);
  }
};

static int create_impl(folp_handle* h, const folp_problem* p, const folp_params* q,
                       const folp_dist* dist) {
  PhaseTimer pt;
  // pinned host scratch leased for this create (one GPU); released on return, after the copies out of it have
  // completed (StreamGuard below is destroyed first)
  HostPoolLease host_pool;
  struct StreamGuard {
    folp_handle* h;
    ~StreamGuard() { if (h->stream) cudaStreamSynchronize(h->stream); }
  } stream_guard{h};
  const int64_t n = p->num_variables, m = p->num_constraints, nnz = p->num_nonzeros;
  if (n < 0 || m < 0 || nnz < 0 || p->num_equalities < 0 || p->num_equalities > m ||
      (p->index_base != 0 && p->index_base != 1)) {
    h->err = "invalid problem dimensions";
    return FOLP_INVALID_ARGUMENT;
  }
  if (nnz > 0 && (!p->colptr || !p->rowval || !p->nzval)) {
    h->err = "null matrix arrays";
    return FOLP_INVALID_ARGUMENT;
  }
  if ((n > 0 && (!p->objective_vector || !p->variable_lower_bound || !p->variable_upper_bound)) ||
      (m > 0 && !p->right_hand_side)) {
    h->err = "null problem vectors";
    return FOLP_INVALID_ARGUMENT;
  }
  if (n + m >= (int64_t{1} << 31) - 64 || nnz >= (int64_t{1} << 31) - 64) {
    h->err = "per-GPU shard exceeds 32-bit indexing (n+m or nnz >= 2^31)";
    return FOLP_UNSUPPORTED;
  }
  bool has_q = false;  // is_linear_programming_problem, quadratic_programming.jl: iszero(Q)
  for (int64_t k = 0; k < p->q_num_nonzeros && !has_q; ++k)
    if (p->q_nzval && p->q_nzval[k] != 0.0) has_q = true;
  if (has_q && (!p->q_colptr || !p->q_rowval)) {
    h->err = "null objective-matrix arrays";
    return FOLP_INVALID_ARGUMENT;
  }
  if (has_q && q->step_size_policy == FOLP_STEP_MALITSKY_POCK) {  // pdhg.jl:560-565
    h->err = "the Malitsky-Pock step size policy supports linear programs only";
    return FOLP_UNSUPPORTED;
  }
  if (has_q && dist && dist->world_size > 1) {
    h->err = "quadratic objectives are not supported in partitioned (multi-GPU) mode";
    return FOLP_UNSUPPORTED;
  }
  if (q->termination_evaluation_frequency < 1) {
    h->err = "termination_evaluation_frequency must be >= 1";
    return FOLP_INVALID_ARGUMENT;
  }
  if (!(q->initial_step_size > 0.0) || !(q->initial_primal_weight > 0.0)) {
    h->err = "initial_step_size and initial_primal_weight must be positive";
    return FOLP_INVALID_ARGUMENT;
  }
  if (dist && (dist->world_size < 1 || dist->rank < 0 || dist->rank >= dist->world_size)) {
    h->err = "invalid folp_dist rank / world_size";
    return FOLP_INVALID_ARGUMENT;
  }
  h->use_graphs = getenv("FOLP_NO_GRAPHS") == nullptr;
  h->world = dist ? dist->world_size : 1;
  h->rank = dist ? dist->rank : 0;
  if (h->world > 1 && !dist->nccl_unique_id && !h->shared) {
    h->err = "folp_dist.nccl_unique_id is required when world_size > 1";
    return FOLP_INVALID_ARGUMENT;
  }
  h->n_glob = n; h->m_glob = m;
  h->prm = *q;
  h->cache[0] = p->l_inf_norm_primal_linear_objective;
  h->cache[1] = p->l_inf_norm_primal_right_hand_side;
  h->cache[2] = p->l2_norm_primal_linear_objective;
  h->cache[3] = p->l2_norm_primal_right_hand_side;
  h->objective_constant = p->objective_constant;

  if (dist) h->device = dist->device;
  else TRY(cudaGetDevice(&h->device));
  TRY(cudaSetDevice(h->device));
  cudaDeviceProp prop;
  TRY(cudaGetDeviceProperties(&prop, h->device));
  if (prop.major != 10) {
    h->err = std::string("libfolp_b200 is built for sm_100a only; device is ") + prop.name;
    return FOLP_UNSUPPORTED;
  }
  h->sm_count = prop.multiProcessorCount;
  // one cooperative kernel per trust-region solve: single GPU, and the partitioned mode once the
  // peer-memory exchange is up (decided after setup_peer_exchange below)
  const bool tr_coop = prop.cooperativeLaunch && getenv("FOLP_TR_MULTIKERNEL") == nullptr;
  if (tr_coop) h->tr_grid = tr_solve_grid(h->sm_count);
  TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  TRY(cudaEventCreate(&h->ev0));
  TRY(cudaEventCreate(&h->ev1));
  TRY(static_cast<cudaError_t>(spmv_configure()));
  {  // the four pinned mirrors in one allocation (each cudaMallocHost costs ~1.5 ms)
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t b0 = up(sizeof(DevState)), b1 = up(sizeof(double) * 4 * kMaxScalars), b2 = up(sizeof(TrState) * kTrSlots),
                 b3 = up(sizeof(double) * h->world * kScBlock);
    char* block = nullptr;
    TRY(cudaMallocHost(reinterpret_cast<void**>(&block), b0 + b1 + b2 + b3));
    h->hs = reinterpret_cast<DevState*>(block);
    h->h_red = reinterpret_cast<double*>(block + b0);
    h->h_trs = reinterpret_cast<TrState*>(block + b0 + b1);
    h->h_sc = reinterpret_cast<double*>(block + b0 + b1 + b2);
  }

  if (h->world > 1 && !h->shared) {
    h->nccl = nccl_api(&h->err);
    if (!h->nccl) return FOLP_NCCL_ERROR;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "folp_dist carries a 128-byte id");
    memcpy(&id, dist->nccl_unique_id, sizeof(id));
    const bool use_cache = getenv("FOLP_NO_COMM_CACHE") == nullptr;
    const auto key = std::make_tuple(h->world, h->rank, h->device);
    if (use_cache) {
      std::lock_guard<std::mutex> lock(g_comm_mutex);
      auto it = g_comm_cache.find(key);
      if (it != g_comm_cache.end()) {
        h->comm = it->second;
        h->comm_cached = true;
      }
    }
    if (!h->comm) {
      NCCL_TRY(h->nccl->CommInitRank(&h->comm, h->world, id, h->rank));
      if (use_cache) {
        std::lock_guard<std::mutex> lock(g_comm_mutex);
        g_comm_cache[key] = h->comm;
        h->comm_cached = true;
      }
    }
  }

  pt.mark("device, stream, pinned state");
  // ---- matrices: A' in CSR is the caller's CSC; A in CSR by counting sort ----
  const int base = p->index_base;
  const int P = h->world;
  h->row_begin.assign(static_cast<size_t>(P) + 1, 0);
  VarOrder vo;  // device numbering of the variables (order_variables)
  {
    IVec rp(static_cast<size_t>(n) + 1);
    {
      std::atomic<int> bad{0};
      rp[n] = nnz ? static_cast<int>(p->colptr[n] - base) : 0;
      parallel_for(0, n, 1 << 16, [&](int64_t lo, int64_t hi, int) {
        for (int64_t j = lo; j < hi; ++j) {
          const int64_t a = nnz ? p->colptr[j] - base : 0, b = nnz ? p->colptr[j + 1] - base : 0;
          rp[j] = static_cast<int>(a);
          if (a > b || a < 0 || b > nnz) bad.store(1, std::memory_order_relaxed);
        }
      });
      if (bad.load()) {
        h->err = "colptr is not monotone";
        return FOLP_INVALID_ARGUMENT;
      }
    }
    const int64_t* src_row = p->rowval;
    const double* src_val = p->nzval;
    auto row_of = [=](int64_t k) { return src_row[k] - base; };  // may be out of range until checked
    auto val_of = [=](int64_t k) { return src_val[k]; };
    const int warps_total = h->sm_count * kSpmvCtasPerSm * kSpmvWarps;
    IVec rp2, ci2;
    DVec v2;
    int rc;
    int rc_arena = FOLP_OK;
    std::thread arena_thread;  // one GPU: the arena's cudaMalloc runs beside the renumbering of the variables
    if (P == 1) arena_thread = std::thread([&] {
      cudaSetDevice(h->device);
      // device memory of the whole handle in one allocation, reserved up front from upper bounds so that
      // A' can be uploaded while A is still being transposed (dev_alloc falls back to cudaMalloc when full)
        auto mat_bound = [&](int64_t rows, int64_t nz) {
          const int64_t tiles = rows / 32 + 2 * (nz / 33) + nz / kChunkNnz + 3 * static_cast<int64_t>(warps_total) + 64;
          return static_cast<size_t>(4 * (rows + 17) + 12 * (nz + 32) + 32 * tiles + 8 * (rows + 16) +
                                     8 * (nz / kChunkNnz + 64) + 4 * (rows + 64) + 8 * 256);
        };
        const int64_t qnnz = has_q ? p->q_num_nonzeros : 0;
        const size_t q_bytes = has_q ? mat_bound(n, qnnz) + static_cast<size_t>(5 * 8 * (n + 48)) : 0;
        const size_t vec_bytes = static_cast<size_t>(8) * (27 * (n + 48) + 21 * (m + 48)) +
                                 sizeof(double) * kNumSlots * kMaxScalars * kMaxPartialBlocks + (1 << 17);
        rc_arena = arena_reserve(h, mat_bound(n, nnz) + mat_bound(m, nnz) + q_bytes + vec_bytes);
    });
    order_variables(n, rp, &vo);
    if (arena_thread.joinable()) arena_thread.join();
    if (rc_arena) return rc_arena;
    if (!vo.identity) h->new2old = vo.new2old;
    const IVec& rp_t = vo.identity ? rp : vo.rp_new;  // column pointers in the device numbering
    if (P == 1) {
      h->row_begin[1] = m;
      h->n_pad = n; h->m_pad = m;
      h->n = n; h->m = m; h->nnz = nnz; h->neq = p->num_equalities;
      HostMatrices hm;
      int rc_at = FOLP_OK;
      // pinned scratch of the process, if it is free: transposition buckets (16 B / nonzero), unpacked CSR of A
      // and both packed matrices (12 B each), the renumbered copies of up to 8 primal vectors
      host_pool.acquire(static_cast<size_t>(nnz) * 52 + static_cast<size_t>(n) * 64 + (1 << 16), h->device);
      HostCarver* carver = host_pool.held ? &host_pool.carver : nullptr;
      const bool ok = prepare_host_matrices(p, rp, vo, warps_total, &hm, [&] {
        cudaSetDevice(h->device);  // second thread: A' goes to the device while the first one transposes
        rc_at = upload_matrix(h, &h->At, static_cast<int>(n), static_cast<int>(m), rp_t, hm.pk_t, hm.atc, hm.atv, !hm.pooled);
      }, carver);
      if (!ok) {
        h->err = "row index out of range";
        return FOLP_INVALID_ARGUMENT;
      }
      if (rc_at) return rc_at;
      pt.mark("transpose + pack (host), A' uploaded");
      if ((rc = upload_matrix(h, &h->A, static_cast<int>(m), static_cast<int>(n), hm.rp2, hm.pk_a, hm.ac, hm.av, !hm.pooled))) return rc;
    } else {
      // every rank transposes the whole matrix and keeps its shard: the transposition's scratch (buckets 16 B,
      // unpacked CSR 12 B per nonzero) comes from the process's pinned pool when it is free (no page faults;
      // one process per GPU only -- the rank threads of a single process would contend for the one pool)
      // ... and so do this rank's packed shards (12 B per local nonzero each) and the remapped slice of A'
      if (!h->shared) host_pool.acquire(static_cast<size_t>(nnz) * 28 + (static_cast<size_t>(nnz) / P + n + m) * 40 + (1 << 20), h->device);
      HostCarver* carver = host_pool.held ? &host_pool.carver : nullptr;
      int* ci2p = nullptr;
      double* v2p = nullptr;
      if (!transpose_to_csr(n, m, nnz, rp, row_of, val_of, &rp2, &ci2, &v2,
                            vo.identity ? nullptr : vo.old2new.data(), carver, &ci2p, &v2p)) {
        h->err = "row index out of range";
        return FOLP_INVALID_ARGUMENT;
      }
      pt.mark("transpose");
      std::vector<int64_t> prefix(static_cast<size_t>(m) + 1, 0);
      for (int64_t i = 0; i < m; ++i) prefix[i + 1] = prefix[i] + (rp2[i + 1] - rp2[i]) + 2;
      partition_rows(m, prefix, P, h->row_begin.data());
      h->n_pad = (n + P - 1) / P;
      h->n_pad += h->n_pad & 1;
      h->m_pad = 0;
      for (int r = 0; r < P; ++r) h->m_pad = std::max(h->m_pad, h->row_begin[r + 1] - h->row_begin[r]);
      h->m_pad += h->m_pad & 1;  // every rank's rows start on a 16-byte boundary of y_full (bulk pushes)
      h->col0 = std::min<int64_t>(n, h->rank * h->n_pad);
      h->n = std::min<int64_t>(n, (h->rank + 1) * h->n_pad) - h->col0;
      h->row0 = h->row_begin[h->rank];
      const int64_t row1 = h->row_begin[h->rank + 1];
      h->m = row1 - h->row0;
      h->neq = std::max<int64_t>(0, std::min<int64_t>(p->num_equalities, row1) - h->row0);
      // A_r: local rows, global columns
      const int k0 = rp2[h->row0], k1 = rp2[row1];
      {  // one allocation for this rank's matrices and vectors (upper bounds; dev_alloc falls back when it is full)
        const int64_t c1_ = std::min<int64_t>(n, h->col0 + h->n);
        const int64_t nz_t = rp_t[c1_] - rp_t[h->col0], nz_a = k1 - k0;
        auto mat_bound = [&](int64_t rows, int64_t nz) {
          const int64_t tiles = rows / 32 + 2 * (nz / 33) + nz / kChunkNnz + 3 * static_cast<int64_t>(warps_total) + 64;
          return static_cast<size_t>(4 * (rows + 17) + 12 * (nz + 32) + 32 * tiles + 8 * (rows + 16) +
                                     8 * (nz / kChunkNnz + 64) + 4 * (rows + 64) + 8 * 256);
        };
        const int64_t na_ = std::max(h->n_pad, h->n), ma_ = std::max(h->m_pad, h->m);
        const size_t vec_bound = static_cast<size_t>(8) * (30 * (na_ + 48) + 24 * (ma_ + 48) + P * (h->n_pad + h->m_pad) +
                                                           2 * (n + m + 64) + P * std::max(na_, ma_) + 8 * (h->n + h->m + 64)) +
                                 sizeof(double) * kNumSlots * kMaxScalars * kMaxPartialBlocks + (1 << 18);
        if ((rc = arena_reserve(h, mat_bound(h->m, nz_a) + mat_bound(h->n, nz_t) + vec_bound))) return rc;
      }
      h->nnz = k1 - k0;
      IVec lrp(static_cast<size_t>(h->m) + 1);
      for (int64_t i = 0; i <= h->m; ++i) lrp[i] = rp2[h->row0 + i] - k0;
      if ((rc = build_matrix(h, &h->A, static_cast<int>(h->m), static_cast<int>(n), lrp, ci2p + k0, v2p + k0, carver))) return rc;
      // (A[:, slice])': the local columns of the caller's CSC, i.e. n_local rows of full length.
      // Their column indices (= global rows of A) are remapped into the padded rank-major
      // layout of y_full: row i of rank q -> q * m_pad + (i - row_begin[q]).
      // (the slice is a range of the DEVICE numbering of the variables: its columns are gathered
      // from the caller's arrays through the renumbering)
      const int64_t c1 = h->col0 + h->n;
      const int t0 = rp_t[h->col0], t1 = rp_t[c1];
      IVec trp(static_cast<size_t>(h->n) + 1);
      for (int64_t j = 0; j <= h->n; ++j) trp[j] = rp_t[h->col0 + j] - t0;
      IVec owner_off(static_cast<size_t>(m) + 1);  // global row -> padded index
      for (int q = 0; q < P; ++q)
        for (int64_t i = h->row_begin[q]; i < h->row_begin[q + 1]; ++i)
          owner_off[i] = static_cast<int>(q * h->m_pad + (i - h->row_begin[q]));
      IVec tci_own;
      DVec tv_own;
      int* tci = carver ? carver->take<int>(static_cast<size_t>(t1 - t0)) : nullptr;
      double* tv = carver ? carver->take<double>(static_cast<size_t>(t1 - t0)) : nullptr;
      if (!tci || !tv) {
        tci_own.resize(static_cast<size_t>(t1 - t0));
        tv_own.resize(static_cast<size_t>(t1 - t0));
        tci = tci_own.data();
        tv = tv_own.data();
      }
      parallel_for(h->col0, c1, 1 << 12, [&](int64_t jlo, int64_t jhi, int) {
        for (int64_t j = jlo; j < jhi; ++j) {
          const int src = vo.identity ? rp[j] : vo.src_start[j];
          for (int k = rp_t[j]; k < rp_t[j + 1]; ++k) {
            tci[k - t0] = owner_off[row_of(src + (k - rp_t[j]))];  // rows were range-checked above
            tv[k - t0] = src_val[src + (k - rp_t[j])];
          }
        }
      });
      if ((rc = build_matrix(h, &h->At, static_cast<int>(h->n), static_cast<int>(P * h->m_pad), trp, tci, tv, carver)))
        return rc;
      h->nnz = (k1 - k0) + (t1 - t0);
    }
  }

  pt.mark("pack + upload matrices");
  // ---- objective matrix: CSR of Q by a stable counting sort of the caller's CSC ----
  if (has_q) {
    const int64_t qnnz = p->q_num_nonzeros;
    if (qnnz >= (int64_t{1} << 31) - 64) {
      h->err = "objective matrix exceeds 32-bit indexing";
      return FOLP_UNSUPPORTED;
    }
    IVec qrp(static_cast<size_t>(n) + 1, 0), qci(static_cast<size_t>(qnnz));
    DVec qv(static_cast<size_t>(qnnz));
    for (int64_t j = 0; j < n; ++j)
      if (p->q_colptr[j] > p->q_colptr[j + 1] || p->q_colptr[j] - base < 0 ||
          p->q_colptr[j + 1] - base > qnnz) {
        h->err = "objective-matrix colptr is not monotone";
        return FOLP_INVALID_ARGUMENT;
      }
    // rows and columns are variables: both in the device numbering; the entries of a row keep the
    // order of their ORIGINAL column numbers (the reference's summation order)
    auto dev_id = [&](int64_t j) { return vo.identity ? static_cast<int>(j) : vo.old2new[j]; };
    for (int64_t k = 0; k < qnnz; ++k) {
      const int64_t r = p->q_rowval[k] - base;
      if (r < 0 || r >= n) { h->err = "objective-matrix row index out of range"; return FOLP_INVALID_ARGUMENT; }
      qrp[dev_id(r) + 1] += 1;
    }
    for (int64_t i = 0; i < n; ++i) qrp[i + 1] += qrp[i];
    IVec cursor(qrp.begin(), qrp.end() - 1);
    for (int64_t j = 0; j < n; ++j)
      for (int64_t k = p->q_colptr[j] - base; k < p->q_colptr[j + 1] - base; ++k) {
        const int pos = cursor[dev_id(p->q_rowval[k] - base)]++;
        qci[pos] = dev_id(j);
        qv[pos] = p->q_nzval[k];
      }
    int rcq;
    if ((rcq = build_matrix(h, &h->Q, static_cast<int>(n), static_cast<int>(n), qrp, qci.data(), qv.data()))) return rcq;
  }

  // ---- vectors: primal-indexed arrays hold the local slice, dual-indexed arrays the local rows ----
  Bufs& B = h->B;
  B.has_q = has_q ? 1 : 0;
  const int64_t nl = h->n, ml = h->m;            // local lengths
  const int64_t na = std::max(h->n_pad, nl);     // allocation lengths (exchange strides)
  const int64_t ma = std::max(h->m_pad, ml);
  const int64_t c0 = h->col0, r0 = h->row0;
  B.n = static_cast<int>(nl); B.m = static_cast<int>(ml); B.neq = static_cast<int>(h->neq);
  B.world = P; B.rank = h->rank;
  B.xbar_off = P > 1 ? static_cast<int>(c0) : 0;
  B.grid_spmv = h->sm_count * kSpmvCtasPerSm;
#ifndef FOLP_VEC_CTAS_PER_SM
#define FOLP_VEC_CTAS_PER_SM 8
#endif
  B.grid_vec = h->sm_count * FOLP_VEC_CTAS_PER_SM;
  auto at = [](const double* v, int64_t off) { return v ? v + off : nullptr; };
  // primal-indexed input: entries [c0, c0 + nl) of the DEVICE numbering, gathered from the caller's
  std::vector<std::vector<double>> gathered;  // kept until the stream is synchronised at the end of folp_create
  gathered.reserve(8);
  auto upload_primal = [&](double** dst, const double* src, double fill) -> int {
    if (!src || vo.identity) return dev_upload(h, dst, at(src, c0), nl, fill);
    double* g = host_pool.held ? host_pool.carver.take<double>(static_cast<size_t>(nl)) : nullptr;
    if (!g) {
      gathered.emplace_back(static_cast<size_t>(nl));
      g = gathered.back().data();
    }
    parallel_for(0, nl, 1 << 16, [&](int64_t lo, int64_t hi, int) {
      for (int64_t j = lo; j < hi; ++j) g[j] = src[vo.new2old[c0 + j]];
    });
    return dev_upload(h, dst, g, nl, fill);
  };
  int rc;
  if ((rc = dev_alloc(h, &B.st, 1))) return rc;
  for (int k = 0; k < 2; ++k) {
    if ((rc = dev_zeros(h, &B.x[k], na))) return rc;
    if ((rc = dev_zeros(h, &B.y[k], ma))) return rc;
    if ((rc = dev_zeros(h, &B.aty[k], na))) return rc;
  }
  if (P == 1) {
    if ((rc = dev_zeros(h, &B.xbar, n))) return rc;
  } else {
    // one allocation, so that one handle (multicast object / CUDA IPC) exposes everything a peer touches
    const size_t fx = static_cast<size_t>(P) * h->n_pad, fy = static_cast<size_t>(P) * h->m_pad;
    const size_t region_doubles = fx + fy + 5 * static_cast<size_t>(P) * kScBlock +
                                  kNumFlagKinds * kMaxWorld + 16;
    if ((rc = dev_alloc(h, &B.dseq, 2))) return rc;
    TRY(cudaMemsetAsync(B.dseq, 0, 2 * sizeof(unsigned long long), h->stream));
    if ((rc = dev_zeros(h, &B.col_tmp, na))) return rc;
    if ((rc = dev_zeros(h, &B.sc_send, kScBlock))) return rc;
    if ((rc = dev_zeros(h, &h->d_rows, fy))) return rc;
    if ((rc = dev_zeros(h, &h->d_cols, fx))) return rc;
    if ((rc = setup_peer_exchange(h, region_doubles * sizeof(double)))) return rc;
    // the cooperative trust-region kernel exchanges its scalars through peer memory
    if (!B.p2p || getenv("FOLP_DIST_TR_MULTIKERNEL") != nullptr) h->tr_grid = 0;
  }
  if ((rc = upload_primal(&B.c, p->objective_vector, 0.0))) return rc;
  if ((rc = upload_primal(&B.l, p->variable_lower_bound, 0.0))) return rc;
  if ((rc = upload_primal(&B.u, p->variable_upper_bound, 0.0))) return rc;
  if ((rc = dev_upload(h, &B.b, at(p->right_hand_side, r0), ml, 0.0))) return rc;
  if ((rc = dev_zeros(h, &B.sum_x, na))) return rc;
  if ((rc = dev_zeros(h, &B.sum_y, ma))) return rc;
  if ((rc = dev_zeros(h, &B.avg_x, na))) return rc;
  if ((rc = dev_zeros(h, &B.avg_y, ma))) return rc;
  if ((rc = dev_zeros(h, &B.ax_avg, ma))) return rc;
  if ((rc = dev_zeros(h, &B.aty_avg, na))) return rc;
  if ((rc = dev_zeros(h, &B.ax_cur, ma))) return rc;
  if ((rc = dev_zeros(h, &B.last_x, na))) return rc;   // create_last_restart_info, sp.jl:199-213
  if ((rc = dev_zeros(h, &B.last_y, ma))) return rc;
  if ((rc = dev_zeros(h, &B.last_ax, ma))) return rc;
  if ((rc = dev_zeros(h, &B.last_aty, na))) return rc;
  if ((rc = upload_primal(&B.D, p->variable_rescaling, 1.0))) return rc;
  if ((rc = dev_upload(h, &B.E, at(p->constraint_rescaling, r0), ml, 1.0))) return rc;
  if ((rc = upload_primal(&B.c_orig, p->orig_objective_vector ? p->orig_objective_vector : p->objective_vector,
                          0.0)))
    return rc;
  if ((rc = upload_primal(&B.l_orig, p->orig_variable_lower_bound ? p->orig_variable_lower_bound
                                                                  : p->variable_lower_bound, 0.0)))
    return rc;
  if ((rc = upload_primal(&B.u_orig, p->orig_variable_upper_bound ? p->orig_variable_upper_bound
                                                                  : p->variable_upper_bound, 0.0)))
    return rc;
  if ((rc = dev_upload(h, &B.b_orig,
                       at(p->orig_right_hand_side ? p->orig_right_hand_side : p->right_hand_side, r0),
                       ml, 0.0)))
    return rc;
  if (has_q) {
    for (int k = 0; k < 2; ++k)
      if ((rc = dev_zeros(h, &B.qx[k], na))) return rc;
    if ((rc = dev_zeros(h, &B.dxv, na))) return rc;
    if ((rc = dev_zeros(h, &B.qx_avg, na))) return rc;
    if ((rc = dev_zeros(h, &B.last_qx, na))) return rc;
  }
  if ((rc = dev_zeros(h, &B.tr_t, std::max(nl + ml, P > 1 ? P * std::max(na, ma) : n + m)))) return rc;
  if ((rc = dev_zeros(h, &B.tr_d, std::max(nl + ml, n + m)))) return rc;
  if (h->tr_grid > 0 && getenv("FOLP_TR_SINGLE") == nullptr) h->trm_grid = tr_multi_grid(h->sm_count);
  if (h->trm_grid > 0) {
    if ((rc = dev_zeros(h, &B.trm_t, static_cast<size_t>(4) * (nl + ml)))) return rc;
    if ((rc = dev_zeros(h, &B.trm_d, static_cast<size_t>(4) * (nl + ml)))) return rc;
  }
  if ((rc = dev_zeros(h, &B.part, static_cast<size_t>(kNumSlots) * kMaxScalars * kMaxPartialBlocks)))
    return rc;
  if ((rc = dev_zeros(h, &B.red, static_cast<size_t>(4) * kMaxScalars))) return rc;
  if ((rc = dev_alloc(h, &B.counters, 8))) return rc;
  TRY(cudaMemsetAsync(B.counters, 0, 8 * sizeof(unsigned), h->stream));
  if ((rc = dev_alloc(h, &h->d_trs, kTrSlots))) return rc;
  // ---- persistent take_step kernel: barrier words, phase timers, co-resident grid ----
  if ((rc = dev_alloc(h, &B.bar_top, kBarL1Stride))) return rc;
  if ((rc = dev_alloc(h, &B.bar_gen, static_cast<size_t>(kBarMaxGroups + 1) * kBarGenStride))) return rc;
  if ((rc = dev_alloc(h, &h->d_timers, 16))) return rc;
  TRY(cudaMemsetAsync(B.bar_top, 0, sizeof(unsigned) * kBarL1Stride, h->stream));
  TRY(cudaMemsetAsync(B.bar_gen, 0, sizeof(unsigned long long) * (kBarMaxGroups + 1) * kBarGenStride, h->stream));
  TRY(cudaMemsetAsync(h->d_timers, 0, sizeof(unsigned long long) * 16, h->stream));
  if (const char* t = getenv("FOLP_P2P_TIMEOUT_MS")) {
    const double ms = atof(t);
    if (ms > 0.0) B.p2p_timeout_ns = static_cast<unsigned long long>(ms * 1e6);
  }
  // k_take_steps (one cooperative launch per batch of attempts) is the form of the partitioned mode,
  // where the peer exchanges ride on its grid barriers. On one GPU the CUDA graph of three kernels per
  // attempt is faster (measured, 1e6 x 1e6 x 1e7: 8 250 against 7 690 take_step iterations/s; a software
  // grid barrier costs ~3 us of fences and atomic round trips against ~1.5 us for a kernel boundary
  // inside a graph) and stays the default; FOLP_PERSISTENT=1 / 0 force either form.
  // On TINY instances (a couple of work items per warp of one 8-CTA cluster: the median Netlib LP) an
  // iteration is three latency chains; the persistent kernel runs as ONE thread-block cluster whose
  // hardware barrier costs ~0.2 us, against ~6 us per kernel of a graph launch.
  {
    const char* pe = getenv("FOLP_PERSISTENT");
    int64_t need = std::max<int64_t>((h->A.ntiles + kSpmvWarps - 1) / kSpmvWarps,
                                     (h->At.ntiles + kSpmvWarps - 1) / kSpmvWarps);
    need = std::max<int64_t>(need, (nl / 2 + kSpmvThreads - 1) / kSpmvThreads);
    if (has_q) need = std::max<int64_t>(need, (h->Q.ntiles + kSpmvWarps - 1) / kSpmvWarps);
    const bool tiny = P == 1 && need <= 2 * kTakeClusterMax && getenv("FOLP_NO_CLUSTER") == nullptr;
    const bool want = pe ? atoi(pe) != 0 : (P > 1 || tiny);
    if (want && prop.cooperativeLaunch && getenv("FOLP_NO_PERSISTENT") == nullptr && (P == 1 || B.p2p)) {
      const int cap = take_steps_grid(h->sm_count, P > 1);
      // no more CTAs than there is work: fewer CTAs make a cheaper barrier on small instances
      if (const char* g = getenv("FOLP_TAKE_GRID")) need = atoi(g);
      h->take_grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(cap, need)));
      if (cap <= 0) h->take_grid = 0;
      if (tiny && h->take_grid > 0) {
        h->take_cluster = true;
        h->take_grid = std::min(h->take_grid, kTakeClusterMax);
      }
    }
  }

  pt.mark("vectors");
  // ---- PdhgSolverState scalars, pdhg.jl:805-819 ----
  DevState s;
  memset(&s, 0, sizeof(s));
  s.step_size = q->initial_step_size;
  s.trial_step = q->initial_step_size;
  s.avg_weight = q->initial_step_size;
  s.primal_weight = q->initial_primal_weight;
  s.kkt_passes = q->initial_kkt_passes;
  s.ratio_step_sizes = 1.0;
  s.mp_need_primal = 1;
  s.policy = q->step_size_policy;
  s.reduction_exponent = q->reduction_exponent;
  s.growth_exponent = q->growth_exponent;
  s.downscaling_factor = q->downscaling_factor;
  s.breaking_factor = q->breaking_factor;
  s.interpolation_coefficient = q->interpolation_coefficient;
  *h->hs = s;
  if ((rc = push_state(h))) return rc;
  TRY(cudaStreamSynchronize(h->stream));
  h->start_time = now_sec();
  pt.mark("state");
  pt.total();
  return FOLP_OK;
}

extern "C" int folp_create(const folp_problem* problem, const folp_params* params,
                           const folp_dist* dist, folp_handle** out) {
  if (!problem || !params || !out) {
    g_create_error = "null argument";
    return FOLP_INVALID_ARGUMENT;
  }
  *out = nullptr;
  folp_handle* h = new (std::nothrow) folp_handle();
  if (!h) {
    g_create_error = "host allocation failed";
    return FOLP_OUT_OF_MEMORY;
  }
  int rc;
  try {
    rc = create_impl(h, problem, params, dist);
  } catch (const std::bad_alloc&) {
    h->err = "host allocation failed";
    rc = FOLP_OUT_OF_MEMORY;
  }
  if (rc) {
    g_create_error = h->err;
    free_handle(h);
    return rc;
  }
  *out = h;
  return FOLP_OK;
}

static void destroy_multi(folp_handle* w) {
  MultiCtx* mc = w->multi;
  if (!mc) return;
  if (!mc->workers.empty()) {
    mc->call([mc](int r) {
      free_handle(mc->sub[r], /*collective=*/true);
      mc->sub[r] = nullptr;
      return 0;
    });
    {
      std::unique_lock<std::mutex> lock(mc->mu);
      mc->stop = true;
      mc->cv_job.notify_all();
    }
    for (auto& t : mc->workers) t.join();
  } else {
    for (folp_handle* q : mc->sub)
      if (q && q->vmm.active) vmm_region_unmap(&q->vmm);
    for (folp_handle* q : mc->sub) free_handle(q);
  }
  delete mc;
  w->multi = nullptr;
}

extern "C" int folp_create_multi(const folp_problem* problem, const folp_params* params, int32_t n_gpus,
                                 const int32_t* device_ids, folp_handle** out) {
  if (!problem || !params || !out || n_gpus < 1 || n_gpus > kMaxWorld) {
    g_create_error = "null argument or n_gpus outside 1..8";
    return FOLP_INVALID_ARGUMENT;
  }
  *out = nullptr;
  if (n_gpus == 1) {
    folp_dist d{0, 1, device_ids ? device_ids[0] : 0, 0, nullptr};
    if (!device_ids && cudaGetDevice(&d.device) != cudaSuccess) d.device = 0;
    return folp_create(problem, params, &d, out);
  }
  for (int a = 0; a < n_gpus; ++a)
    for (int b = 0; b < a; ++b)
      if (device_ids && device_ids[a] == device_ids[b]) {
        g_create_error = "device_ids must be distinct";
        return FOLP_INVALID_ARGUMENT;
      }
  folp_handle* w = new (std::nothrow) folp_handle();
  MultiCtx* mc = new (std::nothrow) MultiCtx();
  if (!w || !mc) {
    delete w;
    delete mc;
    g_create_error = "host allocation failed";
    return FOLP_OUT_OF_MEMORY;
  }
  w->multi = mc;
  mc->world = n_gpus;
  mc->rcs.assign(static_cast<size_t>(n_gpus), 0);
  mc->regions.assign(static_cast<size_t>(n_gpus), nullptr);
  mc->peer_ok.assign(static_cast<size_t>(n_gpus), 0);
  mc->words.assign(static_cast<size_t>(n_gpus) * kVmmWords, 0);
  for (int r = 0; r < n_gpus; ++r) {
    folp_handle* q = new folp_handle();
    q->shared = mc;
    q->device = device_ids ? device_ids[r] : r;
    mc->sub.push_back(q);
  }
  for (int r = 0; r < n_gpus; ++r) mc->workers.emplace_back([mc, r] { mc->worker(r); });
  const int rc = mc->call([&](int r) {
    folp_handle* q = mc->sub[r];
    folp_dist d{r, n_gpus, q->device, 0, nullptr};
    int rc_;
    try {
      rc_ = create_impl(q, problem, params, &d);
    } catch (const std::bad_alloc&) {
      q->err = "host allocation failed";
      rc_ = FOLP_OUT_OF_MEMORY;
    }
    if (rc_ && !q->shared_rendezvous_done) {  // failed before the ranks met: let the others through (they fail too)
      mc->peer_ok[r] = 0;
      q->shared_rendezvous_done = true;
      mc->barrier();
      mc->barrier();
    }
    return rc_;
  });
  if (rc) {
    g_create_error = "folp_create_multi failed";
    for (folp_handle* q : mc->sub)
      if (q && !q->err.empty()) { g_create_error = q->err; break; }
    destroy_multi(w);
    delete w;
    return rc;
  }
  w->prm = *params;
  w->world = n_gpus;
  w->n_glob = problem->num_variables;
  w->m_glob = problem->num_constraints;
  *out = w;
  return FOLP_OK;
}

// runs fn on every rank's thread and surfaces the first failing rank's message
static int multi_dispatch(folp_handle* w, std::function<int(folp_handle*, int)> fn) {
  MultiCtx* mc = w->multi;
  const int rc = mc->call([&](int r) { return fn(mc->sub[r], r); });
  if (rc)
    for (folp_handle* q : mc->sub)
      if (q && !q->err.empty()) { w->err = q->err; break; }
  return rc;
}

extern "C" void folp_destroy(folp_handle* h) {
  if (h && h->multi) {
    destroy_multi(h);
    delete h;
    return;
  }
  if (h && getenv("FOLP_TIMING"))
    fprintf(stderr, "[folp_destroy] iterations %lld, launches %lld, trust-region solves %lld, passes %lld, "
                    "take_step seconds %.4f\n",
            static_cast<long long>(h->hs ? h->hs->iterations : 0), static_cast<long long>(h->launches),
            static_cast<long long>(h->tr_solves), static_cast<long long>(h->tr_passes), h->basic_time);
  free_handle(h);
}

extern "C" const char* folp_last_error(const folp_handle* h) {
  return h ? h->err.c_str() : g_create_error.c_str();
}

extern "C" const char* folp_build_info(void) {
#define FOLP_STR2(x) #x
#define FOLP_STR(x) FOLP_STR2(x)
  return "libfolp_b200;sm_100a;cuda 12.9;fp64;fmad=false;spmv=one row per lane on position-major "
         "32-row groups (length-sorted windows), coalesced direct loads, " FOLP_STR(FOLP_GATHER_UNROLL)
         " gathers per lane and round, " FOLP_STR(FOLP_SPMV_CTAS_PER_SM) " CTAs of 256 per SM;"
         "chunk_nnz=" FOLP_STR(FOLP_CHUNK_NNZ);
}

extern "C" int folp_nccl_unique_id(void* out128) {
  if (!out128) return FOLP_INVALID_ARGUMENT;
  memset(out128, 0, 128);
  const NcclApi* api = nccl_api(&g_create_error);
  if (!api) return FOLP_NCCL_ERROR;
  ncclUniqueId id;
  const ncclResult_t r = api->GetUniqueId(&id);
  if (r != ncclSuccess) {
    g_create_error = std::string("ncclGetUniqueId: ") + api->GetErrorString(r);
    return FOLP_NCCL_ERROR;
  }
  memcpy(out128, &id, sizeof(id));
  return FOLP_OK;
}

// ---------------------------------------------------------------------------
// exchanges of the row-partitioned mode (no-ops / plain copies when world == 1)
// ---------------------------------------------------------------------------
// slice (n_pad readable doubles per rank) -> full (world * n_pad), rank-major
static int allgather_cols(folp_handle* h, const double* slice, double* full) {
  NCCL_TRY(h->nccl->AllGather(slice, full, static_cast<size_t>(h->n_pad), ncclDouble, h->comm,
                              h->stream));
  return FOLP_OK;
}
// B.sc_send -> B.sc_recv (device resident, consumed by k_finalize_dist); NCCL mode of take_step
static int exchange_scalars_dev(folp_handle* h) {
  NCCL_TRY(h->nccl->AllGather(h->B.sc_send, h->B.sc_recv, kScBlock, ncclDouble, h->comm, h->stream));
  return FOLP_OK;
}
// Evaluation-block exchange: every rank contributes src[0..count) (device, kScBlock readable); returns
// the device buffer holding world * kScBlock doubles, rank-major. Peer mode: one tiny kernel pushing
// into alternating receive buffers; otherwise an NCCL allgather.
static int exchange_block(folp_handle* h, const double* src, int count, const double** recv) {
  Bufs& B = h->B;
  if (B.p2p) {
    h->xchg_seq += 1;
    const int parity = static_cast<int>(h->xchg_seq & 1);
    launch_exchange(B, src, count, h->xchg_seq, parity, h->stream);
    CHECK_LAUNCH();
    h->launches += 1;
    *recv = B.hx + static_cast<size_t>(parity) * h->world * kScBlock;
    return FOLP_OK;
  }
  NCCL_TRY(h->nccl->AllGather(src, B.sc_recv, kScBlock, ncclDouble, h->comm, h->stream));
  *recv = B.sc_recv;
  return FOLP_OK;
}
// h->h_red[off .. off+count) <- the global value of the reduced scalars B.red[off ...):
// the first nsum entries are sums, the rest maxima. Ranks are combined in rank order on
// the host, so every rank sees the same bits. Synchronises the stream.
static int pull_red(folp_handle* h, int off, int count, int nsum) {
  if (h->world == 1) {
    TRY(cudaMemcpyAsync(h->h_red + off, h->B.red + off, sizeof(double) * count,
                        cudaMemcpyDeviceToHost, h->stream));
    TRY(cudaStreamSynchronize(h->stream));
    return FOLP_OK;
  }
  const double* recv = nullptr;
  int rc = exchange_block(h, h->B.red + off, kScBlock, &recv);
  if (rc) return rc;
  TRY(cudaMemcpyAsync(h->h_sc, recv, sizeof(double) * h->world * kScBlock, cudaMemcpyDeviceToHost,
                      h->stream));
  TRY(cudaStreamSynchronize(h->stream));
  if (h->B.p2p) {
    unsigned timed_out = 0;
    TRY(cudaMemcpy(&timed_out, h->B.counters + 6, sizeof(unsigned), cudaMemcpyDeviceToHost));
    if (timed_out) {
      h->err = "peer exchange timed out: a rank of the row partition stopped responding";
      return FOLP_CUDA_ERROR;
    }
  }
  for (int k = 0; k < count; ++k) {
    double v = h->h_sc[k];
    for (int r = 1; r < h->world; ++r) {
      const double w = h->h_sc[r * kScBlock + k];
      v = k < nsum ? v + w : fmax(v, w);
    }
    h->h_red[off + k] = v;
  }
  return FOLP_OK;
}
// out = A_r * v where v is a primal-indexed vector held as slices (local rows out)
static int spmv_A(folp_handle* h, const double* v_slice, double* out_rows) {
  const double* in = v_slice;
  if (h->world > 1 && h->B.p2p) {
    // peer memory: the slice is pushed into every rank's xbar (free between batches of attempts: the
    // evaluation block of every rank lies between exchanges that all ranks take part in)
    h->xchg_seq += 1;
    launch_push_vec(h->B, v_slice, 0, h->xchg_seq, h->stream);
    h->launches += 1;
    in = h->B.xbar;
  } else if (h->world > 1) {
    int rc = allgather_cols(h, v_slice, h->d_cols);
    if (rc) return rc;
    in = h->d_cols;
  }
  launch_spmv_plain(h->A, in, out_rows, h->B.grid_spmv, h->stream);
  CHECK_LAUNCH();
  h->launches += 1;
  return FOLP_OK;
}
// out = (A' * w) on the local slice where w is dual-indexed, held as local rows
static int spmv_At(folp_handle* h, const double* w_rows, double* out_slice) {
  const double* in = w_rows;
  if (h->world > 1 && h->B.p2p) {
    h->xchg_seq += 1;
    launch_push_vec(h->B, w_rows, 1, h->xchg_seq, h->stream);
    h->launches += 1;
    in = h->B.y_full;
  } else if (h->world > 1) {  // the column-slice matrix gathers from the padded rank-major full vector
    NCCL_TRY(h->nccl->AllGather(w_rows, h->d_rows, static_cast<size_t>(h->m_pad), ncclDouble, h->comm,
                                h->stream));
    in = h->d_rows;
  }
  launch_spmv_plain(h->At, in, out_slice, h->B.grid_spmv, h->stream);
  CHECK_LAUNCH();
  h->launches += 1;
  return FOLP_OK;
}
// Peer-memory mode: an all-rank rendezvous with no payload. A rank raises its flag from a kernel that is
// stream-ordered behind everything it enqueued before, so once a rank has passed the rendezvous every
// peer has finished reading (or copying out) whatever the previous exchange delivered into its staging.
static int peer_rendezvous(folp_handle* h) {
  h->xchg_seq += 1;
  launch_exchange(h->B, nullptr, 0, h->xchg_seq, 0, h->stream);
  CHECK_LAUNCH();
  h->launches += 1;
  return FOLP_OK;
}
// host_out (global length) <- a primal-indexed device vector held as slices. A collective in
// partitioned mode: every rank takes part in the gather even when it wants no copy (host_out NULL).
// Peer-memory mode stages the gather in xbar / y_full (free outside a batch of attempts); unlike the
// pushes of the evaluation block, which are separated by exchanges all ranks take part in, fetches
// may follow each other directly and are read by the copy engine: a rendezvous before (the peers are
// done with the staging) and after (so is this rank, before anybody's next push).
static int fetch_cols(folp_handle* h, const double* slice, double* host_out) {
  if (h->n_glob == 0) return FOLP_OK;
  if (h->world == 1 && !host_out) return FOLP_OK;
  const double* src = slice;
  int rc;
  if (h->world > 1 && h->B.p2p) {
    if ((rc = peer_rendezvous(h))) return rc;
    h->xchg_seq += 1;
    launch_push_vec(h->B, slice, 0, h->xchg_seq, h->stream);
    h->launches += 1;
    src = h->B.xbar;
  } else if (h->world > 1) {
    if ((rc = allgather_cols(h, slice, h->d_cols))) return rc;
    src = h->d_cols;
  }
  if (host_out && h->new2old.empty()) {
    TRY(cudaMemcpyAsync(host_out, src, sizeof(double) * h->n_glob, cudaMemcpyDeviceToHost, h->stream));
  } else if (host_out) {  // back to the caller's numbering of the variables
    h->h_tmp.resize(static_cast<size_t>(h->n_glob));
    TRY(cudaMemcpyAsync(h->h_tmp.data(), src, sizeof(double) * h->n_glob, cudaMemcpyDeviceToHost, h->stream));
    TRY(cudaStreamSynchronize(h->stream));
    for (int64_t j = 0; j < h->n_glob; ++j) host_out[h->new2old[j]] = h->h_tmp[j];
  }
  if (h->world > 1 && h->B.p2p && (rc = peer_rendezvous(h))) return rc;
  return FOLP_OK;
}
// host_out (global length) <- a dual-indexed device vector held as local rows (collective, as above)
static int fetch_rows(folp_handle* h, const double* rows, double* host_out) {
  if (h->m_glob == 0) return FOLP_OK;
  if (h->world == 1) {
    if (host_out)
      TRY(cudaMemcpyAsync(host_out, rows, sizeof(double) * h->m_glob, cudaMemcpyDeviceToHost, h->stream));
    return FOLP_OK;
  }
  const double* full = h->d_rows;
  int rc;
  if (h->B.p2p) {
    if ((rc = peer_rendezvous(h))) return rc;
    h->xchg_seq += 1;
    launch_push_vec(h->B, rows, 1, h->xchg_seq, h->stream);
    h->launches += 1;
    full = h->B.y_full;
  } else {
    NCCL_TRY(h->nccl->AllGather(rows, h->d_rows, static_cast<size_t>(h->m_pad), ncclDouble, h->comm,
                                h->stream));
  }
  if (host_out)
    for (int r = 0; r < h->world; ++r) {
      const int64_t cnt = h->row_begin[r + 1] - h->row_begin[r];
      if (cnt > 0)
        TRY(cudaMemcpyAsync(host_out + h->row_begin[r], full + r * h->m_pad, sizeof(double) * cnt,
                            cudaMemcpyDeviceToHost, h->stream));
    }
  if (h->B.p2p && (rc = peer_rendezvous(h))) return rc;
  return FOLP_OK;
}

// ---------------------------------------------------------------------------
// take_step batches
// ---------------------------------------------------------------------------
static int launch_attempts_any(folp_handle* h, int attempts) {
  if (h->world == 1) {
    launch_step_attempts(h->B, h->A, h->At, h->Q, attempts, h->stream);
    return FOLP_OK;
  }
  const Bufs& B = h->B;
  for (int a = 0; a < attempts; ++a) {
    int rc;
    launch_dist_primal(B, h->stream);  // peer mode: pushes its slice of xbar to every rank
    if (!B.p2p)
      NCCL_TRY(h->nccl->AllGather(B.xbar + static_cast<size_t>(h->rank) * h->n_pad, B.xbar,
                                  static_cast<size_t>(h->n_pad), ncclDouble, h->comm, h->stream));
    launch_dist_dual(B, h->A, h->stream);  // peer mode: pushes its rows of y+ to every rank
    if (!B.p2p)
      NCCL_TRY(h->nccl->AllGather(B.y_full + static_cast<size_t>(h->rank) * h->m_pad, B.y_full,
                                  static_cast<size_t>(h->m_pad), ncclDouble, h->comm, h->stream));
    launch_dist_trans(B, h->A, h->At, h->stream);
    if (!B.p2p && (rc = exchange_scalars_dev(h))) return rc;
    launch_dist_finalize(B, h->stream);
  }
  return FOLP_OK;
}

static int enqueue_attempts(folp_handle* h, int attempts) {
  if (attempts <= 0) return FOLP_OK;
  int rc;
  if (h->take_grid > 0) {  // the whole batch as one cooperative launch
    const cudaError_t le = static_cast<cudaError_t>(
        h->take_cluster ? launch_take_steps_cluster(h->B, h->A, h->At, h->Q, attempts, h->take_grid, h->stream)
                        : launch_take_steps(h->B, h->A, h->At, h->Q, attempts, h->take_grid, h->stream));
    if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorNotSupported ||
        (h->take_cluster && le != cudaSuccess)) {
      cudaGetLastError();  // the device is shared (MPS partition, ...): kernel-per-phase form from now on
      if (h->world > 1) {  // the ranks must agree on the form: an error rather than a silent divergence
        h->err = "cooperative launch of k_take_steps refused on this rank";
        return FOLP_CUDA_ERROR;
      }
      h->take_grid = 0;
    } else {
      TRY(le);
      h->launches += 1;
      return FOLP_OK;
    }
  }
  if (!h->use_graphs || attempts < 2) {
    if ((rc = launch_attempts_any(h, attempts))) return rc;
    CHECK_LAUNCH();
  } else {
    auto it = h->step_graphs.find(attempts);
    if (it == h->step_graphs.end()) {
      cudaGraph_t g = nullptr;
      cudaGraphExec_t ge = nullptr;
      TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
      rc = launch_attempts_any(h, attempts);
      cudaError_t ce = cudaStreamEndCapture(h->stream, &g);
      if (rc) return rc;
      TRY(ce);
      TRY(cudaGraphInstantiate(&ge, g, 0));
      cudaGraphDestroy(g);
      it = h->step_graphs.emplace(attempts, ge).first;
    }
    TRY(cudaGraphLaunch(it->second, h->stream));
  }
  h->launches += (h->world == 1 ? (h->B.has_q ? 5 : 3) : 4) * static_cast<int64_t>(attempts);
  return FOLP_OK;
}

// Runs take_step until `target` iterations are complete or numerical_error.
static int run_steps(folp_handle* h, int64_t target) {
  DevState* s = h->hs;
  int rc;
  s->target_iterations = target;
  s->active = (s->iterations < target && !s->numerical_error) ? 1 : 0;
  if ((rc = push_state(h))) return rc;
  TRY(cudaEventRecord(h->ev0, h->stream));
  int stalled_batches = 0;
  while (s->active) {
    const int64_t before = s->iterations;
    if (stalled_batches >= 64) {  // 64 batches (thousands of attempts) without one accepted step: not a solve any more
      s->numerical_error = 1;
      s->active = 0;
      if ((rc = push_state(h))) return rc;
      break;
    }
    const int64_t remaining = target - s->iterations;
    int64_t attempts = remaining + (remaining >= 8 ? remaining / 16 + 1 : 0);
    if (s->policy == FOLP_STEP_CONSTANT) attempts = remaining;
    attempts = std::min<int64_t>(attempts, 512);
    if ((rc = enqueue_attempts(h, static_cast<int>(attempts)))) return rc;
    if ((rc = pull_state(h))) return rc;
    if (s->p2p_timeout) {
      h->err = "peer exchange timed out: a rank of the row partition stopped responding";
      return FOLP_CUDA_ERROR;
    }
    stalled_batches = s->iterations == before ? stalled_batches + 1 : 0;
  }
  TRY(cudaEventRecord(h->ev1, h->stream));
  TRY(cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->basic_time += 1e-3 * ms;
  return FOLP_OK;
}

// ---------------------------------------------------------------------------
// evaluation block helpers
// ---------------------------------------------------------------------------
static double jl_max(double a, double b) {
  if (a != a) return a;
  if (b != b) return b;
  return a > b ? a : b;
}

struct BoundResult {
  double lagrangian_value, lower_bound_value, upper_bound_value;
};
static double get_gap(const BoundResult& r) { return r.upper_bound_value - r.lower_bound_value; }

// One trust-region solve on device; *out receives the final TrState.
static int tr_round(folp_handle* h, const TrProblem& P, int passes, bool init) {
  if (h->world == 1) {
    launch_tr(h->B, P, h->d_trs, passes, init, h->stream);
    CHECK_LAUNCH();
    h->launches += passes + 1 + (init ? 1 : 0);
    return FOLP_OK;
  }
  // row-partitioned: every kernel leaves local sums, the ranks exchange them, and a
  // one-thread kernel applies the totals -- identical TrState on every rank
  int rc;
  auto stage = [&](int st) -> int {
    launch_tr_stage(h->B, P, h->d_trs, st, h->stream);
    const double* recv = nullptr;
    if ((rc = exchange_block(h, h->B.sc_send, 16, &recv))) return rc;
    launch_tr_combine(h->B, P, h->d_trs, st, recv, h->stream);
    h->launches += 2;
    return FOLP_OK;
  };
  if (init && (rc = stage(kTrInit))) return rc;
  for (int p = 0; p < passes; ++p)
    if ((rc = stage(kTrPass))) return rc;
  if ((rc = stage(kTrFinal))) return rc;
  CHECK_LAUNCH();
  return FOLP_OK;
}

// The search was abandoned on non-finite data (a diverged iterate: e.g. the constant step size of
// pdhg.jl:826-836 ignores Q). The reference's median search returns NaN bounds there and the solve
// carries on to its limits; so does this one.
static void tr_abandoned(TrState* t) {
  t->tau = NAN;
  t->v_primal = NAN;
  t->v_dual = NAN;
}

// slot: which result of the evaluation block this is (0..kTrSlots-1); when the block was enqueued
// blind the solve has already run and its state was fetched with the block's single read-back.
static int tr_solve(folp_handle* h, const TrProblem& P, TrState* out, int slot) {
  int rc;
  if (h->pre_mask & (1u << slot)) {
    *out = h->pre[slot];
    h->tr_solves += 1;
    h->tr_passes += out->passes;
    if (out->done != 1) tr_abandoned(out);
    return FOLP_OK;
  }
  if (h->tr_grid > 0) {  // single GPU: one cooperative kernel, one host read
    const cudaError_t le = static_cast<cudaError_t>(
        launch_tr_solve(h->B, P, h->d_trs, h->tr_grid, h->xchg_seq + 1, h->stream));
    if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorNotSupported) {
      cudaGetLastError();  // the device is shared (MPS partition, ...): kernel-per-pass form from now on
      h->tr_grid = 0;
    } else {
      TRY(le);
      h->launches += 1;
      h->tr_solves += 1;
      TRY(cudaMemcpyAsync(h->h_trs, h->d_trs, sizeof(TrState), cudaMemcpyDeviceToHost, h->stream));
      TRY(cudaStreamSynchronize(h->stream));
      *out = *h->h_trs;
      h->tr_passes += out->passes;
      // (k_tr_solve numbers its exchanges from the device-resident counter Bufs::dseq)
      if (h->world > 1) {
        unsigned timed_out = 0;
        TRY(cudaMemcpy(&timed_out, h->B.counters + 6, sizeof(unsigned), cudaMemcpyDeviceToHost));
        if (timed_out) {
          h->err = "peer exchange timed out: a rank of the row partition stopped responding";
          return FOLP_CUDA_ERROR;
        }
      }
      if (out->done != 1) tr_abandoned(out);
      return FOLP_OK;
    }
  }
  if ((rc = tr_round(h, P, h->world == 1 ? 10 : 6, true))) return rc;
  h->tr_solves += 1;
  for (int round = 0;; ++round) {
    TRY(cudaMemcpyAsync(h->h_trs, h->d_trs, sizeof(TrState), cudaMemcpyDeviceToHost, h->stream));
    TRY(cudaStreamSynchronize(h->stream));
    if (h->h_trs->done) break;
    if (round > 40) {
      *out = *h->h_trs;
      tr_abandoned(out);
      return FOLP_OK;
    }
    if ((rc = tr_round(h, P, h->world == 1 ? 16 : 6, false))) return rc;
  }
  *out = *h->h_trs;
  h->tr_passes += out->passes;
  return FOLP_OK;
}

// bound_optimal_objective (tr.jl:271-360), EUCLIDEAN_NORM, at a point whose
// products A*x and A'*y are already in HBM.
static int euclidean_gap(folp_handle* h, const double* px, const double* atp, const double* py,
                         const double* axp, const double* qxp, double wp, double wd, double radius,
                         BoundResult* out, int slot) {
  TrProblem P{px, atp, py, axp, wp, wd, radius, 1, 1,
              h->prm.use_approximate_localized_duality_gap, h->B.has_q ? qxp : nullptr};
  TrState t;
  int rc = tr_solve(h, P, &t, slot);
  if (rc) return rc;
  out->lagrangian_value = ((0.5 * t.xqx + t.cx) - t.x_aty) + t.y_b + h->objective_constant;
  out->lower_bound_value = out->lagrangian_value + t.v_primal;
  out->upper_bound_value = out->lagrangian_value - t.v_dual;
  return FOLP_OK;
}

// check_termination_criteria, term.jl:233-273
static int check_termination(const folp_params* c, const double cache[4], const folp_eval* s) {
  const double abs_obj = fabs(s->primal_objective) + fabs(s->dual_objective);
  const double gap = fabs(s->primal_objective - s->dual_objective);
  double perr, pbase, derr, dbase;
  if (c->optimality_norm == FOLP_L_INF) {
    perr = s->l_inf_primal_residual; pbase = cache[1];
    derr = s->l_inf_dual_residual; dbase = cache[0];
  } else {
    perr = s->l2_primal_residual; pbase = cache[3];
    derr = s->l2_dual_residual; dbase = cache[2];
  }
  const double ea = c->eps_optimal_absolute, er = c->eps_optimal_relative;
  if (derr < ea + er * dbase && perr < ea + er * pbase && gap < ea + er * abs_obj)
    return FOLP_TERMINATION_REASON_OPTIMAL;
  if (s->dual_ray_objective > 0.0 &&
      s->max_dual_ray_infeasibility / s->dual_ray_objective <= c->eps_primal_infeasible)
    return FOLP_TERMINATION_REASON_PRIMAL_INFEASIBLE;
  if (s->primal_ray_linear_objective < 0.0 &&
      s->max_primal_ray_infeasibility / (-s->primal_ray_linear_objective) <= c->eps_dual_infeasible &&
      s->primal_ray_quadratic_norm / (-s->primal_ray_linear_objective) <= c->eps_dual_infeasible)
    return FOLP_TERMINATION_REASON_DUAL_INFEASIBLE;
  if (s->iteration_number >= c->iteration_limit) return FOLP_TERMINATION_REASON_ITERATION_LIMIT;
  if (s->cumulative_kkt_matrix_passes >= c->kkt_matrix_pass_limit)
    return FOLP_TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT;
  if (s->cumulative_time_sec >= c->time_sec_limit) return FOLP_TERMINATION_REASON_TIME_LIMIT;
  return 0;
}

// run_restart_scheme, sp.jl:688-846. wp/wd are the constant norm weights of
// define_norms (pdhg.jl:265-276).
static int run_restart_scheme(folp_handle* h, int64_t iterations_completed, double wp, double wd,
                              const BoundResult& gap_at_avg_seed, int* choice_out) {
  (void)gap_at_avg_seed;
  DevState* s = h->hs;
  const folp_params* rp = &h->prm;
  Bufs& B = h->B;
  if (!(s->count_x > 0 && s->count_y > 0)) {
    *choice_out = FOLP_RESTART_CHOICE_NO_RESTART;
    return FOLP_OK;
  }
  int rc;
  const int64_t restart_length = s->count_x;
  int do_restart = 0;
  if (static_cast<double>(restart_length) >=
      rp->artificial_restart_threshold * static_cast<double>(iterations_completed))
    do_restart = 1;
  // distances to the last restart point (needed by every branch that restarts)
  if (!h->pre_dist) {
    double* red = B.red + 2 * kMaxScalars;
    launch_dist(B, red, h->stream);
    CHECK_LAUNCH();
    h->launches += 1;
    if ((rc = pull_red(h, 2 * kMaxScalars, SD_TOTAL, SD_TOTAL))) return rc;
  }
  const double* dist = h->h_red + 2 * kMaxScalars;
  const double avg_px = sqrt(wp * dist[SD_avg_x]), avg_dy = sqrt(wd * dist[SD_avg_y]);
  const double cur_px = sqrt(wp * dist[SD_cur_x]), cur_dy = sqrt(wd * dist[SD_cur_y]);

  int reset_to_average;
  int have_candidate = 0, have_ax_cur = 0;
  BoundResult candidate_gap{0, 0, 0};
  double candidate_distance = 0.0;
  if (rp->restart_scheme == FOLP_NO_RESTARTS) {
    reset_to_average = 0;
  } else {
    // compute_localized_duality_gaps, sp.jl:432-496
    const double d_avg = sqrt(avg_px * avg_px + avg_dy * avg_dy);
    const double d_cur = sqrt(cur_px * cur_px + cur_dy * cur_dy);
    BoundResult g_avg, g_cur;
    if ((rc = euclidean_gap(h, B.avg_x, B.aty_avg, B.avg_y, B.ax_avg, B.qx_avg, wp, wd, d_avg,
                            &g_avg, 2)))
      return rc;
    if (!h->pre_ax_cur && (rc = spmv_A(h, B.x[s->cur], B.ax_cur))) return rc;
    have_ax_cur = 1;
    if ((rc = euclidean_gap(h, B.x[s->cur], B.aty[s->cur], B.y[s->cur], B.ax_cur, B.qx[s->cur], wp,
                            wd, d_cur, &g_cur, 3)))
      return rc;
    // should_reset_to_average, sp.jl:530-547
    const double cur_ng = get_gap(g_cur) / d_cur, avg_ng = get_gap(g_avg) / d_avg;
    if (rp->restart_to_current_metric == FOLP_GAP_OVER_DISTANCE_SQUARED)
      reset_to_average = (cur_ng / d_cur >= avg_ng / d_avg);
    else if (rp->restart_to_current_metric == FOLP_GAP_OVER_DISTANCE)
      reset_to_average = (cur_ng >= avg_ng);
    else
      reset_to_average = 1;
    have_candidate = 1;
    if (reset_to_average) { candidate_gap = g_avg; candidate_distance = d_avg; }
    else { candidate_gap = g_cur; candidate_distance = d_cur; }
  }
  if (!do_restart) {
    const double pw = s->primal_weight;
    if (rp->restart_scheme == FOLP_ADAPTIVE_NORMALIZED) {  // sp.jl:549-593
      const double d_last = sqrt(h->pd_last * h->pd_last * pw + h->dd_last * h->dd_last / pw);
      BoundResult g_last;
      if ((rc = euclidean_gap(h, B.last_x, B.last_aty, B.last_y, B.last_ax, B.last_qx, wp, wd,
                              d_last, &g_last, 4)))
        return rc;
      const double ncg = get_gap(candidate_gap) / candidate_distance;
      const double nlg = get_gap(g_last) / d_last;
      const double ratio = ncg / nlg;
      if (ratio < rp->necessary_reduction_for_restart) {
        if (ratio < rp->sufficient_reduction_for_restart) do_restart = 1;
        else if (ratio > h->gap_reduction_ratio_last_trial) do_restart = 1;
      }
      h->gap_reduction_ratio_last_trial = ratio;
    } else if ((rp->restart_scheme == FOLP_ADAPTIVE_LOCALIZED ||
                rp->restart_scheme == FOLP_ADAPTIVE_DISTANCE) && !h->has_last_gap) {
      do_restart = 1;
    } else if (rp->restart_scheme == FOLP_ADAPTIVE_LOCALIZED) {  // sp.jl:597-620
      const double new_potential = get_gap(candidate_gap) / static_cast<double>(restart_length);
      const double old_potential = h->last_gap / static_cast<double>(h->last_restart_length);
      if (new_potential / old_potential < rp->necessary_reduction_for_restart) do_restart = 1;
    } else if (rp->restart_scheme == FOLP_ADAPTIVE_DISTANCE) {  // sp.jl:623-648
      const double d_last = sqrt(h->pd_last * h->pd_last * pw + h->dd_last * h->dd_last / pw);
      const double new_potential = candidate_distance / static_cast<double>(restart_length);
      const double old_potential = d_last / static_cast<double>(h->last_restart_length);
      if (new_potential / old_potential < rp->necessary_reduction_for_restart) do_restart = 1;
    } else if (rp->restart_scheme == FOLP_FIXED_FREQUENCY &&
               rp->restart_frequency_if_fixed <= restart_length) {
      do_restart = 1;
    }
  }
  if (!do_restart) {
    *choice_out = FOLP_RESTART_CHOICE_NO_RESTART;
    return FOLP_OK;
  }
  launch_apply_restart(B, reset_to_average, have_ax_cur, h->stream);
  CHECK_LAUNCH();
  h->launches += 1;
  s->count_x = s->count_y = 0;  // reset_solution_weighted_average, sp.jl:238-250
  s->sum_w_x = s->sum_w_y = 0.0;
  // update_last_restart_info, sp.jl:893-927
  h->pd_last = avg_px / sqrt(s->primal_weight);
  h->dd_last = avg_dy * sqrt(s->primal_weight);
  h->last_restart_length = restart_length;
  h->has_last_gap = have_candidate;
  h->last_gap = have_candidate ? get_gap(candidate_gap) : 0.0;
  *choice_out = reset_to_average ? FOLP_RESTART_CHOICE_RESTART_TO_AVERAGE
                                 : FOLP_RESTART_CHOICE_WEIGHTED_AVERAGE_RESET;
  return FOLP_OK;
}

// One GPU, cooperative trust-region kernel available: enqueues everything the evaluation block may
// need after the statistics kernels, in the order (and with the operands) evaluate() and
// run_restart_scheme() would launch it, with the weights / radii taken from the device-resident sums.
// Sets pre_mask / pre_dist / pre_ax_cur for what was enqueued; on a refused cooperative launch it
// stops early and the remaining pieces are computed by the synchronous path.
static int evaluate_enqueue_blind(folp_handle* h) {
  if (h->tr_grid <= 0 || getenv("FOLP_EVAL_SYNC") != nullptr) return FOLP_OK;
  if (h->world > 1 && !h->B.p2p) return FOLP_OK;
  DevState* s = h->hs;
  const folp_params* rp = &h->prm;
  Bufs& B = h->B;
  const double wp = 1 / s->step_size * s->primal_weight;  // define_norms, pdhg.jl:265-276
  const double wd = 1 / s->step_size / s->primal_weight;
  auto solve = [&](TrProblem P, int slot) -> bool {
    const cudaError_t le = static_cast<cudaError_t>(
        launch_tr_solve(B, P, h->d_trs + slot, h->tr_grid, h->xchg_seq + 1, h->stream));
    if (le != cudaSuccess) {
      cudaGetLastError();
      if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorNotSupported) h->tr_grid = 0;
      return false;
    }
    h->launches += 1;
    h->pre_mask |= 1u << slot;
    return true;
  };
  // update_objective_bound_estimates, sp.jl:1015-1047 (never approximate)
  TrProblem Pp{B.avg_x, B.aty_avg, B.avg_y, B.ax_avg, wp, wd, 1.0, 1, 0, 0, B.has_q ? B.qx_avg : nullptr};
  Pp.param_src = kTrParamBounds;
  TrProblem Pd = Pp;
  Pd.use_primal = 0;
  Pd.use_dual = 1;
  const bool restart_possible = s->count_x > 0 && s->count_y > 0;  // else run_restart_scheme returns at once
  const bool gaps = restart_possible && rp->restart_scheme != FOLP_NO_RESTARTS;
  const int approx = rp->use_approximate_localized_duality_gap;
  TrProblem Pa{B.avg_x, B.aty_avg, B.avg_y, B.ax_avg, wp, wd, 0.0, 1, 1, approx, B.has_q ? B.qx_avg : nullptr};
  Pa.param_src = kTrParamDistAvg;
  TrProblem Pc{B.x[s->cur], B.aty[s->cur], B.y[s->cur], B.ax_cur, wp, wd, 0.0, 1, 1, approx,
               B.has_q ? B.qx[s->cur] : nullptr};
  Pc.param_src = kTrParamDistCur;
  const double pw = s->primal_weight;
  const double d_last = sqrt(h->pd_last * h->pd_last * pw + h->dd_last * h->dd_last / pw);  // sp.jl:549-593
  TrProblem Pl{B.last_x, B.last_aty, B.last_y, B.last_ax, wp, wd, d_last, 1, 1, approx,
               B.has_q ? B.last_qx : nullptr};
  auto distances = [&]() -> int {
    launch_dist(B, B.red + 2 * kMaxScalars, h->stream);
    h->launches += 1;
    if (h->world > 1) {  // rank-ordered totals of the local sums, back into B.red on every rank
      const double* recv = nullptr;
      int rc = exchange_block(h, B.red + 2 * kMaxScalars, kScBlock, &recv);
      if (rc) return rc;
      launch_combine_red(B, recv, 2 * kMaxScalars, SD_TOTAL, SD_TOTAL, 0, 0, 0, h->stream);
      h->launches += 1;
    }
    h->pre_dist = true;
    return FOLP_OK;
  };
  if (h->trm_grid > 0) {
    // every solve of the block in ONE cooperative kernel, behind everything it reads
    TrMulti M{};
    M.P[0] = Pp;
    M.P[1] = Pd;
    M.nslots = 2;
    int rc;
    if (restart_possible && (rc = distances())) return rc;
    if (gaps) {
      if ((rc = spmv_A(h, B.x[s->cur], B.ax_cur))) return rc;
      h->pre_ax_cur = true;
      M.P[2] = Pa;
      M.P[3] = Pc;
      M.nslots = 4;
      if (rp->restart_scheme == FOLP_ADAPTIVE_NORMALIZED) {
        M.P[4] = Pl;
        M.nslots = 5;
      }
    }
    const cudaError_t le = static_cast<cudaError_t>(launch_tr_multi(B, M, h->d_trs, h->trm_grid, h->stream));
    if (le == cudaSuccess) {
      h->launches += 1;
      h->pre_mask = (1u << M.nslots) - 1u;
      return FOLP_OK;
    }
    cudaGetLastError();
    if (h->world > 1) {
      h->err = std::string("cooperative launch of k_tr_multi refused on this rank: ") + cudaGetErrorString(le);
      return FOLP_CUDA_ERROR;
    }
    h->trm_grid = 0;  // one kernel per solve from now on; dist / ax_cur above stay valid
    if (!solve(Pp, 0) || !solve(Pd, 1)) return FOLP_OK;
    if (!gaps) return FOLP_OK;
    if (!solve(Pa, 2) || !solve(Pc, 3)) return FOLP_OK;
    if (rp->restart_scheme == FOLP_ADAPTIVE_NORMALIZED) solve(Pl, 4);
    return FOLP_OK;
  }
  if (!solve(Pp, 0) || !solve(Pd, 1)) return FOLP_OK;
  if (!restart_possible) return FOLP_OK;
  {
    int rc = distances();
    if (rc) return rc;
  }
  if (!gaps) return FOLP_OK;
  if (!solve(Pa, 2)) return FOLP_OK;
  {
    int rc = spmv_A(h, B.x[s->cur], B.ax_cur);
    if (rc) return rc;
  }
  h->pre_ax_cur = true;
  if (!solve(Pc, 3)) return FOLP_OK;
  if (rp->restart_scheme == FOLP_ADAPTIVE_NORMALIZED) solve(Pl, 4);
  return FOLP_OK;
}

// The evaluation block, pdhg.jl:892-1023. hs holds the current device state.
static int evaluate(folp_handle* h, folp_eval* out) {
  DevState* s = h->hs;
  const folp_params* prm = &h->prm;
  Bufs& B = h->B;
  const int64_t iteration = h->iteration;
  int rc;
  s->kkt_passes += 2.0;  // :899
  const int use_current = (s->numerical_error || s->count_x == 0 || s->count_y == 0) ? 1 : 0;
  if (s->pending_avg) {
    launch_flush_avg(B, h->stream);
    h->launches += 2;
    s->pending_avg = 0;
  }
  launch_make_avg(B, use_current, h->stream);
  if ((rc = spmv_A(h, B.avg_x, B.ax_avg))) return rc;
  if ((rc = spmv_At(h, B.avg_y, B.aty_avg))) return rc;
  if (B.has_q) {  // Q * avg_x: objective values, primal gradient, Lagrangian, ray norm
    launch_spmv_plain(h->Q, B.avg_x, B.qx_avg, B.grid_spmv, h->stream);
    h->launches += 1;
  }
  launch_stats_n(B, B.red, h->stream);
  launch_stats_m(B, B.red + kMaxScalars, h->stream);
  CHECK_LAUNCH();
  h->launches += 3;
  h->pre_mask = 0;
  h->pre_dist = h->pre_ax_cur = false;
  const bool blind = (h->world == 1 || h->B.p2p) && h->tr_grid > 0 && getenv("FOLP_EVAL_SYNC") == nullptr;
  if (h->world == 1 || blind) {
    // The REST of the block's device work -- both bound-estimate solves, the distances to the last
    // restart point, A * x_cur and the localized-gap solves -- is enqueued behind the statistics
    // without a host round trip (their weights / radii are formed on the device from the sums in
    // B.red, see TrParamSrc), and everything comes back with ONE copy and ONE synchronisation. The
    // scalar decisions below then consume the fetched results instead of launching and waiting seven
    // times per evaluation. (A terminating evaluation has computed restart candidates it does not use.)
    // Partitioned mode over peer memory: the ranks' local sums are exchanged and combined in rank order
    // on the device (same bits on every rank), the products gather through peer pushes: no NCCL call.
    if (h->world > 1) {
      static_assert(2 * kMaxScalars <= kScBlock, "both statistics blocks travel in one exchange");
      const double* recv = nullptr;
      if ((rc = exchange_block(h, B.red, kScBlock, &recv))) return rc;
      launch_combine_red(B, recv, 0, SN_TOTAL, SN_NSUM, kMaxScalars, SM_TOTAL, SM_NSUM, h->stream);
      h->launches += 1;
    }
    if ((rc = evaluate_enqueue_blind(h))) return rc;
    TRY(cudaMemcpyAsync(h->h_red, B.red, sizeof(double) * 3 * kMaxScalars, cudaMemcpyDeviceToHost,
                        h->stream));
    if (h->pre_mask)
      TRY(cudaMemcpyAsync(h->h_trs, h->d_trs, sizeof(TrState) * kTrSlots, cudaMemcpyDeviceToHost, h->stream));
    unsigned* timed_out = reinterpret_cast<unsigned*>(h->h_sc);
    *timed_out = 0;
    if (h->world > 1)
      TRY(cudaMemcpyAsync(timed_out, B.counters + 6, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    TRY(cudaStreamSynchronize(h->stream));
    if (*timed_out) {
      h->err = "peer exchange timed out: a rank of the row partition stopped responding";
      return FOLP_CUDA_ERROR;
    }
    for (int k = 0; k < kTrSlots; ++k)
      if (h->pre_mask & (1u << k)) h->pre[k] = h->h_trs[k];
  } else {
    static_assert(2 * kMaxScalars <= kScBlock, "both statistics blocks travel in one exchange");
    if ((rc = pull_red(h, 0, SN_TOTAL, SN_NSUM))) return rc;  // gathers all 64 scalars once
    for (int k = 0; k < SM_TOTAL; ++k) {
      double v = h->h_sc[kMaxScalars + k];
      for (int r = 1; r < h->world; ++r) {
        const double w = h->h_sc[r * kScBlock + kMaxScalars + k];
        v = k < SM_NSUM ? v + w : fmax(v, w);
      }
      h->h_red[kMaxScalars + k] = v;
    }
  }
  const double* sn = h->h_red;
  const double* sm = h->h_red + kMaxScalars;

  folp_eval e;
  memset(&e, 0, sizeof(e));
  e.iteration_number = static_cast<int32_t>(iteration - 1);
  e.candidate_type = FOLP_POINT_TYPE_AVERAGE_ITERATE;
  e.cumulative_kkt_matrix_passes = s->kkt_passes;
  e.cumulative_time_sec = now_sec() - h->start_time;
  const double eps_ratio = prm->eps_optimal_absolute / prm->eps_optimal_relative;
  const double c0 = h->objective_constant;
  // compute_convergence_information, isu.jl:228-280
  const double xqx = B.has_q ? sn[SN_xqx] : 0.0;  // xhat' Q_O xhat
  e.primal_objective = c0 + sn[SN_cx] + 0.5 * xqx;  // isu.jl:67-74
  e.l_inf_primal_residual = jl_max(jl_max(sm[SM_pres_max], sn[SN_lviol_max]), sn[SN_uviol_max]);
  e.l2_primal_residual = sqrt(sm[SM_pres2] + sn[SN_lviol2] + sn[SN_uviol2]);
  e.relative_l_inf_primal_residual = e.l_inf_primal_residual / (eps_ratio + h->cache[1]);
  e.relative_l2_primal_residual = e.l2_primal_residual / (eps_ratio + h->cache[3]);
  e.l_inf_primal_variable = sn[SN_x_max];
  e.l2_primal_variable = sqrt(sn[SN_x2]);
  e.dual_objective = (sm[SM_by] + c0 - 0.5 * xqx) + sn[SN_rcobj];  // isu.jl:186-190
  e.l_inf_dual_residual = jl_max(sm[SM_yneg_max], sn[SN_dres_max]);
  e.l2_dual_residual = sqrt(sm[SM_yneg2] + sn[SN_dres2]);
  e.relative_l_inf_dual_residual = e.l_inf_dual_residual / (eps_ratio + h->cache[0]);
  e.relative_l2_dual_residual = e.l2_dual_residual / (eps_ratio + h->cache[2]);
  e.l_inf_dual_variable = sm[SM_y_max];
  e.l2_dual_variable = sqrt(sm[SM_y2]);
  e.corrected_dual_objective = (e.l_inf_dual_residual == 0.0) ? e.dual_objective : -INFINITY;
  {
    const double gap = fabs(e.primal_objective - e.dual_objective);
    const double abs_obj = fabs(e.primal_objective) + fabs(e.dual_objective);
    e.relative_optimality_gap = gap / (eps_ratio + abs_obj);
  }
  // compute_infeasibility_information, isu.jl:287-349
  {
    const double xs = sn[SN_x_max];
    const double inv = xs != 0.0 ? xs : 1.0;
    e.max_primal_ray_infeasibility =
        jl_max(jl_max(sm[SM_ray_act_max], sn[SN_ray_l_max]), sn[SN_ray_u_max]) / inv;
    e.primal_ray_linear_objective = sn[SN_cx] / inv;
    e.primal_ray_quadratic_norm = B.has_q ? sn[SN_qx_max] / inv : 0.0;  // isu.jl:311-313
    const double dobj = sm[SM_by] + sn[SN_ray_rcobj];
    const double sf = jl_max(sm[SM_y_max], sn[SN_ray_rc_max]);
    if (sf != 0.0) {
      e.max_dual_ray_infeasibility = jl_max(sm[SM_yneg_max], sn[SN_ray_dres_max]) / sf;
      e.dual_ray_objective = dobj / sf;
    }
  }
  e.step_size = s->step_size;
  e.primal_weight = s->primal_weight;
  e.time_spent_doing_basic_algorithm = h->basic_time;  // :929
  // define_norms, pdhg.jl:265-276
  const double wp = 1 / s->step_size * s->primal_weight;
  const double wd = 1 / s->step_size / s->primal_weight;
  BoundResult at_avg{0, 0, 0};
  {  // update_objective_bound_estimates, sp.jl:1015-1047
    const double rp_ = jl_max(1e-8, sqrt(wp * sn[SN_xs2]));
    const double rd_ = jl_max(1e-8, sqrt(wd * sm[SM_ys2]));
    const double L = ((0.5 * (B.has_q ? sn[SN_xs_qxs] : 0.0) + sn[SN_cs_x]) - sn[SN_x_aty]) +
                     sm[SM_bs_y] + c0;  // sp.jl:1109-1120
    TrProblem Pp{B.avg_x, B.aty_avg, B.avg_y, B.ax_avg, wp / (rp_ * rp_), wd / (rd_ * rd_), 1.0,
                 1, 0, 0, B.has_q ? B.qx_avg : nullptr};
    TrProblem Pd = Pp;
    Pd.use_primal = 0; Pd.use_dual = 1;
    TrState tp, td;
    if ((rc = tr_solve(h, Pp, &tp, 0))) return rc;
    if ((rc = tr_solve(h, Pd, &td, 1))) return rc;
    e.lagrangian_value = L;
    e.estimated_lower_bound = L + tp.v_primal;
    e.estimated_upper_bound = L - td.v_dual;
    at_avg.lagrangian_value = L;
  }
  int reason = check_termination(prm, h->cache, &e);  // :947
  if (s->numerical_error && reason == 0) reason = FOLP_TERMINATION_REASON_NUMERICAL_ERROR;
  e.termination_reason = reason;
  e.numerical_error = s->numerical_error;
  e.total_number_iterations = s->total_iterations;
  if (reason != 0) {  // :972-993
    h->terminated = 1;
    e.restart_used = FOLP_RESTART_CHOICE_UNSPECIFIED;
    if ((rc = push_state(h))) return rc;
    h->last_eval = e;
    *out = e;
    return FOLP_OK;
  }
  int choice = FOLP_RESTART_CHOICE_NO_RESTART;
  if ((rc = run_restart_scheme(h, iteration - 1, wp, wd, at_avg, &choice))) return rc;  // :995
  e.restart_used = choice;
  if (choice != FOLP_RESTART_CHOICE_NO_RESTART) {  // :1009-1017, sp.jl:862-891
    const double eps = 2.220446049250313e-16;
    if (h->pd_last > eps && h->dd_last > eps) {
      const double est = h->dd_last / h->pd_last;
      const double th = prm->primal_weight_update_smoothing;
      s->primal_weight = exp(th * log(est) + (1 - th) * log(s->primal_weight));
    }
    s->ratio_step_sizes = 1.0;
  }
  if ((rc = push_state(h))) return rc;
  h->need_step = 1;
  h->last_eval = e;
  *out = e;
  return FOLP_OK;
}

// iterations completed at the next evaluation, pdhg.jl:892-895
static int64_t next_evaluation(const folp_handle* h, int64_t k) {
  const int64_t freq = h->prm.termination_evaluation_frequency;
  int64_t t = (k / freq + 1) * freq;
  if (k + 1 <= 9) t = std::min(t, k + 1);
  const int64_t limit = h->prm.iteration_limit;
  if (limit > k) t = std::min(t, limit);
  return t;
}

extern "C" int folp_run(folp_handle* h, folp_eval* out) {
  if (!h || !out) return FOLP_INVALID_ARGUMENT;
  if (h->multi) {  // every rank returns the same global record; rank 0's goes to the caller
    std::vector<folp_eval> scratch(static_cast<size_t>(h->multi->world));
    const int rc = multi_dispatch(h, [&](folp_handle* q, int r) { return folp_run(q, r == 0 ? out : &scratch[r]); });
    if (!rc) h->terminated = out->termination_reason != 0;
    return rc;
  }
  if (h->terminated) {
    *out = h->last_eval;
    return FOLP_OK;
  }
  cudaSetDevice(h->device);
  int rc;
  if (h->need_step) {  // :1044, as many take_step calls as fit before the next evaluation
    const int64_t k = h->hs->iterations;
    if ((rc = run_steps(h, next_evaluation(h, k)))) return rc;
    h->need_step = 0;
  }
  h->iteration = h->hs->iterations + 1;  // :887
  return evaluate(h, out);
}

extern "C" int folp_get_solution(folp_handle* h, int which, int unscaled, double* x_out,
                                 double* y_out) {
  if (!h) return FOLP_INVALID_ARGUMENT;
  if (h->multi)
    return multi_dispatch(h, [&](folp_handle* q, int r) {
      return folp_get_solution(q, which, unscaled, r == 0 ? x_out : nullptr, r == 0 ? y_out : nullptr);
    });
  cudaSetDevice(h->device);
  DevState* s = h->hs;
  Bufs& B = h->B;
  const double *px, *py;
  if (which == 0) {  // pdhg.jl:902-910
    const int use_current = (s->numerical_error || s->count_x == 0 || s->count_y == 0) ? 1 : 0;
    if (s->pending_avg) {
      launch_flush_avg(B, h->stream);
      h->launches += 2;
      s->pending_avg = 0;
    }
    launch_make_avg(B, use_current, h->stream);
    h->launches += 1;
    px = B.avg_x; py = B.avg_y;
  } else {
    px = B.x[s->cur]; py = B.y[s->cur];
  }
  // sp.jl:65-67; tr_t / tr_d / col_tmp are free scratch outside a trust-region solve / attempt
  double* sx = h->world > 1 ? B.col_tmp : B.tr_t;
  double* sy = B.tr_d;
  launch_scale_div(px, unscaled ? B.D : nullptr, sx, B.n, B.grid_vec, h->stream);
  launch_scale_div(py, unscaled ? B.E : nullptr, sy, B.m, B.grid_vec, h->stream);
  CHECK_LAUNCH();
  h->launches += 2;
  int rc;
  if ((rc = fetch_cols(h, sx, x_out))) return rc;
  if ((rc = fetch_rows(h, sy, y_out))) return rc;
  TRY(cudaStreamSynchronize(h->stream));
  return FOLP_OK;
}

extern "C" int folp_solve(folp_handle* h, folp_eval* evals, int64_t max_evals, int64_t* num_evals,
                          int32_t* termination_reason, int32_t* iteration_count, double* x_out,
                          double* y_out) {
  if (!h) return FOLP_INVALID_ARGUMENT;
  int64_t cnt = 0;
  folp_eval e;
  for (;;) {
    int rc = folp_run(h, &e);
    if (rc) return rc;
    if (h->prm.record_iteration_stats || e.termination_reason != 0) {  // :958-960
      if (evals && cnt < max_evals) evals[cnt] = e;
      else if (evals && max_evals > 0) evals[max_evals - 1] = e;  // full: the last slot follows the latest record
      cnt += 1;
    }
    if (e.termination_reason != 0) break;
  }
  if (num_evals) *num_evals = cnt;  // records produced; more than max_evals = the history was truncated
  if (termination_reason) *termination_reason = e.termination_reason;
  if (iteration_count) *iteration_count = e.iteration_number;
  return folp_get_solution(h, 0, 1, x_out, y_out);
}

// ---------------------------------------------------------------------------
// test hooks
// ---------------------------------------------------------------------------
extern "C" int folp_debug_attempts(folp_handle* h, int64_t attempts) {
  if (!h) return FOLP_INVALID_ARGUMENT;
  if (h->multi) return multi_dispatch(h, [&](folp_handle* q, int) { return folp_debug_attempts(q, attempts); });
  cudaSetDevice(h->device);
  DevState* s = h->hs;
  int rc;
  if (s->policy == FOLP_STEP_ADAPTIVE) {
    s->target_iterations = INT64_MAX / 4;
    s->active = s->numerical_error ? 0 : 1;
    if ((rc = push_state(h))) return rc;
    TRY(cudaEventRecord(h->ev0, h->stream));
    while (attempts > 0) {
      const int chunk = static_cast<int>(std::min<int64_t>(attempts, 256));
      if ((rc = enqueue_attempts(h, chunk))) return rc;
      attempts -= chunk;
    }
    TRY(cudaEventRecord(h->ev1, h->stream));
    if ((rc = pull_state(h))) return rc;
    float ms = 0.f;
    TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->basic_time += 1e-3 * ms;
    return FOLP_OK;
  }
  return run_steps(h, s->iterations + attempts);
}

extern "C" int folp_debug_state(folp_handle* h, double* x, double* y, double* dual_product,
                                double* sum_x, double* sum_y, folp_debug_scalars* out) {
  if (!h) return FOLP_INVALID_ARGUMENT;
  if (h->multi)
    return multi_dispatch(h, [&](folp_handle* q, int r) {
      return r == 0 ? folp_debug_state(q, x, y, dual_product, sum_x, sum_y, out)
                    : folp_debug_state(q, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    });
  cudaSetDevice(h->device);
  DevState* s = h->hs;
  Bufs& B = h->B;
  if (s->pending_avg) {
    launch_flush_avg(B, h->stream);
    h->launches += 2;
    s->pending_avg = 0;
  }
  int rc;
  if ((rc = fetch_cols(h, B.x[s->cur], x))) return rc;
  if (h->world > 1) TRY(cudaStreamSynchronize(h->stream));  // the gather staging is reused below
  if ((rc = fetch_rows(h, B.y[s->cur], y))) return rc;
  if (h->world > 1) TRY(cudaStreamSynchronize(h->stream));
  if ((rc = fetch_cols(h, B.aty[s->cur], dual_product))) return rc;
  if (h->world > 1) TRY(cudaStreamSynchronize(h->stream));
  if ((rc = fetch_cols(h, B.sum_x, sum_x))) return rc;
  if (h->world > 1) TRY(cudaStreamSynchronize(h->stream));
  if ((rc = fetch_rows(h, B.sum_y, sum_y))) return rc;
  TRY(cudaStreamSynchronize(h->stream));
  if (out) {
    memset(out, 0, sizeof(*out));
    out->step_size = s->policy == FOLP_STEP_ADAPTIVE ? s->trial_step : s->step_size;
    out->primal_weight = s->primal_weight;
    out->cumulative_kkt_passes = s->kkt_passes;
    out->sum_primal_solution_weights = s->sum_w_x;
    out->sum_dual_solution_weights = s->sum_w_y;
    out->total_number_iterations = s->total_iterations;
    out->iterations_completed = s->count_x;
    out->sum_primal_solutions_count = s->count_x;
    out->sum_dual_solutions_count = s->count_y;
    out->numerical_error = s->numerical_error;
    out->last_interaction = s->last_interaction;
    out->last_movement = s->last_movement;
  }
  return FOLP_OK;
}

extern "C" int folp_debug_set_state(folp_handle* h, const double* x, const double* y,
                                    double step_size, double primal_weight) {
  if (!h) return FOLP_INVALID_ARGUMENT;
  if (h->multi)
    return multi_dispatch(h, [&](folp_handle* q, int) { return folp_debug_set_state(q, x, y, step_size, primal_weight); });
  cudaSetDevice(h->device);
  DevState* s = h->hs;
  Bufs& B = h->B;
  if (x && B.n) {
    const double* xs = x + h->col0;
    if (!h->new2old.empty()) {  // the caller's numbering -> the device's
      h->h_tmp.resize(static_cast<size_t>(B.n));
      for (int j = 0; j < B.n; ++j) h->h_tmp[j] = x[h->new2old[h->col0 + j]];
      xs = h->h_tmp.data();
    }
    TRY(cudaMemcpyAsync(B.x[s->cur], xs, sizeof(double) * B.n, cudaMemcpyHostToDevice, h->stream));
    if (!h->new2old.empty()) TRY(cudaStreamSynchronize(h->stream));
    if (B.has_q) {
      launch_spmv_plain(h->Q, B.x[s->cur], B.qx[s->cur], B.grid_spmv, h->stream);
      h->launches += 1;
    }
  }
  if (y) {
    if (B.m)
      TRY(cudaMemcpyAsync(B.y[s->cur], y + h->row0, sizeof(double) * B.m, cudaMemcpyHostToDevice,
                          h->stream));
    int rc2 = spmv_At(h, B.y[s->cur], B.aty[s->cur]);
    if (rc2) return rc2;
  }
  if (step_size > 0) s->step_size = s->trial_step = s->avg_weight = step_size;
  if (primal_weight > 0) s->primal_weight = primal_weight;
  int rc = push_state(h);
  if (rc) return rc;
  TRY(cudaStreamSynchronize(h->stream));
  return FOLP_OK;
}

extern "C" int folp_debug_spmv(folp_handle* h, int transpose, const double* in, double* out) {
  if (!h || !in || !out) return FOLP_INVALID_ARGUMENT;
  if (h->multi) {
    const size_t len = static_cast<size_t>(transpose ? h->n_glob : h->m_glob);
    std::vector<std::vector<double>> scratch(static_cast<size_t>(h->multi->world));
    return multi_dispatch(h, [&](folp_handle* q, int r) {
      if (r != 0) scratch[r].resize(len + 1);
      return folp_debug_spmv(q, transpose, in, r == 0 ? out : scratch[r].data());
    });
  }
  cudaSetDevice(h->device);
  Bufs& B = h->B;
  int rc;
  if (!transpose) {  // in: global n -> out: global m
    double* d_in = h->world > 1 ? h->d_cols : B.tr_t;
    const double* src = in;
    if (!h->new2old.empty()) {  // the caller's numbering -> the device's
      h->h_tmp.resize(static_cast<size_t>(h->n_glob));
      for (int64_t j = 0; j < h->n_glob; ++j) h->h_tmp[j] = in[h->new2old[j]];
      src = h->h_tmp.data();
    }
    if (h->n_glob) {
      TRY(cudaMemcpyAsync(d_in, src, sizeof(double) * h->n_glob, cudaMemcpyHostToDevice, h->stream));
      if (!h->new2old.empty()) TRY(cudaStreamSynchronize(h->stream));
    }
    launch_spmv_plain(h->A, d_in, B.tr_d, B.grid_spmv, h->stream);
    CHECK_LAUNCH();
    h->launches += 1;
    if ((rc = fetch_rows(h, B.tr_d, out))) return rc;
  } else {  // in: global m -> out: global n
    if (B.m)
      TRY(cudaMemcpyAsync(B.tr_t, in + h->row0, sizeof(double) * B.m, cudaMemcpyHostToDevice,
                          h->stream));
    double* d_out = h->world > 1 ? B.col_tmp : B.tr_d;
    if ((rc = spmv_At(h, B.tr_t, d_out))) return rc;
    if ((rc = fetch_cols(h, d_out, out))) return rc;
  }
  TRY(cudaStreamSynchronize(h->stream));
  return FOLP_OK;
}

extern "C" int folp_debug_profile_attempts(folp_handle* h, int64_t attempts, double ms_out[8],
                                           int64_t* attempts_run) {
  if (!h || !ms_out) return FOLP_INVALID_ARGUMENT;
  if (h->multi) {
    std::vector<double> scratch(static_cast<size_t>(8 * h->multi->world));
    return multi_dispatch(h, [&](folp_handle* q, int r) {
      return folp_debug_profile_attempts(q, attempts, r == 0 ? ms_out : &scratch[8 * r], r == 0 ? attempts_run : nullptr);
    });
  }
  cudaSetDevice(h->device);
  DevState* s = h->hs;
  int rc;
  s->target_iterations = INT64_MAX / 4;
  s->active = s->numerical_error ? 0 : 1;
  if ((rc = push_state(h))) return rc;
  if (h->take_grid > 0) {  // persistent kernel: per-phase device time from its own phase timers
    TRY(cudaMemsetAsync(h->d_timers, 0, sizeof(unsigned long long) * 16, h->stream));
    Bufs Bt = h->B;
    Bt.timers = h->d_timers;
    for (int64_t left = attempts; left > 0; left -= 256) {
      const int chunk = static_cast<int>(std::min<int64_t>(left, 256));
      const cudaError_t le = static_cast<cudaError_t>(
          h->take_cluster ? launch_take_steps_cluster(Bt, h->A, h->At, h->Q, chunk, h->take_grid, h->stream)
                          : launch_take_steps(Bt, h->A, h->At, h->Q, chunk, h->take_grid, h->stream));
      TRY(le);
      h->launches += 1;
    }
    if ((rc = pull_state(h))) return rc;
    unsigned long long t[8] = {0};
    TRY(cudaMemcpy(t, h->d_timers, sizeof(t), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 8; ++k) ms_out[k] = 0.0;
    for (int k = 0; k < 3; ++k) ms_out[k] = 1e-6 * static_cast<double>(t[1 + k]);
    if (attempts_run) *attempts_run = s->numerical_error ? -1 : static_cast<int64_t>(t[4]);
    return FOLP_OK;
  }
  const int nk = h->world == 1 ? 3 : 4;  // kernels (and, without peer memory, NCCL calls) per attempt
  std::vector<cudaEvent_t> ev(static_cast<size_t>((nk + 1) * attempts));
  for (auto& e : ev) TRY(cudaEventCreate(&e));
  for (int64_t a = 0; a < attempts; ++a) {
    cudaEvent_t* e = ev.data() + (nk + 1) * a;
    if (h->world == 1) {
      launch_step_attempt_timed(h->B, h->A, h->At, h->Q, e, h->stream);
      continue;
    }
    const Bufs& B = h->B;
    TRY(cudaEventRecord(e[0], h->stream));
    launch_dist_primal(B, h->stream);
    if (!B.p2p)
      NCCL_TRY(h->nccl->AllGather(B.xbar + static_cast<size_t>(h->rank) * h->n_pad, B.xbar,
                                  static_cast<size_t>(h->n_pad), ncclDouble, h->comm, h->stream));
    if (B.dbg & 4) launch_copy(B.xbar, h->d_cols, static_cast<int64_t>(h->world) * h->n_pad, h->stream);
    TRY(cudaEventRecord(e[1], h->stream));
    launch_dist_dual(B, h->A, h->stream);
    if (!B.p2p)
      NCCL_TRY(h->nccl->AllGather(B.y_full + static_cast<size_t>(h->rank) * h->m_pad, B.y_full,
                                  static_cast<size_t>(h->m_pad), ncclDouble, h->comm, h->stream));
    if (B.dbg & 4) launch_copy(B.y_full, h->d_rows, static_cast<int64_t>(h->world) * h->m_pad, h->stream);
    TRY(cudaEventRecord(e[2], h->stream));
    launch_dist_trans(B, h->A, h->At, h->stream);
    if (!B.p2p && (rc = exchange_scalars_dev(h))) return rc;
    TRY(cudaEventRecord(e[3], h->stream));
    launch_dist_finalize(B, h->stream);
    TRY(cudaEventRecord(e[4], h->stream));
  }
  CHECK_LAUNCH();
  h->launches += (nk + (h->B.has_q ? 2 : 0)) * attempts;
  if ((rc = pull_state(h))) return rc;
  for (int k = 0; k < 8; ++k) ms_out[k] = 0.0;
  for (int64_t a = 0; a < attempts; ++a)
    for (int k = 0; k < nk; ++k) {
      float ms = 0.f;
      TRY(cudaEventElapsedTime(&ms, ev[(nk + 1) * a + k], ev[(nk + 1) * a + k + 1]));
      ms_out[k] += ms;
    }
  for (auto& e : ev) cudaEventDestroy(e);
  // every attempt does work unless a numerical error stopped the batch early
  if (attempts_run) *attempts_run = s->numerical_error ? -1 : attempts;
  return FOLP_OK;
}

extern "C" int folp_debug_time_spmv(folp_handle* h, int transpose, int reps, double* ms_out) {
  if (!h || !ms_out || reps < 1) return FOLP_INVALID_ARGUMENT;
  if (h->multi) {
    std::vector<double> scratch(static_cast<size_t>(h->multi->world));
    return multi_dispatch(h, [&](folp_handle* q, int r) {
      return folp_debug_time_spmv(q, transpose, reps, r == 0 ? ms_out : &scratch[r]);
    });
  }
  cudaSetDevice(h->device);
  Bufs& B = h->B;
  // input: the live iterate (x or y); output: trust-region scratch
  const double* in = transpose ? B.y[h->hs->cur] : (h->world > 1 ? h->d_cols : B.x[h->hs->cur]);
  for (int w = 0; w < 3; ++w)
    launch_spmv_plain(transpose ? h->At : h->A, in, B.tr_d, B.grid_spmv, h->stream);
  TRY(cudaEventRecord(h->ev0, h->stream));
  for (int r = 0; r < reps; ++r)
    launch_spmv_plain(transpose ? h->At : h->A, in, B.tr_d, B.grid_spmv, h->stream);
  TRY(cudaEventRecord(h->ev1, h->stream));
  CHECK_LAUNCH();
  TRY(cudaEventSynchronize(h->ev1));
  h->launches += reps + 3;
  float ms = 0.f;
  TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  *ms_out = ms / reps;
  return FOLP_OK;
}

// Walks a packed matrix lane by lane with k_spmv's slot arithmetic (host, test hooks only).
static int emulate_packed_spmv(const PackedMatrix& pk, const IVec& rp, const int* ci, const double* v,
                               int64_t rows, const double* x, double* y, int64_t warps_total,
                               int64_t* stats) {
  int64_t n_sorted = 0, n_rounds = 0;
  const int64_t W = warps_total > 0 ? warps_total : 148 * kSpmvCtasPerSm * kSpmvWarps;
  std::vector<int64_t> warp_rounds(static_cast<size_t>(W), 0);
  int64_t item = 0;
  std::vector<double> partial(static_cast<size_t>(std::max(pk.nchunks_total, 1)), 0.0);
  for (int64_t i = 0; i < rows; ++i) y[i] = 0.0;  // empty rows are covered by narrow groups too
  for (const Tile& t : pk.tiles) {
    const int kind = t.rows_kind >> 16, cnt = t.rows_kind & 0xffff;
    if (kind == kTileThreadPerRow || kind == kTileThreadPerRowSorted) {
      n_sorted += kind == kTileThreadPerRowSorted;
      int row[32], len[32];
      double s[32];
      int maxlen = 0;
      for (int lane = 0; lane < 32; ++lane) {
        row[lane] = -1; len[lane] = 0; s[lane] = 0.0;
        if (lane < cnt) {
          row[lane] = kind == kTileThreadPerRowSorted ? pk.rowid[t.row_begin + lane] : t.row_begin + lane;
          len[lane] = rp[row[lane] + 1] - rp[row[lane]];
          maxlen = std::max(maxlen, len[lane]);
        }
      }
      int off = t.nnz_begin;
      for (int p = 0; p < maxlen; ++p) {
        unsigned m = 0;
        for (int lane = 0; lane < 32; ++lane)
          if (len[lane] > p) m |= 1u << lane;
        for (int lane = 0; lane < 32; ++lane)
          if (len[lane] > p) {
            const int k = off + __builtin_popcount(m & ((1u << lane) - 1u));
            s[lane] += v[k] * x[ci[k]];
          }
        off += __builtin_popcount(m);
      }
      if (off != t.nnz_end) return FOLP_INVALID_ARGUMENT;  // the group's range is exactly consumed
      n_rounds += (maxlen + FOLP_GATHER_UNROLL - 1) / FOLP_GATHER_UNROLL;
      warp_rounds[item % W] += (maxlen + FOLP_GATHER_UNROLL - 1) / FOLP_GATHER_UNROLL;
      for (int lane = 0; lane < cnt; ++lane) y[row[lane]] = s[lane];
    } else {
      warp_rounds[item % W] += (t.nnz_end - t.nnz_begin + 127) / 128;
      // four partial sums per lane, as the kernel: trips of 128 entries feed s0..s3, the rest s0
      double lane_sum[32];
      for (int lane = 0; lane < 32; ++lane) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int k = t.nnz_begin + lane;
        for (; k + 96 < t.nnz_end; k += 128) {
          s0 += v[k] * x[ci[k]];
          s1 += v[k + 32] * x[ci[k + 32]];
          s2 += v[k + 64] * x[ci[k + 64]];
          s3 += v[k + 96] * x[ci[k + 96]];
        }
        for (; k < t.nnz_end; k += 32) s0 += v[k] * x[ci[k]];
        lane_sum[lane] = (s0 + s1) + (s2 + s3);
      }
      auto tree = [](double* a) {  // warp_sum's xor-shuffle tree
        for (int o = 16; o > 0; o >>= 1)
          for (int l = 0; l < o; ++l) a[l] += a[l + o];  // lane l of the next level; same pairing as xor
        return a[0];
      };
      const double sum = tree(lane_sum);
      if (kind == kTileWarpPerRow) {
        y[t.row_begin] = sum;
      } else {
        partial[t.chunk_first + t.chunk_index] = sum;
        if (t.chunk_index == t.chunk_count - 1) {
          double ls[32] = {0.0};
          for (int q = 0; q < t.chunk_count; ++q) ls[q & 31] += partial[t.chunk_first + q];
          y[t.row_begin] = tree(ls);
        }
      }
    }
    item += 1;
  }
  if (stats) {
    stats[0] = static_cast<int64_t>(pk.tiles.size());
    stats[1] = n_sorted;
    stats[2] = n_rounds;
    stats[3] = pk.nlong;
    stats[4] = *std::max_element(warp_rounds.begin(), warp_rounds.end());  // the busiest warp's rounds
  }
  return FOLP_OK;
}

// Host emulation of k_spmv's traversal of a packed matrix (test hook, no CUDA call): packs the
// CSR matrix exactly as folp_create does and walks the work items lane by lane with the
// kernel's slot arithmetic (ballot + popcounts), so that the packing can be checked on a
// machine without a GPU. y = A * x; stats = {tiles, sorted groups, narrow rounds, long rows}.
extern "C" int folp_debug_host_spmv(int64_t rows, int64_t cols, const int64_t* rowptr,
                                    const int64_t* colidx, const double* vals, const double* x,
                                    double* y, int64_t warps_total, int64_t* stats) {
  if (rows < 0 || cols < 0 || !rowptr || !y || (rowptr[rows] > 0 && (!colidx || !vals || !x)))
    return FOLP_INVALID_ARGUMENT;
  const int64_t nnz = rowptr[rows];
  if (rows + cols >= (int64_t{1} << 31) - 64 || nnz >= (int64_t{1} << 31) - 64) return FOLP_UNSUPPORTED;
  IVec rp(static_cast<size_t>(rows) + 1), ci(static_cast<size_t>(nnz));
  DVec v(static_cast<size_t>(nnz));
  for (int64_t i = 0; i <= rows; ++i) rp[i] = static_cast<int>(rowptr[i]);
  for (int64_t k = 0; k < nnz; ++k)
    if (colidx[k] < 0 || colidx[k] >= cols) return FOLP_INVALID_ARGUMENT;
  PackedMatrix pk;
  plan_tiles(static_cast<int>(rows), rp, warps_total > 0 ? static_cast<int>(warps_total) : 148 * kSpmvCtasPerSm * kSpmvWarps, &pk);
  // packed straight from the caller's Int64 arrays, as folp_create does for A'
  fill_packed(pk, rp, [=](int k) { return static_cast<int>(colidx[k]); }, [=](int k) { return vals[k]; },
              ci.data(), v.data());
  return emulate_packed_spmv(pk, rp, ci.data(), v.data(), rows, x, y, warps_total, stats);
}

// Test hook without any CUDA call: runs the host half of folp_create on a folp_problem (transposition
// included) and evaluates y = A * x (transpose == 0, x of length n) or y = A' * x (x of length m)
// on the host through the packed layouts.
extern "C" int folp_debug_host_problem_spmv(const folp_problem* p, int transpose, const double* x,
                                            double* y) {
  if (!p || !x || !y) return FOLP_INVALID_ARGUMENT;
  const int64_t n = p->num_variables, m = p->num_constraints, nnz = p->num_nonzeros;
  IVec rp(static_cast<size_t>(n) + 1);
  for (int64_t j = 0; j <= n; ++j) rp[j] = nnz ? static_cast<int>(p->colptr[j] - p->index_base) : 0;
  HostMatrices hm;
  VarOrder vo;
  order_variables(n, rp, &vo);
  if (!prepare_host_matrices(p, rp, vo, 148 * kSpmvCtasPerSm * kSpmvWarps, &hm)) return FOLP_INVALID_ARGUMENT;
  if (vo.identity)
    return transpose ? emulate_packed_spmv(hm.pk_t, rp, hm.atc, hm.atv, n, x, y, 0, nullptr)
                     : emulate_packed_spmv(hm.pk_a, hm.rp2, hm.ac, hm.av, m, x, y, 0, nullptr);
  // the packed matrices live in the device numbering of the variables: x in / A'x out are renumbered here
  std::vector<double> tmp(static_cast<size_t>(n));
  if (transpose) {
    const int rc = emulate_packed_spmv(hm.pk_t, vo.rp_new, hm.atc, hm.atv, n, x, tmp.data(), 0, nullptr);
    for (int64_t j = 0; j < n; ++j) y[vo.new2old[j]] = tmp[j];
    return rc;
  }
  for (int64_t j = 0; j < n; ++j) tmp[j] = x[vo.new2old[j]];
  return emulate_packed_spmv(hm.pk_a, hm.rp2, hm.ac, hm.av, m, tmp.data(), y, 0, nullptr);
}

extern "C" void* folp_debug_stream(folp_handle* h) {
  if (h && h->multi) h = h->multi->sub[0];
  return h ? static_cast<void*>(h->stream) : nullptr;
}

extern "C" int folp_counters(folp_handle* h, int64_t* kernel_launches,
                             double* basic_algorithm_seconds, int64_t* iterations) {
  if (!h) return FOLP_INVALID_ARGUMENT;
  if (h->multi) {  // rank 0's clocks; launches summed over the devices
    int rc = folp_counters(h->multi->sub[0], kernel_launches, basic_algorithm_seconds, iterations);
    if (!rc && kernel_launches)
      for (int r = 1; r < h->multi->world; ++r) *kernel_launches += h->multi->sub[r]->launches;
    return rc;
  }
  if (kernel_launches) *kernel_launches = h->launches;
  if (basic_algorithm_seconds) *basic_algorithm_seconds = h->basic_time;
  if (iterations) *iterations = h->hs->iterations;
  return FOLP_OK;
}

extern "C" int folp_exchange_mode(folp_handle* h) {
  if (!h) return -1;
  if (h->multi) h = h->multi->sub[0];
  return h->world == 1 ? 0 : (h->B.p2p ? (h->B.xbar_mc ? 3 : 2) : 1);
}

extern "C" int folp_shard_info(folp_handle* h, int64_t* row_begin, int64_t* row_end,
                               int64_t* col_begin, int64_t* col_end, int64_t* local_nonzeros) {
  if (!h) return FOLP_INVALID_ARGUMENT;
  if (h->multi) h = h->multi->sub[0];  // rank 0's shard
  if (row_begin) *row_begin = h->row0;
  if (row_end) *row_end = h->row0 + h->m;
  if (col_begin) *col_begin = h->col0;
  if (col_end) *col_end = h->col0 + h->n;
  if (local_nonzeros) *local_nonzeros = h->nnz;
  return FOLP_OK;
}
