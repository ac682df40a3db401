// folp_internal.cuh -- shared declarations of libfolp_b200.so (sm_100a only).
//
// The library replaces the loop of FirstOrderLp.optimize(::PdhgParameters, qp)
// (reference: src/primal_dual_hybrid_gradient.jl:862-1048). Everything is fp64;
// the whole library is compiled with -fmad=false so that element-wise arithmetic
// rounds exactly like the (non-contracting) Julia reference; only the order of
// long reductions (norms, dots) differs, see DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/folp_b200.h"

namespace folp {

// ---------------------------------------------------------------------------
// Device-resident solver scalars: PdhgSolverState (pdhg.jl:205-258) minus the
// vectors, plus the bookkeeping that lets a whole batch of take_step attempts be
// enqueued without a host round trip.
// ---------------------------------------------------------------------------
struct DevState {
  double step_size;        // solver_state.step_size (value at entry of take_step)
  double trial_step;       // the local `step_size` of take_step (pdhg.jl:660)
  double primal_weight;
  double kkt_passes;       // cumulative_kkt_passes
  double avg_weight;       // weight of the next accepted iterate = step at entry (:512)
  double sum_w_x, sum_w_y; // SolutionWeightedAverage weights (sp.jl:215-222)
  double pending_w;        // weight of the accepted iterate not yet added to sum_x/sum_y
  double last_interaction, last_movement;
  double ratio_step_sizes; // Malitsky-Pock
  double mp_theta;         // extrapolation coefficient of the current attempt
  double mp_old_step;      // step_size at entry of the MP take_step
  long long count_x, count_y;
  long long total_iterations;   // total_number_iterations (counts rejected attempts)
  long long iterations;         // completed take_step calls
  long long target_iterations;  // attempts are no-ops once iterations == target
  int cur;                 // parity of the live x/y/aty buffers
  int numerical_error;
  int pending_avg;         // 1: x[cur], y[cur] still have to be added to the sums
  int active;              // 1: the next attempt does work
  int mp_need_primal;      // Malitsky-Pock: next attempt starts a new take_step
  int mp_retries;          // dual retries used in the current MP take_step
  int policy;              // folp_step_size_policy
  int p2p_timeout;         // row-partitioned peer exchange: a wait gave up (reported as an error)
  long long epoch;         // active attempts finalized so far (orders the peer-exchange flags)
  double reduction_exponent, growth_exponent;
  double downscaling_factor, breaking_factor, interpolation_coefficient;
};

// ---------------------------------------------------------------------------
// CSR matrix cut into warp-sized work items ("tiles"):
//   narrow group  up to 32 consecutive rows of <= kNarrowMax nonzeros each, one row per
//                 lane. Their nonzeros are stored POSITION-MAJOR inside the group's own
//                 range [rowptr[g0], rowptr[g0+rows)): first the 1st entry of every row
//                 that has one, then the 2nd entries, ... Each row keeps its ascending
//                 column order, and at every position the lanes of a warp read
//                 CONSECUTIVE values / column indices (the slot of lane l at position p is
//                 base + #entries of earlier positions + #lanes < l that reach p: a ballot
//                 and two popcounts) -- fully coalesced loads with no staging.
//   sorted group  the same, but its rows were picked from a window of kSortWindow consecutive
//                 narrow rows in order of decreasing length (rowid[slot] names the row of
//                 a slot), so that the 32 lanes of a warp run out of nonzeros together.
//                 Only windows whose rounds at least halve are sorted (a few long rows among
//                 many short ones); the others keep the identity order and never read
//                 rowid. A row's own nonzeros keep their ascending column order.
//   wide row      one row of 33 .. kChunkNnz nonzeros, plain CSR order, one warp.
//   long chunk    kChunkNnz nonzeros of a longer row; the partial sums are combined in
//                 chunk order by the last chunk to finish (deterministic).
// ---------------------------------------------------------------------------
#ifndef FOLP_CHUNK_NNZ
#define FOLP_CHUNK_NNZ 1024
#endif
#ifndef FOLP_GATHER_UNROLL
#define FOLP_GATHER_UNROLL 3
#endif
#ifndef FOLP_SPMV_CTAS_PER_SM
#define FOLP_SPMV_CTAS_PER_SM 4
#endif
constexpr int kChunkNnz = FOLP_CHUNK_NNZ;   // a longer row is split into chunks of this many nonzeros
constexpr int kSpmvThreads = 256;           // 8 warps per CTA, one work item per warp at a time
constexpr int kSpmvWarps = kSpmvThreads / 32;
constexpr int kSpmvCtasPerSm = FOLP_SPMV_CTAS_PER_SM;
constexpr int kNarrowMax = 32;              // rows up to this length are "narrow"
constexpr int kTilePad = 8;                 // slack behind the arrays
constexpr int kSortWindow = 256;            // rows per length-sorted window = the 8 groups a CTA works on together
enum TileKind : int { kTileThreadPerRow = 0, kTileWarpPerRow = 1, kTileLongChunk = 2,
                      kTileThreadPerRowSorted = 3 };

struct Tile {
  // hot half: one 16-byte load gives a role everything it needs for the common kinds
  int row_begin;           // first row (long chunk: the single row; sorted group: first slot of rowid)
  int nnz_begin, nnz_end;  // nonzeros [nnz_begin,nnz_end)
  int rows_kind;           // (kind << 16) | number of rows
  // long rows only
  int long_id;             // index of the long row (kTileLongChunk)
  int chunk_first;         // index into long_partials of this row's first chunk
  int chunk_count;         // chunks of this long row
  int chunk_index;         // which chunk of the long row this tile is
};
static_assert(sizeof(Tile) == 32, "two 16-byte halves");

struct SpmvMat {
  int rows = 0, cols = 0;
  int64_t nnz = 0;
  int* rowptr = nullptr;     // rows+1 (device)
  int2* rowid = nullptr;     // rows: slot -> (row, its length) inside length-sorted windows (nullptr: none)
  int* colidx = nullptr;     // nnz + pad
  double* vals = nullptr;    // nnz + pad
  Tile* tiles = nullptr;
  int ntiles = 0;
  int nlong = 0;
  double* long_partials = nullptr;   // one per long chunk
  unsigned* long_tickets = nullptr;  // one per long row
};

// per-kernel partial-reduction scratch
constexpr int kMaxPartialBlocks = 4096;
constexpr int kMaxScalars = 32;

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
#define FOLP_CUDA_TRY(h, expr)                                                    \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);              \
      return _e == cudaErrorMemoryAllocation ? FOLP_OUT_OF_MEMORY : FOLP_CUDA_ERROR; \
    }                                                                             \
  } while (0)

}  // namespace folp

#ifdef __CUDACC__
namespace folp {

// ---------------------------------------------------------------------------
// PTX wrappers: bulk asynchronous copies shared -> global (the TMA engine; SASS UBLKCP). Used by the
// partitioned primal step to push its tile of xbar into every peer's copy over NVLink: one elected
// thread hands whole 4 KB tiles to the copy engine instead of every thread issuing seven 16-byte
// remote stores (k_take_steps / k_primal, DIST).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// generic-proxy writes to shared memory become visible to the async proxy (the copy engine)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// dst: global (possibly a peer mapping), 16-byte aligned; src: shared, 16-byte aligned; bytes % 16 == 0
__device__ __forceinline__ void bulk_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's bulk groups may still be READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// all of this thread's bulk groups have completed (their writes are performed)
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  asm volatile("fence.proxy.async;" ::: "memory");
}

// ---------------------------------------------------------------------------
// programmatic dependent launch (PDL): the three kernels of an attempt are chained with
// cudaLaunchAttributeProgrammaticStreamSerialization, so that the CTAs of the next kernel are
// placed on SMs as the previous kernel's CTAs retire instead of after its last one has: the launch
// latency and the ramp of ~600 CTAs leave the critical path. pdl_wait() returns when the
// previous kernel has completed and its writes are visible -- nothing it produced (DevState
// included) may be read before; pdl_release() lets the next kernel start being placed. Both are
// no-ops for a kernel launched without the attribute.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------
// deterministic block reductions (fixed shape: shuffle tree, then warp 0)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// result valid in every thread; sh must hold 32 doubles
template <bool IS_MAX>
__device__ __forceinline__ double block_reduce(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  v = IS_MAX ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = (lane < nwarps) ? sh[lane] : (IS_MAX ? -INFINITY : 0.0);
  r = IS_MAX ? warp_max(r) : warp_sum(r);
  return r;
}

// "last block done" ticket. Every block calls it after writing its partials.
// Returns true in exactly one block (all threads), with the counter reset.
__device__ __forceinline__ bool last_block_arrive(unsigned* counter) {
  __shared__ int s_is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(counter, 1u);
    s_is_last = (t == gridDim.x - 1);
    if (s_is_last) *counter = 0;
  }
  __syncthreads();
  if (s_is_last) __threadfence();
  return s_is_last != 0;
}

// Sum (or max) `count` partials with the whole block in a fixed order.
template <bool IS_MAX>
__device__ __forceinline__ double reduce_partials(const double* p, int count, double* sh) {
  double v = IS_MAX ? -INFINITY : 0.0;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    double q = __ldcg(p + i);
    v = IS_MAX ? fmax(v, q) : v + q;
  }
  return block_reduce<IS_MAX>(v, sh);
}

}  // namespace folp
#endif  // __CUDACC__
