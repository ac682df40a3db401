// folp_spmv.cuh -- the tiled fp64 CSR SpMV kernel template and its epilogues.
//
// One kernel serves A*xbar fused with the dual step (compute_dual_gradient +
// compute_next_dual_solution + project_dual!, sp.jl:1102-1107, pdhg.jl:472-494,
// sp.jl:110-117), A'*y+ fused with the interaction dot product
// (compute_interaction_and_movement, pdhg.jl:527-549) and the plain products of
// the evaluation block. A' is held as its own CSR (= the caller's CSC), so both
// products are row-gather kernels.
//
// Data movement: a CTA walks its tiles; thread 0 streams each tile's values
// (fp64) and column indices (int32) into shared memory with two 1-D bulk async
// copies (TMA engine, mbarrier completion, L2 evict-first), double buffered so
// the next tile's copy overlaps this tile's arithmetic. Phase A turns the
// staged values into products in place: thread t owns nonzeros t, t+256, ...
// (conflict-free shared-memory reads) and gathers the input vector through L2
// with ld.global.nc, eight independent gathers in flight -- on B200 the random
// 8-byte gather (~0.9 per clock per SM, measured) bounds this kernel, not HBM.
// Phase B sums each row's products: short rows by one thread in ascending
// column order -- the summation order of the reference's stdlib kernels -- so
// such rows are bit-identical to the CPU oracle; rows longer than 32 nonzeros
// use a warp, rows longer than a tile use several CTAs.
#pragma once
#include "folp_internal.cuh"

namespace folp {

// ---- shared-memory stage layout (bytes; every offset a multiple of 16) --------
constexpr int kOffVals = 0;                                   // fp64 values -> products
constexpr int kOffCols = kOffVals + (kTileNnz + kTilePad) * 8;   // int32 column indices
constexpr int kOffRowp = kOffCols + (kTileNnz + kTilePad) * 4;   // int32 row pointers
constexpr int kOffIn = kOffRowp + (kTileRows + 8) * 4;           // up to 3 epilogue input vectors
constexpr int kInStride = (kTileRows + 4) * 8;
constexpr int kOffDesc = kOffIn + 3 * kInStride;                 // the Tile descriptor
constexpr int kSpmvStageBytes = kOffDesc + 64;
constexpr int kSpmvStages = FOLP_STAGES;
constexpr int kSpmvSmemBytes = kSpmvStages * kSpmvStageBytes;
static_assert(kSpmvStageBytes % 16 == 0 && kOffCols % 16 == 0 && kOffRowp % 16 == 0 &&
                  kOffIn % 16 == 0 && kInStride % 16 == 0 && kOffDesc % 16 == 0,
              "bulk copies need 16-byte aligned destinations");
static_assert(kSpmvSmemBytes <= 227 * 1024 - 1024, "stage ring must fit one SM");
constexpr int kGatherThreads = kGatherWarps * 32;
constexpr int kReduceThreads = kReduceWarps * 32;
static_assert(kSpmvThreads == 32 + kGatherThreads + kReduceThreads, "role split");
static_assert(kTileNnz % kGatherThreads == 0, "phase A unroll");
static_assert(kTileRows <= kReduceThreads, "one row per reduce thread");

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier among the reduce warps only (named barrier 1)
__device__ __forceinline__ void reduce_group_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(kReduceThreads) : "memory");
}
__device__ __forceinline__ void bulk_load_plain(void* dst_smem, const void* src_gmem,
                                                uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Warp-specialised persistent kernel, one CTA per SM. Roles:
//   warp 0         producer: one lane issues every global read of a tile as 1-D bulk
//                  async copies (TMA engine, mbarrier completion) into a 6-stage
//                  ring: values + column indices (L2 evict-first), the tile's row
//                  pointers and the epilogue's per-row input vectors
//   warps 1..16    gather: turn the staged values into products in place
//                  (conflict-free shared-memory reads, 4 independent L2 gathers in
//                  flight per thread, 2048 per tile)
//   warps 17..24   reduce: one row per thread, products summed in ascending column
//                  order, fused epilogue fed from shared memory
// The roles are connected by mbarriers only (full -> products -> empty): the
// gather path of the L1/LSU (~0.9 random 8-byte gathers per clock per SM,
// measured) is the binding resource and never waits for a row sum, an epilogue
// operand or a descriptor; nothing in the loop depends on a global-load latency
// except the gathers themselves.
template <class Epi>
__global__ void __launch_bounds__(kSpmvThreads, 1) k_spmv(SpmvMat A, Epi epi) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t s_full[kSpmvStages], s_prod[kSpmvStages], s_empty[kSpmvStages];
  __shared__ double s_red[32];
  __shared__ int s_flag;

  if (!epi.begin()) return;
  const double* __restrict__ xin = epi.input();

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < kSpmvStages; ++st) {
      mbar_init(&s_full[st], 1);
      mbar_init(&s_prod[st], kGatherWarps);
      mbar_init(&s_empty[st], kReduceWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int my_tiles = (A.ntiles > static_cast<int>(blockIdx.x))
                           ? (A.ntiles - static_cast<int>(blockIdx.x) + gridDim.x - 1) / gridDim.x
                           : 0;

  if (warp == 0) {
    // ---------------- producer ----------------
    if (lane == 0 && my_tiles > 0) {
      const uint64_t policy = policy_evict_first();
      Tile t = A.tiles[blockIdx.x];
      for (int i = 0; i < my_tiles; ++i) {
        const int st = i % kSpmvStages;
        unsigned char* sb = smem_raw + st * kSpmvStageBytes;
        Tile tn = t;
        if (i + 1 < my_tiles) tn = A.tiles[blockIdx.x + (i + 1) * gridDim.x];  // prefetch
        if (i >= kSpmvStages) mbar_wait(&s_empty[st], ((i / kSpmvStages) - 1) & 1);
        *reinterpret_cast<Tile*>(sb + kOffDesc) = t;
        const int kb = t.nnz_begin & ~3;
        const uint32_t cnt = static_cast<uint32_t>(((t.nnz_end + 3) & ~3) - kb);
        uint32_t bytes = cnt * 12u;
        uint32_t rp_cnt = 0, in_cnt = 0;
        const int rb4 = t.row_begin & ~3, rb2 = t.row_begin & ~1;
        if (t.kind != kTileLongChunk) {
          rp_cnt = static_cast<uint32_t>(((t.row_end + 1 + 3) & ~3) - rb4);
          in_cnt = static_cast<uint32_t>(((t.row_end + 1) & ~1) - rb2);
          bytes += rp_cnt * 4u + static_cast<uint32_t>(Epi::kNumIn) * in_cnt * 8u;
        }
        mbar_expect_tx(&s_full[st], bytes);
        bulk_load(sb + kOffVals, A.vals + kb, cnt * 8u, &s_full[st], policy);
        bulk_load(sb + kOffCols, A.colidx + kb, cnt * 4u, &s_full[st], policy);
        if (t.kind != kTileLongChunk) {
          bulk_load_plain(sb + kOffRowp, A.rowptr + rb4, rp_cnt * 4u, &s_full[st]);
#pragma unroll
          for (int v = 0; v < Epi::kNumIn; ++v)
            bulk_load_plain(sb + kOffIn + v * kInStride, epi.in_ptr(v) + rb2, in_cnt * 8u,
                            &s_full[st]);
        }
        t = tn;
      }
    }
  } else if (warp <= kGatherWarps) {
    // ---------------- gather ----------------
    // Software-pipelined across tiles: the gathers of tile i+1 are issued before the
    // products of tile i are formed, so every warp keeps kPer..2*kPer L2 gathers in
    // flight and the LSU never drains at a tile boundary.
    const int gt = tid - 32;
    constexpr int kPer = kTileNnz / kGatherThreads;
    double xv[kPer];
    int kb = 0, ke = 0;
    auto fetch = [&](int i) {
      const int st = i % kSpmvStages;
      unsigned char* sb = smem_raw + st * kSpmvStageBytes;
      mbar_wait(&s_full[st], (i / kSpmvStages) & 1);
      const Tile* td = reinterpret_cast<const Tile*>(sb + kOffDesc);
      const int nb = td->nnz_begin, ne = td->nnz_end;
      const int base = nb & ~3;
      kb = nb - base;
      ke = ne - base;
      const int* __restrict__ sc = reinterpret_cast<const int*>(sb + kOffCols);
      int c[kPer];
#pragma unroll
      for (int u = 0; u < kPer; ++u) {
        const int k = kb + gt + u * kGatherThreads;
        c[u] = k < ke ? sc[k] : -1;
      }
#pragma unroll
#ifdef FOLP_ABLATE_GATHER  // development: timing without the L2 gathers (wrong results)
      for (int u = 0; u < kPer; ++u) xv[u] = c[u] >= 0 ? 1.0 + c[u] : 0.0;
#else
      for (int u = 0; u < kPer; ++u) xv[u] = c[u] >= 0 ? __ldg(xin + c[u]) : 0.0;
#endif
    };
    if (my_tiles > 0) fetch(0);
    for (int i = 0; i < my_tiles; ++i) {
      const int st = i % kSpmvStages;
      double* __restrict__ sv = reinterpret_cast<double*>(smem_raw + st * kSpmvStageBytes + kOffVals);
      double xc[kPer];
#pragma unroll
      for (int u = 0; u < kPer; ++u) xc[u] = xv[u];
      const int kbc = kb, kec = ke;
      if (i + 1 < my_tiles) fetch(i + 1);
#pragma unroll
      for (int u = 0; u < kPer; ++u) {
        const int k = kbc + gt + u * kGatherThreads;
        if (k < kec) sv[k] = sv[k] * xc[u];
      }
      // products were written through the generic proxy; the stage is refilled by
      // the async proxy once the reduce warps release it
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_prod[st]);
    }
  } else {
    // ---------------- reduce ----------------
    const int rt = tid - 32 - kGatherThreads;
    const int rwarp = rt >> 5;
    for (int i = 0; i < my_tiles; ++i) {
      const int st = i % kSpmvStages;
      unsigned char* sb = smem_raw + st * kSpmvStageBytes;
      mbar_wait(&s_prod[st], (i / kSpmvStages) & 1);
#ifdef FOLP_ABLATE_REDUCE  // development: timing without row sums / epilogue (wrong results)
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);
      continue;
#endif
      const Tile t = *reinterpret_cast<const Tile*>(sb + kOffDesc);
      const int base = t.nnz_begin & ~3;
      const double* __restrict__ sv = reinterpret_cast<const double*>(sb + kOffVals);
      const int* __restrict__ rp = reinterpret_cast<const int*>(sb + kOffRowp);
      const double* __restrict__ in0 = reinterpret_cast<const double*>(sb + kOffIn);
      const double* __restrict__ in1 = reinterpret_cast<const double*>(sb + kOffIn + kInStride);
      const double* __restrict__ in2 =
          reinterpret_cast<const double*>(sb + kOffIn + 2 * kInStride);
      const int rb4 = t.row_begin & ~3, rb2 = t.row_begin & ~1;
      if (t.kind == kTileThreadPerRow) {
        const int r = t.row_begin + rt;
        if (r < t.row_end) {
          const int k0 = rp[r - rb4] - base, k1 = rp[r - rb4 + 1] - base;
          // ascending column order; the shared-memory loads are issued four at a time
          // so that only the additions are serialised
          double s = 0.0;
          int k = k0;
          for (; k + 4 <= k1; k += 4) {
            const double a0 = sv[k], a1 = sv[k + 1], a2 = sv[k + 2], a3 = sv[k + 3];
            s += a0;
            s += a1;
            s += a2;
            s += a3;
          }
          for (; k < k1; ++k) s += sv[k];
          const int li = r - rb2;
          epi.row(r, s, Epi::kNumIn > 0 ? in0[li] : 0.0, Epi::kNumIn > 1 ? in1[li] : 0.0,
                  Epi::kNumIn > 2 ? in2[li] : 0.0);
        }
      } else if (t.kind == kTileWarpPerRow) {
        for (int r = t.row_begin + rwarp; r < t.row_end; r += kReduceWarps) {
          const int k0 = rp[r - rb4] - base, k1 = rp[r - rb4 + 1] - base;
          double s = 0.0;
          for (int k = k0 + lane; k < k1; k += 32) s += sv[k];
          s = warp_sum(s);
          const int li = r - rb2;
          if (lane == 0)
            epi.row(r, s, Epi::kNumIn > 0 ? in0[li] : 0.0, Epi::kNumIn > 1 ? in1[li] : 0.0,
                    Epi::kNumIn > 2 ? in2[li] : 0.0);
        }
      } else {  // one chunk of a row longer than a tile
        const int k0 = t.nnz_begin - base, k1 = t.nnz_end - base;
        double s = 0.0;
        for (int k = k0 + rt; k < k1; k += kReduceThreads) s += sv[k];
        s = warp_sum(s);
        if (lane == 0) s_red[rwarp] = s;
        reduce_group_sync();
        if (rt == 0) {
          double tot = 0.0;
#pragma unroll
          for (int w = 0; w < kReduceWarps; ++w) tot += s_red[w];
          A.long_partials[t.chunk_first + t.chunk_index] = tot;
          __threadfence();
          const unsigned done = atomicAdd(A.long_tickets + t.long_id, 1u);
          s_flag = (done == static_cast<unsigned>(t.chunk_count) - 1u);
          if (s_flag) A.long_tickets[t.long_id] = 0u;
        }
        reduce_group_sync();
        if (s_flag) {  // last chunk of this row: combine in chunk order
          __threadfence();
          double v = 0.0;
          for (int q = rt; q < t.chunk_count; q += kReduceThreads)
            v += __ldcg(A.long_partials + t.chunk_first + q);
          v = warp_sum(v);
          if (lane == 0) s_red[rwarp] = v;
          reduce_group_sync();
          if (rt == 0) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < kReduceWarps; ++w) tot += s_red[w];
            const int r = t.row_begin;
            epi.row(r, tot, Epi::kNumIn > 0 ? epi.in_ptr(0)[r] : 0.0,
                    Epi::kNumIn > 1 ? epi.in_ptr(1)[r] : 0.0,
                    Epi::kNumIn > 2 ? epi.in_ptr(2)[r] : 0.0);
          }
        }
        reduce_group_sync();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);
    }
  }
  __syncthreads();
  epi.finish(s_red);
}

// ---- epilogues ---------------------------------------------------------------
// An epilogue names up to three per-row input vectors (in_ptr) that the producer
// stages next to the matrix tile; row() receives their values for its row.

// out = A * in
struct EpiPlain {
  static constexpr int kNumIn = 0;
  const double* in;
  double* out;
  __device__ bool begin() { return true; }
  __device__ const double* input() const { return in; }
  __device__ const double* in_ptr(int) const { return nullptr; }
  __device__ void row(int r, double s, double, double, double) { out[r] = s; }
  __device__ void finish(double*) {}
};

}  // namespace folp
