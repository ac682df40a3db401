// folp_spmv.cuh -- the fp64 CSR SpMV kernel template and its epilogues.
//
// One kernel serves A*xbar fused with the dual step (compute_dual_gradient +
// compute_next_dual_solution + project_dual!, sp.jl:1102-1107, pdhg.jl:472-494,
// sp.jl:110-117), A'*y+ fused with the interaction dot product
// (compute_interaction_and_movement, pdhg.jl:527-549) and the plain products of
// the evaluation block. A' is held as its own CSR (= the caller's CSC), so both
// products are row-gather kernels.
//
// What bounds it (measured, DESIGN.md section 5, tools/gather_bench.cu): every
// nonzero costs one random 8-byte gather of the input vector -- one slot of the
// SM's load/store pipe and one L2 sector -- and a B200 sustains ~0.8 of those per
// clock per SM (~0.5 when the vector no longer fits L2), which is less than what
// HBM could stream for this format. The design therefore spends as little else
// of the load/store pipe per nonzero as it can and keeps it saturated:
//   * one lane owns one row; thanks to the position-major group layout
//     (folp_internal.cuh) the 32 lanes of a warp read CONSECUTIVE values and
//     column indices at every position: 3 coalesced wavefronts per 32 nonzeros,
//     straight from global memory (streaming, evict-first) into registers. There is
//     no shared-memory staging: a TMA-fed shared-memory ring (v1-v3 of this kernel, git history:
//     commit 2090fdf and before) costs the same pipe slots to read back, needs a second pass
//     for the row sums and measured 15-40 % slower;
//   * each lane gathers, multiplies and adds in ascending column order in a
//     register and runs the fused epilogue on its row: no products written back;
//   * latency is hidden by occupancy (1024 threads per SM, kGatherUnroll independent
//     gathers per lane and round), not by a software pipeline. Measured on the
//     1e6 x 1e6 x 1e7 workload (K2 / K3 microseconds): unroll 1: 57.6 / 69.7, 2: 55.7 / 68.2,
//     3: 55.4 / 65.3, 4: 58.9 / 71.5, 5: 54.5 / 67.4, 6: 55.6 / 65.9; pipelined variants
//     (FOLP_PIPELINE) 2: 55.9 / 64.7, 3: 57.6 / 66.7, 5: 61.8 / 68.5; 5 CTAs per SM lose.
// Rows of up to 32 nonzeros are summed by one lane in ascending column order --
// the summation order of the reference's stdlib kernels -- so such rows are
// bit-identical to the CPU oracle; longer rows use a warp, rows longer than
// kChunkNnz several warps whose partial sums are combined in chunk order.
#pragma once
#include "folp_internal.cuh"

namespace folp {

#ifndef FOLP_PIPELINE
#define FOLP_PIPELINE 0
#endif
constexpr int kGatherUnroll = FOLP_GATHER_UNROLL;  // independent gathers per lane and round

// streaming loads of the matrix arrays: read once per product, keep them out of the way
__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
// per-row operands of the epilogues. FOLP_EPI_STREAM = 1 (default): streamed past L2 like the matrix;
// 0: default cache policy (development probe)
#ifndef FOLP_EPI_STREAM
#define FOLP_EPI_STREAM 1
#endif
__device__ __forceinline__ double ld_epi(const double* p) {
#if FOLP_EPI_STREAM
  return __ldcs(p);
#else
  return *p;
#endif
}

// hot half of a work-item descriptor
struct TileHot {
  int row_begin, nnz_begin, nnz_end, rows_kind;
  __device__ __forceinline__ int rows() const { return rows_kind & 0xffff; }
  __device__ __forceinline__ int kind() const { return rows_kind >> 16; }
};
__device__ __forceinline__ TileHot load_hot(const Tile* tiles, int idx) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(tiles + idx));
  TileHot h;
  h.row_begin = v.x; h.nnz_begin = v.y; h.nnz_end = v.z; h.rows_kind = v.w;
  return h;
}

// gather of the input vector. COH = false: read-only for the lifetime of the kernel (one product per
// launch), the non-coherent path. COH = true: the persistent take_step kernel, where the gathered
// vector is rewritten between grid barriers of the same launch -- a plain (coherent) load.
template <bool COH>
__device__ __forceinline__ double ld_gather(const double* p) {
  if (COH) return *p;
  return __ldg(p);
}

// The work items `first`, `first + stride`, ... of A, one per warp at a time; every row's sum goes to
// epi.row(). Called by all threads of the CTA.
template <class Epi, bool COH>
__device__ __forceinline__ void spmv_items(const SpmvMat& A, Epi& epi, int first, int stride) {
  const double* xin = epi.input();
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
#ifndef FOLP_TILE_PREFETCH
#define FOLP_TILE_PREFETCH 0
#endif
#if FOLP_TILE_PREFETCH
  // the descriptor of the NEXT work item is requested while this one is processed: one of the three
  // dependent latencies (descriptor -> row extents -> first entries) at the head of an item leaves its
  // critical path
  TileHot t_next;
  if (first < A.ntiles) t_next = load_hot(A.tiles, first);
#endif
  for (int item = first; item < A.ntiles; item += stride) {
#if FOLP_TILE_PREFETCH
    const TileHot t = t_next;
    if (item + stride < A.ntiles) t_next = load_hot(A.tiles, item + stride);
#else
    const TileHot t = load_hot(A.tiles, item);
#endif
    const int kind = t.kind();
    if (kind == kTileThreadPerRow || kind == kTileThreadPerRowSorted) {
      // ---- up to 32 narrow rows, one per lane ----
      const bool valid = lane < t.rows();
      int r = t.row_begin + lane;
      int len = 0;
      double in0 = 0.0, in1 = 0.0, in2 = 0.0;
      if (valid) {
        if (kind == kTileThreadPerRowSorted) {  // (row, length) of this lane's slot
          const int2 rl = __ldg(A.rowid + r);
          r = rl.x;
          len = rl.y;
        } else {
          len = __ldg(A.rowptr + r + 1) - __ldg(A.rowptr + r);
        }
        // per-row operands are read once: stream them past L2 so that the gathered vector stays
        if (Epi::kNumIn > 0) in0 = ld_epi(epi.in_ptr(0) + r);
        if (Epi::kNumIn > 1) in1 = ld_epi(epi.in_ptr(1) + r);
        if (Epi::kNumIn > 2) in2 = ld_epi(epi.in_ptr(2) + r);
      }
      epi.group_begin(t.row_begin, t.rows(), kind == kTileThreadPerRowSorted);  // all lanes
      const int maxlen = __reduce_max_sync(0xffffffffu, len);
      int off = t.nnz_begin;  // first entry of the current position
      double s = 0.0;
      // one round = kGatherUnroll positions: coalesced streaming loads of this lane's columns / values
      auto load_round = [&](int p0, int (&c)[kGatherUnroll], double (&a)[kGatherUnroll]) {
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u) {
          const bool act = len > p0 + u;
          const unsigned m = __ballot_sync(0xffffffffu, act);
          const int k = off + __popc(m & lt_mask);
          off += __popc(m);
          c[u] = act ? ld_stream(A.colidx + k) : -1;
          a[u] = act ? ld_stream(A.vals + k) : 0.0;
        }
      };
#if FOLP_PIPELINE
      // software pipeline: the next round's stream loads are issued before this round's gathers
      // are consumed, so a row costs one exposed stream latency instead of one per round
      int c[kGatherUnroll], cn[kGatherUnroll];
      double a[kGatherUnroll], an[kGatherUnroll], x[kGatherUnroll];
      if (maxlen > 0) load_round(0, c, a);
      for (int p0 = 0; p0 < maxlen; p0 += kGatherUnroll) {
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u) x[u] = c[u] >= 0 ? ld_gather<COH>(xin + c[u]) : 0.0;
        const bool more = p0 + kGatherUnroll < maxlen;  // warp-uniform
        if (more) load_round(p0 + kGatherUnroll, cn, an);
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u)
          if (c[u] >= 0) s += a[u] * x[u];  // ascending column order
        if (more) {
#pragma unroll
          for (int u = 0; u < kGatherUnroll; ++u) {
            c[u] = cn[u];
            a[u] = an[u];
          }
        }
      }
#else
      for (int p0 = 0; p0 < maxlen; p0 += kGatherUnroll) {
        int c[kGatherUnroll];
        double a[kGatherUnroll], x[kGatherUnroll];
        load_round(p0, c, a);
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u) x[u] = c[u] >= 0 ? ld_gather<COH>(xin + c[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u)
          if (c[u] >= 0) s += a[u] * x[u];  // ascending column order
      }
#endif
      if (valid) epi.row(r, s, in0, in1, in2);
      epi.group_end();  // all lanes
    } else {
      epi.group_begin(t.row_begin, 1, true);  // a single row handled by the whole warp: no group staging
      // ---- one warp on (a chunk of) one row, plain CSR order ----
      // Four independent loads / gathers / partial sums per lane and trip: a chunk is a serial
      // chain otherwise (measured on the PageRank LP's dense row: 128 dependent trips of ~0.4 us
      // each put 50 us on the kernel's critical path).
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int k = t.nnz_begin + lane;
      for (; k + 96 < t.nnz_end; k += 128) {
        const int c0 = ld_stream(A.colidx + k), c1 = ld_stream(A.colidx + k + 32);
        const int c2 = ld_stream(A.colidx + k + 64), c3 = ld_stream(A.colidx + k + 96);
        const double a0 = ld_stream(A.vals + k), a1 = ld_stream(A.vals + k + 32);
        const double a2 = ld_stream(A.vals + k + 64), a3 = ld_stream(A.vals + k + 96);
        const double x0 = ld_gather<COH>(xin + c0), x1 = ld_gather<COH>(xin + c1);
        const double x2 = ld_gather<COH>(xin + c2), x3 = ld_gather<COH>(xin + c3);
        s0 += a0 * x0;
        s1 += a1 * x1;
        s2 += a2 * x2;
        s3 += a3 * x3;
      }
      for (; k < t.nnz_end; k += 32) s0 += ld_stream(A.vals + k) * ld_gather<COH>(xin + ld_stream(A.colidx + k));
      double s = (s0 + s1) + (s2 + s3);
      s = warp_sum(s);
      const int r = t.row_begin;
      bool emit = kind == kTileWarpPerRow;
      if (kind == kTileLongChunk) {
        const Tile full = A.tiles[item];
        unsigned done = 0;
        if (lane == 0) {
          A.long_partials[full.chunk_first + full.chunk_index] = s;
          __threadfence();
          done = atomicAdd(A.long_tickets + full.long_id, 1u);
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        if (done == static_cast<unsigned>(full.chunk_count) - 1u) {  // last chunk: combine in order
          __threadfence();
          double v = 0.0;
          for (int q = lane; q < full.chunk_count; q += 32)
            v += __ldcg(A.long_partials + full.chunk_first + q);
          s = warp_sum(v);
          if (lane == 0) A.long_tickets[full.long_id] = 0u;
          emit = true;
        }
      }
      if (emit && lane == 0)
        epi.row(r, s, Epi::kNumIn > 0 ? epi.in_ptr(0)[r] : 0.0,
                Epi::kNumIn > 1 ? epi.in_ptr(1)[r] : 0.0, Epi::kNumIn > 2 ? epi.in_ptr(2)[r] : 0.0);
    }
  }
}

template <class Epi>
__global__ void __launch_bounds__(kSpmvThreads, kSpmvCtasPerSm) k_spmv(SpmvMat A, Epi epi) {
  __shared__ double s_red[32];
  pdl_wait();
  pdl_release();
  if (!epi.begin()) return;
  spmv_items<Epi, false>(A, epi, blockIdx.x * kSpmvWarps + (threadIdx.x >> 5), gridDim.x * kSpmvWarps);
  __syncthreads();
  epi.finish(s_red);
}

// ---- epilogues ---------------------------------------------------------------
// An epilogue names up to three per-row input vectors (in_ptr); row() receives
// their values for its row.

// out = A * in
struct EpiPlain {
  static constexpr int kNumIn = 0;
  const double* in;
  double* out;
  __device__ bool begin() { return true; }
  __device__ const double* input() const { return in; }
  __device__ const double* in_ptr(int) const { return nullptr; }
  __device__ void row(int r, double s, double, double, double) { out[r] = s; }
  __device__ void group_begin(int, int, bool) {}
  __device__ void group_end() {}
  __device__ void finish(double*) {}
};

}  // namespace folp
