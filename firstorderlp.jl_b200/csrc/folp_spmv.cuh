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
// the next tile's copy overlaps this tile's arithmetic. The input vector is
// gathered through L2 with ld.global.nc. Short rows are summed by one thread in
// ascending column order -- the summation order of the reference's stdlib
// kernels -- so such rows are bit-identical to the CPU oracle; rows longer than
// 32 nonzeros use a warp, rows longer than a tile use several CTAs.
#pragma once
#include "folp_internal.cuh"

namespace folp {

constexpr int kSpmvStageBytes = (kTileNnz + kTilePad) * 12;
constexpr int kSpmvSmemBytes = 2 * kSpmvStageBytes;

template <class Epi>
__global__ void __launch_bounds__(kSpmvThreads, 4) k_spmv(SpmvMat A, Epi epi) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t s_bar[2];
  __shared__ double s_red[32];
  __shared__ int s_flag;

  if (!epi.begin()) return;
  const double* __restrict__ xin = epi.input();

  auto stage_vals = [&](int st) {
    return reinterpret_cast<double*>(smem_raw + st * kSpmvStageBytes);
  };
  auto stage_cols = [&](int st) {
    return reinterpret_cast<int*>(smem_raw + st * kSpmvStageBytes + (kTileNnz + kTilePad) * 8);
  };

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kSpmvThreads / 32;

  uint64_t policy = 0;
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    fence_mbar_init();
    policy = policy_evict_first();
  }
  __syncthreads();

  auto issue = [&](int tile, int stage) {
    const Tile t = A.tiles[tile];
    const int kb = t.nnz_begin & ~3;
    const int ke = (t.nnz_end + 3) & ~3;
    const uint32_t cnt = static_cast<uint32_t>(ke - kb);
    mbar_expect_tx(&s_bar[stage], cnt * 12u);
    bulk_load(stage_vals(stage), A.vals + kb, cnt * 8u, &s_bar[stage], policy);
    bulk_load(stage_cols(stage), A.colidx + kb, cnt * 4u, &s_bar[stage], policy);
  };

  int tile = blockIdx.x;
  if (tid == 0 && tile < A.ntiles) issue(tile, 0);
  uint32_t phases = 0u;  // bit s = parity to wait for on stage s
  int stage = 0;

  for (; tile < A.ntiles; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    if (tid == 0 && next < A.ntiles) issue(next, stage ^ 1);
    const Tile t = A.tiles[tile];
    const int base = t.nnz_begin & ~3;
    mbar_wait(&s_bar[stage], (phases >> stage) & 1u);
    phases ^= 1u << stage;
    const double* __restrict__ sv = stage_vals(stage);
    const int* __restrict__ sc = stage_cols(stage);

    if (t.kind == kTileThreadPerRow) {
      for (int r = t.row_begin + tid; r < t.row_end; r += kSpmvThreads) {
        const int k0 = __ldg(A.rowptr + r) - base;
        const int k1 = __ldg(A.rowptr + r + 1) - base;
        double s = 0.0;
        int k = k0;
        for (; k + 4 <= k1; k += 4) {
          const double x0 = __ldg(xin + sc[k]);
          const double x1 = __ldg(xin + sc[k + 1]);
          const double x2 = __ldg(xin + sc[k + 2]);
          const double x3 = __ldg(xin + sc[k + 3]);
          s += sv[k] * x0;
          s += sv[k + 1] * x1;
          s += sv[k + 2] * x2;
          s += sv[k + 3] * x3;
        }
        for (; k < k1; ++k) s += sv[k] * __ldg(xin + sc[k]);
        epi.row(r, s);
      }
    } else if (t.kind == kTileWarpPerRow) {
      for (int r = t.row_begin + warp; r < t.row_end; r += kWarps) {
        const int k0 = __ldg(A.rowptr + r) - base;
        const int k1 = __ldg(A.rowptr + r + 1) - base;
        double s = 0.0;
        for (int k = k0 + lane; k < k1; k += 32) s += sv[k] * __ldg(xin + sc[k]);
        s = warp_sum(s);
        if (lane == 0) epi.row(r, s);
      }
    } else {  // one chunk of a row longer than a tile
      const int k0 = t.nnz_begin - base, k1 = t.nnz_end - base;
      double s = 0.0;
      for (int k = k0 + tid; k < k1; k += kSpmvThreads) s += sv[k] * __ldg(xin + sc[k]);
      s = block_reduce<false>(s, s_red);
      if (tid == 0) {
        A.long_partials[t.chunk_first + t.chunk_index] = s;
        __threadfence();
        const unsigned done = atomicAdd(A.long_tickets + t.long_id, 1u);
        s_flag = (done == static_cast<unsigned>(t.chunk_count) - 1u);
        if (s_flag) A.long_tickets[t.long_id] = 0u;
      }
      __syncthreads();
      if (s_flag) {  // last chunk of this row: combine in chunk order
        __threadfence();
        const double total =
            reduce_partials<false>(A.long_partials + t.chunk_first, t.chunk_count, s_red);
        if (tid == 0) epi.row(t.row_begin, total);
      }
    }
    __syncthreads();  // stage may be refilled by the next iteration's issue()
    stage ^= 1;
  }
  epi.finish(s_red);
}

// ---- epilogues ---------------------------------------------------------------

// out = A * in
struct EpiPlain {
  const double* in;
  double* out;
  __device__ bool begin() { return true; }
  __device__ const double* input() const { return in; }
  __device__ void row(int r, double s) { out[r] = s; }
  __device__ void finish(double*) {}
};

}  // namespace folp
