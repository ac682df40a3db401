// folp_kernels.cu -- device code of libfolp_b200.so (sm_100a, fp64, -fmad=false).
//
// Per take_step attempt (pdhg.jl:653-731) three kernels run back to back on one
// stream, all gated by DevState::active so that a batch can be enqueued blind:
//   k_primal           S1-S3 + extrapolation + deferred primal average + |dx|^2
//   k_spmv<EpiDual>    S4-S6: A*xbar, dual step, sign projection, deferred dual
//                      average, |dy|^2
//   k_spmv<EpiTrans>   A'*y+ (pdhg.jl:492), interaction dot (S7) and, in the last
//                      CTA to finish, the scalar accept/reject + step-size rule (S8)
// The evaluation block (pdhg.jl:892-1023) uses the kernels in the second half.
#include <cooperative_groups.h>
#include <math_constants.h>
#include <stdlib.h>

#include "folp_kernels.cuh"
#include "folp_spmv.cuh"

namespace folp {

constexpr int kVecThreads = 256;

__device__ __forceinline__ double* part_ptr(const Bufs& B, int slot, int scalar) {
  return B.part + (static_cast<size_t>(slot) * kMaxScalars + scalar) * kMaxPartialBlocks;
}

// ---- peer-exchange flags (row-partitioned mode over CUDA IPC memory) --------------------
// flag value of exchange `kind` in the attempt that starts at DevState::epoch
__device__ __forceinline__ unsigned long long p2p_value(const DevState& s, int kind) {
  return 3ull * static_cast<unsigned long long>(s.epoch) + kind + 1ull;
}
// One thread, after the data it publishes was fenced system-wide: raise this rank's flag of
// `kind` on every rank.
// (Loops over ranks run to the constant kMaxWorld and are unrolled: a run-time index into the
// pointer tables of the kernel-parameter struct would force the whole struct into local memory.)
__device__ __forceinline__ void p2p_signal(const Bufs& B, int kind, unsigned long long v) {
#pragma unroll
  for (int r = 0; r < kMaxWorld; ++r)
    if (r < B.world) {
      unsigned long long* f = B.flag_peer[r] + kind * kMaxWorld + B.rank;
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(v) : "memory");
    }
}
// The same after ONE system-scope fence issued by the caller: plain (relaxed) flag stores, so that
// the fence's wait for the acknowledgements of everything pushed is paid once, not once per peer.
__device__ __forceinline__ void p2p_signal_fenced(const Bufs& B, int kind, unsigned long long v) {
#pragma unroll
  for (int r = 0; r < kMaxWorld; ++r)
    if (r < B.world) {
      unsigned long long* f = B.flag_peer[r] + kind * kMaxWorld + B.rank;
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(f), "l"(v) : "memory");
    }
}
__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// One thread: wait until every rank's flag of `kind` has reached v. Gives up after
// Bufs::p2p_timeout_ns of wall clock (FOLP_P2P_TIMEOUT_MS, default 30 s: ranks are separate processes
// and do host work between batches) so that a lost peer surfaces as an error instead of a hung
// device; returns true (and raises the sticky flag the host turns into an error) if it gave up.
__device__ __forceinline__ bool p2p_wait(const Bufs& B, int kind, unsigned long long v) {
  unsigned long long t0 = 0;
  unsigned spins = 0;
  const unsigned long long* f = B.flags + kind * kMaxWorld;
  if (B.dbg & 16) {
    // development probe: all flags polled with relaxed loads, ONE acquire fence at the end. Measured
    // on 8 GPUs (1e6 x 1e6 x 1e7): 78.5 us per attempt against 74.3 us for the acquire loads below
    // -- the closing system-scope fence costs more than eight acquire loads of flags that are
    // already up.
    for (;;) {
      bool all = true;
      for (int r = 0; r < B.world; ++r) {
        unsigned long long cur;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(cur) : "l"(f + r) : "memory");
        if (cur < v) {
          all = false;
          break;
        }
      }
      if (all) break;
      if ((++spins & 255u) == 0u) {
        const unsigned long long now = gtimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > B.p2p_timeout_ns || __ldcg(B.counters + 6)) {
          atomicExch(B.counters + 6, 1u);
          return true;
        }
      }
      if (spins > 64u) __nanosleep(32);
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    return false;
  }
  for (int r = 0; r < B.world; ++r) {
    unsigned long long cur;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(cur) : "l"(f + r) : "memory");
      if (cur >= v) break;
      if ((++spins & 255u) == 0u) {
        const unsigned long long now = gtimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > B.p2p_timeout_ns || __ldcg(B.counters + 6)) {
          atomicExch(B.counters + 6, 1u);  // sticky; the host turns it into an error
          return true;
        }
      }
      if (spins > 64u) __nanosleep(32);
    }
  }
  return false;
}
// "Last block done" for data that peers will read. Every block orders its (remote) stores
// before its ticket with a gpu-scope fence; the last block, having observed every ticket, issues
// the ONE system-scope fence before the flag goes out (causality order is transitive across the
// two scopes). A system-scope fence per block is correct too, but those serialise device-wide
// (~20 ns each, 600-1200 blocks per kernel). True in all threads of exactly one block.
__device__ __forceinline__ bool last_block_arrive_sys(unsigned* counter) {
  const bool last = last_block_arrive(counter);
  if (last && threadIdx.x == 0) asm volatile("fence.acq_rel.sys;" ::: "memory");
  return last;
}
// All threads of a CTA: block until the exchange has arrived. False if the wait gave up (a lost
// peer): the caller returns without touching the half-delivered vector.
__device__ __forceinline__ bool p2p_wait_cta(const Bufs& B, int kind) {
  __shared__ int s_timed_out;
  if (B.dbg & 2) return true;
  if (threadIdx.x == 0) s_timed_out = p2p_wait(B, kind, p2p_value(*B.st, kind)) ? 1 : 0;
  __syncthreads();
  return s_timed_out == 0;
}

// Constant-index selection keeps the kernel parameter struct out of local memory.
template <class T>
__device__ __forceinline__ T* sel(T* const (&a)[2], int k) {
  return k ? a[1] : a[0];
}

// Trial step and extrapolation coefficient of the attempt about to run.
__device__ __forceinline__ void attempt_params(const DevState& s, double& trial, double& theta) {
  if (s.policy == FOLP_STEP_MALITSKY_POCK) {
    if (s.mp_need_primal) {  // pdhg.jl:580-584
      trial = s.step_size +
              s.interpolation_coefficient * (sqrt(1 + s.ratio_step_sizes) - 1) * s.step_size;
    } else {
      trial = s.trial_step;
    }
    theta = trial / s.step_size;  // pdhg.jl:592
  } else {
    trial = s.trial_step;
    theta = 1.0;
  }
}

// ---------------------------------------------------------------------------
// K1: compute_next_primal_solution (pdhg.jl:442-470) + xbar (pdhg.jl:486)
// ---------------------------------------------------------------------------
// One element of K1. Returns dx; writes x+ / xbar / sum_x through the references.
struct PrimalCtx {
  double f, theta, w, w_old;
  bool do_primal, pend, pend_old, has_q;
};
// qx = (Q * x)_j, read only when k.has_q
__device__ __forceinline__ double primal_elem(const PrimalCtx& k, double x, double& xn, double c,
                                              double at, double l, double u, double& sx,
                                              double& xbar, double qx = 0.0) {
  if (k.pend || k.pend_old) {  // deferred add_to_primal_solution_weighted_average (sp.jl:252-263)
    if (k.pend_old) sx += xn * k.w_old;  // xn still holds the previous iterate (pdhg.jl:621-627)
    if (k.pend) sx += x * k.w;
  }
  double xp;
  if (k.do_primal) {
    const double g = k.has_q ? (qx + c) - at : c - at;  // sp.jl:1093-1100
    xp = x - k.f * g;
    xp = fmin(u, fmax(l, xp));  // sp.jl:82-93
    xn = xp;
  } else {
    xp = xn;
  }
  const double d = xp - x;
  xbar = xp + k.theta * d;
  return d;
}

// The primal step of one attempt over the elements t0, t0 + stride, ... (pairs of elements per trip):
// returns this thread's share of |dx|^2. DIST: partitioned mode (world > 1), the slice is also
// pushed into every peer's copy of xbar; the single-GPU instantiation carries no exchange code.
// No __restrict__ / non-coherent loads on the iterates: the persistent kernel rewrites them between
// grid barriers of one launch.
// FOLP_K1_STREAM = 1 (default): K1's read-once operands bypass L2 residency (ld.cs / st.cs);
// 0: default cache policy (development probe: do the vectors stay L2-resident between kernels?)
#ifndef FOLP_K1_STREAM
#define FOLP_K1_STREAM 1
#endif
__device__ __forceinline__ double2 k1_ld(const double2* p) {
#if FOLP_K1_STREAM
  return __ldcs(p);
#else
  return *p;
#endif
}
// One store into the NVSwitch multicast range: the switch replicates it into every rank's region
// (PTX requires the multimem form on multicast addresses; SASS is a plain STG to the multicast mapping).
__device__ __forceinline__ void mc_store(double* p, double v) {
  asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// 16 bytes at once (p 16-byte aligned). multimem.st has no .v2.f64 form; a store does not interpret its
// operand, so the two doubles travel as the four 32-bit halves of a .v4.f32. Two 8-byte stores per lane
// instead would leave every 32-byte sector half-written per instruction (measured on 8 GPUs: the xbar
// phase of the 1e7 x 1e7 workload 151 us against 116 us for unicast double2 stores).
__device__ __forceinline__ void mc_store2(double* p, double2 v) {
  const unsigned long long a = static_cast<unsigned long long>(__double_as_longlong(v.x));
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v.y));
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p),
               "f"(__uint_as_float(static_cast<unsigned>(a))), "f"(__uint_as_float(static_cast<unsigned>(a >> 32))),
               "f"(__uint_as_float(static_cast<unsigned>(b))), "f"(__uint_as_float(static_cast<unsigned>(b >> 32)))
               : "memory");
}
// cta_first: first PAIR of elements of this CTA (blockIdx.x * blockDim.x); the CTA's threads take
// consecutive pairs, so a trip of the CTA covers one contiguous tile of 2 * blockDim.x elements.
// stage (DIST): 2 * blockDim.x double2 of shared memory. With peer memory the tile of xbar is staged
// there and ONE thread hands it to the copy engine once per peer (cp.async.bulk, double-buffered:
// the engine streams 4 KB tiles over NVLink while the CTA computes the next one) instead of every
// thread issuing world-1 remote 16-byte stores. OPT-IN (FOLP_BULK_PUSH=1): verified by the partitioned
// parity tests on 2 and 8 GPUs, but measured 1-2 % slower than the plain posted stores (8 GPUs, 1e6 x
// 1e6 x 1e7: 78.5 against 77.2 us per attempt; 1e7 x 1e7 x 1e8: 437 us either way -- both forms run
// into the same NVLink ingress bound).
template <bool DIST>
__device__ __forceinline__ double primal_range(const Bufs& B, const DevState& s, int cta_first, int stride,
                                               double2* stage) {
  double trial, theta;
  attempt_params(s, trial, theta);
  const bool mp = s.policy == FOLP_STEP_MALITSKY_POCK;
  PrimalCtx k;
  k.do_primal = !mp || s.mp_need_primal;
  k.f = (mp ? s.step_size : trial) / s.primal_weight;
  k.theta = theta;
  k.pend = s.pending_avg & 1;
  k.pend_old = (s.pending_avg & 2) != 0;
  k.w = s.pending_w;
  k.w_old = s.mp_old_step;
  k.has_q = B.has_q != 0;
  const int cur = s.cur;
  const double* xc = sel(B.x, cur);
  double* xn = sel(B.x, cur ^ 1);
  const double* at = sel(B.aty, cur);
  const double* qxc = sel(B.qx, cur);
  const bool avg = k.pend || k.pend_old;
  const bool rd_xn = k.pend_old || !k.do_primal;
  const bool push = DIST && B.p2p && !(B.dbg & 1);
  const bool bulk = push && stage != nullptr && (B.dbg & 8);
  double acc = 0.0;
  // two elements per thread and trip: every stream moves as 16-byte accesses
  const int n2 = B.n >> 1;
  int trip = 0;
  for (int base = cta_first; base < n2; base += stride, ++trip) {
    const int j = base + static_cast<int>(threadIdx.x);
    const bool act = j < n2;
    double2 xb = make_double2(0.0, 0.0);
    if (act) {
      const double2 x = reinterpret_cast<const double2*>(xc)[j];
      // read-once streams go past L2 (evict-first): at n >= 1e7 L2 is needed for the gathered vectors
      const double2 c = k1_ld(reinterpret_cast<const double2*>(B.c) + j);
      const double2 a = k1_ld(reinterpret_cast<const double2*>(at) + j);
      const double2 l = k1_ld(reinterpret_cast<const double2*>(B.l) + j);
      const double2 u = k1_ld(reinterpret_cast<const double2*>(B.u) + j);
      double2 sx = avg ? k1_ld(reinterpret_cast<const double2*>(B.sum_x) + j) : make_double2(0.0, 0.0);
      double2 xp = rd_xn ? reinterpret_cast<const double2*>(xn)[j] : make_double2(0.0, 0.0);
      const double2 q = k.has_q ? reinterpret_cast<const double2*>(qxc)[j] : make_double2(0.0, 0.0);
      const double d0 = primal_elem(k, x.x, xp.x, c.x, a.x, l.x, u.x, sx.x, xb.x, q.x);
      const double d1 = primal_elem(k, x.y, xp.y, c.y, a.y, l.y, u.y, sx.y, xb.y, q.y);
      if (k.has_q) reinterpret_cast<double2*>(B.dxv)[j] = make_double2(d0, d1);
      if (avg) __stcs(reinterpret_cast<double2*>(B.sum_x) + j, sx);
      if (k.do_primal) reinterpret_cast<double2*>(xn)[j] = xp;
      reinterpret_cast<double2*>(B.xbar + B.xbar_off)[j] = xb;
      acc += d0 * d0;
      acc += d1 * d1;
    }
    if (bulk) {
      double2* buf = stage + (trip & 1) * blockDim.x;
      if (trip >= 2) {  // the copies of two trips ago have finished reading this buffer
        if (threadIdx.x == 0) bulk_wait_read<1>();
        __syncthreads();
      }
      if (act) buf[threadIdx.x] = xb;
      fence_proxy_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        const int rest = n2 - base;
        const uint32_t bytes = static_cast<uint32_t>(rest < static_cast<int>(blockDim.x) ? rest : blockDim.x) * 16u;
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
          if (r < B.world && r != B.rank)
            bulk_store(reinterpret_cast<double2*>(B.xbar_peer[r] + B.xbar_off) + base, buf, bytes);
        bulk_commit();
      }
    } else if (push && act) {  // posted NVLink stores from every thread
      if (B.xbar_mc != nullptr) {  // one store each, replicated inside the switch
        mc_store2(B.xbar_mc + B.xbar_off + 2 * j, xb);
      } else {
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
          if (r < B.world && r != B.rank) reinterpret_cast<double2*>(B.xbar_peer[r] + B.xbar_off)[j] = xb;
      }
    }
  }
  if (bulk) {  // everything this CTA handed to the copy engine has been written
    if (threadIdx.x == 0) bulk_wait_all();
    __syncthreads();
  }
  if ((B.n & 1) && cta_first == 0 && threadIdx.x == 0) {
    const int j = B.n - 1;
    double sx = avg ? B.sum_x[j] : 0.0, xp = rd_xn ? xn[j] : 0.0, xb;
    const double d = primal_elem(k, xc[j], xp, B.c[j], at[j], B.l[j], B.u[j], sx, xb,
                                 k.has_q ? qxc[j] : 0.0);
    if (k.has_q) B.dxv[j] = d;
    if (avg) B.sum_x[j] = sx;
    if (k.do_primal) xn[j] = xp;
    B.xbar[B.xbar_off + j] = xb;
    if (DIST && B.p2p) {
      if (B.xbar_mc != nullptr) {
        mc_store(B.xbar_mc + B.xbar_off + j, xb);
      } else {
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
          if (r < B.world && r != B.rank) B.xbar_peer[r][B.xbar_off + j] = xb;
      }
    }
    acc += d * d;
  }
  return acc;
}

template <bool DIST>
__global__ void __launch_bounds__(kVecThreads) k_primal(Bufs B) {
  __shared__ double sh[32];
  pdl_wait();
  pdl_release();
  const DevState& s = *B.st;
  if (!s.active) return;
  __shared__ double2 s_stage[DIST ? 2 * kVecThreads : 1];
  const double acc = primal_range<DIST>(B, s, blockIdx.x * blockDim.x, gridDim.x * blockDim.x,
                                        DIST ? s_stage : nullptr);
  const double t = block_reduce<false>(acc, sh);
  if (threadIdx.x == 0) part_ptr(B, kSlotPrimal, 0)[blockIdx.x] = t;
  // the last block to finish announces this rank's slice on every rank
  if (DIST && B.p2p && last_block_arrive_sys(B.counters + 4) && threadIdx.x == 0)
    p2p_signal(B, 0, p2p_value(s, 0));
}

// ---------------------------------------------------------------------------
// scalar rule at the end of an attempt
// ---------------------------------------------------------------------------
// qd = dx' * Q * dx (0 for an LP)
__device__ __noinline__ void finalize_attempt(DevState* st, double dx2, double dy2, double inter,
                                              double dp2, double qd) {
  DevState s = *st;
  double trial, theta;
  attempt_params(s, trial, theta);
  s.pending_avg = 0;  // K1/K2 of this attempt have applied it
  bool accepted = false;
  if (s.policy == FOLP_STEP_ADAPTIVE) {  // pdhg.jl:653-731
    s.total_iterations += 1;
    const double ndx = sqrt(dx2), ndy = sqrt(dy2);
    const double movement =
        0.5 * s.primal_weight * (ndx * ndx) + (0.5 / s.primal_weight) * (ndy * ndy);
    const double interaction = fabs(inter) + fabs(0.5 * qd);  // pdhg.jl:536-544
    s.last_interaction = interaction;
    s.last_movement = movement;
    s.kkt_passes += 1;
    // movement == 0: the reference's numerical error (:691-695). A NaN movement or interaction (a
    // diverged or non-finite iterate) makes the reference's `step_size <= step_size_limit` false
    // for ever -- its take_step never returns; here it ends the solve as a numerical error too.
    if (movement == 0.0 || movement != movement || interaction != interaction) {
      s.numerical_error = 1;
      s.step_size = trial;  // :730
      s.iterations += 1;    // the failed take_step call still advances `iteration` (:887)
    } else {
      const double limit = interaction > 0 ? movement / interaction : CUDART_INF;
      accepted = trial <= limit;
      const double k1 = static_cast<double>(s.total_iterations + 1);
      const double first = (1 - pow(k1, -s.reduction_exponent)) * limit;
      const double second = (1 + pow(k1, -s.growth_exponent)) * trial;
      const double next = fmin(first, second);
      if (accepted) {
        s.pending_w = s.avg_weight;  // :512, weight = step size at entry
        s.step_size = next;
        s.avg_weight = next;
      }
      s.trial_step = next;
    }
  } else if (s.policy == FOLP_STEP_CONSTANT) {  // pdhg.jl:737-767
    s.kkt_passes += 1;
    accepted = true;
    s.pending_w = s.step_size;
  } else {  // Malitsky-Pock, pdhg.jl:555-647
    if (s.mp_need_primal) {
      s.kkt_passes += 0.5;
      s.mp_retries = 0;
    }
    s.mp_retries += 1;
    s.total_iterations += 1;
    s.kkt_passes += 0.5;
    const double ratio = theta;
    if (trial * sqrt(dp2) <= s.breaking_factor * sqrt(dy2)) {  // :615
      if (s.count_x == 0) {  // :621-627: the previous iterate enters the primal average
        s.mp_old_step = trial * ratio;
        s.pending_avg |= 2;
        s.count_x += 1;
        s.sum_w_x += s.mp_old_step;
      }
      accepted = true;
      s.pending_w = s.step_size;
      s.step_size = trial;
      s.ratio_step_sizes = ratio;
      s.mp_need_primal = 1;
    } else {
      s.trial_step = trial * s.downscaling_factor;
      s.mp_need_primal = 0;
      if (s.mp_retries >= 60) {  // :640-643
        s.numerical_error = 1;
        s.iterations += 1;
      }
    }
  }
  if (accepted) {  // update_solution_in_solver_state, pdhg.jl:500-519
    s.cur ^= 1;
    s.count_x += 1;
    s.count_y += 1;
    s.sum_w_x += s.pending_w;
    s.sum_w_y += s.pending_w;
    s.pending_avg |= 1;
    s.iterations += 1;
  }
  s.epoch += 1;
  if (s.p2p_timeout) s.numerical_error = 1;  // stop the batch; the host reports the error
  s.active = (s.iterations < s.target_iterations && !s.numerical_error) ? 1 : 0;
  *st = s;
}

// ---------------------------------------------------------------------------
// K2 epilogue: dual step on row i given (A*xbar)_i
// ---------------------------------------------------------------------------
// BR = Bufs (an epilogue passed by value as a kernel parameter) or const Bufs& (built inside the
// persistent kernel around its own parameter). setup() / publish() are the halves of begin() /
// finish() that the persistent kernel uses: it holds the state in shared memory and closes an
// attempt behind a grid barrier instead of a last-block ticket.
template <bool DIST, class BR = Bufs>
struct EpiDualT {
  static constexpr int kNumIn = 3;  // y, b, sum_y
  BR B;
  const double* yc;
  double* yn;
  double f, w, acc;
  bool pend;
  // partitioned mode over peer memory, persistent kernel: the 32 new dual values of a group of
  // consecutive rows are staged in this warp's slot of shared memory (stage: 2 x 32 doubles per warp,
  // nullptr = off) and lane 0 hands the 256 bytes to the copy engine once per peer (cp.async.bulk)
  // instead of every lane issuing world-1 remote stores
  double* stage = nullptr;
  int g_row0 = 0, g_rows = 0, g_item = 0;
  bool g_bulk = false;
  __device__ void group_begin(int row0, int rows, bool sorted) {
    if (!DIST) return;
    g_row0 = row0;
    g_rows = rows;
    // a contiguous run of an even number of rows starting on a 16-byte boundary of y_full
    g_bulk = stage != nullptr && B.p2p && (B.dbg & 9) == 8 && !sorted && (rows & 1) == 0 &&
             ((static_cast<size_t>(B.rank) * B.m_pad + row0) & 1) == 0;
    if (g_bulk && (threadIdx.x & 31) == 0 && g_item >= 2) bulk_wait_read<1>();  // this buffer's last copies have read it
    __syncwarp();
  }
  __device__ void group_end() {
    if (!DIST || !g_bulk) return;
    fence_proxy_async_smem();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
      const double* buf = stage + ((threadIdx.x >> 5) * 2 + (g_item & 1)) * 32;
      const size_t at = static_cast<size_t>(B.rank) * B.m_pad + g_row0;
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r)
        if (r < B.world && r != B.rank) bulk_store(B.yfull_peer[r] + at, buf, static_cast<uint32_t>(g_rows) * 8u);
      bulk_commit();
    }
    g_item += 1;
  }
  __device__ void setup(const DevState& s) {
    double trial, theta;
    attempt_params(s, trial, theta);
    f = s.primal_weight * trial;  // pdhg.jl:488
    yc = sel(B.y, s.cur);
    yn = sel(B.y, s.cur ^ 1);
    pend = s.pending_avg & 1;
    w = s.pending_w;
    acc = 0.0;
  }
  __device__ bool begin() {
    const DevState& s = *B.st;
    if (!s.active) return false;
    setup(s);
    if (DIST && B.p2p) return p2p_wait_cta(B, 0);  // every rank's slice of xbar has landed
    return true;
  }
  __device__ const double* input() const { return B.xbar_priv ? B.xbar_priv : B.xbar; }
  __device__ const double* in_ptr(int v) const { return v == 0 ? yc : (v == 1 ? B.b : B.sum_y); }
  // (rows handled by a whole warp -- wide rows, long-row chunks -- come here from lane 0 with g_bulk
  // as the last narrow group left it: spmv_items clears it through group_begin before such an item)
  __device__ void row(int i, double ax, double yv, double bi, double sy) {
    if (pend) __stcs(B.sum_y + i, sy + yv * w);  // deferred add_to_dual_solution_weighted_average
    const double g = bi - ax;        // compute_dual_gradient, sp.jl:1102-1107
    double yp = yv + f * g;
    if (i >= B.neq) yp = fmax(yp, 0.0);  // project_dual!, sp.jl:110-117
    yn[i] = yp;
    if (DIST) {  // the transposed product of every rank gathers the full new dual iterate
      const size_t at = static_cast<size_t>(B.rank) * B.m_pad + i;
      B.y_full[at] = yp;
      if (g_bulk) {
        stage[((threadIdx.x >> 5) * 2 + (g_item & 1)) * 32 + (i - g_row0)] = yp;
      } else if (B.p2p && !(B.dbg & 1)) {
        if (B.yfull_mc != nullptr) {
          mc_store(B.yfull_mc + at, yp);
        } else {
#pragma unroll
          for (int r = 0; r < kMaxWorld; ++r)
            if (r < B.world && r != B.rank) B.yfull_peer[r][at] = yp;
        }
      }
    }
    const double d = yp - yv;
    acc += d * d;
  }
  __device__ void publish(double* sh) {
    if (DIST && stage != nullptr && (threadIdx.x & 31) == 0) bulk_wait_all();  // this warp's copies have been written
    const double t = block_reduce<false>(acc, sh);
    if (threadIdx.x == 0) part_ptr(B, kSlotDual, 0)[blockIdx.x] = t;
  }
  __device__ void finish(double* sh) {
    publish(sh);
    if (DIST && B.p2p && last_block_arrive_sys(B.counters + 5) && threadIdx.x == 0)
      p2p_signal(B, 1, p2p_value(*B.st, 1));
  }
};

// The five totals of an attempt, side by side: warp w < 5 of the calling CTA sums total w over the
// per-block partials (lanes stride the blocks in a fixed order, then a shuffle tree) into
// sh[16 + w]: |dx|^2, |dy|^2, dx . dA'y, |dA'y|^2, dx' Q dx. Ends with a barrier.
__device__ __forceinline__ void attempt_totals(const Bufs& B, int g_primal, int g_dual, int g_trans, int g_q,
                                               double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < 5) {
    const double* src = warp == 0 ? part_ptr(B, kSlotPrimal, 0)
                      : warp == 1 ? part_ptr(B, kSlotDual, 0)
                      : warp == 2 ? part_ptr(B, kSlotTrans, 0)
                      : warp == 3 ? part_ptr(B, kSlotTrans, 1) : part_ptr(B, kSlotPrimal, 1);
    const int count = warp == 0 ? g_primal : warp == 1 ? g_dual : warp == 4 ? g_q : g_trans;
    double t = 0.0;
    for (int j = lane; j < count; j += 32) t += __ldcg(src + j);
    t = warp_sum(t);
    if (lane == 0) sh[16 + warp] = t;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// K3 epilogue: (A'*y+)_j, interaction dot, and the scalar rule in the last CTA
// ---------------------------------------------------------------------------
template <bool DIST, class BR = Bufs>
struct EpiTransT {
  static constexpr int kNumIn = 3;  // x, x+, A'y
  BR B;
  int g_primal, g_dual;  // grids of K1 and K2 (number of partials they wrote)
  int g_q = 0;           // grid of the dx' Q dx kernel (QP only), 0 for an LP
  const double *xc, *xn, *atc, *yin;
  double* atn;
  double inter, dp2;
  __device__ void setup(const DevState& s) {
    xc = sel(B.x, s.cur);
    xn = sel(B.x, s.cur ^ 1);
    atc = sel(B.aty, s.cur);
    atn = sel(B.aty, s.cur ^ 1);
    yin = DIST ? (B.yfull_priv ? B.yfull_priv : B.y_full) : sel(B.y, s.cur ^ 1);
    inter = 0.0;
    dp2 = 0.0;
  }
  __device__ bool begin() {
    const DevState& s = *B.st;
    if (!s.active) return false;
    setup(s);
    if (DIST && B.p2p) return p2p_wait_cta(B, 1);  // every rank's rows of y+ have landed
    return true;
  }
  __device__ const double* input() const { return yin; }
  __device__ void group_begin(int, int, bool) {}
  __device__ void group_end() {}
  __device__ const double* in_ptr(int v) const { return v == 0 ? xc : (v == 1 ? xn : atc); }
  __device__ void row(int j, double at, double xcj, double xnj, double atcj) {
    atn[j] = at;
    const double dx = xnj - xcj;
    const double dat = at - atcj;
    inter += dx * dat;  // pdhg.jl:542-544
    dp2 += dat * dat;   // pdhg.jl:615 (Malitsky-Pock)
  }
  // sh: 32 doubles. One barrier for both sums: warp leaders park them, threads 0/1 combine.
  __device__ void publish(double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double wa = warp_sum(inter), wb = warp_sum(dp2);
    if (lane == 0) {
      sh[warp] = wa;
      sh[kSpmvWarps + warp] = wb;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      double t = sh[threadIdx.x * kSpmvWarps];
      for (int w = 1; w < kSpmvWarps; ++w) t += sh[threadIdx.x * kSpmvWarps + w];
      part_ptr(B, kSlotTrans, threadIdx.x)[blockIdx.x] = t;
    }
  }
  __device__ void finish(double* sh) {
    publish(sh);
    if (!last_block_arrive(B.counters + kSlotTrans)) return;
    // The last CTA closes the attempt; this tail is serial time of every iteration.
    attempt_totals(B, g_primal, g_dual, static_cast<int>(gridDim.x), g_q, sh);
    const double dx2 = sh[16], dy2 = sh[17], it = sh[18], dp = sh[19], qd = sh[20];
    if (threadIdx.x != 0) return;
    if (!DIST) {
      finalize_attempt(B.st, dx2, dy2, it, dp, qd);
    } else if (B.p2p) {  // push the four scalars into every rank's slot for this rank, then announce
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r)
        if (r < B.world) {
          double* slot = B.sc_peer[r] + B.rank * kScBlock;
          slot[0] = dx2;
          slot[1] = dy2;
          slot[2] = it;
          slot[3] = dp;
        }
      __threadfence_system();
      p2p_signal(B, 2, p2p_value(*B.st, 2));
    } else {
      B.sc_send[0] = dx2;
      B.sc_send[1] = dy2;
      B.sc_send[2] = it;
      B.sc_send[3] = dp;
    }
  }
};

// ---------------------------------------------------------------------------
// partitioned take_step (world > 1). Rank r owns a block of rows of A (CSR, for A*xbar and
// the dual step) AND a slice of columns of A (CSR of A[:, slice]', for A'*y+ and the primal
// step), so both products have full-length rows, need no cross-rank reduction and are summed
// in the same order as on one GPU. Same arithmetic as the single-GPU kernels, with the
// exchanges where a product needs the other side's full vector:
//   k_primal (slice, pushes xbar)  -> [xbar on every rank] -> k_spmv<EpiDual> (local rows,
//   pushes y+) -> [y+ on every rank] -> k_spmv<EpiTrans> (local slice, interaction) ->
//   [{|dx|^2, |dy|^2, dx.dA'y, |dA'y|^2} of every rank] -> k_finalize_dist
// Every rank sums the gathered scalars in rank order, so all ranks take the same
// accept/reject decision and step size bit for bit.
// ---------------------------------------------------------------------------
__global__ void k_finalize_dist(Bufs B) {
  if (threadIdx.x != 0 || !B.st->active) return;
  if (B.p2p) {
    p2p_wait(B, 2, p2p_value(*B.st, 2));
    if (__ldcg(B.counters + 6)) B.st->p2p_timeout = 1;  // stops the batch, see finalize_attempt
  }
  double t[4] = {0.0, 0.0, 0.0, 0.0};
  for (int r = 0; r < B.world; ++r)
    for (int k = 0; k < 4; ++k) t[k] += __ldcg(B.sc_recv + r * kScBlock + k);
  finalize_attempt(B.st, t[0], t[1], t[2], t[3], 0.0);  // partitioned mode is LP only
}

// ---------------------------------------------------------------------------
// quadratic objective: two products with Q (CSR) between K2 and K3 of an attempt.
//   EpiQx    qx[next] = Q * x+          the Q*x term of the NEXT primal gradient (sp.jl:1093-1100)
//   EpiQDot  sum_j dx_j (Q * dx)_j      the objective part of the interaction (pdhg.jl:536-541);
//            computed from dx itself, not as a difference of products (no cancellation)
// ---------------------------------------------------------------------------
template <class BR = Bufs>
struct EpiQxT {
  static constexpr int kNumIn = 0;
  BR B;
  const double* in;
  double* out;
  __device__ void setup(const DevState& s) {
    in = sel(B.x, s.cur ^ 1);
    out = sel(B.qx, s.cur ^ 1);
  }
  __device__ bool begin() {
    const DevState& s = *B.st;
    if (!s.active) return false;
    setup(s);
    return true;
  }
  __device__ const double* input() const { return in; }
  __device__ void group_begin(int, int, bool) {}
  __device__ void group_end() {}
  __device__ const double* in_ptr(int) const { return nullptr; }
  __device__ void row(int j, double s, double, double, double) { out[j] = s; }
  __device__ void publish(double*) {}
  __device__ void finish(double*) {}
};
template <class BR = Bufs>
struct EpiQDotT {
  static constexpr int kNumIn = 1;  // dx
  BR B;
  double acc;
  __device__ void setup(const DevState&) { acc = 0.0; }
  __device__ bool begin() {
    acc = 0.0;
    return B.st->active != 0;
  }
  __device__ const double* input() const { return B.dxv; }
  __device__ void group_begin(int, int, bool) {}
  __device__ void group_end() {}
  __device__ const double* in_ptr(int) const { return B.dxv; }
  __device__ void row(int, double s, double dxj, double, double) { acc += dxj * s; }
  __device__ void publish(double* sh) {
    const double t = block_reduce<false>(acc, sh);
    if (threadIdx.x == 0) part_ptr(B, kSlotPrimal, 1)[blockIdx.x] = t;
  }
  __device__ void finish(double* sh) { publish(sh); }
};

using EpiDual = EpiDualT<false>;
using EpiTrans = EpiTransT<false>;
using EpiDualDist = EpiDualT<true>;
using EpiTransDist = EpiTransT<true>;
using EpiQx = EpiQxT<>;
using EpiQDot = EpiQDotT<>;

// ---------------------------------------------------------------------------
// k_take_steps: a whole batch of take_step attempts as ONE cooperative launch.
//
// The three kernels of an attempt become three phases of a persistent grid (one CTA set resident
// for the whole batch, static striding over the same work items as k_spmv), separated by grid
// barriers instead of kernel boundaries: no launch ramp / drain between phases (ncu on the
// per-kernel form: 7 K of k_primal's 27 K cycles, 13 K of 109 K and 22 K of 127 K for the two
// products are spent with SMs waiting to be filled or emptied), the solver scalars live in shared
// memory (every CTA holds a bit-identical copy, refreshed once per attempt), and A'y+ / x+ written
// by one phase are still in L2 when the next one reads them.
//
// The barrier is two-level (groups of kBarGroupSize CTAs arrive on their own counter, the last of a
// group on the top counter) and runs a callback in the LAST CTA to arrive before it releases the
// others. That callback is where an attempt is closed -- the five totals are formed from the
// per-CTA partials in a fixed order and the scalar rule runs -- and where, in partitioned mode, the
// peer exchanges ride: one thread of one CTA per GPU fences system-wide, raises this rank's flag
// on every rank and waits for the peers' flags, instead of every CTA polling eight system-scope
// flags at the head of the next kernel and a one-thread kernel for the step rule (DESIGN.md
// section 6). A wait that gives up (lost peer) aborts the whole grid through the release word.
// ---------------------------------------------------------------------------
constexpr unsigned long long kBarAbortBit = 1ull << 63;
constexpr unsigned long long kBarTimeoutNs = 20000000000ull;  // co-resident CTAs: only a bug gets here

__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// Grid-wide barrier of a cooperative launch with a callback in the LAST CTA to arrive.
// All threads of all CTAs call it. `gen` = generation of the last completed barrier (same in every
// thread). last_fn() runs in every thread of the last CTA to arrive, after all other CTAs' writes are
// visible to it, and returns (uniformly) whether the grid must abort. s_flag: one shared int.
// Returns true in every thread of every CTA if the grid aborts.
//   arrival  ONE acq_rel atomic per CTA on one counter. CTAs finish a phase spread over microseconds,
//            so what counts is the latency of the LAST arrival: one round trip (a two-level tree costs
//            the last CTA two, measured +1.5 us per barrier).
//   release  the last CTA writes the new generation into one word per group of kBarGroupSize CTAs; a
//            CTA polls its own group's word. (With one shared word ~600 pollers keep a single L2 slice
//            busy at about one request per clock while the slowest CTAs still gather through it.)
template <class F>
__device__ __forceinline__ bool grid_barrier(const Bufs& B, unsigned long long& gen, int* s_flag, F&& last_fn) {
  const int group = blockIdx.x / kBarGroupSize;
  const int ngroups = (static_cast<int>(gridDim.x) + kBarGroupSize - 1) / kBarGroupSize;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned prev;
    // release: this CTA's writes (observed through the barrier above) before its arrival;
    // acquire: the last arrival sees every other CTA's
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(B.bar_top) : "memory");
    int last = 0;
    if (prev == gridDim.x - 1u) {
      *B.bar_top = 0u;  // re-armed: nobody arrives again before the release below
      last = 1;
    }
    *s_flag = last;
  }
  __syncthreads();
  bool aborted;
  if (*s_flag == 1) {
    aborted = last_fn();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long v = (gen + 1ull) | (aborted ? kBarAbortBit : 0ull);
      fence_gpu();
      for (int g = 0; g <= ngroups; ++g) st_relaxed_gpu_u64(B.bar_gen + g * kBarGenStride, v);
    }
  } else {
    if (threadIdx.x == 0) {
      const unsigned long long* word = B.bar_gen + (group + 1) * kBarGenStride;
      unsigned long long v, t0 = 0;
      unsigned spins = 0;
      for (;;) {
        v = ld_relaxed_gpu_u64(word);
        if ((v & ~kBarAbortBit) > gen) break;
        if ((++spins & 4095u) == 0u) {
          const unsigned long long now = gtimer_ns();
          if (t0 == 0) t0 = now;
          if (now - t0 > kBarTimeoutNs + B.p2p_timeout_ns) {  // never on a healthy device
            atomicExch(B.counters + 6, 2u);
            v = kBarAbortBit;
            break;
          }
        }
      }
      fence_gpu();
      *s_flag = (v & kBarAbortBit) ? 2 : 0;
    }
    __syncthreads();
    aborted = *s_flag == 2;
  }
  gen += 1ull;
  __syncthreads();  // s_flag is reused by the next barrier
  return aborted;
}

// phase timers (development / bench probe): thread 0 of the last CTA of a barrier
__device__ __forceinline__ void stamp_phase(const Bufs& B, int phase) {
  if (!B.timers) return;
  const unsigned long long now = gtimer_ns();
  const unsigned long long prev = B.timers[0];
  if (prev) B.timers[1 + phase] += now - prev;
  B.timers[0] = now;
  if (phase == 2) B.timers[4] += 1ull;
}

// every thread of the CTA: the CTA's copy of the solver scalars <- global
__device__ __forceinline__ void load_state(DevState* dst, const DevState* src) {
  static_assert(sizeof(DevState) % 8 == 0, "copied as 8-byte words");
  constexpr int kWords = sizeof(DevState) / 8;
  if (threadIdx.x < kWords)
    reinterpret_cast<unsigned long long*>(dst)[threadIdx.x] =
        __ldcg(reinterpret_cast<const unsigned long long*>(src) + threadIdx.x);
}

__device__ __forceinline__ void store_state(DevState* dst, const DevState* src) {
  constexpr int kWords = sizeof(DevState) / 8;
  if (threadIdx.x < kWords)
    reinterpret_cast<unsigned long long*>(dst)[threadIdx.x] =
        reinterpret_cast<const unsigned long long*>(src)[threadIdx.x];
}

// CLUSTER: the whole grid is ONE thread-block cluster of at most 8 CTAs (tiny instances: the median
// Netlib LP has ~1e3 x 2e3 entries and a few thousand nonzeros, an iteration is three latency chains
// and no bandwidth at all). The hardware cluster barrier (release / acquire at cluster scope, ~0.2 us)
// then replaces the atomics and fences of grid_barrier (~3 us), and CTA 0 closes the attempt.
template <bool DIST, bool CLUSTER>
__global__ void __launch_bounds__(kSpmvThreads, kSpmvCtasPerSm)
k_take_steps(const __grid_constant__ Bufs B, const __grid_constant__ SpmvMat A, const __grid_constant__ SpmvMat At,
             const __grid_constant__ SpmvMat Q, int max_attempts) {
  static_assert(!(DIST && CLUSTER), "the cluster form is single-GPU");
  cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
  __shared__ double s_red[32];
  __shared__ DevState st;
  __shared__ int s_flag, s_abort;
  __shared__ unsigned long long s_gen;
  load_state(&st, B.st);
  if (threadIdx.x == 0) s_gen = ld_relaxed_gpu_u64(B.bar_gen) & ~kBarAbortBit;
  __syncthreads();
  unsigned long long gen = s_gen;
  const int G = static_cast<int>(gridDim.x);
  const int warp_first = blockIdx.x * kSpmvWarps + (threadIdx.x >> 5), warp_stride = G * kSpmvWarps;
  const int cta_first = blockIdx.x * kSpmvThreads, tstride = G * kSpmvThreads;
  __shared__ double2 s_stage[DIST ? 2 * kSpmvThreads : 1];
  if (B.timers && blockIdx.x == 0 && threadIdx.x == 0) B.timers[0] = gtimer_ns();

  // one peer exchange (partitioned mode), by thread 0 of the last CTA of a barrier: everything this
  // rank pushed during the phase is fenced system-wide, the rank's flag goes up on every rank, and
  // the peers' flags are awaited. Returns true if a peer was lost.
  auto exchange = [&](int kind) -> bool {
    if (!(DIST && B.p2p)) return false;
    if (threadIdx.x == 0) {
      bool lost = false;
      if (!(B.dbg & 2)) {
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        const unsigned long long v = p2p_value(st, kind);
        p2p_signal_fenced(B, kind, v);
        lost = p2p_wait(B, kind, v);
      }
      s_abort = lost ? 1 : 0;
    }
    __syncthreads();
    return s_abort != 0;
  };

  // grid-wide barrier with a callback in ONE CTA (the last to arrive, or CTA 0 of the cluster) that
  // every CTA waits for; need_fn = false: the callback may be skipped (no exchange, no timers)
  auto barrier = [&](auto&& fn, bool need_fn) -> bool {
    if (CLUSTER) {
      cluster.sync();
      if (need_fn) {
        if (blockIdx.x == 0) (void)fn();
        cluster.sync();
      }
      return false;
    }
    return grid_barrier(B, gen, &s_flag, fn);
  };
  const bool phase_fn = DIST || B.timers != nullptr;

  for (int a = 0; a < max_attempts; ++a) {
    if (!st.active) break;  // the same in every CTA: the copies are bit-identical
    // ---- phase 1: primal step on the (local slice of the) variables, xbar ----
    {
      const double acc = primal_range<DIST>(B, st, cta_first, tstride, DIST ? s_stage : nullptr);
      const double t = block_reduce<false>(acc, s_red);
      if (threadIdx.x == 0) part_ptr(B, kSlotPrimal, 0)[blockIdx.x] = t;
    }
    if (barrier([&]() -> bool {
          const bool lost = exchange(0);
          if (threadIdx.x == 0) stamp_phase(B, 0);
          return lost;
        }, phase_fn))
      break;
    // ---- phase 2: A * xbar, dual step (and, with a quadratic objective, Q * x+ and dx' Q dx) ----
    {
      EpiDualT<DIST, const Bufs&> ed{B};
      ed.setup(st);
      if (DIST) ed.stage = reinterpret_cast<double*>(s_stage);  // 8 warps x 2 x 32 doubles (the primal tiles are done)
      spmv_items<EpiDualT<DIST, const Bufs&>, true>(A, ed, warp_first, warp_stride);
      __syncthreads();
      ed.publish(s_red);
      if (B.has_q) {
        EpiQxT<const Bufs&> ex{B};
        ex.setup(st);
        spmv_items<EpiQxT<const Bufs&>, true>(Q, ex, warp_first, warp_stride);
        EpiQDotT<const Bufs&> eq{B};
        eq.setup(st);
        spmv_items<EpiQDotT<const Bufs&>, true>(Q, eq, warp_first, warp_stride);
        __syncthreads();
        eq.publish(s_red);
      }
    }
    if (barrier([&]() -> bool {
          const bool lost = exchange(1);
          if (threadIdx.x == 0) stamp_phase(B, 1);
          return lost;
        }, phase_fn))
      break;
    // ---- phase 3: A' * y+, interaction; the last CTA of the barrier closes the attempt ----
    {
      EpiTransT<DIST, const Bufs&> et{B};
      et.setup(st);
      spmv_items<EpiTransT<DIST, const Bufs&>, true>(At, et, warp_first, warp_stride);
      __syncthreads();
      et.publish(s_red);
    }
    if (barrier([&]() -> bool {
          attempt_totals(B, G, G, G, B.has_q ? G : 0, s_red);
          if (threadIdx.x == 0) {
            double t[5] = {s_red[16], s_red[17], s_red[18], s_red[19], s_red[20]};
            bool lost = false;
            if (DIST) {  // every rank sums the ranks' totals in rank order: identical decisions everywhere
#pragma unroll
              for (int r = 0; r < kMaxWorld; ++r)
                if (r < B.world) {
                  double* slot = B.sc_peer[r] + B.rank * kScBlock;
                  slot[0] = t[0];
                  slot[1] = t[1];
                  slot[2] = t[2];
                  slot[3] = t[3];
                }
              asm volatile("fence.acq_rel.sys;" ::: "memory");
              const unsigned long long v = p2p_value(st, 2);
              p2p_signal_fenced(B, 2, v);
              lost = (B.dbg & 2) ? false : p2p_wait(B, 2, v);
              for (int k = 0; k < 4; ++k) t[k] = 0.0;
              for (int r = 0; r < B.world; ++r)
                for (int k = 0; k < 4; ++k) t[k] += __ldcg(B.sc_recv + r * kScBlock + k);
              t[4] = 0.0;  // partitioned mode is LP only
            }
            if (lost) {
              st.p2p_timeout = 1;
              st.active = 0;
            } else {
              finalize_attempt(&st, t[0], t[1], t[2], t[3], t[4]);  // on this CTA's copy (shared memory)
            }
            stamp_phase(B, 2);
            s_abort = lost ? 1 : 0;
          }
          __syncthreads();
          store_state(B.st, &st);  // published to the other CTAs (and the host) by the barrier's release
          return s_abort != 0;
        }, true))
      break;
    const bool closing = CLUSTER ? blockIdx.x == 0 : s_flag == 1;
    if (!closing) load_state(&st, B.st);  // the closing CTA already holds the new state
    __syncthreads();
  }
}

int take_steps_grid(int sm_count, bool dist) {
  int per_sm = 0;
  const cudaError_t e = dist ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_take_steps<true, false>, kSpmvThreads, 0)
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_take_steps<false, false>, kSpmvThreads, 0);
  if (e != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    return 0;
  }
  if (per_sm > kSpmvCtasPerSm) per_sm = kSpmvCtasPerSm;
  int g = sm_count * per_sm;
  const int cap = kBarGroupSize * kBarMaxGroups;
  if (g > cap) g = cap;
  if (g > kMaxPartialBlocks) g = kMaxPartialBlocks;
  return g;
}

int launch_take_steps(const Bufs& B, const SpmvMat& A, const SpmvMat& At, const SpmvMat& Q, int attempts,
                      int grid, cudaStream_t s) {
  void* args[] = {const_cast<Bufs*>(&B), const_cast<SpmvMat*>(&A), const_cast<SpmvMat*>(&At),
                  const_cast<SpmvMat*>(&Q), &attempts};
  const void* fn = B.world > 1 ? reinterpret_cast<const void*>(k_take_steps<true, false>)
                               : reinterpret_cast<const void*>(k_take_steps<false, false>);
  return cudaLaunchCooperativeKernel(fn, dim3(static_cast<unsigned>(grid)), dim3(kSpmvThreads), args, 0, s);
}

// The same batch as ONE thread-block cluster of `grid` <= kTakeClusterMax CTAs (single GPU, tiny instances).
int launch_take_steps_cluster(const Bufs& B, const SpmvMat& A, const SpmvMat& At, const SpmvMat& Q, int attempts,
                              int grid, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kSpmvThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(grid);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_take_steps<false, true>, B, A, At, Q, attempts);
}

static int spmv_grid(const SpmvMat& A, int grid_spmv);
// every block of a pushing kernel ends with one system-scope fence, and those serialise:
// the partitioned primal step uses fewer, longer-running blocks
static int dist_primal_grid(const Bufs& B) { return B.grid_vec / 2; }

void launch_dist_primal(const Bufs& B, cudaStream_t s) {
  k_primal<true><<<dist_primal_grid(B), kVecThreads, 0, s>>>(B);
}
void launch_dist_dual(const Bufs& B, const SpmvMat& A, cudaStream_t s) {
  EpiDualDist ed;
  ed.B = B;
  k_spmv<EpiDualDist><<<spmv_grid(A, B.grid_spmv), kSpmvThreads, 0, s>>>(A, ed);
}
void launch_dist_trans(const Bufs& B, const SpmvMat& A, const SpmvMat& At, cudaStream_t s) {
  EpiTransDist et;
  et.B = B;
  et.g_primal = dist_primal_grid(B);
  et.g_dual = spmv_grid(A, B.grid_spmv);
  k_spmv<EpiTransDist><<<spmv_grid(At, B.grid_spmv), kSpmvThreads, 0, s>>>(At, et);
}
void launch_dist_finalize(const Bufs& B, cudaStream_t s) { k_finalize_dist<<<1, 32, 0, s>>>(B); }

int spmv_configure() { return cudaSuccess; }  // no dynamic shared memory to opt into

// grid_spmv = resident CTAs of k_spmv on the device; fewer when there is less work
static int spmv_grid(const SpmvMat& A, int grid_spmv) {
  const int need = (A.ntiles + kSpmvWarps - 1) / kSpmvWarps;
  const int g = need < grid_spmv ? need : grid_spmv;
  return g < 1 ? 1 : g;
}

// launch with (pdl) or without the programmatic-serialization attribute, see pdl_wait()
template <class... P, class... A>
static void launch_chained(void (*kernel)(P...), int grid, int threads, cudaStream_t s, bool pdl,
                           A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(static_cast<unsigned>(threads));
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, args...);
}

static void launch_q_products(const Bufs& B, const SpmvMat& Q, int gq, cudaStream_t s, bool pdl = false) {
  EpiQx ex;
  ex.B = B;
  EpiQDot eq;
  eq.B = B;
  launch_chained(k_spmv<EpiQx>, gq, kSpmvThreads, s, pdl, Q, ex);
  launch_chained(k_spmv<EpiQDot>, gq, kSpmvThreads, s, pdl, Q, eq);
}

void launch_step_attempts(const Bufs& B, const SpmvMat& A, const SpmvMat& At, const SpmvMat& Q,
                          int attempts, cudaStream_t s) {
  const int g1 = B.grid_vec;
  const int g2 = spmv_grid(A, B.grid_spmv), g3 = spmv_grid(At, B.grid_spmv);
  const int gq = B.has_q ? spmv_grid(Q, B.grid_spmv) : 0;
  EpiDual ed;
  ed.B = B;
  EpiTrans et;
  et.B = B;
  et.g_primal = g1;
  et.g_dual = g2;
  et.g_q = gq;
  // Opt-in (FOLP_PDL=1): measured on the 1e6 x 1e6 x 1e7 workload it changes nothing (8 045 vs 8 044
  // take_step iterations/s): inside a CUDA graph the launch gap is already off the critical
  // path, what remains between kernels is the slowest CTA's tail, which a dependent launch
  // cannot start ahead of.
  const bool pdl = getenv("FOLP_PDL") != nullptr;
  for (int a = 0; a < attempts; ++a) {
    launch_chained(k_primal<false>, g1, kVecThreads, s, pdl, B);
    launch_chained(k_spmv<EpiDual>, g2, kSpmvThreads, s, pdl, A, ed);
    if (gq) launch_q_products(B, Q, gq, s, pdl);
    launch_chained(k_spmv<EpiTrans>, g3, kSpmvThreads, s, pdl, At, et);
  }
}

void launch_step_attempt_timed(const Bufs& B, const SpmvMat& A, const SpmvMat& At, const SpmvMat& Q,
                               cudaEvent_t* ev, cudaStream_t s) {
  const int g1 = B.grid_vec;
  const int g2 = spmv_grid(A, B.grid_spmv), g3 = spmv_grid(At, B.grid_spmv);
  const int gq = B.has_q ? spmv_grid(Q, B.grid_spmv) : 0;
  EpiDual ed;
  ed.B = B;
  EpiTrans et;
  et.B = B;
  et.g_primal = g1;
  et.g_dual = g2;
  et.g_q = gq;
  cudaEventRecord(ev[0], s);
  k_primal<false><<<g1, kVecThreads, 0, s>>>(B);
  cudaEventRecord(ev[1], s);
  k_spmv<EpiDual><<<g2, kSpmvThreads, 0, s>>>(A, ed);
  cudaEventRecord(ev[2], s);
  if (gq) launch_q_products(B, Q, gq, s);  // timed with K3
  k_spmv<EpiTrans><<<g3, kSpmvThreads, 0, s>>>(At, et);
  cudaEventRecord(ev[3], s);
}

void launch_spmv_plain(const SpmvMat& A, const double* in, double* out, int grid, cudaStream_t s) {
  EpiPlain ep;
  ep.in = in;
  ep.out = out;
  k_spmv<EpiPlain><<<spmv_grid(A, grid), kSpmvThreads, 0, s>>>(A, ep);
}

// ---------------------------------------------------------------------------
// evaluation block
// ---------------------------------------------------------------------------

// Applies the deferred average update (if any) so that sum_x / sum_y are current.
__global__ void __launch_bounds__(kVecThreads) k_flush_avg(Bufs B) {
  const DevState& s = *B.st;
  const bool pend = s.pending_avg & 1, pend_old = (s.pending_avg & 2) != 0;
  if (!pend && !pend_old) return;
  const double w = s.pending_w, w_old = s.mp_old_step;
  const double* __restrict__ xc = sel(B.x, s.cur);
  const double* __restrict__ xo = sel(B.x, s.cur ^ 1);
  const double* __restrict__ yc = sel(B.y, s.cur);
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < B.n; j += stride) {
    double sx = B.sum_x[j];
    if (pend_old) sx += xo[j] * w_old;
    if (pend) sx += xc[j] * w;
    B.sum_x[j] = sx;
  }
  if (pend)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B.m; i += stride)
      B.sum_y[i] += yc[i] * w;
}
__global__ void k_clear_pending(DevState* st) { st->pending_avg = 0; }

void launch_flush_avg(const Bufs& B, cudaStream_t s) {
  k_flush_avg<<<B.grid_vec, kVecThreads, 0, s>>>(B);
  k_clear_pending<<<1, 1, 0, s>>>(B.st);
}

// compute_average (sp.jl:296-301) or the current iterate (pdhg.jl:902-910)
__global__ void __launch_bounds__(kVecThreads) k_make_avg(Bufs B, int use_current) {
  const DevState& s = *B.st;
  const int stride = gridDim.x * blockDim.x;
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (use_current) {
    const double* __restrict__ xc = sel(B.x, s.cur);
    const double* __restrict__ yc = sel(B.y, s.cur);
    for (int j = t0; j < B.n; j += stride) B.avg_x[j] = xc[j];
    for (int i = t0; i < B.m; i += stride) B.avg_y[i] = yc[i];
  } else {
    const double wx = s.sum_w_x, wy = s.sum_w_y;
    for (int j = t0; j < B.n; j += stride) B.avg_x[j] = B.sum_x[j] / wx;
    for (int i = t0; i < B.m; i += stride) B.avg_y[i] = B.sum_y[i] / wy;
  }
}
void launch_make_avg(const Bufs& B, int use_current, cudaStream_t s) {
  k_make_avg<<<B.grid_vec, kVecThreads, 0, s>>>(B, use_current);
}

// Block-reduces NSUM sums followed by NMAX maxima, and lets the last block
// produce the final values in red_out[0 .. NSUM+NMAX). One barrier for all scalars: every warp
// shuffle-reduces each of them, warp leaders park them in shared memory, thread k combines
// scalar k over the warps in warp order; in the last block warp w finishes scalars w, w+8, ...
// over the per-block partials (lanes stride the blocks, fixed order, then a shuffle tree).
template <int NSUM, int NMAX>
__device__ __forceinline__ void reduce_and_publish(const Bufs& B, double* sums, double* maxs,
                                                   double* red_out, double* /*sh, unused*/) {
  constexpr int K = NSUM + NMAX;
  constexpr int kWarps = kVecThreads / 32;
  __shared__ double sh_w[K * kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* part = B.part + static_cast<size_t>(kSlotEval) * kMaxScalars * kMaxPartialBlocks;
#pragma unroll
  for (int k = 0; k < NSUM; ++k) {
    const double t = warp_sum(sums[k]);
    if (lane == 0) sh_w[k * kWarps + warp] = t;
  }
#pragma unroll
  for (int k = 0; k < NMAX; ++k) {
    const double t = warp_max(maxs[k]);
    if (lane == 0) sh_w[(NSUM + k) * kWarps + warp] = t;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    const int k = threadIdx.x;
    double t = sh_w[k * kWarps];
    for (int w = 1; w < kWarps; ++w)
      t = k < NSUM ? t + sh_w[k * kWarps + w] : fmax(t, sh_w[k * kWarps + w]);
    part[static_cast<size_t>(k) * kMaxPartialBlocks + blockIdx.x] = t;
  }
  if (!last_block_arrive(B.counters + kSlotEval)) return;
  const int G = gridDim.x;
  for (int k = warp; k < K; k += kWarps) {
    const bool is_max = k >= NSUM;
    double t = is_max ? -INFINITY : 0.0;
    for (int j = lane; j < G; j += 32) {
      const double q = __ldcg(part + static_cast<size_t>(k) * kMaxPartialBlocks + j);
      t = is_max ? fmax(t, q) : t + q;
    }
    t = is_max ? warp_max(t) : warp_sum(t);
    if (lane == 0) red_out[k] = t;
  }
}

// compute_convergence_information / compute_infeasibility_information
// (isu.jl:228-349) restricted to the variable-indexed terms, evaluated at
// xhat = avg_x ./ D with A_O' yhat = D .* (A_P' avg_y); plus the pieces of the
// Lagrangian of the scaled problem (sp.jl:1109-1120).
__global__ void __launch_bounds__(kVecThreads) k_stats_n(Bufs B, double* red_out) {
  __shared__ double sh[32];
  constexpr int NMAX = SN_TOTAL - SN_NSUM;
  double s[SN_NSUM], mx[NMAX];
#pragma unroll
  for (int k = 0; k < SN_NSUM; ++k) s[k] = 0.0;
#pragma unroll
  for (int k = 0; k < NMAX; ++k) mx[k] = 0.0;
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < B.n; j += stride) {
    const double xa = B.avg_x[j], at = B.aty_avg[j], Dj = B.D[j];
    const double xh = xa / Dj;  // sp.jl:65-67, isu.jl:436
    const double q = at * Dj;
    // Q_O xhat = D .* (Q_P avg_x): Q_P = D^-1 Q_O D^-1 (preprocess.jl:99-113)
    const double qxa = B.has_q ? B.qx_avg[j] : 0.0;
    const double qh = qxa * Dj;
    const double c = B.c_orig[j], l = B.l_orig[j], u = B.u_orig[j];
    s[SN_cx] += c * xh;  // isu.jl:67-74
    if (B.has_q) {
      s[SN_xqx] += xh * qh;  // x' Q x of isu.jl:67-74 and :186
      s[SN_xs_qxs] += xa * qxa;  // sp.jl:1109-1120 on the scaled problem
      mx[SN_qx_max - SN_NSUM] = fmax(mx[SN_qx_max - SN_NSUM], fabs(qh));  // isu.jl:311-313
    }
    const double lv = fmax(l - xh, 0.0), uv = fmax(xh - u, 0.0);  // isu.jl:52-55
    s[SN_lviol2] += lv * lv;
    s[SN_uviol2] += uv * uv;
    mx[SN_lviol_max - SN_NSUM] = fmax(mx[SN_lviol_max - SN_NSUM], lv);
    mx[SN_uviol_max - SN_NSUM] = fmax(mx[SN_uviol_max - SN_NSUM], uv);
    const double g = B.has_q ? (qh + c) - q : c - q;  // sp.jl:1081-1091
    const double bound = g > 0.0 ? l : u;  // isu.jl:128-147
    const double rc = isfinite(bound) ? g : 0.0;
    const double dres = g - rc;  // isu.jl:171-178
    s[SN_dres2] += dres * dres;
    mx[SN_dres_max - SN_NSUM] = fmax(mx[SN_dres_max - SN_NSUM], fabs(dres));
    if (rc != 0.0) s[SN_rcobj] += bound * rc;  // isu.jl:93-117
    s[SN_x2] += xh * xh;
    mx[SN_x_max - SN_NSUM] = fmax(mx[SN_x_max - SN_NSUM], fabs(xh));
    // primal ray, bounds of the homogeneous problem (isu.jl:301-313); scaled by
    // 1/|xhat|_inf on the host (positively homogeneous)
    if (isfinite(l)) mx[SN_ray_l_max - SN_NSUM] = fmax(mx[SN_ray_l_max - SN_NSUM], -xh);
    if (isfinite(u)) mx[SN_ray_u_max - SN_NSUM] = fmax(mx[SN_ray_u_max - SN_NSUM], xh);
    // dual ray: objective vector and matrix zeroed (isu.jl:319-330)
    const double g2 = -q;
    const double bound2 = g2 > 0.0 ? l : u;
    const double rc2 = isfinite(bound2) ? g2 : 0.0;
    mx[SN_ray_dres_max - SN_NSUM] = fmax(mx[SN_ray_dres_max - SN_NSUM], fabs(g2 - rc2));
    mx[SN_ray_rc_max - SN_NSUM] = fmax(mx[SN_ray_rc_max - SN_NSUM], fabs(rc2));
    if (rc2 != 0.0) s[SN_ray_rcobj] += bound2 * rc2;
    // scaled problem
    s[SN_cs_x] += xa * B.c[j];
    s[SN_x_aty] += xa * at;
    s[SN_xs2] += xa * xa;
  }
  reduce_and_publish<SN_NSUM, NMAX>(B, s, mx, red_out, sh);
}

__global__ void __launch_bounds__(kVecThreads) k_stats_m(Bufs B, double* red_out) {
  __shared__ double sh[32];
  constexpr int NMAX = SM_TOTAL - SM_NSUM;
  double s[SM_NSUM], mx[NMAX];
#pragma unroll
  for (int k = 0; k < SM_NSUM; ++k) s[k] = 0.0;
#pragma unroll
  for (int k = 0; k < NMAX; ++k) mx[k] = 0.0;
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B.m; i += stride) {
    const double ya = B.avg_y[i], ax = B.ax_avg[i], Ei = B.E[i];
    const double yh = ya / Ei;
    const double act = ax * Ei;  // A_O xhat = E .* (A_P avg_x)
    const double b = B.b_orig[i];
    double r = b - act;  // isu.jl:36-50
    const bool ineq = i >= B.neq;
    if (ineq) r = fmax(r, 0.0);
    s[SM_pres2] += r * r;
    mx[SM_pres_max - SM_NSUM] = fmax(mx[SM_pres_max - SM_NSUM], fabs(r));
    s[SM_by] += b * yh;
    s[SM_y2] += yh * yh;
    mx[SM_y_max - SM_NSUM] = fmax(mx[SM_y_max - SM_NSUM], fabs(yh));
    if (ineq) {
      const double yn = fmax(-yh, 0.0);  // isu.jl:171-173
      s[SM_yneg2] += yn * yn;
      mx[SM_yneg_max - SM_NSUM] = fmax(mx[SM_yneg_max - SM_NSUM], yn);
      mx[SM_ray_act_max - SM_NSUM] = fmax(mx[SM_ray_act_max - SM_NSUM], -act);
    } else {
      mx[SM_ray_act_max - SM_NSUM] = fmax(mx[SM_ray_act_max - SM_NSUM], fabs(act));
    }
    s[SM_bs_y] += ya * B.b[i];
    s[SM_ys2] += ya * ya;
  }
  reduce_and_publish<SM_NSUM, NMAX>(B, s, mx, red_out, sh);
}

// squared distances of the average and of the current iterate to the last
// restart point (sp.jl:444-488, :911-920)
__global__ void __launch_bounds__(kVecThreads) k_dist(Bufs B, double* red_out) {
  __shared__ double sh[32];
  double s[SD_TOTAL] = {0.0, 0.0, 0.0, 0.0};
  double mx[1] = {0.0};
  const DevState& st = *B.st;
  const double* __restrict__ xc = sel(B.x, st.cur);
  const double* __restrict__ yc = sel(B.y, st.cur);
  const int stride = gridDim.x * blockDim.x;
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int j = t0; j < B.n; j += stride) {
    const double lx = B.last_x[j];
    const double da = B.avg_x[j] - lx, dc = xc[j] - lx;
    s[SD_avg_x] += da * da;
    s[SD_cur_x] += dc * dc;
  }
  for (int i = t0; i < B.m; i += stride) {
    const double ly = B.last_y[i];
    const double da = B.avg_y[i] - ly, dc = yc[i] - ly;
    s[SD_avg_y] += da * da;
    s[SD_cur_y] += dc * dc;
  }
  reduce_and_publish<SD_TOTAL, 0>(B, s, mx, red_out, sh);
}

// The statistics kernels end in a block reduction of 4-28 values (~300 shuffles per thread): a fixed cost
// per block. FOLP_STATS_CTAS_PER_SM blocks per SM (default 3; was 8: 3.3 elements per thread at n = 1e6),
// never more blocks than there are elements to give every thread one.
#ifndef FOLP_STATS_CTAS_PER_SM
#define FOLP_STATS_CTAS_PER_SM 3
#endif
static int stats_grid(const Bufs& B, int len) {
  int g = B.grid_spmv / kSpmvCtasPerSm * FOLP_STATS_CTAS_PER_SM;
  const int need = (len + kVecThreads - 1) / kVecThreads;
  if (g > need) g = need;
  return g < 1 ? 1 : g;
}
void launch_stats_n(const Bufs& B, double* red_out, cudaStream_t s) {
  k_stats_n<<<stats_grid(B, B.n), kVecThreads, 0, s>>>(B, red_out);
}
void launch_stats_m(const Bufs& B, double* red_out, cudaStream_t s) {
  k_stats_m<<<stats_grid(B, B.m), kVecThreads, 0, s>>>(B, red_out);
}
void launch_dist(const Bufs& B, double* red_out, cudaStream_t s) {
  k_dist<<<stats_grid(B, B.n > B.m ? B.n : B.m), kVecThreads, 0, s>>>(B, red_out);
}

// The vector half of a restart (sp.jl:804-825, :921-923, pdhg.jl:1018-1022).
__global__ void __launch_bounds__(kVecThreads) k_apply_restart(Bufs B, int to_average,
                                                               int have_ax_cur) {
  const DevState& st = *B.st;
  double* __restrict__ xc = sel(B.x, st.cur);
  double* __restrict__ yc = sel(B.y, st.cur);
  double* __restrict__ atc = sel(B.aty, st.cur);
  const int stride = gridDim.x * blockDim.x;
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int j = t0; j < B.n; j += stride) {
    if (to_average) {
      xc[j] = B.avg_x[j];
      atc[j] = B.aty_avg[j];  // = A' * avg_y, what pdhg.jl:1021 recomputes
    }
    B.sum_x[j] = 0.0;
    B.last_x[j] = xc[j];
    B.last_aty[j] = atc[j];
    if (B.has_q) {
      double* __restrict__ qc = sel(B.qx, st.cur);
      if (to_average) qc[j] = B.qx_avg[j];
      B.last_qx[j] = qc[j];
    }
  }
  for (int i = t0; i < B.m; i += stride) {
    if (to_average) yc[i] = B.avg_y[i];
    B.sum_y[i] = 0.0;
    B.last_y[i] = yc[i];
    if (to_average) B.last_ax[i] = B.ax_avg[i];
    else if (have_ax_cur) B.last_ax[i] = B.ax_cur[i];
  }
}
void launch_apply_restart(const Bufs& B, int to_average, int have_ax_cur, cudaStream_t s) {
  k_apply_restart<<<B.grid_vec, kVecThreads, 0, s>>>(B, to_average, have_ax_cur);
}

// ---------------------------------------------------------------------------
// bound-constrained trust region (tr.jl:68-224) by safeguarded Newton passes.
//
// With h_i = w_i d_i^2 and thresholds t_i, rho^2(tau) = L(tau) + tau^2 H(tau),
// L(tau) = sum_{t_i <= tau} h_i t_i^2, H(tau) = sum_{t_i > tau} h_i, is the
// squared weighted radius reached at threshold tau. The reference finds the
// tau with rho^2 = r^2 by repeated medians; as a function of tau^2 it is
// concave piecewise linear, so Newton from the left, tau <- sqrt((r^2-L)/H)
// with the partition induced by the previous tau, increases monotonically and
// stops exactly when the partition stops changing -- at the same tau up to
// rounding. A bisection probe in bit-pattern space bounds the pass count.
// ---------------------------------------------------------------------------
enum TrInit {
  TI_g2 = 0, TI_H0, TI_Hinf, TI_Ltot, TI_cnt0, TI_cx, TI_xaty, TI_yb, TI_norm2, TI_gdp, TI_gdd,
  TI_xqx, TI_NSUM, TI_max_t = TI_NSUM, TI_TOTAL
};
static_assert(TI_TOTAL <= 16, "trust-region sums travel in one 16-scalar exchange");

struct TrElem {
  double x0, g, lb, ub, w;
};
__device__ __forceinline__ TrElem tr_elem(const Bufs& B, const TrProblem& P, int idx) {
  TrElem e;
  if (idx < B.n) {
    e.x0 = P.px[idx];
    e.g = P.qxp ? (P.qxp[idx] + B.c[idx]) - P.atp[idx] : B.c[idx] - P.atp[idx];  // sp.jl:1081-1091
    e.lb = B.l[idx];
    e.ub = B.u[idx];
    e.w = P.wp;
  } else {
    const int i = idx - B.n;
    e.x0 = P.py[i];
    e.g = -(B.b[i] - P.axp[i]);  // tr.jl:291, :312
    e.lb = i < B.neq ? -CUDART_INF : 0.0;  // tr.jl:288-290
    e.ub = CUDART_INF;
    e.w = P.wd;
  }
  return e;
}
__device__ __forceinline__ double jl_clamp(double x, double lo, double hi) {
  return x > hi ? hi : (x < lo ? lo : x);
}
__device__ __forceinline__ double bit_mid(double a, double b) {
  const long long ia = __double_as_longlong(a), ib = __double_as_longlong(b);
  return __longlong_as_double(ia + (ib - ia) / 2);
}

__device__ void tr_next_candidates(TrState& t) {
  double num = t.r2 - t.L_lo;
  if (num < 0.0) num = 0.0;
  double c0 = sqrt(num / t.H_lo);
  if (c0 > t.hi) c0 = t.hi;
  if (c0 < t.lo) c0 = t.lo;
  t.cand[0] = c0;
  t.cand[1] = bit_mid(c0, t.hi);
}

// Starts the search from the global sums of k_tr_init (thread 0 only).
__device__ void tr_setup(TrState* trs, const double* r, const TrProblem& P) {
  TrState t;
  t.radius = P.radius;
  t.r2 = P.radius * P.radius;
  t.lo = 0.0; t.L_lo = 0.0; t.H_lo = r[TI_H0];
  t.cnt_lo = static_cast<long long>(r[TI_cnt0]);
  t.hi = r[TI_max_t];
  t.max_t = r[TI_max_t];
  t.cand[0] = t.cand[1] = 0.0;
  t.tau = 0.0;
  t.done = 0; t.zero_value = 0; t.approx = P.approx; t.passes = 0;
  t.approx_scale = 1.0;
  t.cx = r[TI_cx]; t.x_aty = r[TI_xaty]; t.y_b = r[TI_yb]; t.xqx = r[TI_xqx];
  t.v_primal = 0.0; t.v_dual = 0.0;
  if (P.approx) {
    const double nrm = sqrt(r[TI_norm2]);
    if (nrm > 0.0) t.approx_scale = P.radius / nrm;
    t.v_primal = r[TI_gdp] * t.approx_scale;
    t.v_dual = r[TI_gdd] * t.approx_scale;
    t.done = 1; t.zero_value = 1;  // no final pass needed
  } else if (P.radius == 0.0 || r[TI_g2] == 0.0 || r[TI_H0] == 0.0) {  // tr.jl:88-91
    t.done = 1; t.zero_value = 1;
  } else if (r[TI_Hinf] == 0.0 && r[TI_Ltot] < t.r2) {  // everything reaches its bound, tr.jl:175-177
    t.tau = t.max_t; t.done = 1;
  } else if (r[TI_Hinf] > 0.0 && r[TI_Ltot] <= t.r2 &&
             sqrt((t.r2 - r[TI_Ltot]) / r[TI_Hinf]) >= t.max_t) {
    t.tau = sqrt((t.r2 - r[TI_Ltot]) / r[TI_Hinf]); t.done = 1;
  } else {
    tr_next_candidates(t);
  }
  *trs = t;
}

// One Newton / bisection step from the global sums {L0,H0,cnt0,L1,H1,cnt1} of k_tr_pass.
__device__ void tr_update(TrState* trs, const double* r) {
  TrState t = *trs;
  const double c0 = t.cand[0], c1 = t.cand[1];
  t.passes += 1;
  const long long cnt0 = static_cast<long long>(r[2]), cnt1 = static_cast<long long>(r[5]);
  if (cnt0 == t.cnt_lo) {  // partition unchanged: cand[0] is the fixed point
    t.tau = c0;
    t.done = 1;
  } else {
    t.lo = c0; t.L_lo = r[0]; t.H_lo = r[1]; t.cnt_lo = cnt0;
    if (c1 > c0 && c1 < t.hi) {
      const double F1 = r[3] + c1 * c1 * r[4];
      if (F1 < t.r2) { t.lo = c1; t.L_lo = r[3]; t.H_lo = r[4]; t.cnt_lo = cnt1; }
      else t.hi = c1;
    }
    if (t.H_lo <= 0.0) {  // nothing left above lo
      t.tau = t.max_t;
      t.done = 1;
    } else {
      tr_next_candidates(t);
    }
  }
  *trs = t;
}

__global__ void __launch_bounds__(kVecThreads) k_tr_init(Bufs B, TrProblem P, TrState* trs,
                                                         double* red) {
  __shared__ double sh[32];
  double s[TI_NSUM], mx[1] = {0.0};
#pragma unroll
  for (int k = 0; k < TI_NSUM; ++k) s[k] = 0.0;
  const int stride = gridDim.x * blockDim.x;
  const int total = B.n + B.m;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const bool primal = idx < B.n;
    const TrElem e = tr_elem(B, P, idx);
    if (primal) {  // compute_lagrangian_value, sp.jl:1109-1120
      s[TI_cx] += e.x0 * B.c[idx];
      s[TI_xaty] += e.x0 * P.atp[idx];
      if (P.qxp) s[TI_xqx] += e.x0 * P.qxp[idx];
    } else {
      s[TI_yb] += e.x0 * B.b[idx - B.n];
    }
    if (primal ? !P.use_primal : !P.use_dual) continue;
    s[TI_g2] += e.g * e.g;
    const bool skip = (e.x0 >= e.ub && e.g <= 0.0) || (e.x0 <= e.lb && e.g >= 0.0);  // tr.jl:96-103
    const double d = skip ? 0.0 : -e.g / e.w;
    B.tr_d[idx] = d;
    if (P.approx) {  // tr.jl:194-224
      s[TI_norm2] += e.w * d * d;
      if (primal) s[TI_gdp] += e.g * d;
      else s[TI_gdd] += e.g * d;
      continue;
    }
    double t = 0.0;  // tr.jl:104-116
    if (d > 0.0) t = (e.ub - e.x0) / d;
    else if (d < 0.0) t = (e.lb - e.x0) / d;
    B.tr_t[idx] = t;
    const double h = e.w * d * d;
    if (isinf(t)) {
      s[TI_Hinf] += h;
      s[TI_H0] += h;
    } else {
      if (t > 0.0) s[TI_H0] += h;
      else s[TI_cnt0] += 1.0;
      s[TI_Ltot] += h * t * t;
      mx[0] = fmax(mx[0], t);
    }
  }
  // publish through the generic path into `red`, then thread 0 of the last
  // block (the only one that returns with values) sets up the search
  double* part = B.part + static_cast<size_t>(kSlotEval) * kMaxScalars * kMaxPartialBlocks;
#pragma unroll
  for (int k = 0; k < TI_NSUM; ++k) {
    const double v = block_reduce<false>(s[k], sh);
    if (threadIdx.x == 0) part[static_cast<size_t>(k) * kMaxPartialBlocks + blockIdx.x] = v;
  }
  {
    const double v = block_reduce<true>(mx[0], sh);
    if (threadIdx.x == 0) part[static_cast<size_t>(TI_max_t) * kMaxPartialBlocks + blockIdx.x] = v;
  }
  if (!last_block_arrive(B.counters + kSlotEval)) return;
  double r[TI_TOTAL];
  for (int k = 0; k < TI_NSUM; ++k)
    r[k] = reduce_partials<false>(part + static_cast<size_t>(k) * kMaxPartialBlocks, gridDim.x, sh);
  r[TI_max_t] = reduce_partials<true>(part + static_cast<size_t>(TI_max_t) * kMaxPartialBlocks,
                                      gridDim.x, sh);
  if (threadIdx.x != 0) return;
  if (B.world > 1) {  // local sums only; tr_setup runs after the scalar exchange
    for (int k = 0; k < TI_TOTAL; ++k) B.sc_send[k] = r[k];
    return;
  }
  for (int k = 0; k < TI_TOTAL; ++k) red[k] = r[k];
  tr_setup(trs, r, P);
}

__global__ void __launch_bounds__(kVecThreads) k_tr_pass(Bufs B, TrProblem P, TrState* trs) {
  __shared__ double sh[32];
  if (trs->done) return;
  const double c0 = trs->cand[0], c1 = trs->cand[1];
  double s[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // L0,H0,cnt0,L1,H1,cnt1
  const int begin = P.use_primal ? 0 : B.n;
  const int end = P.use_dual ? B.n + B.m : B.n;
  const int stride = gridDim.x * blockDim.x;
  for (int idx = begin + blockIdx.x * blockDim.x + threadIdx.x; idx < end; idx += stride) {
    const double t = B.tr_t[idx], d = B.tr_d[idx];
    const double h = (idx < B.n ? P.wp : P.wd) * d * d;
    const double lt = h * t * t;
    if (t <= c0) { s[0] += lt; s[2] += 1.0; } else { s[1] += h; }
    if (t <= c1) { s[3] += lt; s[5] += 1.0; } else { s[4] += h; }
  }
  double* part = B.part + static_cast<size_t>(kSlotEval) * kMaxScalars * kMaxPartialBlocks;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double v = block_reduce<false>(s[k], sh);
    if (threadIdx.x == 0) part[static_cast<size_t>(k) * kMaxPartialBlocks + blockIdx.x] = v;
  }
  if (!last_block_arrive(B.counters + kSlotEval)) return;
  double r[6];
  for (int k = 0; k < 6; ++k)
    r[k] = reduce_partials<false>(part + static_cast<size_t>(k) * kMaxPartialBlocks, gridDim.x, sh);
  if (threadIdx.x != 0) return;
  if (B.world > 1) {
    for (int k = 0; k < 6; ++k) B.sc_send[k] = r[k];
    return;
  }
  tr_update(trs, r);
}

__global__ void __launch_bounds__(kVecThreads) k_tr_final(Bufs B, TrProblem P, TrState* trs) {
  __shared__ double sh[32];
  if (!trs->done || trs->zero_value) return;
  const double tau = trs->tau;
  double s[2] = {0.0, 0.0};
  const int begin = P.use_primal ? 0 : B.n;
  const int end = P.use_dual ? B.n + B.m : B.n;
  const int stride = gridDim.x * blockDim.x;
  for (int idx = begin + blockIdx.x * blockDim.x + threadIdx.x; idx < end; idx += stride) {
    const TrElem e = tr_elem(B, P, idx);
    const double sol = jl_clamp(e.x0 + tau * B.tr_d[idx], e.lb, e.ub);  // tr.jl:182-188
    const double v = e.g * (sol - e.x0);
    if (idx < B.n) s[0] += v;
    else s[1] += v;
  }
  double* part = B.part + static_cast<size_t>(kSlotEval) * kMaxScalars * kMaxPartialBlocks;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double v = block_reduce<false>(s[k], sh);
    if (threadIdx.x == 0) part[static_cast<size_t>(k) * kMaxPartialBlocks + blockIdx.x] = v;
  }
  if (!last_block_arrive(B.counters + kSlotEval)) return;
  const double vp = reduce_partials<false>(part, gridDim.x, sh);
  const double vd = reduce_partials<false>(part + kMaxPartialBlocks, gridDim.x, sh);
  if (threadIdx.x == 0) {
    if (B.world > 1) {
      B.sc_send[0] = vp;
      B.sc_send[1] = vd;
    } else {
      trs->v_primal = vp;
      trs->v_dual = vd;
    }
  }
}

// row-partitioned mode: applies the rank-ordered totals of the exchanged local sums
__global__ void k_tr_combine(Bufs B, TrProblem P, TrState* trs, int stage,
                             const double* __restrict__ recv) {
  if (threadIdx.x != 0) return;
  double r[TI_TOTAL];
  const int count = stage == kTrInit ? TI_TOTAL : (stage == kTrPass ? 6 : 2);
  for (int k = 0; k < count; ++k) {
    const bool is_max = stage == kTrInit && k == TI_max_t;
    double v = is_max ? -CUDART_INF : 0.0;
    for (int q = 0; q < B.world; ++q) {
      const double w = __ldcg(recv + q * kScBlock + k);
      v = is_max ? fmax(v, w) : v + w;
    }
    r[k] = v;
  }
  if (stage == kTrInit) {
    tr_setup(trs, r, P);
  } else if (stage == kTrPass) {
    if (!trs->done) tr_update(trs, r);
  } else if (trs->done && !trs->zero_value) {
    trs->v_primal = r[0];
    trs->v_dual = r[1];
  }
}

// ---------------------------------------------------------------------------
// The whole trust-region solve in ONE cooperative kernel (single GPU): init, Newton /
// bisection passes until the partition stops changing, final value -- separated by grid-wide
// barriers instead of kernel boundaries, so that a solve costs one launch and one host read
// instead of ~12 launches (most of them early exits) and the passes stream tr_t / tr_d out of
// L2. Every block reduces the same per-block partials in the same order, so all blocks hold
// bit-identical copies of the TrState and take the same branches.
// ---------------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int kTrThreads = 256;
constexpr int kTrWarps = kTrThreads / 32;
constexpr int kTrMaxK = 16;

// Cross-rank part of grid_totals (partitioned mode over peer memory): block 0 pushes this rank's
// totals into slot `rank` of every rank's receive buffer (parity = seq & 1, the convention of
// k_exchange), raises flag kind 3 to seq and waits for every rank's; after a second grid barrier
// every block of every rank combines the same world x K numbers in rank order -- identical bits
// everywhere. Two receive buffers suffice: a rank cannot start exchange e+2 before every rank has
// consumed exchange e (it needs their flags of e+1, raised after that).
struct TrXchg {
  unsigned long long seq;  // number of the next exchange (host: folp_handle::xchg_seq + 1)
  int count;               // exchanges done by this kernel
};

// totals of K per-thread values over the whole grid -> tot[0..K) (shared memory, valid in every
// thread of every block after the call). Entry k is a maximum if bit k of max_mask is set, a sum
// otherwise. part: two alternating buffers of kTrMaxK * gridDim.x doubles (parity flips per call).
template <int K>
__device__ __forceinline__ void grid_totals(cg::grid_group& grid, const Bufs& B, const double (&v)[K],
                                            unsigned max_mask, double* part, int& parity,
                                            double* sh /* K * kTrWarps */, double* tot /* K */,
                                            TrXchg& xc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = gridDim.x;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double w = (max_mask >> k) & 1u ? warp_max(v[k]) : warp_sum(v[k]);
    if (lane == 0) sh[k * kTrWarps + warp] = w;
  }
  __syncthreads();
  double* buf = part + static_cast<size_t>(parity) * kTrMaxK * G;
  if (threadIdx.x < K) {
    const int k = threadIdx.x;
    const bool is_max = (max_mask >> k) & 1u;
    double t = sh[k * kTrWarps];
    for (int w = 1; w < kTrWarps; ++w) t = is_max ? fmax(t, sh[k * kTrWarps + w]) : t + sh[k * kTrWarps + w];
    buf[static_cast<size_t>(k) * G + blockIdx.x] = t;
  }
  grid.sync();
  for (int k = warp; k < K; k += kTrWarps) {
    const bool is_max = (max_mask >> k) & 1u;
    double t = is_max ? -CUDART_INF : 0.0;
    for (int j = lane; j < G; j += 32) {
      const double q = __ldcg(buf + static_cast<size_t>(k) * G + j);
      t = is_max ? fmax(t, q) : t + q;
    }
    t = is_max ? warp_max(t) : warp_sum(t);
    if (lane == 0) tot[k] = t;
  }
  __syncthreads();
  parity ^= 1;
  if (B.world > 1) {
    const int xpar = static_cast<int>(xc.seq & 1ull);
    const size_t base = static_cast<size_t>(xpar) * B.world * kScBlock;
    if (blockIdx.x == 0) {
      if (threadIdx.x < K) {
        const double mine = tot[threadIdx.x];
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
          if (r < B.world) B.scx_peer[r][base + B.rank * kScBlock + threadIdx.x] = mine;
      }
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) {
        p2p_signal(B, 3, xc.seq);
        p2p_wait(B, 3, xc.seq);
      }
    }
    grid.sync();
    if (threadIdx.x < K) {
      const int k = threadIdx.x;
      const bool is_max = (max_mask >> k) & 1u;
      double t = is_max ? -CUDART_INF : 0.0;
      for (int r = 0; r < B.world; ++r) {
        double q;
        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(q) : "l"(B.scx + base + r * kScBlock + k) : "memory");
        t = is_max ? fmax(t, q) : t + q;
      }
      tot[k] = t;
    }
    __syncthreads();
    xc.seq += 1;
    xc.count += 1;
  }
}

__global__ void __launch_bounds__(kTrThreads) k_tr_solve(Bufs B, TrProblem P, TrState* trs,
                                                         double* part, unsigned long long seq_first) {
  cg::grid_group grid = cg::this_grid();
  if (B.world > 1 && B.dseq) seq_first = __ldcg(B.dseq) + 1ull;  // every block reads it before block 0 rewrites it at the end (grid barriers in between)
  if (P.param_src != kTrParamHost) {  // same arithmetic as the host's (evaluate / run_restart_scheme in folp_api.cu)
    if (P.param_src == kTrParamBounds) {
      const double xs2 = B.red[SN_xs2], ys2 = B.red[kMaxScalars + SM_ys2];
      double rp = sqrt(P.wp * xs2), rd = sqrt(P.wd * ys2);
      rp = (rp != rp) ? rp : (1e-8 > rp ? 1e-8 : rp);  // Julia's max(1e-8, .) propagates NaN
      rd = (rd != rd) ? rd : (1e-8 > rd ? 1e-8 : rd);
      P.wp = P.wp / (rp * rp);
      P.wd = P.wd / (rd * rd);
      P.radius = 1.0;
    } else {
      const double* dist = B.red + 2 * kMaxScalars;
      const bool avg = P.param_src == kTrParamDistAvg;
      const double px = sqrt(P.wp * dist[avg ? SD_avg_x : SD_cur_x]);
      const double dy = sqrt(P.wd * dist[avg ? SD_avg_y : SD_cur_y]);
      P.radius = sqrt(px * px + dy * dy);
    }
  }
  TrXchg xc{seq_first, 0};
  __shared__ double sh[kTrMaxK * kTrWarps];
  __shared__ double tot[kTrMaxK];
  __shared__ TrState st;
  int parity = 0;
  const int stride = gridDim.x * blockDim.x;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = B.n + B.m;
  // ---- init (k_tr_init) ----
  {
    double s[TI_TOTAL];
#pragma unroll
    for (int k = 0; k < TI_TOTAL; ++k) s[k] = 0.0;
    for (int idx = tid; idx < total; idx += stride) {
      const bool primal = idx < B.n;
      const TrElem e = tr_elem(B, P, idx);
      if (primal) {  // compute_lagrangian_value, sp.jl:1109-1120
        s[TI_cx] += e.x0 * B.c[idx];
        s[TI_xaty] += e.x0 * P.atp[idx];
        if (P.qxp) s[TI_xqx] += e.x0 * P.qxp[idx];
      } else {
        s[TI_yb] += e.x0 * B.b[idx - B.n];
      }
      if (primal ? !P.use_primal : !P.use_dual) continue;
      s[TI_g2] += e.g * e.g;
      const bool skip = (e.x0 >= e.ub && e.g <= 0.0) || (e.x0 <= e.lb && e.g >= 0.0);  // tr.jl:96-103
      const double d = skip ? 0.0 : -e.g / e.w;
      B.tr_d[idx] = d;
      if (P.approx) {  // tr.jl:194-224
        s[TI_norm2] += e.w * d * d;
        if (primal) s[TI_gdp] += e.g * d;
        else s[TI_gdd] += e.g * d;
        continue;
      }
      double t = 0.0;  // tr.jl:104-116
      if (d > 0.0) t = (e.ub - e.x0) / d;
      else if (d < 0.0) t = (e.lb - e.x0) / d;
      B.tr_t[idx] = t;
      const double h = e.w * d * d;
      if (isinf(t)) {
        s[TI_Hinf] += h;
        s[TI_H0] += h;
      } else {
        if (t > 0.0) s[TI_H0] += h;
        else s[TI_cnt0] += 1.0;
        s[TI_Ltot] += h * t * t;
        s[TI_max_t] = fmax(s[TI_max_t], t);
      }
    }
    grid_totals<TI_TOTAL>(grid, B, s, 1u << TI_max_t, part, parity, sh, tot, xc);
    if (threadIdx.x == 0) tr_setup(&st, tot, P);
    __syncthreads();
  }
  // ---- passes (k_tr_pass) ----
  const int begin = P.use_primal ? 0 : B.n;
  const int end = P.use_dual ? total : B.n;
  while (!st.done) {
    const double c0 = st.cand[0], c1 = st.cand[1];
    double s[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // L0,H0,cnt0,L1,H1,cnt1
#pragma unroll 4
    for (int idx = begin + tid; idx < end; idx += stride) {
      const double t = B.tr_t[idx], d = B.tr_d[idx];
      const double h = (idx < B.n ? P.wp : P.wd) * d * d;
      const double lt = h * t * t;
      if (t <= c0) { s[0] += lt; s[2] += 1.0; } else { s[1] += h; }
      if (t <= c1) { s[3] += lt; s[5] += 1.0; } else { s[4] += h; }
    }
    grid_totals<6>(grid, B, s, 0u, part, parity, sh, tot, xc);
    if (threadIdx.x == 0) {
      tr_update(&st, tot);
      // non-finite data (a diverged iterate) can keep the partition changing for ever; a lost peer
      // (sticky time-out flag, identical on this rank's blocks after the barrier) ends the search too
      if (!st.done && (st.passes >= 120 || (B.world > 1 && __ldcg(B.counters + 6)))) {
        st.done = 2;
        st.zero_value = 1;
      }
    }
    __syncthreads();
  }
  // ---- final value (k_tr_final) ----
  if (!st.zero_value) {
    const double tau = st.tau;
    double s[2] = {0.0, 0.0};
    for (int idx = begin + tid; idx < end; idx += stride) {
      const TrElem e = tr_elem(B, P, idx);
      const double sol = jl_clamp(e.x0 + tau * B.tr_d[idx], e.lb, e.ub);  // tr.jl:182-188
      const double v = e.g * (sol - e.x0);
      if (idx < B.n) s[0] += v;
      else s[1] += v;
    }
    grid_totals<2>(grid, B, s, 0u, part, parity, sh, tot, xc);
    if (threadIdx.x == 0) {
      st.v_primal = tot[0];
      st.v_dual = tot[1];
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st.exchanges = xc.count;
    *trs = st;
    if (B.world > 1 && B.dseq) *B.dseq = xc.seq - 1ull;
  }
}

// ---------------------------------------------------------------------------
// k_tr_multi: ALL trust-region solves of one evaluation block in one cooperative kernel.
//
// An evaluation needs up to five bound-constrained trust-region problems (slots: bound estimates
// on the primal / on the dual part at the average, localized duality gaps at the average, at the
// current iterate and at the last restart point). As five k_tr_solve launches they cost ~480 us of
// the ~740 us evaluation block on the 1e6 x 1e6 x 1e7 workload, and ~30 grid-wide barriers -- in
// partitioned mode as many cross-rank scalar exchanges. Here the searches advance together: one
// barrier (and one exchange of up to 42 scalars) per stage for all slots, the slots that share a
// centre (the three at the average) share its loads, and the passes of a slot stop as soon as its
// own partition stops changing. Per slot the arithmetic is k_tr_solve's (same per-element
// formulas, tr_setup / tr_update); only the shape of the reductions differs (512 threads).
// Per centre there is at most one slot over both index ranges (a gap), one over the primal range
// and one over the dual range only (the bound estimates); the host checks that.
// Scratch: slots 0 and 1 cover disjoint index ranges and share one (n+m) pair of B.trm_t / B.trm_d,
// slots 2..4 own one each.
// ---------------------------------------------------------------------------
constexpr int kTrmThreads = 512;
constexpr int kTrmWarps = kTrmThreads / 32;
constexpr int kTrmMaxK = 48;   // values per grid-wide reduction (init: 3 centres x 4 + 5 slots x 6 = 42)
constexpr int kTrmV = 6;       // per-slot sums of the init stage and of a pass
constexpr int kTrmLag = 3 * 4; // first value of the per-slot blocks in the init stage
static_assert(kTrmMaxK <= kScBlock, "one exchange block carries a stage's scalars");
static_assert(2 * kTrmMaxK * 1024 <= kMaxScalars * kMaxPartialBlocks, "partials fit the evaluation slot");

// block-level reduction of NV per-thread values -> buf[(kbase + k) * G + blockIdx.x]
template <int NV>
__device__ __forceinline__ void trm_publish(const double (&v)[NV], unsigned max_mask, int kbase, double* buf,
                                            double* sh /* kTrmV * kTrmWarps */) {
  static_assert(NV <= kTrmV, "shared scratch");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = gridDim.x;
  __syncthreads();  // sh may still be read by the previous publish
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double w = (max_mask >> k) & 1u ? warp_max(v[k]) : warp_sum(v[k]);
    if (lane == 0) sh[k * kTrmWarps + warp] = w;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    const int k = threadIdx.x;
    const bool is_max = (max_mask >> k) & 1u;
    double t = sh[k * kTrmWarps];
    for (int w = 1; w < kTrmWarps; ++w) t = is_max ? fmax(t, sh[k * kTrmWarps + w]) : t + sh[k * kTrmWarps + w];
    buf[static_cast<size_t>(kbase + k) * G + blockIdx.x] = t;
  }
}

// grid-wide totals of the K published values -> tot[0..K) in every block (and, partitioned, the
// rank-ordered totals over all ranks). Bit k of max_mask: value k is a maximum. Values nobody
// published this stage hold stale numbers and are ignored by the caller.
__device__ __forceinline__ void trm_totals(cg::grid_group& grid, const Bufs& B, int K, unsigned long long max_mask,
                                           const double* buf, double* tot, TrXchg& xc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = gridDim.x;
  grid.sync();
  for (int k = warp; k < K; k += kTrmWarps) {
    const bool is_max = (max_mask >> k) & 1ull;
    double t = is_max ? -CUDART_INF : 0.0;
    for (int j = lane; j < G; j += 32) {
      const double q = __ldcg(buf + static_cast<size_t>(k) * G + j);
      t = is_max ? fmax(t, q) : t + q;
    }
    t = is_max ? warp_max(t) : warp_sum(t);
    if (lane == 0) tot[k] = t;
  }
  __syncthreads();
  if (B.world > 1) {
    const int xpar = static_cast<int>(xc.seq & 1ull);
    const size_t base = static_cast<size_t>(xpar) * B.world * kScBlock;
    if (blockIdx.x == 0) {
      if (threadIdx.x < K) {
        const double mine = tot[threadIdx.x];
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
          if (r < B.world) B.scx_peer[r][base + B.rank * kScBlock + threadIdx.x] = mine;
      }
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) {
        p2p_signal(B, 3, xc.seq);
        p2p_wait(B, 3, xc.seq);
      }
    }
    grid.sync();
    if (threadIdx.x < K) {
      const int k = threadIdx.x;
      const bool is_max = (max_mask >> k) & 1ull;
      double t = is_max ? -CUDART_INF : 0.0;
      for (int r = 0; r < B.world; ++r) {
        double q;
        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(q) : "l"(B.scx + base + r * kScBlock + k) : "memory");
        t = is_max ? fmax(t, q) : t + q;
      }
      tot[k] = t;
    }
    __syncthreads();
    xc.seq += 1;
    xc.count += 1;
  }
}

__global__ void __launch_bounds__(kTrmThreads, 2) k_tr_multi(Bufs B, TrMulti M, TrState* trs, double* part) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[kTrmV * kTrmWarps];
  __shared__ double tot[kTrmMaxK];
  __shared__ TrState st[kTrSlots];
  __shared__ double s_wp[kTrSlots], s_wd[kTrSlots], s_radius[kTrSlots];
  __shared__ int s_centre[kTrSlots];  // index of the slot's centre among the distinct centres
  __shared__ int s_any;
  TrXchg xc{0ull, 0};
  if (B.world > 1) xc.seq = __ldcg(B.dseq) + 1ull;  // read by every block before block 0 rewrites it at the end
  const int NS = M.nslots;
  const int G = gridDim.x;
  const int stride = G * kTrmThreads;
  const int tid = blockIdx.x * kTrmThreads + threadIdx.x;
  const int n = B.n, total = B.n + B.m;
  auto same_centre = [&](int a, int b) { return M.P[a].px == M.P[b].px && M.P[a].py == M.P[b].py; };
  // weights / radii: from the host, or from the reduced statistics in HBM (TrParamSrc; the same
  // arithmetic as evaluate() / run_restart_scheme() in folp_api.cu)
  if (threadIdx.x < NS) {
    const int s_ = threadIdx.x;
    const TrProblem& P = M.P[s_];
    double wp = P.wp, wd = P.wd, radius = P.radius;
    if (P.param_src == kTrParamBounds) {
      const double xs2 = B.red[SN_xs2], ys2 = B.red[kMaxScalars + SM_ys2];
      double rp = sqrt(P.wp * xs2), rd = sqrt(P.wd * ys2);
      rp = (rp != rp) ? rp : (1e-8 > rp ? 1e-8 : rp);  // Julia's max(1e-8, .) propagates NaN
      rd = (rd != rd) ? rd : (1e-8 > rd ? 1e-8 : rd);
      wp = P.wp / (rp * rp);
      wd = P.wd / (rd * rd);
      radius = 1.0;
    } else if (P.param_src == kTrParamDistAvg || P.param_src == kTrParamDistCur) {
      const double* dist = B.red + 2 * kMaxScalars;
      const bool avg = P.param_src == kTrParamDistAvg;
      const double px = sqrt(P.wp * dist[avg ? SD_avg_x : SD_cur_x]);
      const double dy = sqrt(P.wd * dist[avg ? SD_avg_y : SD_cur_y]);
      radius = sqrt(px * px + dy * dy);
    }
    s_wp[s_] = wp;
    s_wd[s_] = wd;
    s_radius[s_] = radius;
    int ci = 0;  // number of distinct centres among the slots before the first slot with this centre
    int first = s_;
    for (int q = s_ - 1; q >= 0; --q)
      if (same_centre(q, s_)) first = q;
    for (int q = 0; q < first; ++q) {
      bool lead = true;
      for (int r = 0; r < q; ++r) lead = lead && !same_centre(r, q);
      ci += lead ? 1 : 0;
    }
    s_centre[s_] = ci;
  }
  __syncthreads();
  double* const buf0 = part;
  double* const buf1 = part + static_cast<size_t>(kTrmMaxK) * G;
  int parity = 0;
  auto scratch_t = [&](int s_) { return B.trm_t + static_cast<size_t>(s_ <= 1 ? 0 : s_ - 1) * total; };
  auto scratch_d = [&](int s_) { return B.trm_d + static_cast<size_t>(s_ <= 1 ? 0 : s_ - 1) * total; };
  // the slots of the centre led by slot `lead`: over both ranges (gap), primal range only, dual range only
  auto centre_slots = [&](int lead, int& sg, int& sp, int& sd) {
    sg = sp = sd = -1;
    for (int q = lead; q < NS; ++q) {
      if (!same_centre(lead, q)) continue;
      if (M.P[q].use_primal && M.P[q].use_dual) sg = q;
      else if (M.P[q].use_primal) sp = q;
      else if (M.P[q].use_dual) sd = q;
    }
  };
  auto is_lead = [&](int q) {
    bool lead = true;
    for (int r = 0; r < q; ++r) lead = lead && !same_centre(r, q);
    return lead;
  };

  // ================= init: one sweep per distinct centre =================
  {
    double* buf = parity ? buf1 : buf0;
    for (int lead = 0; lead < NS; ++lead) {
      if (!is_lead(lead)) continue;
      const TrProblem& PL = M.P[lead];
      int sg, sp, sd;
      centre_slots(lead, sg, sp, sd);
      double lag[4] = {0.0, 0.0, 0.0, 0.0};  // c.x, x.A'y, y.b, x.Qx of the centre
      double vg[kTrmV] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      // one element of slot s_ (weight w): direction, threshold, sums.
      // value layout: exact {g2, H0, Hinf, Ltot, cnt0, max_t}, approximate {g2, norm2, gdp, gdd, -, -}
      auto element = [&](int s_, double (&v)[kTrmV], int idx, bool primal, double x0, double g, double lb,
                         double ub, bool skip) {
        const double w = primal ? s_wp[s_] : s_wd[s_];
        v[0] += g * g;
        const double d = skip ? 0.0 : -g / w;
        scratch_d(s_)[idx] = d;
        if (M.P[s_].approx) {  // tr.jl:194-224
          v[1] += w * d * d;
          if (primal) v[2] += g * d;
          else v[3] += g * d;
          return;
        }
        double t = 0.0;  // tr.jl:104-116
        if (d > 0.0) t = (ub - x0) / d;
        else if (d < 0.0) t = (lb - x0) / d;
        scratch_t(s_)[idx] = t;
        const double hh = w * d * d;
        if (isinf(t)) {
          v[2] += hh;
          v[1] += hh;
        } else {
          if (t > 0.0) v[1] += hh;
          else v[4] += 1.0;
          v[3] += hh * t * t;
          v[5] = fmax(v[5], t);
        }
      };
      {  // ---- primal range ----
        double vh[kTrmV] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int idx = tid; idx < n; idx += stride) {
          const double x0 = PL.px[idx];
          const double at = PL.atp[idx], cj = B.c[idx];
          const double qx = PL.qxp ? PL.qxp[idx] : 0.0;
          const double g = PL.qxp ? (qx + cj) - at : cj - at;  // sp.jl:1081-1091
          const double lb = B.l[idx], ub = B.u[idx];
          lag[0] += x0 * cj;  // compute_lagrangian_value, sp.jl:1109-1120
          lag[1] += x0 * at;
          if (PL.qxp) lag[3] += x0 * qx;
          const bool skip = (x0 >= ub && g <= 0.0) || (x0 <= lb && g >= 0.0);  // tr.jl:96-103
          if (sg >= 0) element(sg, vg, idx, true, x0, g, lb, ub, skip);
          if (sp >= 0) element(sp, vh, idx, true, x0, g, lb, ub, skip);
        }
        if (sp >= 0) trm_publish<kTrmV>(vh, M.P[sp].approx ? 0u : (1u << 5), kTrmLag + kTrmV * sp, buf, sh);
      }
      {  // ---- dual range ----
        double vh[kTrmV] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int idx = n + tid; idx < total; idx += stride) {
          const int i = idx - n;
          const double x0 = PL.py[i];
          const double bi = B.b[i];
          const double g = -(bi - PL.axp[i]);  // tr.jl:291, :312
          const double lb = i < B.neq ? -CUDART_INF : 0.0, ub = CUDART_INF;  // tr.jl:288-290
          lag[2] += x0 * bi;
          const bool skip = (x0 >= ub && g <= 0.0) || (x0 <= lb && g >= 0.0);
          if (sg >= 0) element(sg, vg, idx, false, x0, g, lb, ub, skip);
          if (sd >= 0) element(sd, vh, idx, false, x0, g, lb, ub, skip);
        }
        if (sd >= 0) trm_publish<kTrmV>(vh, M.P[sd].approx ? 0u : (1u << 5), kTrmLag + kTrmV * sd, buf, sh);
      }
      if (sg >= 0) trm_publish<kTrmV>(vg, M.P[sg].approx ? 0u : (1u << 5), kTrmLag + kTrmV * sg, buf, sh);
      trm_publish<4>(lag, 0u, 4 * s_centre[lead], buf, sh);
    }
    unsigned long long mask = 0ull;
    for (int q = 0; q < NS; ++q)
      if (!M.P[q].approx) mask |= 1ull << (kTrmLag + kTrmV * q + 5);
    trm_totals(grid, B, kTrmLag + kTrmV * NS, mask, buf, tot, xc);
    parity ^= 1;
    if (threadIdx.x < NS) {
      const int s_ = threadIdx.x;
      const double* v = tot + kTrmLag + kTrmV * s_;
      const double* lg = tot + 4 * s_centre[s_];
      double r[TI_TOTAL];
      for (int k = 0; k < TI_TOTAL; ++k) r[k] = 0.0;
      r[TI_g2] = v[0];
      if (M.P[s_].approx) {
        r[TI_norm2] = v[1]; r[TI_gdp] = v[2]; r[TI_gdd] = v[3];
      } else {
        r[TI_H0] = v[1]; r[TI_Hinf] = v[2]; r[TI_Ltot] = v[3]; r[TI_cnt0] = v[4]; r[TI_max_t] = v[5];
      }
      r[TI_cx] = lg[0]; r[TI_xaty] = lg[1]; r[TI_yb] = lg[2]; r[TI_xqx] = lg[3];
      TrProblem P = M.P[s_];
      P.wp = s_wp[s_]; P.wd = s_wd[s_]; P.radius = s_radius[s_];
      tr_setup(&st[s_], r, P);
    }
    __syncthreads();
  }

  // ================= passes: every unfinished slot, one barrier per round =================
  for (;;) {
    if (threadIdx.x == 0) {
      int any = 0;
      for (int q = 0; q < NS; ++q) any |= st[q].done ? 0 : 1;
      s_any = any;
    }
    __syncthreads();
    if (!s_any) break;
    double* buf = parity ? buf1 : buf0;
    for (int q = 0; q < NS; ++q) {
      if (st[q].done) continue;
      const double c0 = st[q].cand[0], c1 = st[q].cand[1];
      const double wp = s_wp[q], wd = s_wd[q];
      const double* pt = scratch_t(q);
      const double* pd = scratch_d(q);
      double v[kTrmV] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // L0,H0,cnt0,L1,H1,cnt1
      const int begin = M.P[q].use_primal ? 0 : n;
      const int end = M.P[q].use_dual ? total : n;
#pragma unroll 2
      for (int idx = begin + tid; idx < end; idx += stride) {
        const double t = pt[idx], d = pd[idx];
        const double hh = (idx < n ? wp : wd) * d * d;
        const double lt = hh * t * t;
        if (t <= c0) { v[0] += lt; v[2] += 1.0; } else { v[1] += hh; }
        if (t <= c1) { v[3] += lt; v[5] += 1.0; } else { v[4] += hh; }
      }
      trm_publish<kTrmV>(v, 0u, kTrmV * q, buf, sh);
    }
    trm_totals(grid, B, kTrmV * NS, 0ull, buf, tot, xc);
    parity ^= 1;
    if (threadIdx.x < NS && !st[threadIdx.x].done) {
      TrState* t_ = &st[threadIdx.x];
      tr_update(t_, tot + kTrmV * threadIdx.x);
      // non-finite data (a diverged iterate) can keep the partition changing for ever; a lost peer ends the search too
      if (!t_->done && (t_->passes >= 120 || (B.world > 1 && __ldcg(B.counters + 6)))) {
        t_->done = 2;
        t_->zero_value = 1;
      }
    }
    __syncthreads();
  }

  // ================= final values: one sweep per distinct centre =================
  {
    if (threadIdx.x == 0) {
      int any = 0;
      for (int q = 0; q < NS; ++q) any |= st[q].zero_value ? 0 : 1;
      s_any = any;
    }
    __syncthreads();
    if (s_any) {
      double* buf = parity ? buf1 : buf0;
      for (int lead = 0; lead < NS; ++lead) {
        if (!is_lead(lead)) continue;
        const TrProblem& PL = M.P[lead];
        int sg, sp, sd;
        centre_slots(lead, sg, sp, sd);
        if (sg >= 0 && st[sg].zero_value) sg = -1;
        if (sp >= 0 && st[sp].zero_value) sp = -1;
        if (sd >= 0 && st[sd].zero_value) sd = -1;
        if (sg < 0 && sp < 0 && sd < 0) continue;
        const double tau_g = sg >= 0 ? st[sg].tau : 0.0, tau_p = sp >= 0 ? st[sp].tau : 0.0,
                     tau_d = sd >= 0 ? st[sd].tau : 0.0;
        double vgap[2] = {0.0, 0.0}, vhp[1] = {0.0}, vhd[1] = {0.0};
        if (sg >= 0 || sp >= 0) {
          for (int idx = tid; idx < n; idx += stride) {
            const double x0 = PL.px[idx];
            const double at = PL.atp[idx], cj = B.c[idx];
            const double g = PL.qxp ? (PL.qxp[idx] + cj) - at : cj - at;
            const double lb = B.l[idx], ub = B.u[idx];
            if (sg >= 0) {
              const double sol = jl_clamp(x0 + tau_g * scratch_d(sg)[idx], lb, ub);  // tr.jl:182-188
              vgap[0] += g * (sol - x0);
            }
            if (sp >= 0) {
              const double sol = jl_clamp(x0 + tau_p * scratch_d(sp)[idx], lb, ub);
              vhp[0] += g * (sol - x0);
            }
          }
        }
        if (sg >= 0 || sd >= 0) {
          for (int idx = n + tid; idx < total; idx += stride) {
            const int i = idx - n;
            const double x0 = PL.py[i];
            const double g = -(B.b[i] - PL.axp[i]);
            const double lb = i < B.neq ? -CUDART_INF : 0.0, ub = CUDART_INF;
            if (sg >= 0) {
              const double sol = jl_clamp(x0 + tau_g * scratch_d(sg)[idx], lb, ub);
              vgap[1] += g * (sol - x0);
            }
            if (sd >= 0) {
              const double sol = jl_clamp(x0 + tau_d * scratch_d(sd)[idx], lb, ub);
              vhd[0] += g * (sol - x0);
            }
          }
        }
        if (sg >= 0) trm_publish<2>(vgap, 0u, 2 * sg, buf, sh);
        if (sp >= 0) trm_publish<1>(vhp, 0u, 2 * sp, buf, sh);
        if (sd >= 0) trm_publish<1>(vhd, 0u, 2 * sd + 1, buf, sh);
      }
      trm_totals(grid, B, 2 * NS, 0ull, buf, tot, xc);
      if (threadIdx.x < NS && !st[threadIdx.x].zero_value) {
        const int s_ = threadIdx.x;
        st[s_].v_primal = M.P[s_].use_primal ? tot[2 * s_] : 0.0;
        st[s_].v_dual = M.P[s_].use_dual ? tot[2 * s_ + 1] : 0.0;
      }
      __syncthreads();
    }
  }
  if (blockIdx.x == 0) {
    if (threadIdx.x < NS) {
      st[threadIdx.x].exchanges = xc.count;
      trs[threadIdx.x] = st[threadIdx.x];
    }
    if (threadIdx.x == 0 && B.world > 1) *B.dseq = xc.seq - 1ull;
  }
}

int tr_multi_grid(int sm_count) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tr_multi, kTrmThreads, 0) != cudaSuccess ||
      per_sm < 1) {
    cudaGetLastError();
    return 0;
  }
  int g = sm_count * (per_sm < 2 ? per_sm : 2);
  if (g > 1024) g = 1024;
  return g;
}

int launch_tr_multi(const Bufs& B, const TrMulti& M, TrState* d_trs, int grid, cudaStream_t s) {
  double* part = B.part + static_cast<size_t>(kSlotEval) * kMaxScalars * kMaxPartialBlocks;
  void* args[] = {const_cast<Bufs*>(&B), const_cast<TrMulti*>(&M), &d_trs, &part};
  return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_tr_multi), dim3(grid), dim3(kTrmThreads), args,
                                     0, s);
}

int tr_solve_grid(int sm_count) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tr_solve, kTrThreads, 0) != cudaSuccess ||
      per_sm < 1)
    return 0;
  return sm_count * (per_sm < 2 ? per_sm : 2);
}

int launch_tr_solve(const Bufs& B, const TrProblem& P, TrState* d_trs, int grid,
                    unsigned long long seq_first, cudaStream_t s) {
  // the scratch of the evaluation slot holds 2 * kTrMaxK * grid doubles
  static_assert(2 * kTrMaxK <= kMaxScalars, "partials fit the evaluation slot");
  double* part = B.part + static_cast<size_t>(kSlotEval) * kMaxScalars * kMaxPartialBlocks;
  void* args[] = {const_cast<Bufs*>(&B), const_cast<TrProblem*>(&P), &d_trs, &part, &seq_first};
  return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_tr_solve), dim3(grid), dim3(kTrThreads),
                                     args, 0, s);
}

void launch_tr(const Bufs& B, const TrProblem& P, TrState* d_trs, int passes, bool init,
               cudaStream_t s) {
  if (init) k_tr_init<<<B.grid_vec, kVecThreads, 0, s>>>(B, P, d_trs, B.red);
  for (int p = 0; p < passes; ++p) k_tr_pass<<<B.grid_vec, kVecThreads, 0, s>>>(B, P, d_trs);
  k_tr_final<<<B.grid_vec, kVecThreads, 0, s>>>(B, P, d_trs);
}

void launch_tr_stage(const Bufs& B, const TrProblem& P, TrState* d_trs, int stage, cudaStream_t s) {
  if (stage == kTrInit) k_tr_init<<<B.grid_vec, kVecThreads, 0, s>>>(B, P, d_trs, B.red);
  else if (stage == kTrPass) k_tr_pass<<<B.grid_vec, kVecThreads, 0, s>>>(B, P, d_trs);
  else k_tr_final<<<B.grid_vec, kVecThreads, 0, s>>>(B, P, d_trs);
}
void launch_tr_combine(const Bufs& B, const TrProblem& P, TrState* d_trs, int stage,
                       const double* recv, cudaStream_t s) {
  k_tr_combine<<<1, 32, 0, s>>>(B, P, d_trs, stage, recv);
}

__global__ void k_exchange(Bufs B, const double* __restrict__ src, int count, unsigned long long seq,
                           int parity) {
  const int k = threadIdx.x;
  if (k < count) {
    const double v = src[k];
    const size_t slot = static_cast<size_t>(parity) * B.world * kScBlock + B.rank * kScBlock + k;
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
      if (r < B.world) B.hx_peer[r][slot] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    p2p_signal(B, 4, seq);
    p2p_wait(B, 4, seq);
  }
}
void launch_exchange(const Bufs& B, const double* src, int count, unsigned long long seq, int parity,
                     cudaStream_t s) {
  k_exchange<<<1, kScBlock, 0, s>>>(B, src, count, seq, parity);
}

__global__ void __launch_bounds__(kVecThreads) k_push_vec(Bufs B, const double* __restrict__ src, int which,
                                                          unsigned long long seq) {
  const int len = which == 0 ? B.n : B.m;
  const size_t off = which == 0 ? static_cast<size_t>(B.xbar_off) : static_cast<size_t>(B.rank) * B.m_pad;
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    const double v = src[i];
    if (B.xbar_mc != nullptr) {
      mc_store((which == 0 ? B.xbar_mc : B.yfull_mc) + off + i, v);
    } else {
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r)
        if (r < B.world) (which == 0 ? B.xbar_peer[r] : B.yfull_peer[r])[off + i] = v;
    }
  }
  if (last_block_arrive_sys(B.counters + 7) && threadIdx.x == 0) {
    p2p_signal(B, 4, seq);
    p2p_wait(B, 4, seq);
  }
}
void launch_push_vec(const Bufs& B, const double* src, int which, unsigned long long seq, cudaStream_t s) {
  const int len = which == 0 ? B.n : B.m;
  int g = (len + kVecThreads - 1) / kVecThreads;
  if (g > B.grid_vec / 2) g = B.grid_vec / 2;
  if (g < 1) g = 1;
  k_push_vec<<<g, kVecThreads, 0, s>>>(B, src, which, seq);
}

__global__ void k_combine_red(Bufs B, const double* __restrict__ recv, int off0, int count0, int nsum0,
                              int off1, int count1, int nsum1) {
  const int t = threadIdx.x;
  int off, k, nsum;
  if (t < count0) { off = off0; k = t; nsum = nsum0; }
  else if (t < count0 + count1) { off = off1; k = t - count0; nsum = nsum1; }
  else return;
  const int col = off - off0 + k;  // position inside the exchanged block (it starts at off0)
  double v = __ldcg(recv + col);
  for (int r = 1; r < B.world; ++r) {
    const double w = __ldcg(recv + r * kScBlock + col);
    v = k < nsum ? v + w : fmax(v, w);
  }
  B.red[off + k] = v;
}
void launch_combine_red(const Bufs& B, const double* recv, int off0, int count0, int nsum0, int off1,
                        int count1, int nsum1, cudaStream_t s) {
  k_combine_red<<<1, 2 * kScBlock, 0, s>>>(B, recv, off0, count0, nsum0, off1, count1, nsum1);
}

// ---------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kVecThreads) k_scale_div(const double* __restrict__ in,
                                                           const double* __restrict__ scale,
                                                           double* __restrict__ out, int len) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride)
    out[i] = scale ? in[i] / scale[i] : in[i];
}
void launch_scale_div(const double* in, const double* scale, double* out, int len, int grid,
                      cudaStream_t s) {
  k_scale_div<<<grid, kVecThreads, 0, s>>>(in, scale, out, len);
}
__global__ void __launch_bounds__(kVecThreads) k_fill(double* p, double v, int64_t len) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride)
    p[i] = v;
}
__global__ void __launch_bounds__(kVecThreads) k_copy(const double* __restrict__ in,
                                                       double* __restrict__ out, int64_t len) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride)
    out[i] = in[i];
}
void launch_copy(const double* in, double* out, int64_t len, cudaStream_t s) {
  if (len <= 0) return;
  int64_t g = (len + kVecThreads - 1) / kVecThreads;
  if (g > 1184) g = 1184;
  k_copy<<<static_cast<int>(g), kVecThreads, 0, s>>>(in, out, len);
}
void launch_fill(double* p, double v, int64_t len, cudaStream_t s) {
  if (len <= 0) return;
  int64_t g = (len + kVecThreads - 1) / kVecThreads;
  if (g > 1184) g = 1184;
  k_fill<<<static_cast<int>(g), kVecThreads, 0, s>>>(p, v, len);
}

}  // namespace folp
