// folp_kernels.cuh -- launch interface between folp_api.cu (host logic) and
// folp_kernels.cu (device code).
#pragma once
#include "folp_internal.cuh"

namespace folp {

constexpr int kMaxWorld = 8;
constexpr int kNumFlagKinds = 5;  // xbar pushed | y+ pushed | step-rule scalars pushed | device-sequenced evaluation scalars (k_tr_solve) | host-sequenced evaluation exchanges

// Device pointers of one rank. Lengths: n_loc primal slice, m_loc dual rows.
struct Bufs {
  int n = 0, m = 0, neq = 0;  // local primal slice length, local rows, local equalities
  DevState* st = nullptr;
  double *x[2] = {nullptr, nullptr}, *y[2] = {nullptr, nullptr}, *aty[2] = {nullptr, nullptr};
  double* xbar = nullptr;                    // extrapolated primal, gathered by A*xbar
  double *c = nullptr, *l = nullptr, *u = nullptr, *b = nullptr;  // scaled problem
  double *sum_x = nullptr, *sum_y = nullptr;
  // evaluation
  double *avg_x = nullptr, *avg_y = nullptr, *ax_avg = nullptr, *aty_avg = nullptr;
  double* ax_cur = nullptr;
  double *last_x = nullptr, *last_y = nullptr, *last_ax = nullptr, *last_aty = nullptr;
  double *D = nullptr, *E = nullptr;         // variable_rescaling, constraint_rescaling
  double *c_orig = nullptr, *l_orig = nullptr, *u_orig = nullptr, *b_orig = nullptr;
  double *tr_t = nullptr, *tr_d = nullptr;   // trust-region scratch, n+m each
  double *trm_t = nullptr, *trm_d = nullptr; // k_tr_multi scratch, 4 * (n+m) each (nullptr: not allocated)
  // ---- quadratic objective (has_q: the scaled objective matrix has a nonzero entry) ----
  // qx[k] = Q * x[k] travels with the iterate buffers (same parity), dxv holds x+ - x of the
  // running attempt, qx_avg / last_qx are Q times the evaluated / the last restart point.
  int has_q = 0;
  double *qx[2] = {nullptr, nullptr}, *dxv = nullptr, *qx_avg = nullptr, *last_qx = nullptr;
  // reductions
  double* part = nullptr;     // kMaxScalars * kMaxPartialBlocks doubles per slot, 4 slots
  double* red = nullptr;      // reduced scalars, kMaxScalars per slot
  unsigned* counters = nullptr;
  int grid_vec = 0, grid_spmv = 0;
  // ---- row-partitioned mode (world > 1); see DESIGN.md "Multi-GPU" ----
  // n / m above are then the local primal slice length and the local row count;
  // xbar is the FULL extrapolated primal (world * n_pad doubles) of which this rank
  // writes [xbar_off, xbar_off + n).
  int world = 1, rank = 0;
  int xbar_off = 0;
  double* y_full = nullptr;    // the FULL new dual iterate, world * m_pad doubles (rank-major, padded)
  int m_pad = 0;               // rows per rank in y_full; this rank writes [rank*m_pad, rank*m_pad + m)
  double* col_tmp = nullptr;   // scratch, one padded primal slice
  double* sc_send = nullptr;   // kScBlock scalars this rank contributes to an exchange
  double* sc_recv = nullptr;   // world * kScBlock scalars, rank-major
  // ---- peer-memory exchange (CUDA IPC over NVLink / NVSwitch), replaces NCCL inside take_step ----
  // Every rank exposes one region {xbar, y_full, sc_recv, scx, flags}; *_peer[r] is rank r's copy
  // (this rank's own entries are the local pointers above).
  int p2p = 0;
  int dbg = 0;  // development probes (FOLP_DEBUG_FLAGS): 1 = skip the remote pushes, 2 = skip flag waits, 4 = private gather copies,
                // 8 = copy-engine (cp.async.bulk) pushes of xbar tiles / y+ groups (also FOLP_BULK_PUSH=1), 16 = relaxed flag polling + one fence
  const double* xbar_priv = nullptr;   // when set: K2 gathers from this private copy of xbar
  const double* yfull_priv = nullptr;  // when set: K3 gathers from this private copy of y_full
  double* xbar_peer[kMaxWorld] = {};
  double* yfull_peer[kMaxWorld] = {};
  // NVSwitch multicast views of xbar / y_full (folp_vmm.h; nullptr = off): ONE multimem.st lands in every
  // rank's copy, this rank's included, instead of world-1 unicast stores
  double* xbar_mc = nullptr;
  double* yfull_mc = nullptr;
  double* sc_peer[kMaxWorld] = {};
  unsigned long long* flag_peer[kMaxWorld] = {};
  unsigned long long* flags = nullptr;  // local flags [kNumFlagKinds][kMaxWorld]: kind k, source rank r
  // evaluation-block scalar exchanges: two alternating receive buffers of world * kScBlock doubles
  double* scx = nullptr;                // local
  double* scx_peer[kMaxWorld] = {};
  // the same for the exchanges the HOST numbers (k_exchange, k_push_vec: flag kind 4); the ones above belong
  // to k_tr_solve, which numbers its own (flag kind 3) from the device-resident counter dseq, so that a
  // solve can be enqueued without the host knowing how many passes the previous one took
  double* hx = nullptr;
  double* hx_peer[kMaxWorld] = {};
  unsigned long long* dseq = nullptr;   // exchanges made by k_tr_solve so far (the same on every rank)
  unsigned long long p2p_timeout_ns = 30000000000ull;  // a peer wait gives up after this long (FOLP_P2P_TIMEOUT_MS)
  // ---- persistent take_step kernel (k_take_steps): grid barrier words and phase timers ----
  unsigned* bar_top = nullptr;          // arrival counter of the grid barrier
  unsigned long long* bar_gen = nullptr;  // release words (kBarGenStride apart): [0] master, [1 + g] group g: generation of the last completed barrier | abort bit
  unsigned long long* timers = nullptr;   // [0] last stamp, [1..3] ns spent in the phases, [4] attempts (nullptr: off)
};
constexpr int kScBlock = 64;

constexpr int kSlotPrimal = 0, kSlotDual = 1, kSlotTrans = 2, kSlotEval = 3;
constexpr int kNumSlots = 4;

// n-pass statistics (isu.jl:228-349 on the original problem + Lagrangian pieces)
enum StatN {
  SN_cx = 0, SN_lviol2, SN_uviol2, SN_dres2, SN_rcobj, SN_x2, SN_ray_rcobj, SN_cs_x, SN_x_aty,
  SN_xs2, SN_xqx, SN_xs_qxs, SN_NSUM,
  SN_lviol_max = SN_NSUM, SN_uviol_max, SN_dres_max, SN_x_max, SN_ray_l_max, SN_ray_u_max,
  SN_ray_dres_max, SN_ray_rc_max, SN_qx_max, SN_TOTAL
};
static_assert(SN_TOTAL <= kMaxScalars, "statistics block fits one reduction slot");
enum StatM {
  SM_pres2 = 0, SM_by, SM_y2, SM_yneg2, SM_bs_y, SM_ys2, SM_NSUM,
  SM_pres_max = SM_NSUM, SM_y_max, SM_yneg_max, SM_ray_act_max, SM_TOTAL
};
enum StatDist { SD_avg_x = 0, SD_cur_x, SD_avg_y, SD_cur_y, SD_TOTAL };

// Trust-region solve state (tr.jl:68-224), device resident.
struct TrState {
  double radius, r2;
  double lo, L_lo, H_lo;
  long long cnt_lo;
  double hi;
  double cand[2];
  double tau;        // final threshold
  double max_t;
  int done;          // 1: tau is final; 2: the search was abandoned (non-finite data)
  int zero_value;    // 1: solution == center (value 0)
  int approx;
  int passes;
  double approx_scale;
  // Lagrangian pieces and results
  double cx, x_aty, y_b;
  double v_primal, v_dual;  // objective_vector . (solution - center), per segment
  double xqx;               // center' * Q * center (0 for an LP)
  int exchanges;            // cross-rank scalar exchanges made by k_tr_solve (partitioned mode)
  int reserved0;
};

// Where k_tr_solve takes the weights / radius from: the struct as the host filled it, or -- so that
// a whole evaluation block can be enqueued without a host round trip -- the reduced statistics of
// the kernels that ran before it on the same stream (Bufs::red):
//   kTrParamBounds  update_objective_bound_estimates (sp.jl:1015-1047): wp, wd are the norm weights;
//                   the kernel divides them by max(1e-8, sqrt(wp |x_avg|^2))^2 resp. the dual analogue
//                   (k_stats_n / k_stats_m sums), radius 1
//   kTrParamDistAvg / kTrParamDistCur  compute_localized_duality_gaps (sp.jl:432-496): radius = weighted
//                   distance of the average / current iterate to the last restart point (k_dist sums)
enum TrParamSrc : int { kTrParamHost = 0, kTrParamBounds = 1, kTrParamDistAvg = 2, kTrParamDistCur = 3 };
struct TrProblem {
  const double *px, *atp;   // primal center and A' * dual center
  const double *py, *axp;   // dual center and A * primal center
  double wp, wd, radius;
  int use_primal, use_dual, approx;
  const double* qxp;        // Q * primal center; nullptr for an LP
  int param_src = kTrParamHost;
  int reserved0 = 0;
};
constexpr int kTrSlots = 5;
// all trust-region problems of one evaluation block for k_tr_multi (slot k's result goes to d_trs[k])
struct TrMulti {
  TrProblem P[kTrSlots];
  int nslots;
  int reserved0;
};
int tr_multi_grid(int sm_count);  // co-resident blocks of k_tr_multi (0: unavailable)
int launch_tr_multi(const Bufs& B, const TrMulti& M, TrState* d_trs, int grid, cudaStream_t s);  // cudaError_t  // results of one evaluation block: bound estimates (primal, dual), gaps at avg / current / last restart

// Q: CSR of the objective matrix, used only when B.has_q
void launch_step_attempts(const Bufs& B, const SpmvMat& A, const SpmvMat& At, const SpmvMat& Q,
                          int attempts, cudaStream_t s);
// The same attempts as ONE cooperative launch (k_take_steps): the three phases of an attempt are
// separated by grid barriers instead of kernel boundaries, the state lives in shared memory, and in
// partitioned mode the peer exchanges ride on the barriers. take_steps_grid: co-resident CTAs the
// kernel may use on this device (0: cooperative launch unavailable). Returns a cudaError_t.
int take_steps_grid(int sm_count, bool dist);
int launch_take_steps(const Bufs& B, const SpmvMat& A, const SpmvMat& At, const SpmvMat& Q, int attempts,
                      int grid, cudaStream_t s);
int launch_take_steps_cluster(const Bufs& B, const SpmvMat& A, const SpmvMat& At, const SpmvMat& Q, int attempts,
                              int grid, cudaStream_t s);
constexpr int kTakeClusterMax = 8;      // portable cluster size: CTAs of the cluster form of k_take_steps
constexpr int kBarGroupSize = 32;       // CTAs polling one release word of the grid barrier
constexpr int kBarL1Stride = 32;        // unsigned words between first-level counters (128 bytes)
constexpr int kBarMaxGroups = 64;
constexpr int kBarGenStride = 16;       // unsigned long long words between release words (128 bytes)
// The pieces of one attempt in partitioned mode; without peer memory folp_api.cu interleaves them
// with the NCCL exchanges (allgather xbar | allgather y+ | allgather of the 4 step-rule scalars).
void launch_dist_primal(const Bufs& B, cudaStream_t s);
void launch_dist_dual(const Bufs& B, const SpmvMat& A, cudaStream_t s);
void launch_dist_trans(const Bufs& B, const SpmvMat& A, const SpmvMat& At, cudaStream_t s);
void launch_dist_finalize(const Bufs& B, cudaStream_t s);
// one attempt with events ev[0..3] recorded before/between/after the three kernels
void launch_step_attempt_timed(const Bufs& B, const SpmvMat& A, const SpmvMat& At, const SpmvMat& Q,
                               cudaEvent_t* ev, cudaStream_t s);
void launch_spmv_plain(const SpmvMat& A, const double* in, double* out, int grid, cudaStream_t s);
void launch_copy(const double* in, double* out, int64_t len, cudaStream_t s);
void launch_flush_avg(const Bufs& B, cudaStream_t s);
void launch_make_avg(const Bufs& B, int use_current, cudaStream_t s);
// results in B.red + kSlotEval*kMaxScalars ... see folp_api.cu
void launch_stats_n(const Bufs& B, double* red_out, cudaStream_t s);
void launch_stats_m(const Bufs& B, double* red_out, cudaStream_t s);
void launch_dist(const Bufs& B, double* red_out, cudaStream_t s);
void launch_apply_restart(const Bufs& B, int to_average, int have_ax_cur, cudaStream_t s);
void launch_tr(const Bufs& B, const TrProblem& P, TrState* d_trs, int passes, bool init,
               cudaStream_t s);
// single GPU: the whole solve (init, passes to convergence, final value) as one cooperative
// kernel on `grid` co-resident blocks; returns a cudaError_t
// (partitioned mode over peer memory: every rank launches it; seq_first = number of the first
// scalar exchange it will make, TrState::exchanges = how many it made)
int launch_tr_solve(const Bufs& B, const TrProblem& P, TrState* d_trs, int grid,
                    unsigned long long seq_first, cudaStream_t s);
int tr_solve_grid(int sm_count);  // co-resident blocks for launch_tr_solve (0: unavailable)
// row-partitioned mode: one kernel of the trust-region solve at a time; each leaves its
// local sums in B.sc_send, and after the host's scalar exchange launch_tr_combine applies
// the rank-ordered totals to the TrState.
enum TrStage : int { kTrInit = 0, kTrPass = 1, kTrFinal = 2 };
void launch_tr_stage(const Bufs& B, const TrProblem& P, TrState* d_trs, int stage, cudaStream_t s);
void launch_tr_combine(const Bufs& B, const TrProblem& P, TrState* d_trs, int stage,
                       const double* recv, cudaStream_t s);
// peer-memory scalar exchange of the evaluation block: pushes src[0..count) into slot `rank` of
// every rank's receive buffer `parity` and waits until all ranks have done the same (`seq` is the
// host-side count of such exchanges, identical on every rank)
void launch_exchange(const Bufs& B, const double* src, int count, unsigned long long seq, int parity,
                     cudaStream_t s);
// peer-memory allgather of the evaluation block: this rank's slice of a primal-indexed vector
// (which = 0: into every rank's xbar staging) or its rows of a dual-indexed one (which = 1: into
// y_full), then the same flag exchange; when the kernel has completed the full vector is local.
void launch_push_vec(const Bufs& B, const double* src, int which, unsigned long long seq, cudaStream_t s);
// Combines the per-rank blocks of reduced scalars an exchange delivered (recv: world * kScBlock,
// rank-major) in rank order into B.red: two segments [off, off + count), the first nsum entries of a
// segment are sums, the rest maxima (count1 = 0: one segment).
void launch_combine_red(const Bufs& B, const double* recv, int off0, int count0, int nsum0, int off1,
                        int count1, int nsum1, cudaStream_t s);
void launch_scale_div(const double* in, const double* scale, double* out, int len, int grid,
                      cudaStream_t s);
void launch_fill(double* p, double v, int64_t len, cudaStream_t s);
int spmv_configure();  // sets the dynamic shared memory attribute; returns cudaError_t

}  // namespace folp
