/*
 * folp_b200.h -- C ABI of libfolp_b200.so, the B200 (sm_100a) PDHG inner loop
 * that replaces the body of
 *     FirstOrderLp.optimize(::PdhgParameters, ::QuadraticProgrammingProblem)
 * (reference: src/primal_dual_hybrid_gradient.jl:782-1049).
 *
 * The host (Julia, or the Python mirror in firstorderlp.jl_b200/) keeps doing
 * what primal_dual_hybrid_gradient.jl:786-859 does -- validate, cached norms,
 * rescale_problem, initial step size, initial primal weight -- and hands the
 * SCALED problem plus the scaling vectors to folp_create(). Everything from
 * primal_dual_hybrid_gradient.jl:862 to :1048 (iterates, take_step, weighted
 * averages, KKT statistics, restarts, primal-weight updates, termination
 * checks) runs behind this interface on the GPU(s).
 *
 * Plain C: only pointers, sizes and PODs cross the boundary. No exceptions, no
 * callbacks, no global mutable state (two handles may coexist). All input
 * pointers are borrowed for the duration of the call only.
 *
 * The same PODs (folp_problem, folp_params, folp_eval) are used by the CPU
 * oracle in oracle/ so that parity tests feed both sides identical bytes.
 */
#ifndef FOLP_B200_H
#define FOLP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (reference raises Julia error(); we return codes) ------ */
enum folp_status {
  FOLP_OK = 0,
  FOLP_INVALID_ARGUMENT = 1,
  FOLP_CUDA_ERROR = 2,
  FOLP_NCCL_ERROR = 3,
  FOLP_OUT_OF_MEMORY = 4,
  FOLP_UNSUPPORTED = 5
};

/* ---- enums: ordinals equal the Julia @enum ordinals ---------------------- */
/* src/solve_log.jl:32-37 */
enum folp_restart_choice {
  FOLP_RESTART_CHOICE_UNSPECIFIED = 0,
  FOLP_RESTART_CHOICE_NO_RESTART = 1,
  FOLP_RESTART_CHOICE_WEIGHTED_AVERAGE_RESET = 2,
  FOLP_RESTART_CHOICE_RESTART_TO_AVERAGE = 3
};
/* src/solve_log.jl:52-58 */
enum folp_point_type {
  FOLP_POINT_TYPE_UNSPECIFIED = 0,
  FOLP_POINT_TYPE_CURRENT_ITERATE = 1,
  FOLP_POINT_TYPE_ITERATE_DIFFERENCE = 2,
  FOLP_POINT_TYPE_AVERAGE_ITERATE = 3,
  FOLP_POINT_TYPE_NONE = 4
};
/* src/solve_log.jl:336-347 */
enum folp_termination_reason {
  FOLP_TERMINATION_REASON_UNSPECIFIED = 0, /* == "false": keep going */
  FOLP_TERMINATION_REASON_OPTIMAL = 1,
  FOLP_TERMINATION_REASON_PRIMAL_INFEASIBLE = 2,
  FOLP_TERMINATION_REASON_DUAL_INFEASIBLE = 3,
  FOLP_TERMINATION_REASON_TIME_LIMIT = 4,
  FOLP_TERMINATION_REASON_ITERATION_LIMIT = 5,
  FOLP_TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT = 6,
  FOLP_TERMINATION_REASON_NUMERICAL_ERROR = 7,
  FOLP_TERMINATION_REASON_INVALID_PROBLEM = 8,
  FOLP_TERMINATION_REASON_OTHER = 9
};
/* src/saddle_point.jl:325 */
enum folp_restart_scheme {
  FOLP_NO_RESTARTS = 0,
  FOLP_FIXED_FREQUENCY = 1,
  FOLP_ADAPTIVE_NORMALIZED = 2,
  FOLP_ADAPTIVE_LOCALIZED = 3,
  FOLP_ADAPTIVE_DISTANCE = 4
};
/* src/saddle_point.jl:340 */
enum folp_restart_to_current_metric {
  FOLP_NO_RESTART_TO_CURRENT = 0,
  FOLP_GAP_OVER_DISTANCE = 1,
  FOLP_GAP_OVER_DISTANCE_SQUARED = 2
};
/* src/termination.jl:15 */
enum folp_optimality_norm { FOLP_L_INF = 0, FOLP_L2 = 1 };
/* the Union at src/primal_dual_hybrid_gradient.jl:194-198 */
enum folp_step_size_policy {
  FOLP_STEP_ADAPTIVE = 0,      /* AdaptiveStepsizeParams        :60-63  */
  FOLP_STEP_MALITSKY_POCK = 1, /* MalitskyPockStepsizeParameters :19-41 */
  FOLP_STEP_CONSTANT = 2       /* ConstantStepsizeParams         :68    */
};

/* ---- problem ------------------------------------------------------------- */
/*
 * The SCALED QuadraticProgrammingProblem (src/quadratic_programming.jl:34-76,
 * produced by rescale_problem, src/preprocess.jl:631) in Julia's own CSC
 * layout, plus what evaluate_unscaled_iteration_stats
 * (src/iteration_stats_utils.jl:413-451) needs from the ORIGINAL problem.
 * index_base = 1 when the arrays come straight from a Julia SparseMatrixCSC
 * (colptr/rowval are 1-based Int64), 0 for C/NumPy callers.
 */
typedef struct folp_problem {
  int64_t num_variables;   /* n */
  int64_t num_constraints; /* m */
  int64_t num_nonzeros;    /* nnz(constraint_matrix) */
  int64_t num_equalities;  /* rows [0, num_equalities) are equalities */
  int32_t index_base;      /* 0 or 1 */
  int32_t reserved0;
  const int64_t* colptr;   /* n+1 */
  const int64_t* rowval;   /* nnz, ascending inside each column */
  const double* nzval;     /* nnz, scaled matrix E^-1 A D^-1 */
  const double* objective_vector;     /* n, scaled c ./ D */
  const double* variable_lower_bound; /* n, scaled D .* l, -Inf allowed */
  const double* variable_upper_bound; /* n, scaled D .* u, +Inf allowed */
  const double* right_hand_side;      /* m, scaled b ./ E */
  double objective_constant;
  const double* variable_rescaling;   /* n, D (ScaledQpProblem :293-298) */
  const double* constraint_rescaling; /* m, E */
  /* original problem (scaled_problem.original_qp) */
  const double* orig_objective_vector;     /* n */
  const double* orig_variable_lower_bound; /* n */
  const double* orig_variable_upper_bound; /* n */
  const double* orig_right_hand_side;      /* m */
  /* nnz values of the ORIGINAL matrix on the same sparsity pattern. Only the
   * CPU oracle reads it (it restates the reference literally); the GPU
   * library derives A_orig * x_hat = E .* (A_scaled * x_scaled) instead and
   * accepts NULL here. */
  const double* orig_nzval;
  /* objective matrix Q of the scaled problem (CSC, same index_base) and the
   * values of the original Q on the same pattern. q_num_nonzeros = 0 for an
   * LP. */
  int64_t q_num_nonzeros;
  const int64_t* q_colptr;
  const int64_t* q_rowval;
  const double* q_nzval;
  const double* q_orig_nzval;
  /* CachedQuadraticProgramInfo of the ORIGINAL problem,
   * src/termination.jl:144-158 */
  double l_inf_norm_primal_linear_objective;
  double l_inf_norm_primal_right_hand_side;
  double l2_norm_primal_linear_objective;
  double l2_norm_primal_right_hand_side;
} folp_problem;

/* ---- parameters ---------------------------------------------------------- */
/* Flat mirror of PdhgParameters (src/primal_dual_hybrid_gradient.jl:128-199),
 * TerminationCriteria (src/termination.jl:29-98) and RestartParameters
 * (src/saddle_point.jl:342-400). The rescaling fields of PdhgParameters are
 * consumed by the host; their results arrive as initial_step_size /
 * initial_primal_weight / initial_kkt_passes (:821-857). */
typedef struct folp_params {
  int32_t step_size_policy;      /* folp_step_size_policy */
  int32_t termination_evaluation_frequency;
  double reduction_exponent;     /* adaptive  */
  double growth_exponent;        /* adaptive  */
  double downscaling_factor;     /* Malitsky-Pock */
  double breaking_factor;        /* Malitsky-Pock */
  double interpolation_coefficient; /* Malitsky-Pock */
  double initial_step_size;      /* :823/:826/:836 */
  double initial_primal_weight;  /* :847-857 */
  double initial_kkt_passes;     /* 0.5 or #power iterations, :822/:825/:838 */
  /* TerminationCriteria */
  int32_t optimality_norm;       /* folp_optimality_norm */
  int32_t iteration_limit;       /* Int32 in the reference */
  double eps_optimal_absolute;
  double eps_optimal_relative;
  double eps_primal_infeasible;
  double eps_dual_infeasible;
  double time_sec_limit;
  double kkt_matrix_pass_limit;
  /* RestartParameters */
  int32_t restart_scheme;            /* folp_restart_scheme */
  int32_t restart_to_current_metric; /* folp_restart_to_current_metric */
  int64_t restart_frequency_if_fixed;
  double artificial_restart_threshold;
  double sufficient_reduction_for_restart;
  double necessary_reduction_for_restart;
  double primal_weight_update_smoothing;
  int32_t use_approximate_localized_duality_gap;
  int32_t record_iteration_stats;
  int32_t verbosity;                 /* host-side only; carried for logging */
  int32_t reserved0;
} folp_params;

/* ---- one evaluation = one IterationStats --------------------------------- */
/* Field order follows src/solve_log.jl: ConvergenceInformation :64-164,
 * InfeasibilityInformation :174-221, IterationStats :232-300. */
typedef struct folp_eval {
  int32_t iteration_number;          /* = loop counter - 1 (isu.jl:442) */
  int32_t candidate_type;            /* folp_point_type; PDHG: AVERAGE */
  /* ConvergenceInformation */
  double primal_objective;
  double dual_objective;
  double corrected_dual_objective;   /* dual_objective if l_inf_dual_residual == 0.0 exactly, else -Inf (isu.jl:203-212).
                                      * The residual is formed from the SCALED products times D/E, not from the original
                                      * matrix: where the reference's residual is a rounding-sized nonzero (or the other
                                      * way round) this exact test can fall on the other side. Tolerance-sensitive. */
  double l_inf_primal_residual;
  double l2_primal_residual;
  double l_inf_dual_residual;
  double l2_dual_residual;
  double relative_l_inf_primal_residual;
  double relative_l2_primal_residual;
  double relative_l_inf_dual_residual;
  double relative_l2_dual_residual;
  double relative_optimality_gap;
  double l_inf_primal_variable;
  double l2_primal_variable;
  double l_inf_dual_variable;
  double l2_dual_variable;
  /* InfeasibilityInformation */
  double max_primal_ray_infeasibility;
  double primal_ray_linear_objective;
  double primal_ray_quadratic_norm;
  double max_dual_ray_infeasibility;
  double dual_ray_objective;
  /* IterationStats scalars */
  double cumulative_kkt_matrix_passes;
  double cumulative_time_sec;
  double step_size;                  /* before this evaluation's restart */
  double primal_weight;              /* before this evaluation's restart */
  /* method_specific_stats (pdhg.jl:929, sp.jl:1041-1046) */
  double time_spent_doing_basic_algorithm;
  double lagrangian_value;
  double estimated_lower_bound;
  double estimated_upper_bound;
  int32_t cumulative_rejected_steps; /* declared, never written: always 0 */
  int32_t restart_used;              /* folp_restart_choice */
  int32_t termination_reason;        /* folp_termination_reason; 0 = go on */
  int32_t numerical_error;           /* solver_state.numerical_error */
  int64_t total_number_iterations;   /* includes rejected inner attempts */
} folp_eval;

typedef struct folp_handle folp_handle;

/* Multi-GPU: one process per GPU (torchrun). Each rank passes the FULL scaled
 * problem; the library keeps only its shard: a contiguous block of constraint
 * rows balanced by nonzeros (A[rows,:] in CSR, y, b, dual averages) for A*xbar
 * and the dual step, and an equal slice of the columns (A[:,slice]' in CSR, x,
 * c, l, u, A'y, primal averages) for A'*y and the primal step. Per take_step
 * attempt the ranks exchange the extrapolated primal (allgather), the new dual
 * iterate (allgather) and four step-rule scalars (summed in rank order so that
 * every rank takes the same decision); no product needs a cross-rank reduction. Every rank must make the same
 * sequence of calls; records and solutions returned are the global ones on
 * every rank. nccl_unique_id is the 128-byte ncclUniqueId created on rank 0 by
 * folp_nccl_unique_id() and broadcast by the host's own plumbing
 * (torch.distributed); NULL when world_size == 1. NCCL is bound at run time
 * (dlopen of libnccl.so.2), so a single-GPU host needs none. */
typedef struct folp_dist {
  int32_t rank;
  int32_t world_size;
  int32_t device;          /* CUDA device ordinal for this rank */
  int32_t reserved0;
  const void* nccl_unique_id; /* 128 bytes or NULL */
} folp_dist;

/* Fills 128 bytes. */
int folp_nccl_unique_id(void* out128);

/* The partition folp_create applies, as pure host arithmetic (no device needed):
 * row_begin_out[world_size + 1] and col_begin_out[world_size + 1] receive the
 * first global row / column of every rank. rowval: the CSC row indices. */
int folp_partition(int64_t num_constraints, int64_t num_variables, int64_t num_nonzeros,
                   const int64_t* rowval, int32_t index_base, int32_t world_size,
                   int64_t* row_begin_out, int64_t* col_begin_out);

/* How take_step exchanges data between ranks: 0 = single GPU (none), 1 = NCCL
 * collectives, 2 = peer memory (CUDA IPC over NVLink: K1 pushes its slice of
 * xbar, K2 its rows of y+ into every rank's copy; no NCCL call per attempt),
 * 3 = peer memory bound to an NVSwitch multicast object (NVLS): one multimem.st
 * per element lands in every rank's copy (FOLP_NO_MULTICAST=1 switches it off). */
int folp_exchange_mode(folp_handle* h);

/* What this rank holds: rows [row_begin,row_end), primal slice [col_begin,col_end). */
int folp_shard_info(folp_handle* h, int64_t* row_begin, int64_t* row_end, int64_t* col_begin,
                    int64_t* col_end, int64_t* local_nonzeros);

/* Replaces pdhg.jl:805-873: uploads, builds the device layouts (tiled CSR of A
 * and of A'), zero iterates, RestartInfo, averages. dist may be NULL
 * (single GPU, current device). */
int folp_create(const folp_problem* problem, const folp_params* params,
                const folp_dist* dist, folp_handle** out);

/* Single-process multi-GPU entry: ONE call from ONE host
 * thread -- what FirstOrderLp.optimize (pdhg.jl:782-785) is -- drives n_gpus
 * devices of this process (device_ids[0..n_gpus), NULL = devices 0..n_gpus-1;
 * n_gpus <= 8). The problem is partitioned exactly as with one process per GPU
 * (folp_partition); each device is driven by a host thread of the library, the
 * devices reach each other's exchange regions through direct peer access
 * (cudaDeviceEnablePeerAccess: no CUDA IPC, no NCCL, no communicator to set up),
 * and the returned handle is used with folp_run / folp_solve / folp_get_solution /
 * folp_destroy like any other: records and solutions are the global ones.
 * FOLP_UNSUPPORTED if some pair of the devices lacks peer access. n_gpus == 1 is
 * folp_create on device_ids[0]. */
int folp_create_multi(const folp_problem* problem, const folp_params* params,
                      int32_t n_gpus, const int32_t* device_ids, folp_handle** out);

/* Replaces one trip round the while-loop of pdhg.jl:886-1048 up to and
 * including the next evaluation block (:892-1023): runs take_step until the
 * trigger of :892-895 fires, evaluates the KKT statistics of the average
 * iterate on the original problem, the objective-bound estimates, the
 * termination criteria and -- if not terminating -- the restart scheme and
 * primal-weight update. Fills *out. out->termination_reason != 0 means the
 * reference would return here (:972-993). */
int folp_run(folp_handle* h, folp_eval* out);

/* Loops folp_run until termination. evals may be NULL; otherwise up to
 * max_evals records are stored following record_iteration_stats semantics
 * (pdhg.jl:958-960). *num_evals receives the number of records PRODUCED: if it
 * exceeds max_evals the history was truncated (the first max_evals - 1 records
 * and, in the last slot, the final one are kept) -- size the buffer from
 * iteration_limit / termination_evaluation_frequency + 16 when
 * record_iteration_stats is set, 1 otherwise. x_out (n) / y_out (m) receive the
 * UNSCALED average iterate (sp.jl:55-77). */
int folp_solve(folp_handle* h, folp_eval* evals, int64_t max_evals,
               int64_t* num_evals, int32_t* termination_reason,
               int32_t* iteration_count, double* x_out, double* y_out);

/* which: 0 = average iterate as the reference would return it (current when no
 * average exists yet), 1 = current iterate. unscaled != 0 divides by the
 * rescaling vectors (sp.jl:65-67). */
int folp_get_solution(folp_handle* h, int which, int unscaled, double* x_out,
                      double* y_out);

/* Test hooks: run exactly `attempts` inner attempts of take_step (accepted or
 * not) without any evaluation, and read back the raw solver state
 * (PdhgSolverState, pdhg.jl:205-258). dual_product is A' y. */
int folp_debug_attempts(folp_handle* h, int64_t attempts);
typedef struct folp_debug_scalars {
  double step_size, primal_weight, cumulative_kkt_passes;
  double sum_primal_solution_weights, sum_dual_solution_weights;
  int64_t total_number_iterations, iterations_completed;
  int64_t sum_primal_solutions_count, sum_dual_solutions_count;
  int32_t numerical_error, reserved0;
  double last_interaction, last_movement;
} folp_debug_scalars;
int folp_debug_state(folp_handle* h, double* x, double* y, double* dual_product,
                     double* sum_x, double* sum_y, folp_debug_scalars* s);
int folp_debug_set_state(folp_handle* h, const double* x, const double* y,
                         double step_size, double primal_weight);

/* y_out(m) = A x (transpose == 0) or x_out(n) = A' y (transpose != 0) with the
 * device layouts; the SpMV kernels in isolation (parity + roofline probes). */
int folp_debug_spmv(folp_handle* h, int transpose, const double* in,
                    double* out);

/* Counters for bench.py: kernels launched so far, device seconds spent in
 * take_step (the reference's time_spent_doing_basic_algorithm), and the
 * number of outer iterations completed. */
int folp_counters(folp_handle* h, int64_t* kernel_launches,
                  double* basic_algorithm_seconds, int64_t* iterations);

/* Measurement hook for bench.py: runs `attempts` real take_step attempts
 * (un-graphed) with CUDA events on the library's stream around each kernel, and
 * returns the accumulated device milliseconds in ms_out[8]: single GPU
 * {primal step, A*xbar + dual step, A'*y + interaction/step rule};
 * partitioned {primal slice (+ xbar exchange), A[rows,:]*xbar + dual step (+ y
 * exchange), A[:,slice]'*y + interaction (+ scalar exchange), step rule}; the
 * rest 0. *attempts_run = attempts that did work. The solver state advances. */
int folp_debug_profile_attempts(folp_handle* h, int64_t attempts, double ms_out[8],
                                int64_t* attempts_run);

/* Measurement hook: average device milliseconds of `reps` launches of the plain
 * SpMV kernel (A * x when transpose == 0, A' * y otherwise) on the live iterate. */
int folp_debug_time_spmv(folp_handle* h, int transpose, int reps, double* ms_out);

/* Test hook without any CUDA call: packs a 0-based CSR matrix into the library's
 * work-item layout exactly as folp_create does and evaluates y = A * x on the
 * HOST by walking that layout with the kernel's own slot arithmetic, so that the
 * packing (position-major groups, length-sorted windows, wide rows, long-row
 * chunks) can be verified on a machine without a GPU. warps_total = warps of
 * the grid the kernel would run on (<= 0: a full B200). stats (5 entries, may be
 * NULL) = {work items, length-sorted groups, narrow-group rounds, long rows,
 * rounds of the busiest warp}. It is not a solver path: nothing in the library
 * calls it. */
int folp_debug_host_spmv(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colidx,
                         const double* vals, const double* x, double* y, int64_t warps_total,
                         int64_t* stats);

/* rescale_problem (src/preprocess.jl:631-687) computed on the current CUDA device
 * (SURVEY.md section 8f-1): Ruiz rescaling in the infinity norm (ruiz_p = 0) or
 * the 2-norm (ruiz_p = 2) (:412-477), l2-norm rescaling (:358-372) and
 * Pock-Chambolle rescaling (:508-539; pock_chambolle_alpha < 0 = nothing), applied
 * in that order. Same contract as the reference function on a copied problem: the
 * CSC values of A (and Q, q_colptr may be NULL for an LP) and the vectors c, l, u, b
 * are rescaled IN PLACE (scale_problem, :555-573), and the cumulative
 * constraint_rescaling (m) / variable_rescaling (n) come back, from which the
 * host forms ScaledQpProblem(original, scaled, constraint_rescaling,
 * variable_rescaling). Indices are Int64 with index_base 0 or 1. Bit-identical to
 * the reference arithmetic for rows and columns of up to 2048 entries and
 * pock_chambolle_alpha = 1 (other exponents go through pow); see folp_rescale.cu. */
int folp_rescale_problem(int64_t num_constraints, int64_t num_variables, int32_t index_base,
                         const int64_t* colptr, const int64_t* rowval, double* nzval,
                         const int64_t* q_colptr, const int64_t* q_rowval, double* q_nzval,
                         double* objective_vector, double* variable_lower_bound,
                         double* variable_upper_bound, double* right_hand_side,
                         int32_t l_inf_ruiz_iterations, int32_t ruiz_p, int32_t l2_norm_rescaling,
                         double pock_chambolle_alpha, double* constraint_rescaling,
                         double* variable_rescaling);

/* Test / measurement hook without any CUDA call: the host half of folp_create on
 * one GPU (transposition of the caller's CSC into the CSR of A, work-item planning
 * and position-major packing of both matrices). *milliseconds = its wall-clock
 * time. FOLP_INVALID_ARGUMENT if a row index is out of range. */
int folp_debug_host_prepare(const folp_problem* problem, double* milliseconds);

/* Test hook without any CUDA call: the host half of folp_create (transposition
 * included) followed by y = A * x (transpose == 0; x has n entries, y has m) or
 * y = A' * x evaluated on the HOST through the packed layouts, as
 * folp_debug_host_spmv does. */
int folp_debug_host_problem_spmv(const folp_problem* problem, int transpose, const double* x,
                                 double* y);

/* The CUDA stream (cudaStream_t) all kernels of this handle are launched on,
 * so a caller can bracket calls with its own events. */
void* folp_debug_stream(folp_handle* h);

void folp_destroy(folp_handle* h);

/* Last error message of this handle (or of the failed folp_create when h is
 * NULL). Never NULL. */
const char* folp_last_error(const folp_handle* h);

/* "sm_100a;cuda 12.9;..." build description. */
const char* folp_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* FOLP_B200_H */
