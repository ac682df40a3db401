/*
 * folp_oracle.h -- CPU oracle for the PDHG path of FirstOrderLp.jl.
 *
 * TEST INFRASTRUCTURE ONLY. This is a plain-C, single-threaded restatement of
 * the reference algorithm (file:line citations in folp_oracle.c). Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it; the product (libfolp_b200.so and the Python host mirror) never
 * does.
 *
 * Parity pinning: the reference itself cannot run here (no Julia in the image)
 * so the oracle is pinned against every known-answer test the reference holds
 * for this path (tests/test_oracle_*.py port test/test_primal_dual_hybrid_
 * gradient.jl, test_iteration_stats.jl, test_termination.jl,
 * test_trust_region_utils.jl, test_saddle_point.jl, test_qp_processing.jl).
 * Those pin answers, not bit patterns: BLAS nrm2/dot summation order of the
 * Julia stdlib is not reproducible, so sums here are plain left-to-right.
 */
#ifndef FOLP_ORACLE_H
#define FOLP_ORACLE_H

#include "../include/folp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_handle oracle_handle;

/* Same semantics as folp_create / folp_run / folp_solve / folp_get_solution /
 * folp_debug_* in include/folp_b200.h, computed the reference's way. */
int oracle_create(const folp_problem* problem, const folp_params* params,
                  oracle_handle** out);
int oracle_run(oracle_handle* h, folp_eval* out);
int oracle_solve(oracle_handle* h, folp_eval* evals, int64_t max_evals,
                 int64_t* num_evals, int32_t* termination_reason,
                 int32_t* iteration_count, double* x_out, double* y_out);
int oracle_get_solution(oracle_handle* h, int which, int unscaled,
                        double* x_out, double* y_out);
int oracle_debug_attempts(oracle_handle* h, int64_t attempts);
int oracle_debug_state(oracle_handle* h, double* x, double* y,
                       double* dual_product, double* sum_x, double* sum_y,
                       folp_debug_scalars* s);
int oracle_debug_set_state(oracle_handle* h, const double* x, const double* y,
                           double step_size, double primal_weight);
int oracle_debug_spmv(oracle_handle* h, int transpose, const double* in,
                      double* out);
void oracle_destroy(oracle_handle* h);
/* seconds spent inside take_step so far (time_spent_doing_basic_algorithm) */
double oracle_basic_seconds(const oracle_handle* h);

/* ---- unit-level entry points for the reference's known-answer tests ------ */

/* solve_bound_constrained_trust_region, src/trust_region_utils.jl:68-224 */
int oracle_trust_region(int64_t len, const double* center_point,
                        const double* objective_vector,
                        const double* variable_lower_bounds,
                        const double* variable_upper_bounds,
                        const double* norm_weights, double target_radius,
                        int solve_approximately, double* solution_out,
                        double* value_out);

/* bound_optimal_objective, src/trust_region_utils.jl:271-360, on the "scaled"
 * arrays of *problem. norm_kind: 0 MAX_NORM, 1 EUCLIDEAN_NORM.
 * out3 = {lagrangian_value, lower_bound_value, upper_bound_value}. */
int oracle_bound_optimal_objective(const folp_problem* problem,
                                   const double* primal_solution,
                                   const double* dual_solution,
                                   const double* primal_norm_weights,
                                   const double* dual_norm_weights,
                                   double distance_to_optimality, int norm_kind,
                                   int solve_approximately, double* out3,
                                   double* primal_tr_solution,
                                   double* dual_tr_solution);

/* compute_iteration_stats, src/iteration_stats_utils.jl:356-406, evaluated on
 * the "scaled" arrays of *problem taken as THE problem (no unscaling). Fills
 * the convergence + infeasibility fields of *out. */
int oracle_iteration_stats(const folp_problem* problem,
                           const double* primal_iterate,
                           const double* dual_iterate,
                           const double* primal_ray_estimate,
                           const double* dual_ray_estimate,
                           double eps_optimal_absolute,
                           double eps_optimal_relative, folp_eval* out);

/* compute_dual_stats, src/iteration_stats_utils.jl:157-197. dual_residual has
 * (m - neq) + n entries, reduced_costs n. */
int oracle_dual_stats(const folp_problem* problem, const double* primal,
                      const double* dual, double* dual_objective,
                      double* dual_residual, double* reduced_costs);
/* max_primal_violation :17-22, primal_obj :67-74 */
double oracle_max_primal_violation(const folp_problem* problem,
                                   const double* primal);
double oracle_primal_obj(const folp_problem* problem, const double* primal);
/* compute_lagrangian_value, src/saddle_point.jl:1109-1120 */
double oracle_lagrangian_value(const folp_problem* problem,
                               const double* primal, const double* dual);
/* select_initial_primal_weight, src/saddle_point.jl:1049-1075 */
double oracle_select_initial_primal_weight(const folp_problem* problem,
                                           const double* primal_norm_params,
                                           const double* dual_norm_params,
                                           double primal_importance);
/* check_termination_criteria, src/termination.jl:233-273 (uses the cached
 * norms stored in *problem). Returns a folp_termination_reason (0 = false). */
int oracle_check_termination(const folp_params* params,
                             const folp_problem* problem,
                             const folp_eval* stats);

/* rescale_problem, src/preprocess.jl:631-687 (Ruiz :412, l2 :358,
 * Pock-Chambolle :508, scale_problem :555). Works in place on the mutable
 * copies the caller passes (nzval, q_nzval, c, l, u, b) and writes the
 * cumulative scalings. pock_chambolle_alpha < 0 means `nothing`. ruiz_p: 0 for
 * Inf (what rescale_problem uses) or 2. Index arrays are 0-based here. */
int oracle_rescale_problem(int64_t m, int64_t n, const int64_t* colptr,
                           const int64_t* rowval, double* nzval,
                           const int64_t* q_colptr, const int64_t* q_rowval,
                           double* q_nzval, double* c, double* l, double* u,
                           double* b, int l_inf_ruiz_iterations, int ruiz_p,
                           int l2_norm_rescaling, double pock_chambolle_alpha,
                           double* constraint_rescaling,
                           double* variable_rescaling);
/* l2_norm(matrix, dimension), src/preprocess.jl:99-113; dimension 1 -> n
 * column norms, 2 -> m row norms. */
int oracle_l2_norm(int64_t m, int64_t n, const int64_t* colptr,
                   const int64_t* rowval, const double* nzval, int dimension,
                   double* out);

/* Timing mode for bench.py's "all cores" CPU baseline: the two sparse
 * products run on `threads` OpenMP threads, bit-identical to the serial kernels (see
 * folp_oracle.c). Process-wide; 1 (the default) is the reference's serial behaviour. */
void oracle_set_threads(int threads);
int oracle_get_threads(void);
int oracle_openmp_enabled(void); /* 0: built without OpenMP, the timing mode runs serially */

#ifdef __cplusplus
}
#endif
#endif
