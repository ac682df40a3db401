"""The reference's test fixtures (test/shared_test_qp_problems.jl:30-277) and the
PDHG parameter factory (test/test_primal_dual_hybrid_gradient.jl:15-74,
test/utilities.jl:87-99), restated for the Python host mirror."""
import numpy as np

import folp_b200 as F
from folp_b200 import (
    AdaptiveStepsizeParams,
    ConstantStepsizeParams,
    MalitskyPockStepsizeParameters,
    OptimalityNorm,
    PdhgParameters,
    QuadraticProgrammingProblem,
    RestartScheme,
    RestartToCurrentMetric,
    construct_restart_parameters,
    construct_termination_criteria,
    linear_programming_problem,
)

INF = np.inf


def example_lp():
    return linear_programming_problem(
        [0.0, 0.0, 0.0, 0.0], [2.0, 4.0, 6.0, 3.0], [5.0, 2.0, 1.0, 1.0], -14.0,
        np.array([[2.0, 1.0, 1.0, 2.0], [1.0, 0.0, 1.0, 0.0], [0.0, 0.0, 1.0, -1.0]]),
        [12.0, 7.0, 1.0], 1)


def example_lp_without_bounds():
    return linear_programming_problem([-INF], [INF], [-1.0], 0.0, np.array([[-1.0]]), [-2.0], 0)


def example_qp():
    return QuadraticProgrammingProblem(
        [0.0, 0.0], [1.0, 1.0], np.array([[4.0, 0.0], [0.0, 1.0]]), [-1.0, -1.0], -0.0,
        np.array([[-1.0, -1.0]]), [-1.0], 0)


def example_qp2():
    return QuadraticProgrammingProblem(
        [0.0, 0.0], [1.0, 1.0], np.array([[4.0, 0.0], [0.0, 1.0]]), [-1.0, 1.0], -0.0,
        np.array([[-1.0, -1.0]]), [-1.0], 0)


def example_cc_lp():
    return linear_programming_problem(
        [0.0] * 6, [1.0] * 6, [-1.0, -1.0, 1.0, -1.0, 1.0, -1.0], 4.0,
        np.array([[0.0, -1.0, 1.0, 0.0, 0.0, -1.0], [0.0, 0.0, 0.0, -1.0, 1.0, -1.0],
                  [-1.0, -1.0, 0.0, 1.0, 0.0, 0.0]]),
        [-1.0, -1.0, -1.0], 0)


def example_cc_star_lp():
    return linear_programming_problem(
        [0.0] * 6, [1.0] * 6, [-1.0, -1.0, -1.0, 1.0, 1.0, 1.0], 3.0,
        np.array([[-1.0, -1.0, 0.0, 1.0, 0.0, 0.0], [-1.0, 0.0, -1.0, 0.0, 1.0, 0.0],
                  [0.0, -1.0, -1.0, 0.0, 0.0, 1.0]]),
        [-1.0, -1.0, -1.0], 0)


def example_lp_dependent_rows():
    return linear_programming_problem(
        [0.0] * 4, [INF] * 4, [1.0, 2.0, 3.0, 4.0], 0.0,
        np.array([[1.0, 1.0, 1.0, 1.0], [1.0, 1.0, 1.0, 1.0], [1.0, 0.0, 0.0, 1.0]]),
        [2.0, 2.0, 1.0], 3)


def example_lp_easy_primal_infeasible():
    return linear_programming_problem([0.0, 0.0], [INF, INF], [1.0, 0.5], 0.0,
                                      np.array([[-1.0, -1.0]]), [1.0], 1)


def example_lp_hard_primal_infeasible(tol):
    assert tol > 0.0
    return linear_programming_problem(
        [0.0] * 4, [INF] * 4, [1.0, 2.0, 3.0, 4.0], 0.0,
        np.array([[1.0, 1.0, 0.0, 0.0], [0.0, 1.0, 1.0, 0.0], [0.0, 0.0, 1.0, 1.0],
                  [1.0, 1.0, 1.0, 1.0]]),
        [1.0, 1.0, 1.0, 2 + tol], 4)


def example_lp_dual_infeasible():
    return linear_programming_problem([0.0, 0.0], [INF, INF], [-1.0, 0.4], 0.0,
                                      np.array([[1.0, -2.0]]), [1.0], 1)


def terminate_on_iteration_limit(n):
    return construct_termination_criteria(
        optimality_norm=OptimalityNorm.L_INF, eps_optimal_absolute=0.0, eps_optimal_relative=0.0,
        eps_primal_infeasible=0.0, eps_dual_infeasible=0.0, time_sec_limit=100.0,
        iteration_limit=n, kkt_matrix_pass_limit=INF)


def generate_pdhg_params(
    l_inf_ruiz_iterations=0, l2_norm_rescaling=False, pock_chambolle_alpha=None,
    iteration_limit=200, primal_importance=1.0, scale_invariant_initial_primal_weight=True,
    verbosity=0, record_iteration_stats=True, restart_scheme=RestartScheme.NO_RESTARTS,
    restart_frequency_if_fixed=100, artificial_restart_threshold=0.5,
    sufficient_reduction_for_restart=0.1, necessary_reduction_for_restart=0.8,
    primal_weight_update_smoothing=0.5, termination_evaluation_frequency=5,
    use_approximate_localized_duality_gap=False,
    restart_to_current_metric=RestartToCurrentMetric.GAP_OVER_DISTANCE_SQUARED,
    step_size_policy="adaptive",
):
    if step_size_policy == "malitsky-pock":
        sp = MalitskyPockStepsizeParameters(0.7, 0.99, 1.0)
    elif step_size_policy == "constant":
        sp = ConstantStepsizeParams()
    else:
        sp = AdaptiveStepsizeParams(0.3, 0.6)
    rp = construct_restart_parameters(
        restart_scheme, restart_to_current_metric, restart_frequency_if_fixed,
        artificial_restart_threshold, sufficient_reduction_for_restart,
        necessary_reduction_for_restart, primal_weight_update_smoothing,
        use_approximate_localized_duality_gap)
    return PdhgParameters(
        l_inf_ruiz_iterations, l2_norm_rescaling, pock_chambolle_alpha, primal_importance,
        scale_invariant_initial_primal_weight, verbosity, record_iteration_stats,
        termination_evaluation_frequency, terminate_on_iteration_limit(iteration_limit), rp, sp)
