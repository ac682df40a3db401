"""The reference's end-to-end PDHG tests (test/test_primal_dual_hybrid_gradient.jl:76-424),
same parameters and tolerances, parameterised on the `optimize` implementation:
tests/test_oracle_pdhg.py runs them on the CPU oracle (pinning it), and
tests/test_gpu_pdhg.py runs them on libfolp_b200.so through folp_b200.optimize."""
import numpy as np

from folp_b200 import RestartChoice, RestartScheme, RestartToCurrentMetric, TerminationReason
from shared_problems import (
    example_cc_lp, example_cc_star_lp, example_lp, example_lp_without_bounds, example_qp,
    example_qp2, generate_pdhg_params,
)

X_LP = np.array([1.0, 0.0, 6.0, 2.0])
Y_LP = np.array([0.5, 4.0, 0.0])


def _close(a, b, atol):
    assert np.max(np.abs(np.asarray(a) - np.asarray(b))) <= atol, (a, b)


def case_low_precision(optimize):  # :77-87
    out = optimize(generate_pdhg_params(iteration_limit=300), example_lp())
    _close(out.primal_solution, X_LP, 1e-4)
    _close(out.dual_solution, Y_LP, 1e-4)
    assert out.termination_reason == TerminationReason.TERMINATION_REASON_ITERATION_LIMIT
    assert out.iteration_count == 300


def case_terminate_with_optimal_solution(optimize):  # :88-98
    params = generate_pdhg_params(iteration_limit=1000)
    params.termination_criteria.eps_optimal_absolute = 1e-8
    out = optimize(params, example_lp())
    assert out.termination_reason == TerminationReason.TERMINATION_REASON_OPTIMAL


def case_fixed_frequency_restart(optimize):  # :116-129
    params = generate_pdhg_params(iteration_limit=500, restart_scheme=RestartScheme.FIXED_FREQUENCY,
                                  restart_frequency_if_fixed=30)
    out = optimize(params, example_lp())
    _close(out.primal_solution, X_LP, 1e-9)
    _close(out.dual_solution, Y_LP, 1e-9)


def _has_restart_to_average(out):
    return any(s.restart_used == RestartChoice.RESTART_CHOICE_RESTART_TO_AVERAGE
               for s in out.iteration_stats)


def case_adaptive_restart_heuristic(optimize):  # :130-147
    params = generate_pdhg_params(iteration_limit=600, restart_scheme=RestartScheme.ADAPTIVE_NORMALIZED)
    out = optimize(params, example_lp())
    _close(out.primal_solution, X_LP, 1e-9)
    _close(out.dual_solution, Y_LP, 1e-9)
    assert _has_restart_to_average(out)


def case_constant_step_no_smoothing(optimize):  # :149-172 (initial step: parity unpinned, see oracle.py)
    params = generate_pdhg_params(iteration_limit=700, primal_weight_update_smoothing=0.0,
                                  restart_scheme=RestartScheme.ADAPTIVE_NORMALIZED,
                                  step_size_policy="constant")
    out = optimize(params, example_lp())
    _close(out.primal_solution, X_LP, 1e-9)
    _close(out.dual_solution, Y_LP, 1e-9)
    assert _has_restart_to_average(out)
    step = out.iteration_stats[0].step_size
    assert all(s.step_size == step for s in out.iteration_stats)


def case_restart_to_current_metrics(optimize, metric):  # :174-212
    params = generate_pdhg_params(iteration_limit=600, restart_scheme=RestartScheme.ADAPTIVE_NORMALIZED,
                                  restart_to_current_metric=metric)
    out = optimize(params, example_lp())
    _close(out.primal_solution, X_LP, 1e-9)
    _close(out.dual_solution, Y_LP, 1e-9)
    assert _has_restart_to_average(out)


def case_adaptive_restart_zero_objective(optimize, approx, limit):  # :214-243
    params = generate_pdhg_params(iteration_limit=limit, restart_scheme=RestartScheme.ADAPTIVE_NORMALIZED,
                                  use_approximate_localized_duality_gap=approx)
    problem = example_lp()
    problem.objective_vector = np.zeros(4)
    params.termination_criteria.eps_optimal_absolute = 1e-8
    out = optimize(params, problem)
    assert out.termination_reason == TerminationReason.TERMINATION_REASON_OPTIMAL


def case_malitsky_pock(optimize, smoothing):  # :245-274
    params = generate_pdhg_params(iteration_limit=700, primal_weight_update_smoothing=smoothing,
                                  restart_scheme=RestartScheme.ADAPTIVE_NORMALIZED,
                                  step_size_policy="malitsky-pock")
    out = optimize(params, example_lp())
    _close(out.primal_solution, X_LP, 1e-9)
    _close(out.dual_solution, Y_LP, 1e-9)


def case_quadratic_programming_1(optimize):  # :276-286
    out = optimize(generate_pdhg_params(iteration_limit=200), example_qp())
    _close(out.primal_solution, [0.2, 0.8], 1e-4)
    _close(out.dual_solution, [0.2], 1e-4)


def case_quadratic_programming_2(optimize):  # :287-297
    out = optimize(generate_pdhg_params(iteration_limit=200), example_qp2())
    _close(out.primal_solution, [0.25, 0.0], 1e-4)
    _close(out.dual_solution, [0.0], 1e-4)


def case_preprocessing_qp2(optimize, kw):  # :298-322
    out = optimize(generate_pdhg_params(iteration_limit=200, **kw), example_qp2())
    _close(out.primal_solution, [0.25, 0.0], 1e-4)
    _close(out.dual_solution, [0.0], 1e-4)


def case_pock_chambolle_rescaling(optimize):  # :323-335
    out = optimize(generate_pdhg_params(pock_chambolle_alpha=1.0, iteration_limit=3000), example_lp())
    _close(out.primal_solution, X_LP, 1e-4)
    _close(out.dual_solution, Y_LP, 1e-4)


def case_high_precision(optimize):  # :337-347
    out = optimize(generate_pdhg_params(iteration_limit=800), example_lp())
    _close(out.primal_solution, X_LP, 1e-9)
    _close(out.dual_solution, Y_LP, 1e-9)


def case_infeasible_instance(optimize):  # :348-360
    problem = example_lp()
    problem.right_hand_side[2] = 8
    out = optimize(generate_pdhg_params(iteration_limit=800), problem)
    assert out.termination_reason == TerminationReason.TERMINATION_REASON_PRIMAL_INFEASIBLE


def case_lp_without_bounds(optimize):  # :361-371
    out = optimize(generate_pdhg_params(iteration_limit=400), example_lp_without_bounds())
    _close(out.primal_solution, [2.0], 1e-9)
    _close(out.dual_solution, [1.0], 1e-9)


def _check_cc(out):
    tol = 1e-14
    _close(out.primal_solution, [1.0, 1.0, 0.0, 1.0, 0.0, 0.0], tol)
    final = out.iteration_stats[-1]
    assert abs(final.convergence_information[0].dual_objective - 1.0) <= tol
    assert all(out.dual_solution >= 0.0)
    assert out.dual_solution[0] + out.dual_solution[1] >= 1.0 - tol


def case_correlation_clustering_triangle_plus(optimize):  # :372-390
    _check_cc(optimize(generate_pdhg_params(iteration_limit=15), example_cc_lp()))


def case_numerical_error(optimize):  # :391-412
    out = optimize(generate_pdhg_params(iteration_limit=150), example_cc_lp())
    assert out.termination_reason == TerminationReason.TERMINATION_REASON_NUMERICAL_ERROR
    _check_cc(out)


def case_correlation_clustering_star(optimize):  # :413-423
    out = optimize(generate_pdhg_params(iteration_limit=100), example_cc_star_lp())
    _close(out.primal_solution, [0.5, 0.5, 0.5, 0.0, 0.0, 0.0], 1e-6)
    _close(out.dual_solution, [0.5, 0.5, 0.5], 1e-6)


def case_artificial_restart_cadence(optimize):
    """with NO_RESTARTS and threshold 0.5 the artificial
    restarts fire at evaluated iterations 1,2,4,8,20,40,80,... (sp.jl:719-727)."""
    out = optimize(generate_pdhg_params(iteration_limit=100), example_lp())
    resets = [s.iteration_number for s in out.iteration_stats
              if s.restart_used == RestartChoice.RESTART_CHOICE_WEIGHTED_AVERAGE_RESET]
    assert resets == [1, 2, 4, 8, 20, 40, 80]
    evaluated = [s.iteration_number for s in out.iteration_stats]
    assert evaluated[:11] == list(range(10)) + [10] and evaluated[-1] == 100


# (id, function, kwargs, needs a quadratic objective)
def case_config1_trivial_mps(optimize):
    """BASELINE.json configs[0]: the two-variable MPS model read from disk and solved with the CLI
    defaults (scripts/solve_qp.jl:193-472, --iteration_limit 5000 as in README.md:54-57), the summary
    written the way scripts/solve_qp.jl:115-141 writes it. min 2x - y, x + y <= 3, 0 <= x <= 1,
    1 <= y <= 2  ->  (x, y) = (0, 2), objective -2, the constraint is slack."""
    import json
    import os
    import tempfile

    import folp_b200
    from folp_b200 import io as fio
    here = os.path.dirname(os.path.abspath(__file__))
    lp = fio.qps_reader_to_standard_form(os.path.join(here, "golden", "trivial_lp_model.mps"))
    params = folp_b200.PdhgParameters(verbosity=0)
    params.termination_criteria.iteration_limit = 5000
    out = optimize(params, lp)
    assert out.termination_reason == TerminationReason.TERMINATION_REASON_OPTIMAL
    _close(out.primal_solution, [0.0, 2.0], 1e-5)
    _close(out.dual_solution, [0.0], 1e-5)
    with tempfile.TemporaryDirectory() as d:
        summary, _ = fio.write_solve_log_json(d, "trivial_lp_model", out, 0.0)
        log = json.load(open(summary))
    assert log["termination_reason"] == "TERMINATION_REASON_OPTIMAL" and log["termination_string"] == "OPTIMAL"
    assert abs(log["solution_stats"]["convergence_information"][0]["primal_objective"] + 2.0) <= 1e-5


CASES = [
    ("config1_trivial_mps", case_config1_trivial_mps, {}, False),
    ("low_precision", case_low_precision, {}, False),
    ("terminate_with_optimal_solution", case_terminate_with_optimal_solution, {}, False),
    ("fixed_frequency_restart", case_fixed_frequency_restart, {}, False),
    ("adaptive_restart_heuristic", case_adaptive_restart_heuristic, {}, False),
    ("constant_step_no_smoothing", case_constant_step_no_smoothing, {}, False),
    ("restart_metric_none", case_restart_to_current_metrics,
     {"metric": RestartToCurrentMetric.NO_RESTART_TO_CURRENT}, False),
    ("restart_metric_gap_over_distance", case_restart_to_current_metrics,
     {"metric": RestartToCurrentMetric.GAP_OVER_DISTANCE}, False),
    ("zero_objective_exact", case_adaptive_restart_zero_objective, {"approx": False, "limit": 200}, False),
    ("zero_objective_approx", case_adaptive_restart_zero_objective, {"approx": True, "limit": 300}, False),
    ("malitsky_pock_no_smoothing", case_malitsky_pock, {"smoothing": 0.0}, False),
    ("malitsky_pock_smoothing", case_malitsky_pock, {"smoothing": 0.5}, False),
    ("quadratic_programming_1", case_quadratic_programming_1, {}, True),
    ("quadratic_programming_2", case_quadratic_programming_2, {}, True),
    ("preprocessing_qp2_l2", case_preprocessing_qp2, {"kw": dict(l2_norm_rescaling=True)}, True),
    ("preprocessing_qp2_ruiz", case_preprocessing_qp2, {"kw": dict(l_inf_ruiz_iterations=10)}, True),
    ("pock_chambolle_rescaling", case_pock_chambolle_rescaling, {}, False),
    ("high_precision", case_high_precision, {}, False),
    ("infeasible_instance", case_infeasible_instance, {}, False),
    ("lp_without_bounds", case_lp_without_bounds, {}, False),
    ("correlation_clustering_triangle_plus", case_correlation_clustering_triangle_plus, {}, False),
    ("numerical_error", case_numerical_error, {}, False),
    ("correlation_clustering_star", case_correlation_clustering_star, {}, False),
    ("artificial_restart_cadence", case_artificial_restart_cadence, {}, False),
]
