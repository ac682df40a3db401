"""Console output of the PDHG path: the host mirror display.py against the
formats of src/iteration_stats_utils.jl:459-619, src/saddle_point.jl:947-1013 and
src/primal_dual_hybrid_gradient.jl:281-370."""
import io

import numpy as np

from folp_b200 import PointType, TerminationReason, display
from folp_b200.solve_log import ConvergenceInformation, InfeasibilityInformation, IterationStats
from shared_problems import example_lp


def test_print_to_screen_this_iteration():  # isu.jl:459-490
    f = display.print_to_screen_this_iteration
    assert not f(False, 1, 1, 40) and not f(True, 1, 1, 40)           # verbosity < 2: never
    assert f(True, 123, 2, 40)                                        # terminating: always
    assert f(False, 1, 2, 40) and not f(False, 41, 2, 40)             # verbosity 2: first iteration only
    assert f(False, 1, 3, 40) and f(False, 2001, 3, 40) and not f(False, 41, 3, 40)   # every 50 evaluations
    assert f(False, 801, 4, 40) and not f(False, 401, 4, 40)          # every 20
    assert f(False, 401, 5, 40) and f(False, 121, 6, 40) and not f(False, 41, 6, 40)  # 10, 3
    assert f(False, 41, 9, 40) and not f(False, 42, 9, 40)            # every evaluation


def test_heading_layout():  # isu.jl:499-540
    h = display.iteration_stats_heading(False).split("\n")
    assert h[0] == "runtime                  | residuals                  |  solution information      | relative residuals      |"
    assert h[1] == "#iter   #kkt     seconds | pr norm  du norm   gap     |  pr obj   pr norm  du norm | rel pr  rel du  rel gap |"
    h7 = display.iteration_stats_heading(True).split("\n")
    assert h7[0].endswith("| primal ray                  | dual ray           |")
    assert h7[1].endswith("| pr norm   linear   qu norm  | du norm   dual obj |")


def _stats():
    ci = ConvergenceInformation(candidate_type=PointType.POINT_TYPE_AVERAGE_ITERATE, primal_objective=-12.5,
                                dual_objective=-12.25, corrected_dual_objective=-np.inf, l2_primal_residual=1.5e-3,
                                l2_dual_residual=2.5e-4, l2_primal_variable=3.0, l2_dual_variable=40.0,
                                relative_l2_primal_residual=1e-4, relative_l2_dual_residual=2e-5,
                                relative_optimality_gap=9.8e-3, l_inf_primal_residual=1e-3, l_inf_dual_residual=2e-4,
                                l_inf_primal_variable=2.0, l_inf_dual_variable=30.0)
    ii = InfeasibilityInformation(candidate_type=PointType.POINT_TYPE_AVERAGE_ITERATE, max_primal_ray_infeasibility=0.5,
                                  primal_ray_linear_objective=-1.0, primal_ray_quadratic_norm=0.0,
                                  max_dual_ray_infeasibility=0.25, dual_ray_objective=2.0)
    return IterationStats(iteration_number=80, convergence_information=[ci], infeasibility_information=[ii],
                          cumulative_kkt_matrix_passes=164.5, cumulative_time_sec=0.0123,
                          method_specific_stats={"lagrangian_value": -12.4, "estimated_lower_bound": -13.0,
                                                 "estimated_upper_bound": -12.0,
                                                 "time_spent_doing_basic_algorithm": 0.01})


def test_row_layout():  # isu.jl:562-611
    row = display.iteration_stats_row(_stats(), False)
    assert row == ("80      1.6e+02  1.2e-02 | 1.5e-03  2.5e-04  -2.5e-01 | -1.2e+01  3.0e+00  4.0e+01 | "
                   "1.0e-04 2.0e-05 9.8e-03 |")
    row7 = display.iteration_stats_row(_stats(), True)
    assert row7 == row + " 5.0e-01  -1.0e+00  0.0e+00  | 2.5e-01   2.0e+00  |"
    # the heading and a row have the same column structure
    assert [len(c) for c in row.split("|")] == [len(c) for c in display.iteration_stats_heading(False).split("\n")[1].split("|")]


def test_final_logs():  # sp.jl:947-1013, pdhg.jl:324-370
    out = io.StringIO()
    display.generic_final_log(_stats(), 4, 81, TerminationReason.TERMINATION_REASON_OPTIMAL, file=out)
    text = out.getvalue().splitlines()
    assert text[0] == "Terminated after 81 iterations: OPTIMAL"
    assert text[1] == "For average candidate:"
    assert text[2] == "Primal objective: -12.500000, dual objective: -12.250000, corrected dual objective: -inf "
    assert text[3] == "Estimated optimal objective range: [-13.000000, -12.000000] "
    assert text[4] == "Lagrangian value: -12.400000 "
    assert text[5:] == ["Time (seconds):", " - Basic algorithm: 1.00e-02", " - Full algorithm:  1.23e-02"]
    out = io.StringIO()
    display.generic_final_log(_stats(), 0, 81, TerminationReason.TERMINATION_REASON_OPTIMAL, file=out)
    assert out.getvalue() == ""
    # "Avg solution" at the known optimum of example_lp: feasible, zero dual residual, objective -1
    lp = example_lp()
    out = io.StringIO()
    display.pdhg_final_log(lp, np.array([1.0, 0.0, 6.0, 2.0]), np.array([0.5, 4.0, 0.0]), 2, 301,
                           TerminationReason.TERMINATION_REASON_ITERATION_LIMIT, _stats(), file=out)
    lines = out.getvalue().splitlines()
    assert lines[0] == "Avg solution:"
    assert lines[1].split() == ["pr_infeas=", "0", "pr_obj=", "-1", "dual_infeas=", "0", "dual_obj=", "-1"]
    assert lines[2].startswith("  primal norms: L1=") and "Linf=              6" in lines[2]
    assert lines[4] == "Terminated after 301 iterations: ITERATION_LIMIT"
    out = io.StringIO()
    display.pdhg_specific_log(lp, 41, np.array([1.0, 0.0, 6.0, 2.0]), np.array([0.5, 4.0, 0.0]), 0.25, None, 2.0,
                              file=out)
    assert out.getvalue().startswith("      41 norms=(  6.40312,   4.03113) inv_step_size=        4 ")
    assert "dual_obj=                -1" in out.getvalue()
