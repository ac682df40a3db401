"""Verbosity-gated console output of the PDHG path: host-side
mirror of print_to_screen_this_iteration / display_iteration_stats(_heading) / lpad_float /
print_infinity_norms (src/iteration_stats_utils.jl:459-640), pdhg_specific_log / pdhg_final_log
(src/primal_dual_hybrid_gradient.jl:281-370) and generic_final_log (src/saddle_point.jl:947-1013).
Same column layout and printf formats as the reference; the numbers come from the folp_eval
records the library returns. The "Avg solution" block of pdhg_final_log and the per-iteration
pdhg_specific_log evaluate a few norms of the (scaled) iterate on the host with NumPy: they run
only at verbosity >= 2 (once, at termination) and >= 6 (debugging).
"""
from __future__ import annotations

import sys

import numpy as np

from ._abi import PointType
from .solve_log import IterationStats, termination_reason_to_string


def _e1(x: float) -> str:
    """C printf %.1e (Python's %e pads the exponent the same way)."""
    return "%.1e" % x


def print_to_screen_this_iteration(termination_reason, iteration: int, verbosity: int,
                                   termination_evaluation_frequency: int) -> bool:
    """isu.jl:459-490"""
    if verbosity < 2:
        return False
    if termination_reason:
        return True
    num_of_evaluations = (iteration - 1) / termination_evaluation_frequency
    if verbosity >= 9:
        display_frequency = 1
    elif verbosity >= 6:
        display_frequency = 3
    elif verbosity >= 5:
        display_frequency = 10
    elif verbosity >= 4:
        display_frequency = 20
    elif verbosity >= 3:
        display_frequency = 50
    else:
        return iteration == 1
    return num_of_evaluations % display_frequency == 0


def iteration_stats_heading(show_infeasibility: bool) -> str:
    """isu.jl:499-540"""
    line1 = "%s | %s | %s | %s |" % ("runtime".ljust(24), "residuals".ljust(26),
                                    " solution information".ljust(26), "relative residuals".ljust(23))
    if show_infeasibility:
        line1 += " %s | %s |" % ("primal ray".ljust(27), "dual ray".ljust(18))
    line2 = "%s %s %s | %s %s  %s | %s %s %s | %s %s %s |" % (
        "#iter".ljust(7), "#kkt".ljust(8), "seconds".ljust(7), "pr norm".ljust(8), "du norm".ljust(8),
        "gap".ljust(7), " pr obj".ljust(9), "pr norm".ljust(8), "du norm".ljust(7), "rel pr".ljust(7),
        "rel du".ljust(7), "rel gap".ljust(7))
    if show_infeasibility:
        line2 += " %s %s %s | %s %s |" % ("pr norm".ljust(9), "linear".ljust(8), "qu norm".ljust(8),
                                          "du norm".ljust(9), "dual obj".ljust(8))
    return line1 + "\n" + line2


def display_iteration_stats_heading(verbosity: int, file=None) -> None:
    """isu.jl:543-549"""
    if verbosity >= 2:
        print(iteration_stats_heading(verbosity >= 7), file=file or sys.stdout)


def lpad_float(number: float) -> str:
    """isu.jl:555-557"""
    return _e1(number).rjust(8)


def iteration_stats_row(stats: IterationStats, show_infeasibility: bool) -> str:
    """isu.jl:562-611"""
    head = str(stats.iteration_number).ljust(6)
    if stats.convergence_information:
        ci = stats.convergence_information[0]
        row = "%s  %s  %s | %s  %s  %s | %s  %s  %s | %s %s %s |" % (
            head, _e1(stats.cumulative_kkt_matrix_passes), _e1(stats.cumulative_time_sec),
            _e1(ci.l2_primal_residual), _e1(ci.l2_dual_residual),
            lpad_float(ci.primal_objective - ci.dual_objective), lpad_float(ci.primal_objective),
            _e1(ci.l2_primal_variable), _e1(ci.l2_dual_variable), _e1(ci.relative_l2_primal_residual),
            _e1(ci.relative_l2_dual_residual), _e1(ci.relative_optimality_gap))
    else:
        row = "%s  %s  %s" % (head, _e1(stats.cumulative_kkt_matrix_passes), _e1(stats.cumulative_time_sec))
    if show_infeasibility and stats.infeasibility_information:
        ii = stats.infeasibility_information[0]
        row += " %s  %s  %s  | %s  %s  |" % (
            _e1(ii.max_primal_ray_infeasibility), lpad_float(ii.primal_ray_linear_objective),
            _e1(ii.primal_ray_quadratic_norm), _e1(ii.max_dual_ray_infeasibility),
            lpad_float(ii.dual_ray_objective))
    return row


def display_iteration_stats(stats: IterationStats, verbosity: int, file=None) -> None:
    """isu.jl:613-619"""
    print(iteration_stats_row(stats, verbosity >= 7), file=file or sys.stdout)


def point_type_label(point_type) -> str:
    """sp.jl:929-945"""
    return {PointType.POINT_TYPE_CURRENT_ITERATE: "current", PointType.POINT_TYPE_AVERAGE_ITERATE: "average",
            PointType.POINT_TYPE_ITERATE_DIFFERENCE: "difference"}.get(PointType(point_type), "unknown PointType")


def generic_final_log(last_iteration_stats: IterationStats, verbosity: int, iteration: int,
                      termination_reason, file=None) -> None:
    """sp.jl:947-1013 (the variable / constraint hardness dump of verbosity >= 7 is not mirrored)."""
    out = file or sys.stdout
    if verbosity >= 1:
        print("Terminated after %d iterations: %s" % (iteration, termination_reason_to_string(termination_reason)),
              file=out)
    mss = last_iteration_stats.method_specific_stats
    if verbosity >= 3:
        for ci in last_iteration_stats.convergence_information:
            print("For %s candidate:" % point_type_label(ci.candidate_type), file=out)
            print("Primal objective: %f, dual objective: %f, corrected dual objective: %f " % (
                ci.primal_objective, ci.dual_objective, ci.corrected_dual_objective), file=out)
        if "estimated_lower_bound" in mss and "estimated_upper_bound" in mss:
            print("Estimated optimal objective range: [%f, %f] " % (mss["estimated_lower_bound"],
                                                                    mss["estimated_upper_bound"]), file=out)
        print("Lagrangian value: %f " % mss["lagrangian_value"], file=out)
    if verbosity >= 4:
        print("Time (seconds):\n - Basic algorithm: %.2e\n - Full algorithm:  %.2e" % (
            mss["time_spent_doing_basic_algorithm"], last_iteration_stats.cumulative_time_sec), file=out)
    if verbosity >= 7:
        for ci in last_iteration_stats.convergence_information:
            print("l_inf: primal_res = %.3e, dual_res = %.3e, primal_var = %.3e, dual_var = %.3e" % (
                ci.l_inf_primal_residual, ci.l_inf_dual_residual, ci.l_inf_primal_variable,
                ci.l_inf_dual_variable), file=out)


def _dual_residual_inf_and_objective(problem, x, y):
    """compute_dual_stats (isu.jl:157-197) of `problem` at (x, y): (|dual residual|_inf, dual objective)."""
    A, Q = problem.constraint_matrix, problem.objective_matrix
    qx = Q @ x
    g = qx + problem.objective_vector - A.T @ y
    l, u = problem.variable_lower_bound, problem.variable_upper_bound
    bound = np.where(g > 0.0, l, u)
    rc = np.where(np.isfinite(bound), g, 0.0)
    neq = problem.num_equalities
    dres = np.concatenate([np.maximum(-y[neq:], 0.0), g - rc])
    nz = rc != 0.0
    contrib = float(np.sum(bound[nz] * rc[nz])) if np.all(np.isfinite(bound[nz])) else -np.inf
    dobj = float(problem.right_hand_side @ y) + problem.objective_constant - 0.5 * float(qx @ x) + contrib
    return (float(np.max(np.abs(dres))) if dres.size else 0.0), dobj


def pdhg_final_log(problem, avg_primal_solution, avg_dual_solution, verbosity: int, iteration: int,
                   termination_reason, last_iteration_stats: IterationStats, file=None) -> None:
    """pdhg.jl:324-370; `problem` and the solutions are the SCALED ones, as in the reference."""
    out = file or sys.stdout
    if verbosity >= 2:
        x, y = np.asarray(avg_primal_solution), np.asarray(avg_dual_solution)
        A = problem.constraint_matrix
        act = A @ x
        neq = problem.num_equalities
        viol = np.concatenate([problem.right_hand_side[:neq] - act[:neq],
                               np.maximum(problem.right_hand_side[neq:] - act[neq:], 0.0),
                               np.maximum(problem.variable_lower_bound - x, 0.0),
                               np.maximum(x - problem.variable_upper_bound, 0.0)])
        infeas = float(np.max(np.abs(viol))) if viol.size else 0.0  # max_primal_violation, isu.jl:17-22
        pobj = problem.objective_constant + float(problem.objective_vector @ x) + \
            0.5 * float(x @ (problem.objective_matrix @ x))
        dinf, dobj = _dual_residual_inf_and_objective(problem, x, y)
        print("Avg solution:", file=out)
        print("  pr_infeas=%12g pr_obj=%15.10g dual_infeas=%12g dual_obj=%15.10g" % (infeas, pobj, dinf, dobj),
              file=out)
        for label, v in (("primal norms:", x), ("dual norms:  ", y)):
            print("  %s L1=%15.10g, L2=%15.10g, Linf=%15.10g" % (
                label, float(np.sum(np.abs(v))), float(np.linalg.norm(v)),
                float(np.max(np.abs(v))) if v.size else 0.0), file=out)
    generic_final_log(last_iteration_stats, verbosity, iteration, termination_reason, file=out)


def pdhg_specific_log(problem, iteration: int, current_primal_solution, current_dual_solution,
                      step_size: float, required_ratio, primal_weight: float, file=None) -> None:
    """pdhg.jl:281-319 (verbosity >= 6)."""
    out = file or sys.stdout
    x, y = np.asarray(current_primal_solution), np.asarray(current_dual_solution)
    dinf, dobj = _dual_residual_inf_and_objective(problem, x, y)
    corrected = dobj if dinf == 0.0 else -np.inf  # corrected_dual_obj, isu.jl:203-221
    line = "   %5d norms=(%9g, %9g) inv_step_size=%9g " % (iteration, float(np.linalg.norm(x)),
                                                          float(np.linalg.norm(y)), 1 / step_size)
    if required_ratio is not None:
        line += "   primal_weight=%18g dual_obj=%18g  inverse_ss=%18g" % (primal_weight, corrected, required_ratio)
    else:
        line += "   primal_weight=%18g dual_obj=%18g" % (primal_weight, corrected)
    print(line, file=out)
