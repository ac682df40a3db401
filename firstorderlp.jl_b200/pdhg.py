"""optimize(params::PdhgParameters, qp) -- the entry point the B200 path keeps.

Mirrors src/primal_dual_hybrid_gradient.jl:782-1049: the host half (:786-859:
validate, cached norms, rescale_problem, initial step size, initial primal
weight) runs here exactly as the reference's Julia host does; the loop
(:862-1048) runs in libfolp_b200.so through the C ABI. No CPU fallback.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from . import _abi, _marshal
from .lib import Solver
from .params import (
    AdaptiveStepsizeParams,
    MalitskyPockStepsizeParameters,
    PdhgParameters,
)
from .preprocess import rescale_problem
from .problem import (
    QuadraticProgrammingProblem,
    ScaledQpProblem,
    cached_quadratic_program_info,
    validate,
)
from .solve_log import SaddlePointOutput, iteration_stats_from_eval, termination_reason_to_string


def power_method_failure_probability(dimension: int, epsilon: float, k: int) -> float:
    """src/primal_dual_hybrid_gradient.jl:379-391"""
    if k < 2 or epsilon <= 0.0:
        return 1.0
    return (min(0.824, 0.354 / math.sqrt(epsilon * (k - 1))) * math.sqrt(dimension)
            * (1.0 - epsilon) ** (k - 1 / 2))


def estimate_maximum_singular_value(matrix, probability_of_failure=0.001,
                                    desired_relative_error=0.2, seed=1):
    """src/primal_dual_hybrid_gradient.jl:414-440. The start vector of the
    reference comes from Julia's MersenneTwister(seed) randn stream, which is
    not reproducible outside Julia; NumPy's legacy MT19937 normal stream is used
    (parity of this value is unpinned by the reference's tests)."""
    n = matrix.shape[1]
    epsilon = 1.0 - (1.0 - desired_relative_error) ** 2
    x = np.random.RandomState(seed).randn(n)
    At = matrix.T.tocsr()
    k = 0
    while power_method_failure_probability(n, epsilon, k) > probability_of_failure:
        x = x / np.linalg.norm(x, 2)
        x = At @ (matrix @ x)
        k += 1
    sigma = math.sqrt(float(x @ (At @ (matrix @ x))) / float(np.linalg.norm(x, 2) ** 2))
    return sigma, k


def host_setup(params: PdhgParameters, original_problem: QuadraticProgrammingProblem,
               scaled: Optional[ScaledQpProblem] = None, device_rescaling: bool = False):
    """pdhg.jl:786-859 -> (ProblemHolder, FolpParams, ScaledQpProblem).

    device_rescaling: rescale_problem (preprocess.jl:631-687) runs on the GPU through
    folp_rescale_problem instead of the host mirror (same arithmetic)."""
    validate(original_problem)
    cache = cached_quadratic_program_info(original_problem)
    if scaled is None and device_rescaling:
        from .lib import rescale_problem as device_rescale_problem
        scaled = device_rescale_problem(params.l_inf_ruiz_iterations, params.l2_norm_rescaling,
                                        params.pock_chambolle_alpha, original_problem)
    if scaled is None:
        scaled = rescale_problem(params.l_inf_ruiz_iterations, params.l2_norm_rescaling,
                                 params.pock_chambolle_alpha, params.verbosity, original_problem)
    problem = scaled.scaled_qp
    if params.primal_importance <= 0 or not math.isfinite(params.primal_importance):
        raise ValueError("primal_importance must be positive and finite")  # :798-801
    pol = params.step_size_policy_params
    if isinstance(pol, (AdaptiveStepsizeParams, MalitskyPockStepsizeParameters)):
        kkt0 = 0.5
        step0 = _marshal.initial_step_size_inf_norm(problem)  # :823, :826
    else:
        sigma, k = estimate_maximum_singular_value(problem.constraint_matrix, 0.001, 0.2)
        step0 = (1 - 0.2) / sigma  # :836
        kkt0 = float(k)
    if params.scale_invariant_initial_primal_weight:
        pw0 = _marshal.select_initial_primal_weight(problem, params.primal_importance)
    else:
        pw0 = params.primal_importance
    holder = _marshal.make_problem(scaled, cache, with_original_matrix=False)
    fparams = _marshal.make_params(params, step0, pw0, kkt0)
    return holder, fparams, scaled


def optimize(params: PdhgParameters, original_problem: QuadraticProgrammingProblem,
             device_rescaling: bool = False) -> SaddlePointOutput:
    holder, fparams, scaled = host_setup(params, original_problem, device_rescaling=device_rescaling)
    if params.verbosity <= 0:  # nothing to print: the whole loop in one C call
        with Solver(holder, fparams) as solver:
            x, y, reason, iters, evals = solver.solve()
        stats = [iteration_stats_from_eval(e) for e in evals]
        reason = _abi.TerminationReason(reason)
        return SaddlePointOutput(x, y, reason, termination_reason_to_string(reason), iters, stats)
    return _optimize_with_log(params, holder, fparams, scaled)


def _optimize_with_log(params: PdhgParameters, holder, fparams, scaled) -> SaddlePointOutput:
    """The loop of pdhg.jl:886-1048 driven evaluation by evaluation (folp_run), so that the
    verbosity-gated table of isu.jl:459-619 and the final logs are printed where the reference
    prints them."""
    from . import display
    verbosity = params.verbosity
    freq = params.termination_evaluation_frequency
    stats = []
    display.display_iteration_stats_heading(verbosity)  # pdhg.jl:883
    with Solver(holder, fparams) as solver:
        while True:
            e = solver.run()
            iteration = e.iteration_number + 1  # the reference's loop counter at the evaluation
            st = iteration_stats_from_eval(e)
            terminated = e.termination_reason != 0
            if params.record_iteration_stats or terminated:  # :958-960
                stats.append(st)
            if display.print_to_screen_this_iteration(terminated, iteration, verbosity, freq):
                display.display_iteration_stats(st, verbosity)
            if terminated:
                reason = _abi.TerminationReason(e.termination_reason)
                x, y = solver.get_solution(which=0, unscaled=True)
                if verbosity >= 2:  # the reference logs the SCALED average on the scaled problem
                    xs, ys = solver.get_solution(which=0, unscaled=False)
                else:
                    xs, ys = x, y
                display.pdhg_final_log(scaled.scaled_qp, xs, ys, verbosity, iteration, reason, st)
                return SaddlePointOutput(x, y, reason, termination_reason_to_string(reason),
                                         int(e.iteration_number), stats)
            if verbosity >= 6 and display.print_to_screen_this_iteration(False, iteration, verbosity, freq):
                xc, yc = solver.get_solution(which=1, unscaled=False)  # :1027-1041
                display.pdhg_specific_log(scaled.scaled_qp, iteration, xc, yc, e.step_size, None,
                                          e.primal_weight)
