"""Solver parameters of the host side.

Mirrors the reference structs field for field:
PdhgParameters (src/primal_dual_hybrid_gradient.jl:128-199),
AdaptiveStepsizeParams (:60-63), MalitskyPockStepsizeParameters (:19-41),
ConstantStepsizeParams (:68), TerminationCriteria + construct_termination_criteria
(src/termination.jl:29-120), RestartParameters + construct_restart_parameters
(src/saddle_point.jl:342-430).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Union

from ._abi import OptimalityNorm, RestartScheme, RestartToCurrentMetric

INT32_MAX = 2**31 - 1


@dataclass
class AdaptiveStepsizeParams:
    reduction_exponent: float = 0.3
    growth_exponent: float = 0.6


@dataclass
class MalitskyPockStepsizeParameters:
    downscaling_factor: float = 0.7
    breaking_factor: float = 0.99
    interpolation_coefficient: float = 1.0


@dataclass
class ConstantStepsizeParams:
    pass


@dataclass
class TerminationCriteria:
    optimality_norm: OptimalityNorm = OptimalityNorm.L2
    eps_optimal_absolute: float = 1.0e-6
    eps_optimal_relative: float = 1.0e-6
    eps_primal_infeasible: float = 1.0e-8
    eps_dual_infeasible: float = 1.0e-8
    time_sec_limit: float = math.inf
    iteration_limit: int = INT32_MAX
    kkt_matrix_pass_limit: float = math.inf


def construct_termination_criteria(**kwargs) -> TerminationCriteria:
    return TerminationCriteria(**kwargs)


def validate_termination_criteria(criteria: TerminationCriteria) -> None:
    """src/termination.jl:122-138"""
    if criteria.eps_primal_infeasible < 0:
        raise ValueError("eps_primal_infeasible must be nonnegative")
    if criteria.eps_dual_infeasible < 0:
        raise ValueError("eps_dual_infeasible must be nonnegative")
    if criteria.time_sec_limit <= 0:
        raise ValueError("time_sec_limit must be positive")
    if criteria.iteration_limit <= 0:
        raise ValueError("iteration_limit must be positive")
    if criteria.kkt_matrix_pass_limit <= 0:
        raise ValueError("kkt_matrix_pass_limit must be positive")


@dataclass
class RestartParameters:
    restart_scheme: RestartScheme = RestartScheme.ADAPTIVE_NORMALIZED
    restart_to_current_metric: RestartToCurrentMetric = (
        RestartToCurrentMetric.GAP_OVER_DISTANCE_SQUARED
    )
    restart_frequency_if_fixed: int = 1000
    artificial_restart_threshold: float = 0.5
    sufficient_reduction_for_restart: float = 0.1
    necessary_reduction_for_restart: float = 0.9
    primal_weight_update_smoothing: float = 0.5
    use_approximate_localized_duality_gap: bool = False


def construct_restart_parameters(
    restart_scheme,
    restart_to_current_metric,
    restart_frequency_if_fixed,
    artificial_restart_threshold,
    sufficient_reduction_for_restart,
    necessary_reduction_for_restart,
    primal_weight_update_smoothing,
    use_approximate_localized_duality_gap,
) -> RestartParameters:
    """src/saddle_point.jl:402-430 (same asserts)."""
    assert restart_frequency_if_fixed > 1
    assert 0.0 < artificial_restart_threshold <= 1.0
    assert 0.0 < sufficient_reduction_for_restart <= necessary_reduction_for_restart <= 1.0
    assert 0.0 <= primal_weight_update_smoothing <= 1.0
    return RestartParameters(
        RestartScheme(restart_scheme),
        RestartToCurrentMetric(restart_to_current_metric),
        int(restart_frequency_if_fixed),
        float(artificial_restart_threshold),
        float(sufficient_reduction_for_restart),
        float(necessary_reduction_for_restart),
        float(primal_weight_update_smoothing),
        bool(use_approximate_localized_duality_gap),
    )


@dataclass
class PdhgParameters:
    """Defaults are the `scripts/solve_qp.jl` command-line defaults (:193-472)."""

    l_inf_ruiz_iterations: int = 10
    l2_norm_rescaling: bool = False
    pock_chambolle_alpha: Optional[float] = 1.0
    primal_importance: float = 1.0
    scale_invariant_initial_primal_weight: bool = True
    verbosity: int = 2
    record_iteration_stats: bool = True
    termination_evaluation_frequency: int = 40
    termination_criteria: TerminationCriteria = field(default_factory=TerminationCriteria)
    restart_params: RestartParameters = field(default_factory=RestartParameters)
    step_size_policy_params: Union[
        MalitskyPockStepsizeParameters, AdaptiveStepsizeParams, ConstantStepsizeParams
    ] = field(default_factory=AdaptiveStepsizeParams)
