"""Output record types of the host side.

Mirrors src/solve_log.jl (ConvergenceInformation :64-168,
InfeasibilityInformation :174-225, IterationStats :232-315, SolveLog :349-406)
and SaddlePointOutput (src/saddle_point.jl:22-53). Field order follows the
reference so that JSON written from these records matches the reference's
`StructTypes.Mutable` serialisation (solve_log.jl:423-426).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

from ._abi import (
    CONVERGENCE_FIELDS,
    INFEASIBILITY_FIELDS,
    FolpEval,
    PointType,
    RestartChoice,
    TerminationReason,
)


@dataclass
class ConvergenceInformation:
    candidate_type: PointType = PointType.POINT_TYPE_UNSPECIFIED
    primal_objective: float = 0.0
    dual_objective: float = 0.0
    corrected_dual_objective: float = 0.0
    l_inf_primal_residual: float = 0.0
    l2_primal_residual: float = 0.0
    l_inf_dual_residual: float = 0.0
    l2_dual_residual: float = 0.0
    relative_l_inf_primal_residual: float = 0.0
    relative_l2_primal_residual: float = 0.0
    relative_l_inf_dual_residual: float = 0.0
    relative_l2_dual_residual: float = 0.0
    relative_optimality_gap: float = 0.0
    l_inf_primal_variable: float = 0.0
    l2_primal_variable: float = 0.0
    l_inf_dual_variable: float = 0.0
    l2_dual_variable: float = 0.0


@dataclass
class InfeasibilityInformation:
    candidate_type: PointType = PointType.POINT_TYPE_UNSPECIFIED
    max_primal_ray_infeasibility: float = 0.0
    primal_ray_linear_objective: float = 0.0
    primal_ray_quadratic_norm: float = 0.0
    max_dual_ray_infeasibility: float = 0.0
    dual_ray_objective: float = 0.0


@dataclass
class IterationStats:
    iteration_number: int = 0
    convergence_information: List[ConvergenceInformation] = field(default_factory=list)
    infeasibility_information: List[InfeasibilityInformation] = field(default_factory=list)
    cumulative_kkt_matrix_passes: float = 0.0
    cumulative_rejected_steps: int = 0
    cumulative_time_sec: float = 0.0
    restart_used: RestartChoice = RestartChoice.RESTART_CHOICE_UNSPECIFIED
    step_size: float = 0.0
    primal_weight: float = 0.0
    method_specific_stats: Dict[str, float] = field(default_factory=dict)


def iteration_stats_from_eval(e: FolpEval) -> IterationStats:
    """One folp_eval POD -> the IterationStats the reference would have built
    (pdhg.jl:912-945, :995)."""
    ptype = PointType(e.candidate_type)
    ci = ConvergenceInformation(candidate_type=ptype)
    for name in CONVERGENCE_FIELDS:
        setattr(ci, name, getattr(e, name))
    ii = InfeasibilityInformation(candidate_type=ptype)
    for name in INFEASIBILITY_FIELDS:
        setattr(ii, name, getattr(e, name))
    return IterationStats(
        iteration_number=e.iteration_number,
        convergence_information=[ci],
        infeasibility_information=[ii],
        cumulative_kkt_matrix_passes=e.cumulative_kkt_matrix_passes,
        cumulative_rejected_steps=e.cumulative_rejected_steps,
        cumulative_time_sec=e.cumulative_time_sec,
        restart_used=RestartChoice(e.restart_used),
        step_size=e.step_size,
        primal_weight=e.primal_weight,
        method_specific_stats={
            "time_spent_doing_basic_algorithm": e.time_spent_doing_basic_algorithm,
            "lagrangian_value": e.lagrangian_value,
            "estimated_lower_bound": e.estimated_lower_bound,
            "estimated_upper_bound": e.estimated_upper_bound,
        },
    )


def termination_reason_to_string(reason: TerminationReason) -> str:
    """src/termination.jl:275-278: strip the TERMINATION_REASON_ prefix."""
    return TerminationReason(reason).name[len("TERMINATION_REASON_"):]


@dataclass
class SaddlePointOutput:
    primal_solution: np.ndarray
    dual_solution: np.ndarray
    termination_reason: TerminationReason
    termination_string: str
    iteration_count: int
    iteration_stats: List[IterationStats]


@dataclass
class SolveLog:
    """src/solve_log.jl:349-406; assembled by the caller (scripts/solve_qp.jl:115-137)."""

    instance_name: str = ""
    command_line_invocation: str = ""
    termination_reason: TerminationReason = TerminationReason.TERMINATION_REASON_UNSPECIFIED
    termination_string: str = ""
    iteration_count: int = 0
    solve_time_sec: float = 0.0
    solution_stats: IterationStats = field(default_factory=IterationStats)
    solution_type: PointType = PointType.POINT_TYPE_UNSPECIFIED
    iteration_stats: List[IterationStats] = field(default_factory=list)
