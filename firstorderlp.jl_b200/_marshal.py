"""Python objects -> the PODs of include/folp_b200.h.

This is the host half of primal_dual_hybrid_gradient.jl:786-859 as far as data
marshalling goes: given the ScaledQpProblem and PdhgParameters it fills
folp_problem / folp_params. The returned holder keeps every NumPy buffer alive
for as long as the struct is in use (the C side only borrows pointers).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np

from ._abi import FolpParams, FolpProblem, StepSizePolicy
from .params import (
    AdaptiveStepsizeParams,
    ConstantStepsizeParams,
    MalitskyPockStepsizeParameters,
    PdhgParameters,
)
from .problem import (
    CachedQuadraticProgramInfo,
    QuadraticProgrammingProblem,
    ScaledQpProblem,
    cached_quadratic_program_info,
)

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int64)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


class ProblemHolder:
    """Owns the buffers a FolpProblem points into."""

    def __init__(self):
        self.struct = FolpProblem()
        self._keep = []

    def _d(self, a) -> _pd:
        arr = _f64(a)
        self._keep.append(arr)
        return arr.ctypes.data_as(_pd)

    def _i(self, a) -> _pi:
        arr = _i64(a)
        self._keep.append(arr)
        return arr.ctypes.data_as(_pi)

    def byref(self):
        return C.byref(self.struct)


def make_problem(
    scaled: ScaledQpProblem,
    cache: Optional[CachedQuadraticProgramInfo] = None,
    with_original_matrix: bool = True,
) -> ProblemHolder:
    """ScaledQpProblem -> folp_problem (index_base 0)."""
    P = scaled.scaled_qp
    O = scaled.original_qp
    if cache is None:
        cache = cached_quadratic_program_info(O)
    A = P.constraint_matrix
    Q = P.objective_matrix
    h = ProblemHolder()
    s = h.struct
    s.num_variables = A.shape[1]
    s.num_constraints = A.shape[0]
    s.num_nonzeros = A.nnz
    s.num_equalities = P.num_equalities
    s.index_base = 0
    s.colptr = h._i(A.indptr)
    s.rowval = h._i(A.indices)
    s.nzval = h._d(A.data)
    s.objective_vector = h._d(P.objective_vector)
    s.variable_lower_bound = h._d(P.variable_lower_bound)
    s.variable_upper_bound = h._d(P.variable_upper_bound)
    s.right_hand_side = h._d(P.right_hand_side)
    s.objective_constant = P.objective_constant
    s.variable_rescaling = h._d(scaled.variable_rescaling)
    s.constraint_rescaling = h._d(scaled.constraint_rescaling)
    s.orig_objective_vector = h._d(O.objective_vector)
    s.orig_variable_lower_bound = h._d(O.variable_lower_bound)
    s.orig_variable_upper_bound = h._d(O.variable_upper_bound)
    s.orig_right_hand_side = h._d(O.right_hand_side)
    same_pattern = (
        O.constraint_matrix.nnz == A.nnz
        and np.array_equal(O.constraint_matrix.indptr, A.indptr)
        and np.array_equal(O.constraint_matrix.indices, A.indices)
    )
    if with_original_matrix and same_pattern:
        s.orig_nzval = h._d(O.constraint_matrix.data)
    else:
        s.orig_nzval = None
    s.q_num_nonzeros = Q.nnz
    if Q.nnz:
        s.q_colptr = h._i(Q.indptr)
        s.q_rowval = h._i(Q.indices)
        s.q_nzval = h._d(Q.data)
        OQ = O.objective_matrix
        if OQ.nnz == Q.nnz and np.array_equal(OQ.indptr, Q.indptr):
            s.q_orig_nzval = h._d(OQ.data)
    s.l_inf_norm_primal_linear_objective = cache.l_inf_norm_primal_linear_objective
    s.l_inf_norm_primal_right_hand_side = cache.l_inf_norm_primal_right_hand_side
    s.l2_norm_primal_linear_objective = cache.l2_norm_primal_linear_objective
    s.l2_norm_primal_right_hand_side = cache.l2_norm_primal_right_hand_side
    return h


def unscaled_as_scaled(problem: QuadraticProgrammingProblem) -> ScaledQpProblem:
    """A problem with identity rescaling (what rescale_problem(0,false,nothing) gives)."""
    m, n = problem.constraint_matrix.shape
    return ScaledQpProblem(problem, problem.copy(), np.ones(m), np.ones(n))


def make_params(
    params: PdhgParameters,
    initial_step_size: float,
    initial_primal_weight: float,
    initial_kkt_passes: float,
) -> FolpParams:
    p = FolpParams()
    sp_ = params.step_size_policy_params
    if isinstance(sp_, AdaptiveStepsizeParams):
        p.step_size_policy = StepSizePolicy.ADAPTIVE
        p.reduction_exponent = sp_.reduction_exponent
        p.growth_exponent = sp_.growth_exponent
    elif isinstance(sp_, MalitskyPockStepsizeParameters):
        p.step_size_policy = StepSizePolicy.MALITSKY_POCK
        p.downscaling_factor = sp_.downscaling_factor
        p.breaking_factor = sp_.breaking_factor
        p.interpolation_coefficient = sp_.interpolation_coefficient
    elif isinstance(sp_, ConstantStepsizeParams):
        p.step_size_policy = StepSizePolicy.CONSTANT
    else:
        raise TypeError("unknown step_size_policy_params")
    p.termination_evaluation_frequency = int(params.termination_evaluation_frequency)
    p.initial_step_size = float(initial_step_size)
    p.initial_primal_weight = float(initial_primal_weight)
    p.initial_kkt_passes = float(initial_kkt_passes)
    tc = params.termination_criteria
    p.optimality_norm = int(tc.optimality_norm)
    p.iteration_limit = int(min(tc.iteration_limit, 2**31 - 1))
    p.eps_optimal_absolute = tc.eps_optimal_absolute
    p.eps_optimal_relative = tc.eps_optimal_relative
    p.eps_primal_infeasible = tc.eps_primal_infeasible
    p.eps_dual_infeasible = tc.eps_dual_infeasible
    p.time_sec_limit = tc.time_sec_limit
    p.kkt_matrix_pass_limit = tc.kkt_matrix_pass_limit
    rp = params.restart_params
    p.restart_scheme = int(rp.restart_scheme)
    p.restart_to_current_metric = int(rp.restart_to_current_metric)
    p.restart_frequency_if_fixed = int(rp.restart_frequency_if_fixed)
    p.artificial_restart_threshold = rp.artificial_restart_threshold
    p.sufficient_reduction_for_restart = rp.sufficient_reduction_for_restart
    p.necessary_reduction_for_restart = rp.necessary_reduction_for_restart
    p.primal_weight_update_smoothing = rp.primal_weight_update_smoothing
    p.use_approximate_localized_duality_gap = int(rp.use_approximate_localized_duality_gap)
    p.record_iteration_stats = int(params.record_iteration_stats)
    p.verbosity = int(params.verbosity)
    return p


def select_initial_primal_weight(
    problem: QuadraticProgrammingProblem, primal_importance: float
) -> float:
    """src/saddle_point.jl:1049-1075 with unit norm weights (pdhg.jl:847-854).
    weighted_norm is a serial loop in the reference; math.fsum-free plain
    accumulation keeps the same order."""
    rhs = math.sqrt(_serial_sumsq(problem.right_hand_side))
    obj = math.sqrt(_serial_sumsq(problem.objective_vector))
    if obj > 0.0 and rhs > 0.0:
        return primal_importance * (obj / rhs)
    return primal_importance


def _serial_sumsq(v: np.ndarray) -> float:
    # np.cumsum accumulates strictly left to right (no pairwise blocking).
    if v.size == 0:
        return 0.0
    return float(np.cumsum(v * v)[-1])


def initial_step_size_inf_norm(problem: QuadraticProgrammingProblem) -> float:
    """1 / norm(constraint_matrix, Inf) = 1 / max |A_ij| (pdhg.jl:823, :826)."""
    data = problem.constraint_matrix.data
    mx = float(np.max(np.abs(data))) if data.size else 0.0
    return 1.0 / mx if mx != 0.0 else math.inf
