"""Problem types of the host side.

Mirrors src/quadratic_programming.jl of the reference:
QuadraticProgrammingProblem (:34-76), linear_programming_problem (:255-277),
is_linear_programming_problem (:282-284), ScaledQpProblem (:293-298),
equality_range / inequality_range (:300-304); plus validate
(src/preprocess.jl:18-84) and cached_quadratic_program_info
(src/termination.jl:144-158).
"""
from __future__ import annotations

import copy
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp


def _as_csc(matrix, shape=None) -> sp.csc_matrix:
    if sp.issparse(matrix):
        out = sp.csc_matrix(matrix, dtype=np.float64)
    else:
        arr = np.asarray(matrix, dtype=np.float64)
        if arr.ndim == 1 and shape is not None:
            arr = arr.reshape(shape)
        out = sp.csc_matrix(arr)
    out.sort_indices()
    return out


@dataclass
class QuadraticProgrammingProblem:
    """min 1/2 x'Qx + c'x + c0  s.t.  A[:neq] x = b[:neq], A[neq:] x >= b[neq:], l <= x <= u."""

    variable_lower_bound: np.ndarray
    variable_upper_bound: np.ndarray
    objective_matrix: sp.csc_matrix
    objective_vector: np.ndarray
    objective_constant: float
    constraint_matrix: sp.csc_matrix
    right_hand_side: np.ndarray
    num_equalities: int

    def __post_init__(self):
        self.variable_lower_bound = np.array(self.variable_lower_bound, dtype=np.float64)
        self.variable_upper_bound = np.array(self.variable_upper_bound, dtype=np.float64)
        self.objective_vector = np.array(self.objective_vector, dtype=np.float64)
        self.right_hand_side = np.array(self.right_hand_side, dtype=np.float64)
        n = self.variable_lower_bound.shape[0]
        m = self.right_hand_side.shape[0]
        self.constraint_matrix = _as_csc(self.constraint_matrix, (m, n))
        self.objective_matrix = _as_csc(self.objective_matrix, (n, n))
        self.objective_constant = float(self.objective_constant)
        self.num_equalities = int(self.num_equalities)

    @property
    def num_variables(self) -> int:
        return self.constraint_matrix.shape[1]

    @property
    def num_constraints(self) -> int:
        return self.constraint_matrix.shape[0]

    def copy(self) -> "QuadraticProgrammingProblem":
        return copy.deepcopy(self)


def linear_programming_problem(
    variable_lower_bound,
    variable_upper_bound,
    objective_vector,
    objective_constant,
    constraint_matrix,
    right_hand_side,
    num_equalities,
) -> QuadraticProgrammingProblem:
    n = len(variable_lower_bound)
    return QuadraticProgrammingProblem(
        variable_lower_bound,
        variable_upper_bound,
        sp.csc_matrix((n, n), dtype=np.float64),
        objective_vector,
        objective_constant,
        constraint_matrix,
        right_hand_side,
        num_equalities,
    )


def is_linear_programming_problem(problem: QuadraticProgrammingProblem) -> bool:
    """quadratic_programming.jl: iszero(objective_matrix) -- a test on VALUES: a Q that stores only
    explicit zeros (QUADOBJ entries of 0, cancelled duplicates) is a linear program."""
    return not np.any(problem.objective_matrix.data != 0.0)


def equality_range(problem):
    return range(0, problem.num_equalities)


def inequality_range(problem):
    return range(problem.num_equalities, problem.num_constraints)


@dataclass
class ScaledQpProblem:
    original_qp: QuadraticProgrammingProblem
    scaled_qp: QuadraticProgrammingProblem
    constraint_rescaling: np.ndarray
    variable_rescaling: np.ndarray


@dataclass
class CachedQuadraticProgramInfo:
    l_inf_norm_primal_linear_objective: float
    l_inf_norm_primal_right_hand_side: float
    l2_norm_primal_linear_objective: float
    l2_norm_primal_right_hand_side: float


def cached_quadratic_program_info(qp: QuadraticProgrammingProblem) -> CachedQuadraticProgramInfo:
    def inf(v):
        return float(np.max(np.abs(v))) if v.size else 0.0

    return CachedQuadraticProgramInfo(
        inf(qp.objective_vector),
        inf(qp.right_hand_side),
        float(np.linalg.norm(qp.objective_vector, 2)),
        float(np.linalg.norm(qp.right_hand_side, 2)),
    )


def validate(p: QuadraticProgrammingProblem) -> bool:
    """src/preprocess.jl:18-84; raises ValueError where the reference calls error()."""
    problems = []
    n = len(p.variable_lower_bound)
    if n != len(p.variable_upper_bound):
        problems.append("length(variable_lower_bound) != length(variable_upper_bound)")
    if n != len(p.objective_vector):
        problems.append("length(variable_lower_bound) != length(objective_vector)")
    if p.constraint_matrix.shape[0] != len(p.right_hand_side):
        problems.append("size(constraint_matrix,1) != length(right_hand_side)")
    if p.constraint_matrix.shape[1] != len(p.objective_vector):
        problems.append("size(constraint_matrix,2) != length(objective_vector)")
    if p.objective_matrix.shape != (len(p.objective_vector), len(p.objective_vector)):
        problems.append("objective_matrix is not square with length(objective_vector)")
    if np.any(p.variable_lower_bound == np.inf):
        problems.append("variable_lower_bound contains +Inf")
    if np.any(p.variable_upper_bound == -np.inf):
        problems.append("variable_upper_bound contains -Inf")
    if np.any(np.isnan(p.variable_lower_bound)) or np.any(np.isnan(p.variable_upper_bound)):
        problems.append("NaN found in variable bounds")
    if not np.all(np.isfinite(p.right_hand_side)):
        problems.append("NaN or Inf found in right hand side")
    if not np.all(np.isfinite(p.objective_vector)):
        problems.append("NaN or Inf found in objective vector")
    if not np.all(np.isfinite(p.constraint_matrix.data)):
        problems.append("NaN or Inf found in constraint matrix")
    if not np.all(np.isfinite(p.objective_matrix.data)):
        problems.append("NaN or Inf found in objective matrix")
    if problems:
        raise ValueError(
            "Error found when validating QuadraticProgrammingProblem: " + "; ".join(problems)
        )
    return True
