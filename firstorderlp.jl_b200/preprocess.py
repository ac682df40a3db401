"""Host-side preprocessing, kept on the host exactly as in the reference.

Mirrors src/preprocess.jl: l2_norm (:99-113), remove_empty_rows (:122-144),
remove_empty_columns (:155-186), presolve / undo_presolve (:236-340),
l2_norm_rescaling (:358-372), ruiz_rescaling (:412-477),
pock_chambolle_rescaling (:508-539), scale_problem (:555-573),
rescale_problem (:631-687).

Arithmetic is written so that every element sees the same floating-point
operations, in the same order, as the Julia expressions (e.g. the matrix entry
becomes ((1/E_i) * a) * (1/D_j), sums run in CSC storage order), so the CPU
oracle's C restatement of the same functions agrees bit for bit.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import scipy.sparse as sp

from .problem import QuadraticProgrammingProblem, ScaledQpProblem, is_linear_programming_problem


def _col_index(matrix: sp.csc_matrix) -> np.ndarray:
    return np.repeat(np.arange(matrix.shape[1], dtype=np.int64), np.diff(matrix.indptr))


def _max_abs(matrix: sp.csc_matrix, dims: int) -> np.ndarray:
    """maximum(abs, matrix, dims = dims): dims=1 -> per column, dims=2 -> per row."""
    m, n = matrix.shape
    out = np.zeros(n if dims == 1 else m)
    if matrix.nnz == 0:
        return out
    absdata = np.abs(matrix.data)
    if dims == 1:
        counts = np.diff(matrix.indptr)
        nonempty = counts > 0
        starts = matrix.indptr[:-1][nonempty]
        out[nonempty] = np.maximum.reduceat(absdata, starts)
    else:
        np.maximum.at(out, matrix.indices, absdata)
    return out


def l2_norm(matrix: sp.csc_matrix, dimension: int) -> np.ndarray:
    """src/preprocess.jl:99-113 (dimension 1: column norms, 2: row norms)."""
    m, n = matrix.shape
    scale_factor = _max_abs(matrix, dimension)
    scale_factor[scale_factor == 0.0] = 1.0
    inv = 1.0 / scale_factor
    if dimension == 1:
        idx = _col_index(matrix)
        t = matrix.data * inv[idx]
        length = n
    else:
        idx = matrix.indices
        t = inv[idx] * matrix.data
        length = m
    sums = np.bincount(idx, weights=t * t, minlength=length) if matrix.nnz else np.zeros(length)
    return scale_factor * np.sqrt(sums)


def scale_problem(
    problem: QuadraticProgrammingProblem,
    constraint_rescaling: np.ndarray,
    variable_rescaling: np.ndarray,
) -> None:
    """src/preprocess.jl:555-573, in place."""
    assert np.all(constraint_rescaling > 0)
    assert np.all(variable_rescaling > 0)
    inv_var = 1.0 / variable_rescaling
    inv_con = 1.0 / constraint_rescaling
    problem.objective_vector /= variable_rescaling
    Q = problem.objective_matrix
    if Q.nnz:
        Q.data = (inv_var[Q.indices] * Q.data) * inv_var[_col_index(Q)]
    problem.variable_upper_bound *= variable_rescaling
    problem.variable_lower_bound *= variable_rescaling
    problem.right_hand_side /= constraint_rescaling
    A = problem.constraint_matrix
    if A.nnz:
        A.data = (inv_con[A.indices] * A.data) * inv_var[_col_index(A)]


def unscale_problem(problem, constraint_rescaling, variable_rescaling) -> None:
    """src/preprocess.jl:580-587"""
    scale_problem(problem, 1.0 / constraint_rescaling, 1.0 / variable_rescaling)


def l2_norm_rescaling(problem: QuadraticProgrammingProblem) -> Tuple[np.ndarray, np.ndarray]:
    """src/preprocess.jl:358-372"""
    norm_of_rows = l2_norm(problem.constraint_matrix, 2)
    norm_of_columns = l2_norm(problem.constraint_matrix, 1)
    norm_of_rows[norm_of_rows == 0.0] = 1.0
    norm_of_columns[norm_of_columns == 0.0] = 1.0
    column_rescale_factor = np.sqrt(norm_of_columns)
    row_rescale_factor = np.sqrt(norm_of_rows)
    scale_problem(problem, row_rescale_factor, column_rescale_factor)
    return row_rescale_factor, column_rescale_factor


def ruiz_rescaling(
    problem: QuadraticProgrammingProblem, num_iterations: int, p: float = np.inf
) -> Tuple[np.ndarray, np.ndarray]:
    """src/preprocess.jl:412-477"""
    num_constraints, num_variables = problem.constraint_matrix.shape
    cum_constraint_rescaling = np.ones(num_constraints)
    cum_variable_rescaling = np.ones(num_variables)
    for _ in range(num_iterations):
        A = problem.constraint_matrix
        Q = problem.objective_matrix
        if p == np.inf:
            variable_rescaling = np.sqrt(np.maximum(_max_abs(A, 1), _max_abs(Q, 1)))
        else:
            assert p == 2
            variable_rescaling = np.sqrt(np.sqrt(l2_norm(A, 1) ** 2 + l2_norm(Q, 1) ** 2))
        variable_rescaling[variable_rescaling == 0.0] = 1.0
        if num_constraints == 0:
            constraint_rescaling = np.zeros(0)
        else:
            if p == np.inf:
                constraint_rescaling = np.sqrt(_max_abs(A, 2))
            else:
                norm_of_rows = l2_norm(A, 2)
                if not np.any(Q.data != 0.0):
                    target_row_norm = np.sqrt(num_variables / num_constraints)
                else:
                    target_row_norm = np.sqrt(num_variables / (num_constraints + num_variables))
                constraint_rescaling = np.sqrt(norm_of_rows / target_row_norm)
            constraint_rescaling[constraint_rescaling == 0.0] = 1.0
        scale_problem(problem, constraint_rescaling, variable_rescaling)
        cum_constraint_rescaling *= constraint_rescaling
        cum_variable_rescaling *= variable_rescaling
    return cum_constraint_rescaling, cum_variable_rescaling


def pock_chambolle_rescaling(
    problem: QuadraticProgrammingProblem, alpha: float
) -> Tuple[np.ndarray, np.ndarray]:
    """src/preprocess.jl:508-539. Julia's mapreduce over a sparse matrix with
    dims also folds f(0) for each structural zero (|0|^0 == 1)."""
    assert 0 <= alpha <= 2
    A = problem.constraint_matrix
    m, n = A.shape
    absdata = np.abs(A.data)
    cols = _col_index(A)
    col_pow = np.power(absdata, 2 - alpha) if alpha != 1.0 else absdata
    row_pow = np.power(absdata, alpha) if alpha != 1.0 else absdata
    col_sums = np.bincount(cols, weights=col_pow, minlength=n) if A.nnz else np.zeros(n)
    row_sums = np.bincount(A.indices, weights=row_pow, minlength=m) if A.nnz else np.zeros(m)
    zero_col = 0.0 ** (2 - alpha)
    zero_row = 0.0 ** alpha
    col_sums = col_sums + zero_col * (m - np.diff(A.indptr)).astype(np.float64)
    row_counts = np.bincount(A.indices, minlength=m) if A.nnz else np.zeros(m, dtype=np.int64)
    row_sums = row_sums + zero_row * (n - row_counts).astype(np.float64)
    variable_rescaling = np.sqrt(col_sums)
    constraint_rescaling = np.sqrt(row_sums)
    variable_rescaling[variable_rescaling == 0.0] = 1.0
    constraint_rescaling[constraint_rescaling == 0.0] = 1.0
    scale_problem(problem, constraint_rescaling, variable_rescaling)
    return constraint_rescaling, variable_rescaling


def row_permute_in_place(matrix: sp.csc_matrix, old_row_to_new) -> None:
    """src/preprocess.jl:590-623: permutes the rows of a CSC matrix in place according to the map
    `old_row_to_new` (0-based here), keeping the row indices sorted inside every column."""
    old_row_to_new = np.asarray(old_row_to_new, dtype=np.int64)
    new_rows = old_row_to_new[matrix.indices]
    # stable sort by (column, new row): columns are contiguous, so one lexsort does every column
    order = np.lexsort((new_rows, _col_index(matrix)))
    matrix.indices[:] = new_rows[order].astype(matrix.indices.dtype)
    matrix.data[:] = matrix.data[order]
    matrix.has_sorted_indices = True


def rescale_problem(
    l_inf_ruiz_iterations: int,
    l2_norm_rescaling_flag: bool,
    pock_chambolle_alpha: Optional[float],
    verbosity: int,
    original_problem: QuadraticProgrammingProblem,
) -> ScaledQpProblem:
    """src/preprocess.jl:631-687. `original_problem` is not modified."""
    problem = original_problem.copy()
    num_constraints, num_variables = problem.constraint_matrix.shape
    constraint_rescaling = np.ones(num_constraints)
    variable_rescaling = np.ones(num_variables)
    if l_inf_ruiz_iterations > 0:
        con, var = ruiz_rescaling(problem, l_inf_ruiz_iterations, np.inf)
        constraint_rescaling *= con
        variable_rescaling *= var
    if l2_norm_rescaling_flag:
        con, var = l2_norm_rescaling(problem)
        constraint_rescaling *= con
        variable_rescaling *= var
    if pock_chambolle_alpha is not None:
        con, var = pock_chambolle_rescaling(problem, pock_chambolle_alpha)
        constraint_rescaling *= con
        variable_rescaling *= var
    if verbosity >= 3:
        if l_inf_ruiz_iterations == 0 and not l2_norm_rescaling_flag:
            print("No rescaling.")
        else:
            print(
                f"Problem after rescaling (Ruiz iterations = {l_inf_ruiz_iterations}, "
                f"l2_norm_rescaling = {str(l2_norm_rescaling_flag).lower()}):"
            )
    return ScaledQpProblem(original_problem, problem, constraint_rescaling, variable_rescaling)


# ---------------------------------------------------------------------------
# presolve (empty rows / columns)
# ---------------------------------------------------------------------------
@dataclass
class PresolveInfo:
    original_primal_size: int
    original_dual_size: int
    empty_rows: List[int]
    empty_columns: List[int]
    variable_lower_bound: np.ndarray
    variable_upper_bound: np.ndarray


def remove_empty_rows(problem: QuadraticProgrammingProblem) -> List[int]:
    """src/preprocess.jl:122-144 (0-based row ids)."""
    m = problem.constraint_matrix.shape[0]
    seen_row = np.zeros(m, dtype=bool)
    seen_row[problem.constraint_matrix.indices] = True
    empty_rows = np.flatnonzero(~seen_row)
    for row in empty_rows:
        if row >= problem.num_equalities and problem.right_hand_side[row] > 0.0:
            raise ValueError("The problem is infeasible.")
        if row < problem.num_equalities and problem.right_hand_side[row] != 0.0:
            raise ValueError("The problem is infeasible.")
    if empty_rows.size:
        problem.constraint_matrix = sp.csc_matrix(problem.constraint_matrix.tocsr()[seen_row, :])
        problem.constraint_matrix.sort_indices()
        problem.right_hand_side = problem.right_hand_side[seen_row]
        problem.num_equalities -= int(np.sum(empty_rows < problem.num_equalities))
    return [int(r) for r in empty_rows]


def remove_empty_columns(problem: QuadraticProgrammingProblem) -> List[int]:
    """src/preprocess.jl:155-186 (0-based column ids)."""
    assert is_linear_programming_problem(problem)
    counts = np.diff(problem.constraint_matrix.indptr)
    is_empty = counts == 0
    empty_columns = np.flatnonzero(is_empty)
    if empty_columns.size == 0:
        return []
    for col in empty_columns:
        coef = problem.objective_vector[col]
        if coef >= 0:
            problem.objective_constant += problem.variable_lower_bound[col] * coef
        else:
            problem.objective_constant += problem.variable_upper_bound[col] * coef
    keep = ~is_empty
    problem.constraint_matrix = sp.csc_matrix(problem.constraint_matrix[:, keep])
    problem.constraint_matrix.sort_indices()
    problem.objective_vector = problem.objective_vector[keep]
    problem.variable_lower_bound = problem.variable_lower_bound[keep]
    problem.variable_upper_bound = problem.variable_upper_bound[keep]
    problem.objective_matrix = sp.csc_matrix(problem.objective_matrix[keep, :][:, keep])
    return [int(c) for c in empty_columns]


def presolve(problem: QuadraticProgrammingProblem, verbosity: int = 1) -> PresolveInfo:
    """src/preprocess.jl:236-271 (transform_bounds is not used by the PDHG path)."""
    saved_l = problem.variable_lower_bound.copy()
    saved_u = problem.variable_upper_bound.copy()
    original_dual_size, original_primal_size = problem.constraint_matrix.shape
    empty_rows = remove_empty_rows(problem)
    if is_linear_programming_problem(problem):
        empty_columns = remove_empty_columns(problem)
    else:
        empty_columns = []
    if verbosity >= 1:
        nz_by_row = np.bincount(
            problem.constraint_matrix.indices, minlength=problem.constraint_matrix.shape[0]
        )
        num_single = int(np.sum(nz_by_row == 1))
        if num_single > 0:
            print(f"{num_single} constraints involving exactly a single variable")
    return PresolveInfo(
        original_primal_size, original_dual_size, empty_rows, empty_columns, saved_l, saved_u
    )


def recover_original_solution(solution, empty_indices, original_size) -> np.ndarray:
    """src/preprocess.jl:299-313"""
    mask = np.ones(original_size, dtype=bool)
    mask[np.asarray(empty_indices, dtype=np.int64)] = False
    out = np.zeros(original_size)
    out[mask] = solution[: int(mask.sum())]
    return out


def undo_presolve(presolve_info: PresolveInfo, primal_solution, dual_solution):
    """src/preprocess.jl:315-340"""
    primal = recover_original_solution(
        primal_solution, presolve_info.empty_columns, presolve_info.original_primal_size
    )
    primal = np.minimum(
        presolve_info.variable_upper_bound, np.maximum(presolve_info.variable_lower_bound, primal)
    )
    dual = recover_original_solution(
        dual_solution, presolve_info.empty_rows, presolve_info.original_dual_size
    )
    return primal, dual
