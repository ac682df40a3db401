"""ctypes wrapper of the CPU oracle (oracle/libfolp_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py. Nothing under
firstorderlp.jl_b200/ imports this module.

`optimize()` below restates FirstOrderLp.optimize(::PdhgParameters, qp)
(src/primal_dual_hybrid_gradient.jl:782-1049) end to end on the CPU: the host
half (:786-859) uses the oracle's own C restatement of rescale_problem, the
loop (:862-1048) is oracle_solve.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
import sys
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import folp_b200  # noqa: E402
from folp_b200 import _abi, _marshal  # noqa: E402
from folp_b200._abi import FolpDebugScalars, FolpEval, FolpParams, FolpProblem  # noqa: E402
from folp_b200.params import (  # noqa: E402
    AdaptiveStepsizeParams,
    ConstantStepsizeParams,
    MalitskyPockStepsizeParameters,
    PdhgParameters,
)
from folp_b200.problem import (  # noqa: E402
    QuadraticProgrammingProblem,
    ScaledQpProblem,
    cached_quadratic_program_info,
    validate,
)
from folp_b200.solve_log import (  # noqa: E402
    SaddlePointOutput,
    iteration_stats_from_eval,
    termination_reason_to_string,
)

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int64)
_LIB = None


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "libfolp_oracle.so")
    src = os.path.join(_HERE, "folp_oracle.c")
    stale = (not os.path.exists(path)) or any(
        os.path.getmtime(p) > os.path.getmtime(path)
        for p in (src, os.path.join(_HERE, "folp_oracle.h"),
                  os.path.join(_ROOT, "include", "folp_b200.h"))
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfolp_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return path


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oracle_create.argtypes = [C.POINTER(FolpProblem), C.POINTER(FolpParams), C.POINTER(C.c_void_p)]
        L.oracle_run.argtypes = [C.c_void_p, C.POINTER(FolpEval)]
        L.oracle_solve.argtypes = [C.c_void_p, C.POINTER(FolpEval), C.c_int64, C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32), _pd, _pd]
        L.oracle_get_solution.argtypes = [C.c_void_p, C.c_int, C.c_int, _pd, _pd]
        L.oracle_debug_attempts.argtypes = [C.c_void_p, C.c_int64]
        L.oracle_debug_state.argtypes = [C.c_void_p, _pd, _pd, _pd, _pd, _pd, C.POINTER(FolpDebugScalars)]
        L.oracle_debug_set_state.argtypes = [C.c_void_p, _pd, _pd, C.c_double, C.c_double]
        L.oracle_debug_spmv.argtypes = [C.c_void_p, C.c_int, _pd, _pd]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_destroy.restype = None
        L.oracle_basic_seconds.argtypes = [C.c_void_p]
        L.oracle_basic_seconds.restype = C.c_double
        L.oracle_trust_region.argtypes = [C.c_int64, _pd, _pd, _pd, _pd, _pd, C.c_double, C.c_int, _pd, _pd]
        L.oracle_bound_optimal_objective.argtypes = [C.POINTER(FolpProblem), _pd, _pd, _pd, _pd,
                                                     C.c_double, C.c_int, C.c_int, _pd, _pd, _pd]
        L.oracle_iteration_stats.argtypes = [C.POINTER(FolpProblem), _pd, _pd, _pd, _pd, C.c_double,
                                             C.c_double, C.POINTER(FolpEval)]
        L.oracle_dual_stats.argtypes = [C.POINTER(FolpProblem), _pd, _pd, _pd, _pd, _pd]
        L.oracle_max_primal_violation.argtypes = [C.POINTER(FolpProblem), _pd]
        L.oracle_max_primal_violation.restype = C.c_double
        L.oracle_primal_obj.argtypes = [C.POINTER(FolpProblem), _pd]
        L.oracle_primal_obj.restype = C.c_double
        L.oracle_lagrangian_value.argtypes = [C.POINTER(FolpProblem), _pd, _pd]
        L.oracle_lagrangian_value.restype = C.c_double
        L.oracle_select_initial_primal_weight.argtypes = [C.POINTER(FolpProblem), _pd, _pd, C.c_double]
        L.oracle_select_initial_primal_weight.restype = C.c_double
        L.oracle_check_termination.argtypes = [C.POINTER(FolpParams), C.POINTER(FolpProblem), C.POINTER(FolpEval)]
        L.oracle_rescale_problem.argtypes = [C.c_int64, C.c_int64, _pi, _pi, _pd, _pi, _pi, _pd, _pd, _pd,
                                             _pd, _pd, C.c_int, C.c_int, C.c_int, C.c_double, _pd, _pd]
        L.oracle_l2_norm.argtypes = [C.c_int64, C.c_int64, _pi, _pi, _pd, C.c_int, _pd]
        _LIB = L
    return _LIB


def set_threads(threads: int) -> None:
    """All-cores timing mode of the two sparse products (bit-identical results); 1 = serial, the
    reference's behaviour and the default."""
    lib().oracle_set_threads(int(threads))


def get_threads() -> int:
    return int(lib().oracle_get_threads())


def openmp_enabled() -> bool:
    return bool(lib().oracle_openmp_enabled())


def _d(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_pd)


# ---------------------------------------------------------------------------
# unit-level wrappers
# ---------------------------------------------------------------------------
def problem_struct(problem: QuadraticProgrammingProblem):
    """QuadraticProgrammingProblem taken as-is (identity rescaling) -> holder."""
    return _marshal.make_problem(_marshal.unscaled_as_scaled(problem))


def trust_region(center, objective, lb, ub, weights, radius, approx=False) -> Tuple[np.ndarray, float]:
    c, g, l, u, w = map(_d, (center, objective, lb, ub, weights))
    sol = np.zeros_like(c)
    val = C.c_double()
    lib().oracle_trust_region(len(c), _p(c), _p(g), _p(l), _p(u), _p(w), float(radius), int(approx),
                              _p(sol), C.cast(C.byref(val), _pd))
    return sol, val.value


def bound_optimal_objective(problem, x, y, wp, wd, radius, norm_kind, approx=False):
    h = problem_struct(problem)
    x, y, wp, wd = map(_d, (x, y, wp, wd))
    out3 = np.zeros(3)
    xs = np.zeros_like(x)
    ys = np.zeros_like(y)
    lib().oracle_bound_optimal_objective(h.byref(), _p(x), _p(y), _p(wp), _p(wd), float(radius),
                                         int(norm_kind), int(approx), _p(out3), _p(xs), _p(ys))
    return {"lagrangian_value": out3[0], "lower_bound_value": out3[1], "upper_bound_value": out3[2],
            "primal_solution": xs, "dual_solution": ys}


def iteration_stats(problem, x, y, xray, yray, eps_abs, eps_rel, candidate_type=1) -> FolpEval:
    h = problem_struct(problem)
    e = FolpEval()
    e.candidate_type = candidate_type
    x, y, xray, yray = map(_d, (x, y, xray, yray))
    lib().oracle_iteration_stats(h.byref(), _p(x), _p(y), _p(xray), _p(yray), eps_abs, eps_rel, C.byref(e))
    return e


def dual_stats(problem, x, y):
    h = problem_struct(problem)
    x, y = _d(x), _d(y)
    n, m = problem.num_variables, problem.num_constraints
    dres = np.zeros(m - problem.num_equalities + n)
    rc = np.zeros(n)
    dobj = C.c_double()
    lib().oracle_dual_stats(h.byref(), _p(x), _p(y), C.cast(C.byref(dobj), _pd), _p(dres), _p(rc))
    return dobj.value, dres, rc


def max_primal_violation(problem, x) -> float:
    h = problem_struct(problem)
    x = _d(x)
    return lib().oracle_max_primal_violation(h.byref(), _p(x))


def primal_obj(problem, x) -> float:
    h = problem_struct(problem)
    x = _d(x)
    return lib().oracle_primal_obj(h.byref(), _p(x))


def lagrangian_value(problem, x, y) -> float:
    h = problem_struct(problem)
    x, y = _d(x), _d(y)
    return lib().oracle_lagrangian_value(h.byref(), _p(x), _p(y))


def select_initial_primal_weight(problem, wp, wd, primal_importance) -> float:
    h = problem_struct(problem)
    wp, wd = _d(wp), _d(wd)
    return lib().oracle_select_initial_primal_weight(h.byref(), _p(wp), _p(wd), primal_importance)


def check_termination(params: FolpParams, problem_holder, stats: FolpEval) -> int:
    return lib().oracle_check_termination(C.byref(params), problem_holder.byref(), C.byref(stats))


def l2_norm(matrix, dimension) -> np.ndarray:
    import scipy.sparse as sp
    A = sp.csc_matrix(matrix, dtype=np.float64)
    A.sort_indices()
    m, n = A.shape
    ip = np.ascontiguousarray(A.indptr, dtype=np.int64)
    ix = np.ascontiguousarray(A.indices, dtype=np.int64)
    out = np.zeros(n if dimension == 1 else m)
    lib().oracle_l2_norm(m, n, ip.ctypes.data_as(_pi), ix.ctypes.data_as(_pi), _p(_d(A.data)), dimension, _p(out))
    return out


def rescale_problem(l_inf_ruiz_iterations, l2_norm_rescaling, pock_chambolle_alpha,
                    original_problem: QuadraticProgrammingProblem, ruiz_p: int = 0) -> ScaledQpProblem:
    """C restatement of src/preprocess.jl:631-687. original_problem is untouched."""
    P = original_problem.copy()
    A = P.constraint_matrix
    Q = P.objective_matrix
    m, n = A.shape
    ip = np.ascontiguousarray(A.indptr, dtype=np.int64)
    ix = np.ascontiguousarray(A.indices, dtype=np.int64)
    data = _d(A.data).copy()
    if Q.nnz:
        qip = np.ascontiguousarray(Q.indptr, dtype=np.int64)
        qix = np.ascontiguousarray(Q.indices, dtype=np.int64)
        qdata = _d(Q.data).copy()
        qargs = (qip.ctypes.data_as(_pi), qix.ctypes.data_as(_pi), _p(qdata))
    else:
        qdata = None
        qargs = (None, None, None)
    c, l, u, b = (_d(v).copy() for v in (P.objective_vector, P.variable_lower_bound,
                                         P.variable_upper_bound, P.right_hand_side))
    con = np.zeros(m)
    var = np.zeros(n)
    alpha = -1.0 if pock_chambolle_alpha is None else float(pock_chambolle_alpha)
    lib().oracle_rescale_problem(m, n, ip.ctypes.data_as(_pi), ix.ctypes.data_as(_pi), _p(data), *qargs,
                                 _p(c), _p(l), _p(u), _p(b), int(l_inf_ruiz_iterations), int(ruiz_p),
                                 int(bool(l2_norm_rescaling)), alpha, _p(con), _p(var))
    A.data[:] = data
    if qdata is not None:
        Q.data[:] = qdata
    P.objective_vector, P.variable_lower_bound, P.variable_upper_bound, P.right_hand_side = c, l, u, b
    return ScaledQpProblem(original_problem, P, con, var)


# ---------------------------------------------------------------------------
# solver handle
# ---------------------------------------------------------------------------
class OracleSolver:
    """Same surface as folp_b200.lib.Solver, computed by the CPU oracle."""

    def __init__(self, problem_holder, params: FolpParams):
        self._holder = problem_holder
        self.params = params
        self.n = problem_holder.struct.num_variables
        self.m = problem_holder.struct.num_constraints
        self._h = C.c_void_p()
        rc = lib().oracle_create(problem_holder.byref(), C.byref(params), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"oracle_create failed with status {rc}")

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self) -> FolpEval:
        e = FolpEval()
        rc = lib().oracle_run(self._h, C.byref(e))
        if rc != 0:
            raise RuntimeError(f"oracle_run failed with status {rc}")
        return e

    def solve(self, max_evals: int = 1 << 16):
        evals = (FolpEval * max_evals)()
        n_ev = C.c_int64()
        reason = C.c_int32()
        iters = C.c_int32()
        x = np.zeros(self.n)
        y = np.zeros(self.m)
        rc = lib().oracle_solve(self._h, evals, max_evals, C.byref(n_ev), C.byref(reason), C.byref(iters),
                                _p(x), _p(y))
        if rc != 0:
            raise RuntimeError(f"oracle_solve failed with status {rc}")
        return x, y, reason.value, iters.value, [evals[i] for i in range(n_ev.value)]

    def get_solution(self, which=0, unscaled=True):
        x = np.zeros(self.n)
        y = np.zeros(self.m)
        lib().oracle_get_solution(self._h, which, int(unscaled), _p(x), _p(y))
        return x, y

    def debug_attempts(self, k: int):
        lib().oracle_debug_attempts(self._h, k)

    def debug_state(self):
        x = np.zeros(self.n); y = np.zeros(self.m); aty = np.zeros(self.n)
        sx = np.zeros(self.n); sy = np.zeros(self.m)
        s = FolpDebugScalars()
        lib().oracle_debug_state(self._h, _p(x), _p(y), _p(aty), _p(sx), _p(sy), C.byref(s))
        return {"x": x, "y": y, "dual_product": aty, "sum_x": sx, "sum_y": sy, **s.as_dict()}

    def debug_set_state(self, x, y, step_size=-1.0, primal_weight=-1.0):
        x = None if x is None else _d(x)
        y = None if y is None else _d(y)
        lib().oracle_debug_set_state(self._h, _p(x), _p(y), step_size, primal_weight)

    def spmv(self, v, transpose=False):
        v = _d(v)
        out = np.zeros(self.n if transpose else self.m)
        lib().oracle_debug_spmv(self._h, int(transpose), _p(v), _p(out))
        return out

    def basic_seconds(self) -> float:
        return lib().oracle_basic_seconds(self._h)


def estimate_maximum_singular_value(matrix, probability_of_failure=0.001,
                                    desired_relative_error=0.2, seed=1):
    """src/primal_dual_hybrid_gradient.jl:414-440. PARITY UNPINNED: the start
    vector comes from Julia's MersenneTwister randn stream, which cannot be
    reproduced here; a NumPy normal stream is used instead (the reference's own
    test only pins the converged solution and a constant step across stats)."""
    n = matrix.shape[1]
    epsilon = 1.0 - (1.0 - desired_relative_error) ** 2
    x = np.random.RandomState(seed).randn(n)

    def failure_probability(k):
        if k < 2 or epsilon <= 0.0:
            return 1.0
        return min(0.824, 0.354 / math.sqrt(epsilon * (k - 1))) * math.sqrt(n) * (1.0 - epsilon) ** (k - 0.5)

    k = 0
    At = matrix.T.tocsr()
    while failure_probability(k) > probability_of_failure:
        x = x / np.linalg.norm(x, 2)
        x = At @ (matrix @ x)
        k += 1
    sigma = math.sqrt(float(x @ (At @ (matrix @ x))) / float(np.linalg.norm(x, 2) ** 2))
    return sigma, k


def host_setup(params: PdhgParameters, original_problem: QuadraticProgrammingProblem,
               scaled: Optional[ScaledQpProblem] = None):
    """The host half of optimize(): pdhg.jl:786-859. Returns (holder, FolpParams, scaled)."""
    validate(original_problem)
    cache = cached_quadratic_program_info(original_problem)
    if scaled is None:
        scaled = rescale_problem(params.l_inf_ruiz_iterations, params.l2_norm_rescaling,
                                 params.pock_chambolle_alpha, original_problem)
    problem = scaled.scaled_qp
    if params.primal_importance <= 0 or not math.isfinite(params.primal_importance):
        raise ValueError("primal_importance must be positive and finite")
    pol = params.step_size_policy_params
    if isinstance(pol, (AdaptiveStepsizeParams, MalitskyPockStepsizeParameters)):
        kkt0 = 0.5
        step0 = _marshal.initial_step_size_inf_norm(problem)
    else:
        sigma, k = estimate_maximum_singular_value(problem.constraint_matrix, 0.001, 0.2)
        step0 = (1 - 0.2) / sigma
        kkt0 = float(k)
    if params.scale_invariant_initial_primal_weight:
        n, m = problem.num_variables, problem.num_constraints
        pw0 = select_initial_primal_weight(problem, np.ones(n), np.ones(m), params.primal_importance)
    else:
        pw0 = params.primal_importance
    holder = _marshal.make_problem(scaled, cache)
    fparams = _marshal.make_params(params, step0, pw0, kkt0)
    return holder, fparams, scaled


def optimize(params: PdhgParameters, original_problem: QuadraticProgrammingProblem) -> SaddlePointOutput:
    holder, fparams, _ = host_setup(params, original_problem)
    solver = OracleSolver(holder, fparams)
    x, y, reason, iters, evals = solver.solve()
    solver.close()
    stats = [iteration_stats_from_eval(e) for e in evals]
    reason = _abi.TerminationReason(reason)
    return SaddlePointOutput(x, y, reason, termination_reason_to_string(reason), iters, stats)
