"""ctypes binding of libfolp_b200.so (the C ABI of include/folp_b200.h).

This is what the Julia host would do with `ccall`; see INTEGRATION.md. There
is no CPU fallback: if the shared library is missing or no sm_100 device is
present, `Solver` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from ._abi import FolpDebugScalars, FolpDist, FolpEval, FolpParams, FolpProblem, Status

_HERE = os.path.dirname(os.path.abspath(__file__))
# FOLP_B200_LIB selects another build of the same library (kernel variants under development)
LIB_PATH = os.environ.get("FOLP_B200_LIB") or os.path.join(_HERE, "libfolp_b200.so")
_pd = C.POINTER(C.c_double)
_LIB: Optional[C.CDLL] = None

EXPORTS = [
    "folp_nccl_unique_id", "folp_create_multi", "folp_partition", "folp_rescale_problem", "folp_shard_info", "folp_exchange_mode", "folp_create", "folp_run", "folp_solve", "folp_get_solution",
    "folp_debug_attempts", "folp_debug_state", "folp_debug_set_state", "folp_debug_spmv",
    "folp_debug_profile_attempts", "folp_debug_time_spmv", "folp_debug_host_spmv", "folp_debug_host_prepare", "folp_debug_host_problem_spmv", "folp_debug_stream", "folp_counters", "folp_destroy", "folp_last_error", "folp_build_info",
]


class FolpError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libfolp_b200: {Status(status).name}: {message}")
        self.status = Status(status)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles csrc/*.cu for sm_100a with nvcc (cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "folp_b200.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(LIB_PATH) for p in srcs)
    if force or stale:
        cmd = ["make", "-C", csrc, "-j4"] + (["-B"] if force else [])
        subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL,
                              stderr=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


def lib() -> C.CDLL:
    """Loads libfolp_b200.so; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "folp_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.folp_nccl_unique_id.argtypes = [C.c_void_p]
        _pi64 = C.POINTER(C.c_int64)
        L.folp_partition.argtypes = [C.c_int64, C.c_int64, C.c_int64, _pi64, C.c_int32, C.c_int32, _pi64, _pi64]
        L.folp_shard_info.argtypes = [C.c_void_p, _pi64, _pi64, _pi64, _pi64, _pi64]
        L.folp_exchange_mode.argtypes = [C.c_void_p]
        L.folp_create.argtypes = [C.POINTER(FolpProblem), C.POINTER(FolpParams), C.POINTER(FolpDist),
                                  C.POINTER(C.c_void_p)]
        L.folp_create_multi.argtypes = [C.POINTER(FolpProblem), C.POINTER(FolpParams), C.c_int32,
                                        C.POINTER(C.c_int32), C.POINTER(C.c_void_p)]
        L.folp_run.argtypes = [C.c_void_p, C.POINTER(FolpEval)]
        L.folp_solve.argtypes = [C.c_void_p, C.POINTER(FolpEval), C.c_int64, C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_int32), _pd, _pd]
        L.folp_get_solution.argtypes = [C.c_void_p, C.c_int, C.c_int, _pd, _pd]
        L.folp_debug_attempts.argtypes = [C.c_void_p, C.c_int64]
        L.folp_debug_state.argtypes = [C.c_void_p, _pd, _pd, _pd, _pd, _pd, C.POINTER(FolpDebugScalars)]
        L.folp_debug_set_state.argtypes = [C.c_void_p, _pd, _pd, C.c_double, C.c_double]
        L.folp_debug_spmv.argtypes = [C.c_void_p, C.c_int, _pd, _pd]
        L.folp_debug_profile_attempts.argtypes = [C.c_void_p, C.c_int64, _pd, C.POINTER(C.c_int64)]
        L.folp_debug_time_spmv.argtypes = [C.c_void_p, C.c_int, C.c_int, _pd]
        L.folp_debug_host_spmv.argtypes = [C.c_int64, C.c_int64, _pi64, _pi64, _pd, _pd, _pd, C.c_int64, _pi64]
        L.folp_rescale_problem.argtypes = [C.c_int64, C.c_int64, C.c_int32, _pi64, _pi64, _pd, _pi64, _pi64, _pd,
                                           _pd, _pd, _pd, _pd, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                           _pd, _pd]
        L.folp_debug_host_prepare.argtypes = [C.POINTER(FolpProblem), _pd]
        L.folp_debug_host_problem_spmv.argtypes = [C.POINTER(FolpProblem), C.c_int, _pd, _pd]
        L.folp_debug_stream.argtypes = [C.c_void_p]
        L.folp_debug_stream.restype = C.c_void_p
        L.folp_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64), _pd, C.POINTER(C.c_int64)]
        L.folp_destroy.argtypes = [C.c_void_p]
        L.folp_destroy.restype = None
        L.folp_last_error.argtypes = [C.c_void_p]
        L.folp_last_error.restype = C.c_char_p
        L.folp_build_info.restype = C.c_char_p
        _LIB = L
    return _LIB


def build_info() -> str:
    return lib().folp_build_info().decode()


def _d(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(_pd)


def host_packed_spmv(A_csr, x, warps_total: int = 0):
    """y = A * x evaluated on the HOST through the library's packed work-item layout
    (folp_debug_host_spmv): checks the packing without a GPU. Returns (y, stats)."""
    import scipy.sparse as sp
    A = sp.csr_matrix(A_csr)
    A.sort_indices()
    _pi64 = C.POINTER(C.c_int64)
    rp = np.ascontiguousarray(A.indptr, dtype=np.int64)
    ci = np.ascontiguousarray(A.indices, dtype=np.int64)
    v = _d(A.data)
    xx = _d(x)
    y = np.empty(A.shape[0], dtype=np.float64)
    stats = np.zeros(5, dtype=np.int64)
    rc = lib().folp_debug_host_spmv(A.shape[0], A.shape[1], rp.ctypes.data_as(_pi64),
                                    ci.ctypes.data_as(_pi64), _p(v), _p(xx), _p(y), warps_total,
                                    stats.ctypes.data_as(_pi64))
    if rc != 0:
        raise FolpError(rc, "folp_debug_host_spmv")
    return y, dict(zip(("tiles", "sorted_groups", "narrow_rounds", "long_rows", "busiest_warp_rounds"),
                       stats.tolist()))


def rescale_problem(l_inf_ruiz_iterations, l2_norm_rescaling, pock_chambolle_alpha, original_problem,
                    ruiz_p: int = 0):
    """rescale_problem (src/preprocess.jl:631-687) on the GPU through folp_rescale_problem.
    `original_problem` is untouched; returns a ScaledQpProblem like preprocess.rescale_problem."""
    from .problem import ScaledQpProblem
    P = original_problem.copy()
    A, Q = P.constraint_matrix, P.objective_matrix
    m, n = A.shape
    _pi64 = C.POINTER(C.c_int64)
    ip = np.ascontiguousarray(A.indptr, dtype=np.int64)
    ix = np.ascontiguousarray(A.indices, dtype=np.int64)
    data = _d(A.data).copy()
    if Q.nnz:
        qip = np.ascontiguousarray(Q.indptr, dtype=np.int64)
        qix = np.ascontiguousarray(Q.indices, dtype=np.int64)
        qdata = _d(Q.data).copy()
        qargs = (qip.ctypes.data_as(_pi64), qix.ctypes.data_as(_pi64), _p(qdata))
    else:
        qdata = None
        qargs = (None, None, None)
    c, l, u, b = (_d(v).copy() for v in (P.objective_vector, P.variable_lower_bound,
                                         P.variable_upper_bound, P.right_hand_side))
    con, var = np.zeros(m), np.zeros(n)
    alpha = -1.0 if pock_chambolle_alpha is None else float(pock_chambolle_alpha)
    rc = lib().folp_rescale_problem(m, n, 0, ip.ctypes.data_as(_pi64), ix.ctypes.data_as(_pi64), _p(data),
                                    *qargs, _p(c), _p(l), _p(u), _p(b), int(l_inf_ruiz_iterations),
                                    int(ruiz_p), int(bool(l2_norm_rescaling)), alpha, _p(con), _p(var))
    if rc != 0:
        raise FolpError(rc, lib().folp_last_error(None).decode())
    A.data[:] = data
    if qdata is not None:
        Q.data[:] = qdata
    P.objective_vector, P.variable_lower_bound, P.variable_upper_bound, P.right_hand_side = c, l, u, b
    return ScaledQpProblem(original_problem, P, con, var)


def host_problem_spmv(holder, x, transpose: bool = False) -> np.ndarray:
    """A * x (or A' * x) on the HOST through folp_create's own host preparation (transposition,
    planning, packing) of a marshalled problem: no CUDA call."""
    pr = holder.struct
    xx = _d(x)
    y = np.empty(pr.num_variables if transpose else pr.num_constraints, dtype=np.float64)
    rc = lib().folp_debug_host_problem_spmv(holder.byref(), 1 if transpose else 0, _p(xx), _p(y))
    if rc != 0:
        raise FolpError(rc, "folp_debug_host_problem_spmv")
    return y


def host_prepare_ms(holder) -> float:
    """Wall-clock milliseconds of folp_create's host half (no CUDA call) on a marshalled problem."""
    ms = C.c_double(0.0)
    rc = lib().folp_debug_host_prepare(holder.byref(), C.byref(ms))
    if rc != 0:
        raise FolpError(rc, "folp_debug_host_prepare")
    return ms.value


def nccl_unique_id() -> bytes:
    """128-byte ncclUniqueId (rank 0 creates it, the host broadcasts it)."""
    buf = C.create_string_buffer(128)
    rc = lib().folp_nccl_unique_id(buf)
    if rc != 0:
        raise FolpError(rc, lib().folp_last_error(None).decode())
    return buf.raw


def partition(constraint_matrix, world_size: int):
    """The 1-D partition folp_create applies: (row_begin[world+1], col_begin[world+1])."""
    import scipy.sparse as sp
    A = sp.csc_matrix(constraint_matrix)
    m, n = A.shape
    rowval = np.ascontiguousarray(A.indices, dtype=np.int64)
    rb = np.zeros(world_size + 1, dtype=np.int64)
    cb = np.zeros(world_size + 1, dtype=np.int64)
    _pi64 = C.POINTER(C.c_int64)
    rc = lib().folp_partition(m, n, A.nnz, rowval.ctypes.data_as(_pi64), 0, world_size,
                              rb.ctypes.data_as(_pi64), cb.ctypes.data_as(_pi64))
    if rc != 0:
        raise FolpError(rc, "folp_partition: invalid argument")
    return rb, cb


def make_dist(rank: int, world_size: int, device: int, unique_id: Optional[bytes]):
    """folp_dist for one rank; keep the returned object alive until Solver() returns."""
    d = FolpDist()
    d.rank, d.world_size, d.device = rank, world_size, device
    if unique_id is not None:
        d._id = C.create_string_buffer(unique_id, 128)
        d.nccl_unique_id = C.cast(d._id, C.c_void_p)
    return d


class Solver:
    """One folp_handle: the device-resident PDHG state of one optimize() call."""

    def __init__(self, problem_holder, params: FolpParams, dist: Optional[FolpDist] = None,
                 devices: Optional[list] = None):
        """dist: one rank of a one-process-per-GPU solve (folp_dist; default under torchrun after
        distributed.init()). devices: single-process multi-GPU (folp_create_multi) on these CUDA
        devices; the environment variable FOLP_DEVICES="0,1,..." selects it for every Solver."""
        self._holder = problem_holder
        self.params = params
        self.n = problem_holder.struct.num_variables
        self.m = problem_holder.struct.num_constraints
        self._h = C.c_void_p()
        L = lib()
        if devices is None and dist is None and os.environ.get("FOLP_DEVICES"):
            devices = [int(d) for d in os.environ["FOLP_DEVICES"].split(",") if d.strip() != ""]
        self.devices = devices
        if devices is not None:
            self.dist = None
            ids = (C.c_int32 * len(devices))(*devices)
            rc = L.folp_create_multi(problem_holder.byref(), C.byref(params), len(devices), ids,
                                     C.byref(self._h))
            if rc != 0:
                raise FolpError(rc, L.folp_last_error(None).decode())
            return
        if dist is None:  # under torchrun after distributed.init(): join all ranks
            from . import distributed
            dist = distributed.new_dist()
        self.dist = dist
        rc = L.folp_create(problem_holder.byref(), C.byref(params),
                           C.byref(dist) if dist is not None else None, C.byref(self._h))
        if rc != 0:
            raise FolpError(rc, L.folp_last_error(None).decode())

    def _check(self, rc: int):
        if rc != 0:
            raise FolpError(rc, lib().folp_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().folp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def run(self) -> FolpEval:
        e = FolpEval()
        self._check(lib().folp_run(self._h, C.byref(e)))
        return e

    def solve(self, max_evals: Optional[int] = None):
        """folp_solve. The record buffer is sized from the parameters: every evaluation when
        record_iteration_stats is set (iteration_limit / termination_evaluation_frequency + the ten
        per-iteration evaluations of the start), the final record only otherwise."""
        if max_evals is None:
            if self.params.record_iteration_stats:
                freq = max(1, int(self.params.termination_evaluation_frequency))
                max_evals = min(1 << 22, max(0, int(self.params.iteration_limit)) // freq + 16)
            else:
                max_evals = 1
        evals = (FolpEval * max_evals)()
        n_ev = C.c_int64()
        reason = C.c_int32()
        iters = C.c_int32()
        x = np.zeros(self.n)
        y = np.zeros(self.m)
        self._check(lib().folp_solve(self._h, evals, max_evals, C.byref(n_ev), C.byref(reason),
                                     C.byref(iters), _p(x), _p(y)))
        if n_ev.value > max_evals:  # folp_solve reports how many records there were, not how many fit
            import warnings
            warnings.warn(f"folp_solve produced {n_ev.value} records, only the first {max_evals - 1} and the "
                          "last one were kept")
        return x, y, reason.value, iters.value, [evals[i] for i in range(min(n_ev.value, max_evals))]

    def get_solution(self, which=0, unscaled=True):
        x = np.zeros(self.n)
        y = np.zeros(self.m)
        self._check(lib().folp_get_solution(self._h, which, int(unscaled), _p(x), _p(y)))
        return x, y

    def debug_attempts(self, k: int):
        self._check(lib().folp_debug_attempts(self._h, k))

    def debug_state(self):
        x = np.zeros(self.n); y = np.zeros(self.m); aty = np.zeros(self.n)
        sx = np.zeros(self.n); sy = np.zeros(self.m)
        s = FolpDebugScalars()
        self._check(lib().folp_debug_state(self._h, _p(x), _p(y), _p(aty), _p(sx), _p(sy), C.byref(s)))
        return {"x": x, "y": y, "dual_product": aty, "sum_x": sx, "sum_y": sy, **s.as_dict()}

    def debug_set_state(self, x, y, step_size=-1.0, primal_weight=-1.0):
        x = None if x is None else _d(x)
        y = None if y is None else _d(y)
        self._check(lib().folp_debug_set_state(self._h, _p(x), _p(y), step_size, primal_weight))

    def spmv(self, v, transpose=False):
        v = _d(v)
        out = np.zeros(self.n if transpose else self.m)
        self._check(lib().folp_debug_spmv(self._h, int(transpose), _p(v), _p(out)))
        return out

    def profile_attempts(self, attempts: int):
        """Device ms spent in {primal, A*xbar+dual, A'*y+rule} over `attempts` attempts."""
        ms = (C.c_double * 8)()
        ran = C.c_int64()
        self._check(lib().folp_debug_profile_attempts(self._h, attempts, ms, C.byref(ran)))
        return list(ms), ran.value

    def time_spmv(self, transpose=False, reps=20) -> float:
        """Average device milliseconds of the plain SpMV kernel."""
        ms = C.c_double()
        self._check(lib().folp_debug_time_spmv(self._h, int(transpose), int(reps), C.cast(C.byref(ms), _pd)))
        return ms.value

    def shard_info(self):
        v = [C.c_int64() for _ in range(5)]
        self._check(lib().folp_shard_info(self._h, *[C.byref(x) for x in v]))
        d = dict(zip(("row_begin", "row_end", "col_begin", "col_end", "local_nonzeros"),
                     (x.value for x in v)))
        d["exchange"] = ("none", "nccl", "peer", "multicast")[lib().folp_exchange_mode(self._h)]
        return d

    def stream(self) -> int:
        return int(lib().folp_debug_stream(self._h) or 0)

    def counters(self):
        k = C.c_int64(); t = C.c_double(); it = C.c_int64()
        self._check(lib().folp_counters(self._h, C.byref(k), C.cast(C.byref(t), _pd), C.byref(it)))
        return {"kernel_launches": k.value, "basic_algorithm_seconds": t.value, "iterations": it.value}
