"""On-disk formats either side of the PDHG path.

* `qps_reader_to_standard_form`, `transform_to_standard_form`, `two_sided_rows_to_slacks`,
  `TwoSidedQpProblem` mirror src/quadratic_programming_io.jl:15-197. The reference delegates the
  parsing itself to the QPSReader package (pinned 0.2.1 in Manifest.toml, absent from
  /root/reference); `read_mps` restates the MPS / QPS format as that reader documents it: sections
  NAME, OBJSENSE, ROWS, COLUMNS, RHS, RANGES, BOUNDS, QUADOBJ / QMATRIX, ENDATA; free format
  (whitespace-separated fields) or fixed format (fields at columns 2-3, 5-12, 15-22, 25-36, 40-47,
  50-61); the first N row is the objective, further N rows are dropped; an RHS entry on the
  objective row is minus the objective constant; variables default to [0, +Inf); only the lower
  triangle (or either triangle, once) of the objective matrix is listed. Pinned on the reference's
  two fixtures (test/test_qp_io.jl:15-64); quirks of Netlib's fixed-format files beyond that are
  unpinned, as SURVEY 8c notes.
* `solve_log_to_json`, `write_solve_log_json` write a SolveLog the way scripts/solve_qp.jl:115-137
  does with `JSON3.write(log, allow_inf = true)`: fields in declaration order, enums as their
  names, Infinity / -Infinity / NaN literals, `<instance>_summary.json` without and
  `<instance>_full_log.json.gz` with the iteration history -- the files
  benchmarking/process_json_to_csv.jl:54-110 consumes.
"""
from __future__ import annotations

import dataclasses
import enum
import gzip
import json
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import scipy.sparse as sp

from .preprocess import row_permute_in_place
from .problem import QuadraticProgrammingProblem
from .solve_log import IterationStats, SolveLog


# ---------------------------------------------------------------------------
# standard form, quadratic_programming_io.jl:15-131
# ---------------------------------------------------------------------------
@dataclass
class TwoSidedQpProblem:
    """quadratic_programming_io.jl:15-32"""
    variable_lower_bound: np.ndarray
    variable_upper_bound: np.ndarray
    constraint_lower_bound: np.ndarray
    constraint_upper_bound: np.ndarray
    constraint_matrix: sp.csc_matrix
    objective_offset: float
    objective_vector: np.ndarray
    objective_matrix: sp.csc_matrix

    def __post_init__(self):
        f = lambda v: np.array(v, dtype=np.float64)  # noqa: E731
        self.variable_lower_bound = f(self.variable_lower_bound)
        self.variable_upper_bound = f(self.variable_upper_bound)
        self.constraint_lower_bound = f(self.constraint_lower_bound)
        self.constraint_upper_bound = f(self.constraint_upper_bound)
        self.objective_vector = f(self.objective_vector)
        self.constraint_matrix = sp.csc_matrix(self.constraint_matrix, dtype=np.float64)
        self.objective_matrix = sp.csc_matrix(self.objective_matrix, dtype=np.float64)


def two_sided_rows_to_slacks(qp: TwoSidedQpProblem) -> None:
    """quadratic_programming_io.jl:92-131: l <= a'x <= u becomes a'x - s = 0, l <= s <= u (in place)."""
    lo, up = qp.constraint_lower_bound, qp.constraint_upper_bound
    rows = np.flatnonzero(np.isfinite(lo) & np.isfinite(up) & (lo != up))
    if rows.size == 0:
        return
    m = lo.size
    slack = sp.csc_matrix((np.full(rows.size, -1.0), (rows, np.arange(rows.size))), shape=(m, rows.size))
    qp.variable_lower_bound = np.concatenate([qp.variable_lower_bound, lo[rows]])
    qp.variable_upper_bound = np.concatenate([qp.variable_upper_bound, up[rows]])
    qp.objective_vector = np.concatenate([qp.objective_vector, np.zeros(rows.size)])
    qp.constraint_matrix = sp.hstack([qp.constraint_matrix, slack], format="csc")
    lo[rows] = 0.0
    up[rows] = 0.0
    n = qp.variable_lower_bound.size
    Q = qp.objective_matrix.tocoo()
    qp.objective_matrix = sp.csc_matrix((Q.data, (Q.row, Q.col)), shape=(n, n))


def transform_to_standard_form(qp: TwoSidedQpProblem) -> QuadraticProgrammingProblem:
    """quadratic_programming_io.jl:43-90: equalities first, `>=` rows as they are, `<=` rows negated."""
    two_sided_rows_to_slacks(qp)
    lo, up = qp.constraint_lower_bound, qp.constraint_upper_bound
    is_eq = lo == up
    is_geq = ~is_eq & np.isfinite(lo)
    is_leq = ~is_eq & np.isfinite(up)
    assert not np.any(is_geq & is_leq)
    num_equalities = int(is_eq.sum())
    if num_equalities + int(is_geq.sum()) + int(is_leq.sum()) != lo.size:
        raise ValueError("Not all constraints have finite bounds on at least one side.")
    A = sp.csc_matrix(qp.constraint_matrix, dtype=np.float64)
    A.sort_indices()
    flip = is_leq[A.indices]
    A.data[flip] *= -1
    new_row_to_old = np.concatenate([np.flatnonzero(is_eq), np.flatnonzero(~is_eq)])
    if not np.array_equal(new_row_to_old, np.arange(lo.size)):
        old_row_to_new = np.empty_like(new_row_to_old)       # invperm, :73
        old_row_to_new[new_row_to_old] = np.arange(lo.size)
        row_permute_in_place(A, old_row_to_new)
    rhs = lo.copy()
    rhs[is_leq] = -up[is_leq]
    rhs = rhs[new_row_to_old]
    return QuadraticProgrammingProblem(
        variable_lower_bound=qp.variable_lower_bound,
        variable_upper_bound=qp.variable_upper_bound,
        objective_matrix=qp.objective_matrix,
        objective_vector=qp.objective_vector,
        objective_constant=float(qp.objective_offset),
        constraint_matrix=A,
        right_hand_side=rhs,
        num_equalities=num_equalities,
    )


# ---------------------------------------------------------------------------
# MPS / QPS reader (what QPSReader.readqps returns to quadratic_programming_io.jl:160)
# ---------------------------------------------------------------------------
@dataclass
class QpsData:
    """The fields of QPSReader.QPSData that the reference reads (:163-196)."""
    name: str
    nvar: int
    ncon: int
    objsense: str  # "notset", "min" or "max"
    c0: float
    c: np.ndarray
    qrows: List[int]
    qcols: List[int]
    qvals: List[float]
    arows: List[int]
    acols: List[int]
    avals: List[float]
    lcon: np.ndarray
    ucon: np.ndarray
    lvar: np.ndarray
    uvar: np.ndarray
    varnames: List[str]
    connames: List[str]


_SECTIONS = {"NAME", "OBJSENSE", "OBJSENSE_MAX", "OBJSENSE_MIN", "ROWS", "COLUMNS", "RHS", "RANGES",
             "BOUNDS", "QUADOBJ", "QMATRIX", "ENDATA", "OBJECT", "SOS", "QSECTION", "QCMATRIX"}


def _fixed_fields(line: str) -> List[str]:
    """Fixed-format card: fields at columns 2-3, 5-12, 15-22, 25-36, 40-47, 50-61 (1-based)."""
    cols = [(1, 3), (4, 12), (14, 22), (24, 36), (39, 47), (49, 61)]
    out = [line[a:b].strip() for a, b in cols]
    while out and out[-1] == "":
        out.pop()
    return out


def read_mps(source, fixed_format: bool = False) -> QpsData:
    """Parses an MPS / QPS model from a path (".gz" is gunzipped) or an open text stream."""
    if isinstance(source, (str, os.PathLike)):
        path = os.fspath(source)
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "rt") as f:
            lines = f.read().splitlines()
    else:
        lines = source.read().splitlines()

    name = ""
    objsense = "notset"
    section = None
    row_kind: Dict[str, str] = {}
    row_index: Dict[str, int] = {}
    connames: List[str] = []
    objective_row: Optional[str] = None
    ignored_rows = set()
    var_index: Dict[str, int] = {}
    varnames: List[str] = []
    c_entries: Dict[int, float] = {}
    arows: List[int] = []
    acols: List[int] = []
    avals: List[float] = []
    rhs: Dict[int, float] = {}
    ranges: Dict[int, float] = {}
    c0 = 0.0
    bounds: List[Tuple[str, int, float]] = []
    integer_marker = False
    qrows: List[int] = []
    qcols: List[int] = []
    qvals: List[float] = []
    qmatrix_section = False

    def var_of(nm: str) -> int:
        if nm not in var_index:
            var_index[nm] = len(varnames)
            varnames.append(nm)
        return var_index[nm]

    for raw in lines:
        if not raw.strip() or raw.lstrip().startswith("*"):
            continue
        head = raw.split()
        if not raw[0].isspace() and head[0].upper() in _SECTIONS:  # section header: starts in column 1
            key = head[0].upper()
            if key == "NAME":
                name = head[1] if len(head) > 1 else ""
                section = None
            elif key == "OBJSENSE":
                section = "OBJSENSE"
                if len(head) > 1:
                    objsense = {"MAX": "max", "MAXIMIZE": "max", "MIN": "min", "MINIMIZE": "min"}[head[1].upper()]
            elif key in ("OBJSENSE_MAX", "OBJSENSE_MIN"):
                objsense = "max" if key.endswith("MAX") else "min"
                section = None
            elif key == "ENDATA":
                break
            else:
                section = key
                qmatrix_section = key == "QMATRIX"
            continue
        f = _fixed_fields(raw) if fixed_format else head
        if fixed_format and section in ("COLUMNS", "RHS", "RANGES", "QUADOBJ", "QMATRIX", "QSECTION"):
            f = f[1:]  # field 1 (columns 2-3) is blank on these cards
        if section == "OBJSENSE":
            objsense = {"MAX": "max", "MAXIMIZE": "max", "MIN": "min", "MINIMIZE": "min"}[f[0].upper()]
        elif section == "ROWS":
            kind, nm = f[0].upper(), f[1]
            if kind == "N":
                if objective_row is None:
                    objective_row = nm
                else:
                    ignored_rows.add(nm)  # only the first free row is the objective
            else:
                row_kind[nm] = kind
                row_index[nm] = len(connames)
                connames.append(nm)
        elif section == "COLUMNS":
            if len(f) >= 3 and f[1].upper() == "'MARKER'":
                integer_marker = "INTORG" in f[2].upper()
                continue
            j = var_of(f[0])
            for k in range(1, len(f) - 1, 2):
                nm, v = f[k], float(f[k + 1])
                if nm == objective_row:
                    c_entries[j] = c_entries.get(j, 0.0) + v
                elif nm in ignored_rows:
                    continue
                elif nm in row_index:
                    arows.append(row_index[nm])
                    acols.append(j)
                    avals.append(v)
                else:
                    raise ValueError(f"COLUMNS: unknown row {nm!r}")
            if integer_marker:
                pass  # integrality is dropped: the solver treats the LP relaxation
        elif section in ("RHS", "RANGES"):
            # the set name (first field) is optional in free format: an odd number of fields has it
            start = 1 if len(f) % 2 == 1 else 0
            for k in range(start, len(f) - 1, 2):
                nm, v = f[k], float(f[k + 1])
                if section == "RHS" and nm == objective_row:
                    c0 = -v  # an RHS entry on the objective row is minus the objective constant
                elif nm in ignored_rows:
                    continue
                elif nm in row_index:
                    (rhs if section == "RHS" else ranges)[row_index[nm]] = v
                else:
                    raise ValueError(f"{section}: unknown row {nm!r}")
        elif section == "BOUNDS":
            kind = f[0].upper()
            if kind in ("FR", "MI", "PL", "BV"):
                nm = f[2] if len(f) >= 3 else f[1]
                bounds.append((kind, var_of(nm), 0.0))
            else:
                if len(f) >= 4:
                    nm, v = f[2], float(f[3])
                else:  # no set name
                    nm, v = f[1], float(f[2])
                bounds.append((kind, var_of(nm), v))
        elif section in ("QUADOBJ", "QMATRIX", "QSECTION"):
            i, j, v = var_of(f[0]), var_of(f[1]), float(f[2])
            if qmatrix_section:
                if i < j:
                    continue  # QMATRIX lists both triangles: keep the lower one
            elif i < j:
                i, j = j, i  # QUADOBJ lists one triangle: store it as the lower one
            qrows.append(i)
            qcols.append(j)
            qvals.append(v)
        elif section in ("SOS", "QCMATRIX", "OBJECT"):
            raise ValueError(f"unsupported MPS section {section}")
        else:
            raise ValueError(f"data line outside a section: {raw!r}")

    nvar, ncon = len(varnames), len(connames)
    c = np.zeros(nvar)
    for j, v in c_entries.items():
        c[j] = v
    lcon = np.full(ncon, -np.inf)
    ucon = np.full(ncon, np.inf)
    for nm, i in row_index.items():
        kind = row_kind[nm]
        b = rhs.get(i, 0.0)
        if kind == "E":
            lcon[i] = ucon[i] = b
        elif kind == "G":
            lcon[i] = b
        elif kind == "L":
            ucon[i] = b
        else:
            raise ValueError(f"ROWS: unknown row type {kind!r}")
    for i, r in ranges.items():  # the standard RANGES table
        kind = row_kind[connames[i]]
        if kind == "G":
            ucon[i] = lcon[i] + abs(r)
        elif kind == "L":
            lcon[i] = ucon[i] - abs(r)
        elif r >= 0:
            ucon[i] = lcon[i] + r
        else:
            lcon[i] = ucon[i] + r
    lvar = np.zeros(nvar)
    uvar = np.full(nvar, np.inf)
    for kind, j, v in bounds:
        if kind == "LO":
            lvar[j] = v
        elif kind == "UP":
            uvar[j] = v
            if v < 0 and lvar[j] == 0.0:
                lvar[j] = -np.inf  # the classic convention for a negative upper bound on a default lower bound
        elif kind == "FX":
            lvar[j] = uvar[j] = v
        elif kind == "FR":
            lvar[j], uvar[j] = -np.inf, np.inf
        elif kind == "MI":
            lvar[j] = -np.inf
        elif kind == "PL":
            uvar[j] = np.inf
        elif kind == "BV":
            lvar[j], uvar[j] = 0.0, 1.0
        elif kind == "LI":
            lvar[j] = v
        elif kind == "UI":
            uvar[j] = v
        else:
            raise ValueError(f"BOUNDS: unknown bound type {kind!r}")
    return QpsData(name, nvar, ncon, objsense, c0, c, qrows, qcols, qvals, arows, acols, avals, lcon, ucon,
                   lvar, uvar, varnames, connames)


def qps_reader_to_standard_form(filename, fixed_format: bool = False) -> QuadraticProgrammingProblem:
    """quadratic_programming_io.jl:147-197."""
    mps = read_mps(filename, fixed_format=fixed_format)
    A = sp.csc_matrix((mps.avals, (mps.arows, mps.acols)), shape=(mps.ncon, mps.nvar), dtype=np.float64)
    A.sum_duplicates()
    # the reader returns one triangle of the objective matrix: symmetrize (:171-184)
    r, c, v = [], [], []
    for i, j, x in zip(mps.qrows, mps.qcols, mps.qvals):
        r.append(i); c.append(j); v.append(x)
        if i != j:
            r.append(j); c.append(i); v.append(x)
    Q = sp.csc_matrix((v, (r, c)), shape=(mps.nvar, mps.nvar), dtype=np.float64)
    Q.sum_duplicates()
    assert mps.objsense == "notset"  # :185
    return transform_to_standard_form(TwoSidedQpProblem(mps.lvar, mps.uvar, mps.lcon, mps.ucon, A, mps.c0,
                                                        mps.c, Q))


# ---------------------------------------------------------------------------
# SolveLog JSON, scripts/solve_qp.jl:115-137 + solve_log.jl:423-426
# ---------------------------------------------------------------------------
def _jsonable(x):
    if dataclasses.is_dataclass(x) and not isinstance(x, type):
        return {f.name: _jsonable(getattr(x, f.name)) for f in dataclasses.fields(x)}  # declaration order
    if isinstance(x, enum.Enum):
        return x.name  # JSON3 writes an @enum as its name
    if isinstance(x, dict):
        return {str(k): _jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    if isinstance(x, np.ndarray):
        return [_jsonable(v) for v in x.tolist()]
    if isinstance(x, (np.floating,)):
        return float(x)
    if isinstance(x, (np.integer,)):
        return int(x)
    return x


def solve_log_to_json(log: SolveLog, include_iteration_stats: bool = True) -> str:
    """`JSON3.write(log, allow_inf = true)`: Infinity / -Infinity / NaN are written as such."""
    d = _jsonable(log)
    if not include_iteration_stats:
        d["iteration_stats"] = []
    return json.dumps(d, allow_nan=True, separators=(",", ":"))


def solve_log_from_output(instance_name: str, output, solve_time_sec: float,
                          command_line_invocation: str = "") -> SolveLog:
    """scripts/solve_qp.jl:115-128 (iteration_stats stay empty, as in the summary file)."""
    from ._abi import PointType
    log = SolveLog()
    log.instance_name = instance_name
    log.command_line_invocation = command_line_invocation
    log.termination_reason = output.termination_reason
    log.termination_string = output.termination_string
    log.iteration_count = int(output.iteration_count)
    log.solve_time_sec = float(solve_time_sec)
    log.solution_stats = output.iteration_stats[-1] if output.iteration_stats else IterationStats()
    log.solution_type = PointType.POINT_TYPE_AVERAGE_ITERATE
    return log


def write_solve_log_json(output_dir: str, instance_name: str, output, solve_time_sec: float,
                         command_line_invocation: str = "") -> Tuple[str, str]:
    """Writes `<instance>_summary.json` and `<instance>_full_log.json.gz` (solve_qp.jl:130-141);
    returns the two paths."""
    log = solve_log_from_output(instance_name, output, solve_time_sec, command_line_invocation)
    summary = os.path.join(output_dir, instance_name + "_summary.json")
    with open(summary, "w") as f:
        f.write(solve_log_to_json(log))
    log.iteration_stats = list(output.iteration_stats)
    full = os.path.join(output_dir, instance_name + "_full_log.json.gz")
    with gzip.open(full, "wt") as f:
        f.write(solve_log_to_json(log))
    return summary, full
