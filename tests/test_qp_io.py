"""The reference's I/O tests (test/test_qp_io.jl:15-94) on the host mirror io.py: reading the two
fixture models (plain and gzipped), two_sided_rows_to_slacks, plus the standard-form rules of
quadratic_programming_io.jl:43-90 on a fixed-format file exercising every row and bound type, and
the SolveLog JSON layout scripts/solve_qp.jl:115-141 writes for benchmarking/process_json_to_csv.jl."""
import gzip
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

import folp_b200
from folp_b200 import PointType, TerminationReason
from folp_b200 import io as fio
from folp_b200.solve_log import (ConvergenceInformation, InfeasibilityInformation, IterationStats,
                                 SaddlePointOutput)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INF = np.inf


def _assert_qp(qp, l, u, Q, c, c0, A, b, neq):
    assert np.array_equal(qp.variable_lower_bound, l)
    assert np.array_equal(qp.variable_upper_bound, u)
    assert np.array_equal(qp.objective_matrix.toarray(), np.array(Q, dtype=float))
    assert np.array_equal(qp.objective_vector, c)
    assert qp.objective_constant == c0
    assert np.array_equal(qp.constraint_matrix.toarray(), np.array(A, dtype=float))
    assert np.array_equal(qp.right_hand_side, b)
    assert qp.num_equalities == neq


def test_read_mps_lp():  # :37-42
    qp = fio.qps_reader_to_standard_form(os.path.join(GOLDEN, "trivial_lp_model.mps"))
    _assert_qp(qp, [0.0, 1.0], [1.0, 2.0], np.zeros((2, 2)), [2.0, -1.0], 0.0, [[-1.0, -1.0]], [-3.0], 0)
    assert folp_b200.validate(qp)


def test_read_mps_qp():  # :44-49
    qp = fio.qps_reader_to_standard_form(os.path.join(GOLDEN, "trivial_qp_model.mps"))
    _assert_qp(qp, [0.0, 1.0], [1.0, 2.0], [[2.0, 2.0], [2.0, 4.0]], [2.0, -1.0], 0.0, [[-1.0, -1.0]], [-3.0], 0)


REFERENCE_TESTS = "/root/reference/test"  # present in the build container only; never on the GPU box


@pytest.mark.skipif(not os.path.isdir(REFERENCE_TESTS), reason="the reference tree is not here")
@pytest.mark.parametrize("name, Q", [("trivial_lp_model.mps", np.zeros((2, 2))),
                                     ("trivial_qp_model.mps", [[2.0, 2.0], [2.0, 4.0]])])
def test_read_the_reference_s_own_fixtures(name, Q):
    """The goldens above are re-written files; where the reference is at hand its OWN fixtures
    (section order, QUADOBJ triangle, spacing as QPSReader's authors wrote them) must parse to the
    structs of test/test_qp_io.jl:15-35 too, plain and gzipped."""
    path = os.path.join(REFERENCE_TESTS, name)
    qp = fio.qps_reader_to_standard_form(path)
    _assert_qp(qp, [0.0, 1.0], [1.0, 2.0], Q, [2.0, -1.0], 0.0, [[-1.0, -1.0]], [-3.0], 0)
    # and our golden of the same model describes the same problem
    ours = fio.qps_reader_to_standard_form(os.path.join(GOLDEN, name))
    _assert_qp(ours, qp.variable_lower_bound, qp.variable_upper_bound, qp.objective_matrix.toarray(),
               qp.objective_vector, qp.objective_constant, qp.constraint_matrix.toarray(), qp.right_hand_side,
               qp.num_equalities)


def test_read_mps_gz(tmp_path):  # :51-64
    src = open(os.path.join(GOLDEN, "trivial_qp_model.mps")).read()
    path = tmp_path / "model.mps.gz"
    with gzip.open(path, "wt") as f:
        f.write(src)
    qp = fio.qps_reader_to_standard_form(str(path))
    _assert_qp(qp, [0.0, 1.0], [1.0, 2.0], [[2.0, 2.0], [2.0, 4.0]], [2.0, -1.0], 0.0, [[-1.0, -1.0]], [-3.0], 0)


def test_two_sided_rows_to_slacks():  # :66-94
    qp = fio.TwoSidedQpProblem([-INF, -INF], [INF, INF], [-3.0, -2.0], [1.0, INF], [[1.0, 1.0], [1.0, 1.0]], 2.0,
                               [0.0, 1.0], np.diag([1.0, 3.0]))
    fio.two_sided_rows_to_slacks(qp)
    assert np.array_equal(qp.variable_lower_bound, [-INF, -INF, -3.0])
    assert np.array_equal(qp.variable_upper_bound, [INF, INF, 1.0])
    assert np.array_equal(qp.constraint_lower_bound, [0.0, -2.0])
    assert np.array_equal(qp.constraint_upper_bound, [0.0, INF])
    assert np.array_equal(qp.constraint_matrix.toarray(), [[1.0, 1.0, -1.0], [1.0, 1.0, 0.0]])
    assert qp.objective_offset == 2.0
    assert np.array_equal(qp.objective_vector, [0.0, 1.0, 0.0])
    assert np.array_equal(qp.objective_matrix.toarray(), np.diag([1.0, 3.0, 0.0]))


def test_row_permute_in_place():  # test/test_sparse_linalg.jl:15-35 (1-based maps there)
    from folp_b200.preprocess import row_permute_in_place
    mat = sp.csc_matrix(np.array([[1.0, 0.0], [0.0, 1.0]]))
    row_permute_in_place(mat, [1, 0])
    assert np.array_equal(mat.toarray(), [[0.0, 1.0], [1.0, 0.0]])
    mat = sp.csc_matrix(np.array([[1.0, 0.0], [0.0, 1.0], [2.0, 3.0]]))
    row_permute_in_place(mat, [2, 0, 1])
    assert np.array_equal(mat.toarray(), [[0.0, 1.0], [2.0, 3.0], [1.0, 0.0]])
    assert mat.has_sorted_indices and np.all(np.diff(mat.indices[mat.indptr[0]:mat.indptr[1]]) > 0)


def test_fixed_format_every_row_and_bound_type():
    path = os.path.join(GOLDEN, "ranged_fixed_format.mps")
    mps = fio.read_mps(path, fixed_format=True)
    assert mps.varnames == ["X 1", "X 2", "X 3", "X 4", "X 5", "X 6"]
    assert mps.connames == ["EQ 1", "GE 1", "LE 1", "EQ R"]      # the second N row is dropped
    assert mps.c0 == 7.5                                          # RHS on the objective row is -constant
    assert np.array_equal(mps.c, [1.0, -2.0, 0.0, 0.0, 0.5, 0.0])
    # RANGES: G: [b, b + |r|], L: [b - |r|, b], E with r < 0: [b + r, b]
    assert np.array_equal(mps.lcon, [4.0, 1.0, 6.0, -1.0])
    assert np.array_equal(mps.ucon, [4.0, 3.5, 10.0, 2.0])
    assert np.array_equal(mps.lvar, [0.0, -INF, 1.5, -INF, -INF, 0.0])  # UP < 0 on a default lower bound
    assert np.array_equal(mps.uvar, [5.0, INF, 1.5, INF, -1.0, 1.0])
    qp = fio.qps_reader_to_standard_form(path, fixed_format=True)
    # three two-sided rows became equalities with slack columns 7, 8, 9; equalities come first
    assert qp.num_equalities == 4 and qp.constraint_matrix.shape == (4, 9)
    assert np.array_equal(qp.right_hand_side, [4.0, 0.0, 0.0, 0.0])
    assert np.array_equal(qp.variable_lower_bound[6:], [1.0, 6.0, -1.0])
    assert np.array_equal(qp.variable_upper_bound[6:], [3.5, 10.0, 2.0])
    A = qp.constraint_matrix.toarray()
    assert np.array_equal(A[:, 6:], -np.eye(4)[:, 1:])
    assert np.array_equal(A[1, :6], [2.0, 0.0, 1.0, 0.0, 0.0, 0.0])
    assert qp.objective_constant == 7.5
    assert folp_b200.validate(qp)


def test_standard_form_orders_and_negates_rows():
    # rows: <=, =, >=  ->  =, (<= negated), >=   (quadratic_programming_io.jl:60-78)
    qp = fio.transform_to_standard_form(fio.TwoSidedQpProblem(
        [0.0, 0.0], [INF, INF], [-INF, 2.0, 1.0], [5.0, 2.0, INF], [[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]], 0.0,
        [1.0, 1.0], sp.csc_matrix((2, 2))))
    assert qp.num_equalities == 1
    assert np.array_equal(qp.constraint_matrix.toarray(), [[3.0, 4.0], [-1.0, -2.0], [5.0, 6.0]])
    assert np.array_equal(qp.right_hand_side, [2.0, -5.0, 1.0])
    with pytest.raises(ValueError):  # a row with no finite bound, :56-58
        fio.transform_to_standard_form(fio.TwoSidedQpProblem(
            [0.0], [INF], [-INF], [INF], [[1.0]], 0.0, [1.0], sp.csc_matrix((1, 1))))


def test_solve_log_json_layout(tmp_path):
    ci = ConvergenceInformation(candidate_type=PointType.POINT_TYPE_AVERAGE_ITERATE, primal_objective=1.5,
                                corrected_dual_objective=-INF, relative_optimality_gap=float("nan"))
    ii = InfeasibilityInformation(candidate_type=PointType.POINT_TYPE_AVERAGE_ITERATE)
    st = [IterationStats(iteration_number=k, convergence_information=[ci], infeasibility_information=[ii],
                         method_specific_stats={"lagrangian_value": 1.0}) for k in (0, 40)]
    out = SaddlePointOutput(np.zeros(2), np.zeros(1), TerminationReason.TERMINATION_REASON_OPTIMAL, "OPTIMAL", 40, st)
    summary, full = fio.write_solve_log_json(str(tmp_path), "inst", out, 0.25, "solve_qp --method pdhg")
    text = open(summary).read()
    assert "-Infinity" in text and "NaN" in text          # JSON3.write(...; allow_inf = true)
    d = json.loads(text)
    assert list(d) == ["instance_name", "command_line_invocation", "termination_reason", "termination_string",
                       "iteration_count", "solve_time_sec", "solution_stats", "solution_type", "iteration_stats"]
    assert d["termination_reason"] == "TERMINATION_REASON_OPTIMAL" and d["solution_type"] == "POINT_TYPE_AVERAGE_ITERATE"
    assert d["iteration_stats"] == [] and d["solution_stats"]["iteration_number"] == 40
    assert list(d["solution_stats"])[:3] == ["iteration_number", "convergence_information", "infeasibility_information"]
    # what benchmarking/process_json_to_csv.jl:54-110 reads
    assert d["solution_stats"]["convergence_information"][0]["candidate_type"] == d["solution_type"]
    for key in ("primal_objective", "dual_objective", "relative_optimality_gap", "l2_primal_residual",
                "l_inf_dual_variable"):
        assert key in d["solution_stats"]["convergence_information"][0]
    with gzip.open(full, "rt") as f:
        dfull = json.load(f)
    assert [s["iteration_number"] for s in dfull["iteration_stats"]] == [0, 40]
    assert dfull["iteration_stats"][0]["restart_used"] == "RESTART_CHOICE_UNSPECIFIED"
