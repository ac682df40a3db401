"""Pins the CPU oracle's KKT statistics against the reference's exact known
answers (test/test_iteration_stats.jl:16-308): max_primal_violation, primal_obj,
compute_dual_stats and whole IterationStats records at an optimal, a primal
infeasible and a dual infeasible point. The reference asserts `==`; so do we
(`approx` only where the reference writes `≈`)."""
import numpy as np
import pytest

from folp_b200 import QuadraticProgrammingProblem, linear_programming_problem
from oracle import oracle
from shared_problems import example_qp

INF = np.inf

CONVERGENCE_FIELDS = (
    "primal_objective", "dual_objective", "corrected_dual_objective", "l_inf_primal_residual",
    "l2_primal_residual", "l_inf_dual_residual", "l2_dual_residual",
    "relative_l_inf_primal_residual", "relative_l2_primal_residual",
    "relative_l_inf_dual_residual", "relative_l2_dual_residual", "relative_optimality_gap",
    "l_inf_primal_variable", "l2_primal_variable", "l_inf_dual_variable", "l2_dual_variable")
INFEASIBILITY_FIELDS = (
    "max_primal_ray_infeasibility", "primal_ray_linear_objective", "primal_ray_quadratic_norm",
    "max_dual_ray_infeasibility", "dual_ray_objective")


def _check_record(stats, convergence, infeasibility):
    """test/utilities.jl:22-60 test_fields_equal: every field exactly equal; fields the
    reference leaves at the zero-constructor value must be exactly 0."""
    for f in CONVERGENCE_FIELDS:
        assert getattr(stats, f) == convergence.get(f, 0.0), f
    for f in INFEASIBILITY_FIELDS:
        assert getattr(stats, f) == infeasibility.get(f, 0.0), f
    assert stats.candidate_type == 1  # POINT_TYPE_CURRENT_ITERATE


def test_max_primal_violation():  # :16-37
    lp = linear_programming_problem(
        [-1.0, -INF, -INF], [1.0, INF, INF], np.zeros(3), 0.0,
        np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]), [10.0, 11.0], 1)
    assert oracle.max_primal_violation(lp, [0.0, 10.0, 11.0]) == 0.0
    assert oracle.max_primal_violation(lp, [-2.0, 10.0, 11.0]) == pytest.approx(1.0)
    assert oracle.max_primal_violation(lp, [3.0, 10.0, 11.0]) == pytest.approx(2.0)
    assert oracle.max_primal_violation(lp, [0.0, 11.0, 11.0]) == pytest.approx(1.0)
    assert oracle.max_primal_violation(lp, [0.0, 9.0, 11.0]) == pytest.approx(1.0)
    assert oracle.max_primal_violation(lp, [0.0, 11.0, 0.0]) == pytest.approx(11.0)


def test_primal_obj():  # :39-46
    qp = example_qp()
    assert oracle.primal_obj(qp, [0.0, 0.0]) == 0.0
    assert oracle.primal_obj(qp, [1.0, 1.0]) == 0.5
    assert oracle.primal_obj(qp, [1.0, 0.0]) == 1.0
    assert oracle.primal_obj(qp, [0.0, 1.0]) == -0.5
    assert oracle.primal_obj(qp, [0.0, -1.0]) == 1.5


def _small_lp(lower=(-1.0, -INF), upper=(1.0, INF)):
    # min x + 2y  s.t. x + y >= 1, -1 <= x <= 1   (:49-66)
    return linear_programming_problem(list(lower), list(upper), [1.0, 2.0], 0.0,
                                      np.array([[1.0, 1.0]]), [1.0], 0)


def test_dual_stats():  # :48-116
    lp = _small_lp()
    dobj, dres, _ = oracle.dual_stats(lp, [0.0, 0.0], [0.0])
    assert dobj == -1.0
    assert np.max(np.abs(dres)) == 2.0
    assert np.array_equal(dres, [0.0, 0.0, 2.0])

    dobj, dres, _ = oracle.dual_stats(lp, [0.0, 0.0], [1.0])
    assert dobj == 1.0
    assert np.array_equal(dres, [0.0, 0.0, 1.0])

    dobj, dres, _ = oracle.dual_stats(lp, [0.0, 0.0], [2.0])
    assert dobj == 1.0
    assert np.max(np.abs(dres)) == 0.0

    dobj, dres, _ = oracle.dual_stats(lp, [0.0, 0.0], [3.0])
    assert dobj == 1.0
    assert np.max(np.abs(dres)) == 1.0

    dobj, dres, _ = oracle.dual_stats(lp, [0.0, 1.0], [-1.0])
    assert dobj == -3.0
    assert np.array_equal(dres, [1.0, 0.0, 3.0])

    lp = _small_lp(lower=(INF, -INF), upper=(INF, INF))
    dobj, dres, _ = oracle.dual_stats(lp, [0.0, 1.0], [-1.0])
    assert dobj == -1.0
    assert np.array_equal(dres, [1.0, 2.0, 3.0])

    qp = example_qp()
    dobj, dres, _ = oracle.dual_stats(qp, [0.0, 0.0], [3.0])
    assert dobj == -3.0
    assert np.max(np.abs(dres)) == 0.0
    dobj, dres, _ = oracle.dual_stats(qp, [0.0, 0.0], [1.0])
    assert dobj == -1.0
    assert np.max(np.abs(dres)) == 0.0
    dobj, dres, _ = oracle.dual_stats(qp, [0.5, 0.5], [1.0])
    assert dobj == -1.625
    assert np.max(np.abs(dres)) == 0.0


def test_iteration_stats_primal_dual_optimal():  # :118-178
    stats = oracle.iteration_stats(_small_lp(), [1.0, 0.0], [2.0], [0.0, 0.0], [0.0], 1e-6, 1e-6)
    _check_record(stats, {
        "primal_objective": 1.0, "dual_objective": 1.0, "corrected_dual_objective": 1.0,
        "l_inf_primal_variable": 1.0, "l2_primal_variable": 1.0,
        "l_inf_dual_variable": 2.0, "l2_dual_variable": 2.0}, {})


def test_iteration_stats_primal_infeasible():  # :180-246
    lp = linear_programming_problem([0.0], [1.0], [1.0], 2.0, np.array([[1.0]]), [10.0], 1)
    stats = oracle.iteration_stats(lp, [2.0], [1.0], [0.0], [1.0], 1e-6, 1e-6)
    l2 = float(np.linalg.norm([8.0, 1.0], 2))
    _check_record(stats, {
        "primal_objective": 4.0, "dual_objective": 10.0 + 2.0, "corrected_dual_objective": 12.0,
        "l_inf_primal_residual": 8.0, "l2_primal_residual": l2,
        "relative_l_inf_primal_residual": 8.0 / (1.0 + 10.0),
        "relative_l2_primal_residual": l2 / (1.0 + 10.0),
        "relative_optimality_gap": 8.0 / (1.0 + 16.0),
        "l_inf_primal_variable": 2.0, "l2_primal_variable": 2.0,
        "l_inf_dual_variable": 1.0, "l2_dual_variable": 1.0}, {"dual_ray_objective": 9.0})


def test_iteration_stats_dual_infeasible():  # :248-308
    lp = linear_programming_problem([-INF], [INF], [-1.0], 0.0, np.array([[1.0]]), [10.0], 0)
    stats = oracle.iteration_stats(lp, [10.0], [0.0], [1.0], [0.0], 1e-6, 1e-6)
    _check_record(stats, {
        "primal_objective": -10.0, "corrected_dual_objective": -INF,
        "l_inf_dual_residual": 1.0, "l2_dual_residual": 1.0,
        "relative_l_inf_dual_residual": 1.0 / (1.0 + 1.0),
        "relative_l2_dual_residual": 1.0 / (1.0 + 1.0),
        "relative_optimality_gap": 10.0 / (1.0 + 10.0),
        "l_inf_primal_variable": 10.0, "l2_primal_variable": 10.0},
        {"primal_ray_linear_objective": -1.0})
