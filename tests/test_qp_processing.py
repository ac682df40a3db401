"""Host preprocessing known answers of the reference (test/test_qp_processing.jl:
14-803) on (a) the Python host mirror firstorderlp.jl_b200/preprocess.py and
(b) the CPU oracle's C restatement of rescale_problem. The two must also agree
with each other bit for bit: both feed the solvers the same scaled problem."""
import numpy as np
import pytest
import scipy.sparse as sp

import folp_b200
from folp_b200 import QuadraticProgrammingProblem, linear_programming_problem, preprocess
from oracle import oracle

INF = np.inf
FIELDS = ("variable_lower_bound", "variable_upper_bound", "objective_vector", "right_hand_side")


def _lp(l, u, c, c0, A, b, neq):
    return linear_programming_problem(l, u, c, c0, np.array(A, dtype=float), b, neq)


def _dense(M):
    return np.asarray(M.todense())


def _assert_problem(p, q, approx):
    """test/utilities.jl:22-85 test_fields_equal / test_fields_approx_equal."""
    cmp = (lambda a, b: np.allclose(a, b, rtol=1e-8, atol=0, equal_nan=True) if approx
           else np.array_equal(a, b))
    for f in FIELDS:
        a, b = np.asarray(getattr(p, f)), np.asarray(getattr(q, f))
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin), f
        assert np.array_equal(a[~fin], b[~fin]), f
        assert cmp(a[fin], b[fin]), (f, a, b)
    assert cmp(p.objective_constant, q.objective_constant)
    assert p.num_equalities == q.num_equalities
    assert p.constraint_matrix.shape == q.constraint_matrix.shape
    assert cmp(_dense(p.constraint_matrix), _dense(q.constraint_matrix))
    assert cmp(_dense(p.objective_matrix), _dense(q.objective_matrix))


def test_l2_norm():  # :15-19
    M = sp.csc_matrix(np.array([[3.0, 0.0, -4.0], [4.0, 3.0, 0.0]]))
    for fn in (preprocess.l2_norm, oracle.l2_norm):
        assert np.allclose(fn(M, 1), [5.0, 3.0, 4.0], rtol=0, atol=1e-10)
        assert np.allclose(fn(M, 2), [5.0, 5.0], rtol=0, atol=1e-10)


def test_remove_empty_rows():  # :21-112
    p = _lp([0.0, 0.0], [1.0, 2.0], [1.0, 2.0], 0.0, [[2.0, 0.0], [1.0, 0.0], [0.0, 0.0]], [1.0, 1.0, 0.0], 1)
    preprocess.remove_empty_rows(p)
    _assert_problem(p, _lp([0.0, 0.0], [1.0, 2.0], [1.0, 2.0], 0.0, [[2.0, 0.0], [1.0, 0.0]], [1.0, 1.0], 1),
                    approx=False)
    p = _lp([0.0, 0.0], [1.0, 2.0], [1.0, 2.0], 0.0, [[0.0, 0.0], [1.0, 0.0], [1.0, 0.0]], [0.0, 1.0, 0.0], 1)
    preprocess.remove_empty_rows(p)
    _assert_problem(p, _lp([0.0, 0.0], [1.0, 2.0], [1.0, 2.0], 0.0, [[1.0, 0.0], [1.0, 0.0]], [1.0, 0.0], 0),
                    approx=False)
    p = _lp([0.0, 0.0], [1.0, 2.0], [1.0, 2.0], 0.0, [[1.0, 0.0], [1.0, 0.0], [0.0, 0.0]], [1.0, 1.0, 1.0], 1)
    with pytest.raises(ValueError):
        preprocess.remove_empty_rows(p)
    p = _lp([0.0, 0.0], [1.0, 2.0], [1.0, 2.0], 0.0, [[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]], [1.0, 1.0, 1.0], 1)
    with pytest.raises(ValueError):
        preprocess.remove_empty_rows(p)


@pytest.mark.parametrize("c0, expect_constant", [(3.0, -3.0), (-3.0, -6.0)])
def test_remove_empty_columns(c0, expect_constant):  # :114-160
    p = _lp([-1.0, -1.0], [2.0, 2.0], [c0, 2.0], 0.0, [[0.0, 1.0], [0.0, -1.0]], [1.0, 1.0], 0)
    preprocess.remove_empty_columns(p)
    _assert_problem(p, _lp([-1.0], [2.0], [2.0], expect_constant, [[1.0], [-1.0]], [1.0, 1.0], 0), approx=False)


def test_recover_original_solution_and_presolve():  # :162-209
    out = preprocess.recover_original_solution(np.array([1.0, 1.0, 1.0, 5.0]), [0, 3], 5)
    assert np.array_equal(out, [0.0, 1.0, 1.0, 0.0, 1.0])
    p = _lp([0.0, 0.0, 1.0], [1.0, 2.0, 2.0], [1.0, 2.0, 0.0], 0.0,
            [[1.0, 1.0, 0.0], [1.0, -1.0, 0.0], [0.0, 0.0, 0.0]], [1.0, 1.0, 0.0], 1)
    info = preprocess.presolve(p, verbosity=0)
    x, y = preprocess.undo_presolve(info, np.array([1.0, 0.0]), np.array([1.0, 1.0]))
    assert np.array_equal(x, [1.0, 0.0, 1.0])
    assert np.array_equal(y, [1.0, 1.0, 0.0])
    qp = QuadraticProgrammingProblem(
        [0.0, 0.0, 0.0], [1.0, 2.0, 1.0], np.array([[4.0, 2.0, 0.0], [2.0, 1.0, 0.0], [0.0, 0.0, 1.0]]),
        [1.0, 2.0, 1.0], 0.0, np.array([[1.0, 1.0, 0.0], [1.0, -1.0, 0.0], [1.0, 0.0, 0.0]]), [1.0, 1.0, 2.0], 1)
    preprocess.presolve(qp, verbosity=0)
    assert qp.constraint_matrix.shape == (3, 3)


# further rescale_problem implementations that must agree bit for bit with the two above:
# tests/test_gpu_rescale.py runs every test of this file with folp_rescale_problem in the list
EXTRA_RESCALERS = []


def _both(problem, ruiz, l2, alpha, ruiz_p=np.inf):
    """(scaled problem, con, var) from the host mirror and from the oracle; asserts they agree
    bit for bit."""
    p = problem.copy()
    con, var = np.ones(p.num_constraints), np.ones(p.num_variables)
    if ruiz:
        c, v = preprocess.ruiz_rescaling(p, ruiz, ruiz_p)
        con, var = con * c, var * v
    if l2:
        c, v = preprocess.l2_norm_rescaling(p)
        con, var = con * c, var * v
    if alpha is not None:
        c, v = preprocess.pock_chambolle_rescaling(p, alpha)
        con, var = con * c, var * v
    o = oracle.rescale_problem(ruiz, l2, alpha, problem, ruiz_p=0 if ruiz_p == np.inf else 2)
    assert np.array_equal(o.constraint_rescaling, con)
    assert np.array_equal(o.variable_rescaling, var)
    _assert_problem(o.scaled_qp, p, approx=False)
    for fn in EXTRA_RESCALERS:
        g = fn(ruiz, l2, alpha, problem, ruiz_p=0 if ruiz_p == np.inf else 2)
        assert np.array_equal(g.constraint_rescaling, con)
        assert np.array_equal(g.variable_rescaling, var)
        _assert_problem(g.scaled_qp, p, approx=False)
    return p, con, var


def test_l2_norm_rescaling():  # :234-337
    q = 0.25
    p, _, _ = _both(_lp([0.0, 0.0], [1.0, 2.0], [1.0, 2.0], 0.0, [[1.0, 1.0], [1.0, -1.0], [1.0, 0.0]],
                        [1.0, 1.0, 2.0], 1), 0, True, None)
    _assert_problem(p, _lp([0.0, 0.0], [3 ** q, 2.0 * 2 ** q], [1.0 / 3 ** q, 2.0 / 2 ** q], 0.0,
                           [[6 ** -q, 4 ** -q], [6 ** -q, -(4 ** -q)], [3 ** -q, 0.0]],
                           [2 ** -q, 2 ** -q, 2.0], 1), approx=True)
    p, _, _ = _both(_lp([0.0, 0.0], [1.0, 2.0], [1.0, 2.0], 0.0, [[1.0, 1.0], [1.0, -1.0], [0.0, 0.0]],
                        [1.0, 1.0, 0.0], 1), 0, True, None)
    _assert_problem(p, _lp([0.0, 0.0], [2 ** q, 2.0 * 2 ** q], [1.0 / 2 ** q, 2.0 / 2 ** q], 0.0,
                           [[4 ** -q, 4 ** -q], [4 ** -q, -(4 ** -q)], [0.0, 0.0]],
                           [2 ** -q, 2 ** -q, 0.0], 1), approx=True)
    p, _, _ = _both(_lp([0.0, 0.0], [1.0, 2.0], [1.0, 2.0], 0.0, [[1.0, 0.0], [1.0, 0.0], [2.0, 0.0]],
                        [1.0, 1.0, 2.0], 1), 0, True, None)
    _assert_problem(p, _lp([0.0, 0.0], [6 ** q, 2.0], [1.0 / 6 ** q, 2.0], 0.0,
                           [[6 ** -q, 0.0], [6 ** -q, 0.0], [24 ** -q * 2.0, 0.0]],
                           [1.0, 1.0, 2.0 / np.sqrt(2)], 1), approx=True)


@pytest.mark.parametrize("alpha, con_sq, var_sq", [(0.0, [2, 2, 2], [6, 2]), (1.0, [2, 3, 1], [4, 2]),
                                                   (2.0, [2, 5, 1], [3, 3])])
def test_pock_chambolle_rescaling(alpha, con_sq, var_sq):  # :339-397
    lp = _lp([-1.0, -1.0], [1.0, 2.0], [1.0, 2.0], 0.0, [[1.0, 1.0], [2.0, -1.0], [1.0, 0.0]], [1.0, 1.0, 2.0], 1)
    _, con, var = _both(lp, 0, False, alpha)
    assert np.allclose(con, np.sqrt(con_sq), rtol=1e-8, atol=0)
    assert np.allclose(var, np.sqrt(var_sq), rtol=1e-8, atol=0)


def test_ruiz_rescaling_lp():  # :399-474
    s2, s3 = np.sqrt(2), np.sqrt(3)
    orig = _lp([0.0, 0.0], [1.0, 2.0], [1.0, 2.0], 0.0, [[1.0, 3.0], [1.0, -2.0], [2.0, 0.0]], [1.0, 1.0, 2.0], 1)
    p, con, var = _both(orig, 1, False, None)
    _assert_problem(p, _lp([0.0, 0.0], [s2, 2.0 * s3], [1.0 / s2, 2.0 / s3], 0.0,
                           [[1 / np.sqrt(6), 1.0], [0.5, -s2 / s3], [1.0, 0.0]], [1 / s3, 1 / s2, s2], 1),
                    approx=True)
    assert np.allclose(var, [s2, s3], rtol=1e-8, atol=0)
    assert np.allclose(con, [s3, s2, s2], rtol=1e-8, atol=0)
    preprocess.unscale_problem(p, con, var)
    _assert_problem(p, orig, approx=True)

    orig = _lp([-1.0, -1.0], [1.0, 2.0], [1.0, 2.0], 0.0, [[2.0, 0.0], [0.0, 0.0]], [1.0, 1.0], 1)
    p, con, var = _both(orig, 1, False, None)
    _assert_problem(p, _lp([-s2, -1.0], [s2, 2.0], [1 / s2, 2.0], 0.0, [[1.0, 0.0], [0.0, 0.0]], [1 / s2, 1.0], 1),
                    approx=True)
    assert np.allclose(var, [s2, 1.0], rtol=1e-8, atol=0)
    assert np.allclose(con, [s2, 1.0], rtol=1e-8, atol=0)
    preprocess.unscale_problem(p, con, var)
    _assert_problem(p, orig, approx=True)


def test_ruiz_rescaling_convergence_and_round_trip():  # :476-547
    orig = _lp([0.0, 0.0], [1.0, 2.0], [1.0, 2.0], 0.0, [[1.0, 3.0], [1.0, -2.0], [2.0, 0.0]], [1.0, 1.0, 3.0], 1)
    p, con, var = _both(orig, 30, False, None)
    A = np.abs(_dense(p.constraint_matrix))
    assert np.allclose(np.sqrt(A.max(axis=0)), 1.0, rtol=1e-8)
    assert np.allclose(np.sqrt(A.max(axis=1)), 1.0, rtol=1e-8)
    preprocess.unscale_problem(p, con, var)
    _assert_problem(p, orig, approx=True)
    for scaled in (preprocess.rescale_problem(10, True, None, 0, orig), oracle.rescale_problem(10, True, None, orig)):
        preprocess.unscale_problem(scaled.scaled_qp, scaled.constraint_rescaling, scaled.variable_rescaling)
        _assert_problem(scaled.scaled_qp, scaled.original_qp, approx=True)


def _qp(l=(-INF, -2.0)):
    return QuadraticProgrammingProblem(
        list(l), [1.0, 2.0], np.array([[4.0, 2.0], [2.0, 1.0]]), [1.0, 2.0], 0.0,
        np.array([[1.0, 3.0], [1.0, -2.0], [2.0, 0.0]]), [1.0, 1.0, 2.0], 1)


def test_ruiz_rescaling_qp():  # :549-640
    s2, s3 = np.sqrt(2), np.sqrt(3)
    orig = _qp()
    p, con, var = _both(orig, 1, False, None)
    expect = QuadraticProgrammingProblem(
        [-INF, -2.0 * s3], [2.0, 2.0 * s3], np.array([[1.0, 1.0 / s3], [1.0 / s3, 1 / 3]]), [0.5, 2.0 / s3], 0.0,
        np.array([[0.5 / s3, 1.0], [0.5 / s2, -s2 / s3], [1.0 / s2, 0.0]]), [1 / s3, 1 / s2, s2], 1)
    _assert_problem(p, expect, approx=True)
    assert np.allclose(var, [2.0, s3], rtol=1e-8, atol=0)
    assert np.allclose(con, [s3, s2, s2], rtol=1e-8, atol=0)
    preprocess.unscale_problem(p, con, var)
    _assert_problem(p, orig, approx=True)

    orig = _qp(l=(-1.0, -2.0))
    p, con, var = _both(orig, 30, False, None)
    A, Q = np.abs(_dense(p.constraint_matrix)), np.abs(_dense(p.objective_matrix))
    assert np.allclose(np.sqrt(np.maximum(A.max(axis=0), Q.max(axis=0))), 1.0, rtol=1e-8)
    assert np.allclose(np.sqrt(A.max(axis=1)), 1.0, rtol=1e-8)
    preprocess.unscale_problem(p, con, var)
    _assert_problem(p, orig, approx=True)


def test_l2_ruiz_rescaling():  # :642-803
    q = 0.25
    orig = _lp([0.0, 0.0], [1.0, 2.0], [1.0, 2.0], 0.0, [[1.0, 3.0], [1.0, -2.0], [2.0, 0.0]], [1.0, 1.0, 3.0], 1)
    p, con, var = _both(orig, 1, False, None, ruiz_p=2)
    _assert_problem(p, _lp([0.0, 0.0], [6 ** q, 2 * 13 ** q], [1 / 6 ** q, 2 / 13 ** q], 0.0,
                           [[1 / (6 * 15) ** q, 3 / (13 * 15) ** q], [1 / (7.5 * 6) ** q, -2 / (13 * 7.5) ** q],
                            [2 / (6 * 6) ** q, 0]], [1 / 15 ** q, 1 / 7.5 ** q, 3 / 6 ** q], 1), approx=True)
    assert np.allclose(var, [6 ** q, 13 ** q], rtol=1e-8, atol=0)
    assert np.allclose(con, [15 ** q, 7.5 ** q, 6 ** q], rtol=1e-8, atol=0)

    p, _, _ = _both(orig, 60, False, None, ruiz_p=2)
    assert np.allclose(preprocess.l2_norm(p.constraint_matrix, 1), [1, 1], rtol=0, atol=1e-5)
    assert np.allclose(preprocess.l2_norm(p.constraint_matrix, 2), [np.sqrt(2 / 3)] * 3, rtol=0, atol=1e-5)

    p, con, var = _both(_qp(), 1, False, None, ruiz_p=2)
    expect = QuadraticProgrammingProblem(
        [-INF, -2 * 18 ** q], [26 ** q, 2 * 18 ** q],
        np.array([[4 / 26 ** 0.5, 2 / (26 * 18) ** q], [2 / (26 * 18) ** q, 1 / 18 ** 0.5]]),
        [1 / 26 ** q, 2 / 18 ** q], 0.0,
        np.array([[1 / (25 * 26) ** q, 3 / (18 * 25) ** q], [1 / (12.5 * 26) ** q, -2 / (18 * 12.5) ** q],
                  [2 / (10 * 26) ** q, 0]]), [1 / 25 ** q, 1 / 12.5 ** q, 2 / 10 ** q], 1)
    _assert_problem(p, expect, approx=True)
    assert np.allclose(var, [26 ** q, 18 ** q], rtol=1e-8, atol=0)
    assert np.allclose(con, [25 ** q, 12.5 ** q, 10 ** q], rtol=1e-8, atol=0)

    p, _, _ = _both(_qp(l=(-1.0, -2.0)), 100, False, None, ruiz_p=2)
    cols = np.sqrt(np.sqrt(preprocess.l2_norm(p.constraint_matrix, 1) ** 2 +
                           preprocess.l2_norm(p.objective_matrix, 1) ** 2))
    assert np.allclose(cols, [1, 1], rtol=0, atol=1e-5)
    assert np.allclose(preprocess.l2_norm(p.constraint_matrix, 2), [np.sqrt(2 / 5)] * 3, rtol=0, atol=1e-5)

    p, _, _ = _both(_lp([0.0, 0.0], [1.0, 2.0], [1.0, 2.0], 0.0, [[1.0, 1.0], [1.0, -1.0], [1.0, 1.0]],
                        [1.0, 1.0, 3.0], 1), 10, False, None, ruiz_p=2)
    s = 1 / np.sqrt(3)
    _assert_problem(p, _lp([0.0, 0.0], [3 ** q, 2 * 3 ** q], [1 / 3 ** q, 2 / 3 ** q], 0.0,
                           [[s, s], [s, -s], [s, s]], [1 / 3 ** q, 1 / 3 ** q, 3 / 3 ** q], 1), approx=True)
