"""Pins the CPU oracle's trust-region solver and localized-duality-gap bound
against the reference's known answers (test/test_trust_region_utils.jl:17-326).
`==` where the reference asserts `==`, atol 1e-8 where it writes `≈ ... atol`."""
import math

import numpy as np
import pytest

from oracle import oracle
from shared_problems import example_cc_star_lp, example_lp

INF = np.inf
MAX_NORM, EUCLIDEAN_NORM = 0, 1


@pytest.mark.parametrize("approx", [True, False])
def test_trust_region_unbounded(approx):  # :17-60
    sol, val = oracle.trust_region([0.0], [-1.0], [-INF], [INF], [1.0], 5.0, approx)
    assert val == -5.0
    assert np.array_equal(sol, [5.0])
    sol, val = oracle.trust_region([0.0, 0.0], [1.0, 1.0], [-INF, -INF], [INF, INF], [2.0, 1.0],
                                   math.sqrt(6.0), approx)
    assert np.allclose(sol, [-1.0, -2.0], rtol=0, atol=1e-8)
    assert abs(val - (-3.0)) <= 1e-8


def test_trust_region_bounded():  # :63-181
    sol, _ = oracle.trust_region([0.0], [-1.0], [-INF], [INF], [1.0], 5.0)
    assert np.array_equal(sol, [5.0])
    sol, _ = oracle.trust_region([0.0], [-1.0], [-INF], [0.0], [1.0], 5.0)
    assert np.array_equal(sol, [0.0])
    sol, _ = oracle.trust_region([0.0], [-1.0], [-INF], [2.0], [1.0], 5.0)
    assert np.array_equal(sol, [2.0])
    sol, _ = oracle.trust_region([0.0, 0.0], [-2.0, -1.0], [-INF, -INF], [3.0, INF], [1.0, 1.0], 5.0)
    assert np.allclose(sol, [3.0, 4.0], rtol=0, atol=1e-8)
    sol, _ = oracle.trust_region([0.0, 0.0], [-1.0, 0.0], [-INF, -INF], [2.0, INF], [1.0, 1.0], 5.0)
    assert np.array_equal(sol, [2.0, 0.0])
    w = np.array([4.0 ** 2, 3.0 ** 2])
    sol, _ = oracle.trust_region([0.0, 0.0], [-4.0, -3.0], [-INF, -INF], [INF, INF], w, math.sqrt(2.0))
    assert abs(math.sqrt(float(np.sum(w * sol * sol))) - math.sqrt(2.0)) <= 1e-8
    assert np.allclose(sol, [1.0 / 4.0, 1.0 / 3.0], rtol=0, atol=1e-8)


@pytest.mark.parametrize("m", [10.0, 50.0])
def test_trust_region_hundred_dimensional(m):  # :183-209
    n = 100
    ub = 1.0 * np.arange(1, n + 1)
    expect = np.minimum(ub, m)
    radius = math.sqrt(float(np.sum(expect ** 2)))
    sol, _ = oracle.trust_region(np.zeros(n), -np.ones(n), np.zeros(n), ub, np.ones(n), radius)
    assert np.allclose(sol, expect, rtol=0, atol=1e-8)


def _corrected_dual_obj(lp, x, y):
    return oracle.iteration_stats(lp, x, y, np.zeros_like(x), np.zeros_like(y), 1.0, 1.0).corrected_dual_objective


def test_bound_primal_and_dual_objective():  # :212-326
    lp = example_lp()
    wp, wd = np.ones(4), np.ones(3)
    r = oracle.bound_optimal_objective(lp, [1.0, 0.0, 6.0, 2.0], [0.5, 4.0, 0.0], wp, wd, 10.0, MAX_NORM)
    assert r["lower_bound_value"] == -1.0
    assert r["upper_bound_value"] == -1.0
    r = oracle.bound_optimal_objective(lp, [1.0, 0.0, 5.99999, 2.0], [0.50001, 4.0, 0.0], wp, wd, 10.0,
                                       MAX_NORM)
    assert -1.01 < r["lower_bound_value"] < -1.0
    assert -1.0 < r["upper_bound_value"] < -0.99

    x, y = np.array([1.0, 0.0, 6.0, 1.0]), np.array([0.0, 4.0, 0.0])
    r = oracle.bound_optimal_objective(lp, x, y, wp, wd, 2.0, MAX_NORM)
    assert r["lower_bound_value"] == -4.0
    assert r["upper_bound_value"] == 2.0
    assert r["lower_bound_value"] == _corrected_dual_obj(lp, x, y)

    x, y = np.array([3.0, 0.0, 6.0, 0.0]), np.array([0.0, 4.0, 0.0])
    r = oracle.bound_optimal_objective(lp, x, y, wp, wd, 5.0, EUCLIDEAN_NORM)
    assert r["lower_bound_value"] == -4.0
    assert r["lagrangian_value"] == -1.0
    assert 5.0 ** 2 == np.linalg.norm(r["primal_solution"] - x, 2) ** 2 + \
        np.linalg.norm(r["dual_solution"] - y, 2) ** 2
    assert r["upper_bound_value"] == 7.0

    x, y = np.array([1.0, 1.0, 4.0, 1.0]), np.zeros(3)
    r = oracle.bound_optimal_objective(lp, x, y, wp, wd, 10.0, MAX_NORM)
    assert r["lower_bound_value"] == _corrected_dual_obj(lp, x, y)

    x = np.array([0.5, 0.5, 0.5, 1.0, 1.0, 1.0])  # interior point
    r = oracle.bound_optimal_objective(example_cc_star_lp(), x, np.zeros(3), np.ones(6), np.ones(3), 10.0,
                                       MAX_NORM)
    assert r["lagrangian_value"] == r["upper_bound_value"]
    assert r["lower_bound_value"] < r["lagrangian_value"]
