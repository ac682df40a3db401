"""rescale_problem on the device (folp_rescale_problem) against the
reference's own known-answer tests (test/test_qp_processing.jl:234-803, as ported in
tests/test_qp_processing.py) and against the CPU oracle's restatement of
src/preprocess.jl:358-687 on random, Netlib-shaped, PageRank and quadratic instances.

Parity bar: BIT-IDENTICAL scaled problem and rescaling vectors wherever every row and column has
at most 2048 entries and the Pock-Chambolle exponent is 1 (the CLI default; element-wise
arithmetic and the summation order are the reference's); 1e-13 relative otherwise (block-tree
sums of longer rows; |a|^e for e != 1 goes through pow, which libm and the device round
differently in the last place)."""
import numpy as np
import pytest

import folp_b200
import test_qp_processing as Tq
from folp_b200 import TerminationReason
from folp_b200.lib import rescale_problem as device_rescale
from folp_b200.synthetic import netlib_shaped_lp, pagerank_lp, random_sparse_lp, random_sparse_qp
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _device_in_the_loop(monkeypatch):
    monkeypatch.setattr(Tq, "EXTRA_RESCALERS", [device_rescale])


def test_reference_l2_norm_rescaling():
    Tq.test_l2_norm_rescaling()


@pytest.mark.parametrize("alpha, con_sq, var_sq", [(0.0, [2, 2, 2], [6, 2]), (1.0, [2, 3, 1], [4, 2]),
                                                   (2.0, [2, 5, 1], [3, 3])])
def test_reference_pock_chambolle_rescaling(alpha, con_sq, var_sq):
    Tq.test_pock_chambolle_rescaling(alpha, con_sq, var_sq)


def test_reference_ruiz_rescaling_lp():
    Tq.test_ruiz_rescaling_lp()


def test_reference_ruiz_convergence_and_round_trip():
    Tq.test_ruiz_rescaling_convergence_and_round_trip()


def test_reference_ruiz_rescaling_qp():
    Tq.test_ruiz_rescaling_qp()


def test_reference_l2_ruiz_rescaling():
    Tq.test_l2_ruiz_rescaling()


def _compare(problem, ruiz, l2, alpha, ruiz_p=0, exact=True):
    o = oracle.rescale_problem(ruiz, l2, alpha, problem, ruiz_p=ruiz_p)
    g = device_rescale(ruiz, l2, alpha, problem, ruiz_p=ruiz_p)
    pairs = [(g.constraint_rescaling, o.constraint_rescaling), (g.variable_rescaling, o.variable_rescaling),
             (g.scaled_qp.constraint_matrix.data, o.scaled_qp.constraint_matrix.data),
             (g.scaled_qp.objective_matrix.data, o.scaled_qp.objective_matrix.data),
             (g.scaled_qp.objective_vector, o.scaled_qp.objective_vector),
             (g.scaled_qp.right_hand_side, o.scaled_qp.right_hand_side),
             (g.scaled_qp.variable_lower_bound, o.scaled_qp.variable_lower_bound),
             (g.scaled_qp.variable_upper_bound, o.scaled_qp.variable_upper_bound)]
    for a, b in pairs:
        if exact:
            assert np.array_equal(a, b)
        else:
            fin = np.isfinite(b)
            assert np.array_equal(a[~fin], b[~fin])
            assert np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-300), initial=0.0) <= 1e-13
    # the original problem is untouched
    assert g.original_qp is problem


@pytest.mark.parametrize("ruiz,l2,alpha,ruiz_p", [(10, False, 1.0, 0), (10, True, None, 0), (3, True, 1.0, 0),
                                                  (4, False, None, 2), (0, False, None, 0)])
@pytest.mark.parametrize("make", [lambda: random_sparse_lp(6000, 5000, 8, seed=3, upper_fraction=0.1),
                                  lambda: random_sparse_qp(3000, 2000, 6, seed=4),
                                  lambda: netlib_shaped_lp(seed=5)])
def test_device_rescaling_bit_identical_to_oracle(make, ruiz, l2, alpha, ruiz_p):
    _compare(make(), ruiz, l2, alpha, ruiz_p, exact=True)


def test_device_rescaling_long_row_and_general_exponent():
    _compare(pagerank_lp(5000), 10, False, 1.0, exact=False)          # a dense row of 5000 entries
    _compare(pagerank_lp(5000), 2, True, None, exact=False)
    _compare(random_sparse_lp(3000, 2500, 8, seed=6), 5, False, 0.5, exact=False)  # pow(a, 0.5), pow(a, 1.5)
    _compare(random_sparse_lp(3000, 2500, 8, seed=6), 3, True, 2.0, exact=False)   # a^2 and the count of zeros
    _compare(random_sparse_lp(3000, 2500, 8, seed=6), 0, False, 0.0, exact=False)


def test_device_rescaling_empty_and_degenerate():
    import scipy.sparse as sp
    lp = folp_b200.linear_programming_problem(np.zeros(3), np.full(3, np.inf), np.array([1.0, 2.0, 3.0]), 0.0,
                                              sp.csc_matrix((2, 3)), np.zeros(2), 1)
    _compare(lp, 10, True, 1.0, exact=True)   # no nonzeros: every rescaling factor is 1
    lp = folp_b200.linear_programming_problem(np.zeros(3), np.full(3, np.inf), np.array([1.0, 2.0, 3.0]), 0.0,
                                              sp.csc_matrix((0, 3)), np.zeros(0), 0)
    _compare(lp, 10, False, 1.0, exact=True)  # no constraints


def test_solve_with_device_rescaling_matches_host_rescaling():
    problem = random_sparse_lp(2000, 1500, 8, seed=41)
    params = folp_b200.PdhgParameters(verbosity=0)
    params.termination_criteria.eps_optimal_absolute = 1e-6
    params.termination_criteria.eps_optimal_relative = 1e-6
    a = folp_b200.optimize(params, problem)
    b = folp_b200.optimize(params, problem, device_rescaling=True)
    assert a.termination_reason == b.termination_reason == TerminationReason.TERMINATION_REASON_OPTIMAL
    assert a.iteration_count == b.iteration_count
    assert np.array_equal(a.primal_solution, b.primal_solution)   # same scaled problem bit for bit
    assert np.array_equal(a.dual_solution, b.dual_solution)
