"""Pins the CPU oracle against the reference's own end-to-end PDHG tests
(test/test_primal_dual_hybrid_gradient.jl:76-424), same parameters and
tolerances. The cases live in tests/pdhg_cases.py; `optimize` here is the
whole-path restatement `oracle.oracle.optimize`."""
import pytest

from oracle import oracle
from pdhg_cases import CASES


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reference_pdhg_case_on_oracle(case):
    _, fn, kwargs, _ = case
    fn(oracle.optimize, **kwargs)


def test_all_cores_timing_mode_is_bit_identical():
    """oracle.set_threads(n > 1) (bench.py's stronger CPU baseline) must not change a single bit:
    A' * y is one independent dot product per column, A * x walks a CSR view in the order of the
    serial column scatter."""
    import numpy as np
    import folp_b200
    from folp_b200.synthetic import pagerank_lp, random_sparse_qp
    for problem in (random_sparse_qp(1500, 1000, 6, seed=4), pagerank_lp(1500)):
        params = folp_b200.PdhgParameters(verbosity=0)
        params.termination_criteria.iteration_limit = 160
        a = oracle.optimize(params, problem)
        oracle.set_threads(4)
        try:
            b = oracle.optimize(params, problem)
        finally:
            oracle.set_threads(1)
        assert np.array_equal(a.primal_solution, b.primal_solution)
        assert np.array_equal(a.dual_solution, b.dual_solution)
        fa, fb = (o.iteration_stats[-1].convergence_information[0] for o in (a, b))
        assert fa.primal_objective == fb.primal_objective and fa.l2_dual_residual == fb.l2_dual_residual
