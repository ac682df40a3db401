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
