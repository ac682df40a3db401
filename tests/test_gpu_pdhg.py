"""The reference's own end-to-end PDHG tests
(test/test_primal_dual_hybrid_gradient.jl:76-424) run on the B200 path:
folp_b200.optimize -> C ABI -> libfolp_b200.so. Same parameters and tolerances
as on the oracle (tests/pdhg_cases.py), quadratic-objective cases included.
What the reference rejects with error() (Malitsky-Pock on a QP, pdhg.jl:560-565)
fails loudly with FOLP_UNSUPPORTED; nothing falls back to a CPU path."""
import pytest

import folp_b200
from folp_b200.lib import FolpError
from pdhg_cases import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reference_pdhg_case_on_gpu(case):
    _, fn, kwargs, _needs_qp = case
    fn(folp_b200.optimize, **kwargs)


def test_malitsky_pock_rejects_qp():
    from shared_problems import example_qp, generate_pdhg_params
    params = generate_pdhg_params(iteration_limit=10, step_size_policy="malitsky-pock")
    with pytest.raises(FolpError) as err:
        folp_b200.optimize(params, example_qp())
    assert err.value.status == folp_b200.Status.UNSUPPORTED
