"""The reference's own end-to-end PDHG tests
(test/test_primal_dual_hybrid_gradient.jl:76-424) run on the B200 path:
folp_b200.optimize -> C ABI -> libfolp_b200.so. Same parameters and tolerances
as on the oracle (tests/pdhg_cases.py). Quadratic objectives are not on the
B200 path yet: those cases must fail loudly with FOLP_UNSUPPORTED, never fall
back to a CPU path."""
import pytest

import folp_b200
from folp_b200.lib import FolpError
from pdhg_cases import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reference_pdhg_case_on_gpu(case):
    _, fn, kwargs, needs_qp = case
    if needs_qp:
        with pytest.raises(FolpError) as err:
            fn(folp_b200.optimize, **kwargs)
        assert err.value.status == folp_b200.Status.UNSUPPORTED
    else:
        fn(folp_b200.optimize, **kwargs)
