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


def test_verbosity_table_and_final_log(capsys):
    """Row E37: with verbosity >= 2 optimize() drives the loop evaluation by evaluation and prints
    the reference's table (isu.jl:499-619) and final logs (pdhg.jl:324-370, sp.jl:947-1013); the
    result is the one of the silent single-call path, bit for bit."""
    import numpy as np
    from folp_b200 import display
    from shared_problems import example_lp, generate_pdhg_params
    quiet = folp_b200.optimize(generate_pdhg_params(iteration_limit=300, verbosity=0), example_lp())
    capsys.readouterr()
    loud = folp_b200.optimize(generate_pdhg_params(iteration_limit=300, verbosity=9), example_lp())
    text = capsys.readouterr().out.splitlines()
    assert np.array_equal(quiet.primal_solution, loud.primal_solution)
    assert np.array_equal(quiet.dual_solution, loud.dual_solution)
    assert quiet.iteration_count == loud.iteration_count == 300
    assert len(quiet.iteration_stats) == len(loud.iteration_stats)
    heading = display.iteration_stats_heading(True).split("\n")
    at = text.index(heading[0])        # rescale_problem's own verbosity >= 3 line comes first
    assert text[at + 1] == heading[1] and text[:at] == ["No rescaling."]
    rows = [t for t in text if t[:1].isdigit()]
    shown = [s_ for k, s_ in enumerate(loud.iteration_stats)
             if display.print_to_screen_this_iteration(k == len(loud.iteration_stats) - 1,
                                                       s_.iteration_number + 1, 9, 5)]
    assert len(rows) == len(shown) and len(shown) >= 60    # verbosity 9: every fifth iteration here
    assert rows[-1] == display.iteration_stats_row(loud.iteration_stats[-1], True)
    assert "Avg solution:" in text and "Terminated after 301 iterations: ITERATION_LIMIT" in text
    assert any(t.lstrip().startswith("41 norms=(") for t in text)   # pdhg_specific_log, verbosity >= 6
