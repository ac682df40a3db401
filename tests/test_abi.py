"""The drop-in boundary without a GPU: libfolp_b200.so loads, exports every function
include/folp_b200.h declares (and nothing is declared that the Python binding does not
know), the ctypes mirrors have the C struct sizes, and argument errors come back as status
codes + messages, never as crashes. No compute call is made."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import folp_b200
from folp_b200 import _abi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "folp_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(folp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = lib.lib()
    declared = _declared_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/folp_b200.h but not exported"
    assert sorted(lib.EXPORTS) == declared, "firstorderlp.jl_b200/lib.py binds exactly the declared functions"


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of every POD as the C compiler sees them."""
    src = tmp_path / "sizes.c"
    fields = {
        "folp_problem": ["num_variables", "index_base", "colptr", "objective_constant", "orig_nzval",
                         "q_num_nonzeros", "l2_norm_primal_right_hand_side"],
        "folp_params": ["step_size_policy", "initial_step_size", "optimality_norm", "iteration_limit",
                        "restart_scheme", "restart_frequency_if_fixed", "use_approximate_localized_duality_gap",
                        "verbosity"],
        "folp_eval": ["iteration_number", "primal_objective", "dual_ray_objective", "step_size",
                      "estimated_upper_bound", "cumulative_rejected_steps", "termination_reason",
                      "total_number_iterations"],
        "folp_dist": ["rank", "device", "nccl_unique_id"],
        "folp_debug_scalars": ["step_size", "total_number_iterations", "numerical_error", "last_movement"],
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for s, fs in fields.items():
        lines.append(f'printf("{s} %zu\\n", sizeof({s}));')
        for f in fs:
            lines.append(f'printf("{s}.{f} %zu\\n", offsetof({s}, {f}));')
    lines.append("return 0;}")
    src.write_text("\n".join(lines))
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c11", "-o", str(exe), str(src)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    mirrors = {"folp_problem": _abi.FolpProblem, "folp_params": _abi.FolpParams, "folp_eval": _abi.FolpEval,
               "folp_dist": _abi.FolpDist, "folp_debug_scalars": _abi.FolpDebugScalars}
    for s, fs in fields.items():
        assert C.sizeof(mirrors[s]) == int(out[s]), s
        for f in fs:
            assert getattr(mirrors[s], f).offset == int(out[f"{s}.{f}"]), (s, f)


def test_enum_ordinals_follow_the_reference():
    """solve_log.jl:32-58,336-347; saddle_point.jl:325,340; termination.jl:15 (0-based @enum order)."""
    assert [e.value for e in _abi.RestartChoice] == [0, 1, 2, 3]
    assert _abi.RestartChoice.RESTART_CHOICE_RESTART_TO_AVERAGE == 3
    assert _abi.PointType.POINT_TYPE_AVERAGE_ITERATE == 3
    assert _abi.TerminationReason.TERMINATION_REASON_OPTIMAL == 1
    assert _abi.TerminationReason.TERMINATION_REASON_NUMERICAL_ERROR == 7
    assert _abi.RestartScheme.ADAPTIVE_NORMALIZED == 2
    assert _abi.RestartToCurrentMetric.GAP_OVER_DISTANCE_SQUARED == 2
    assert _abi.OptimalityNorm.L2 == 1
    assert _abi.StepSizePolicy.MALITSKY_POCK == 1


def test_argument_errors_are_status_codes():
    L = lib.lib()
    assert L.folp_create(None, None, None, None) == _abi.Status.INVALID_ARGUMENT
    assert b"null" in L.folp_last_error(None)
    assert L.folp_run(None, None) == _abi.Status.INVALID_ARGUMENT
    assert L.folp_get_solution(None, 0, 1, None, None) == _abi.Status.INVALID_ARGUMENT
    assert L.folp_exchange_mode(None) == -1
    L.folp_destroy(None)  # must be a no-op
    info = lib.build_info()
    assert "sm_100a" in info and "fp64" in info
    rb = np.zeros(3, dtype=np.int64)
    assert L.folp_partition(-1, 1, 0, None, 0, 2, rb.ctypes.data_as(C.POINTER(C.c_int64)),
                            rb.ctypes.data_as(C.POINTER(C.c_int64))) == _abi.Status.INVALID_ARGUMENT


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU the product path must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from shared_problems import example_lp, generate_pdhg_params
    with pytest.raises(lib.FolpError) as err:
        folp_b200.optimize(generate_pdhg_params(iteration_limit=10), example_lp())
    assert err.value.status in (_abi.Status.CUDA_ERROR, _abi.Status.UNSUPPORTED)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under firstorderlp.jl_b200/ may reference it."""
    pkg = os.path.join(ROOT, "firstorderlp.jl_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "folp_oracle" not in text, f
