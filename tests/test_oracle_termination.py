"""Pins the oracle's check_termination_criteria against test/test_termination.jl
(:33-69 infeasibility predicates, :109-143 optimality in both norms, :156-191
each terminal reason in isolation). The predicates are reached through
check_termination_criteria, whose precedence is optimal -> primal infeasible ->
dual infeasible -> iteration -> KKT passes -> time (src/termination.jl:238-272)."""
import copy

import numpy as np
import pytest

from folp_b200 import (OptimalityNorm, PdhgParameters, TerminationReason, _marshal,
                       construct_termination_criteria)
from folp_b200._abi import FolpEval
from oracle import oracle
from shared_problems import example_qp, generate_pdhg_params

R = TerminationReason


def _params(criteria):
    p = generate_pdhg_params()
    p.termination_criteria = criteria
    return _marshal.make_params(p, 1.0, 1.0, 0.0)


def _stats(conv=None, infeas=None, iteration_number=5, kkt=100.5, time_sec=5.0):
    e = FolpEval()
    e.iteration_number = iteration_number
    e.cumulative_kkt_matrix_passes = kkt
    e.cumulative_time_sec = time_sec
    for k, v in {**(conv or {}), **(infeas or {})}.items():
        setattr(e, k, v)
    return e


NO_INFEAS1 = {}
NO_INFEAS2 = {"primal_ray_linear_objective": -1.0, "primal_ray_quadratic_norm": 1.0,
              "max_dual_ray_infeasibility": 1.0}
DUAL_INFEASIBLE = {"primal_ray_linear_objective": -1.0}
PRIMAL_INFEASIBLE = {"dual_ray_objective": 1.0}

OPTIMAL = {"primal_objective": 1.0, "dual_objective": 1.0, "l_inf_primal_variable": 1.0,
           "l2_primal_variable": 1.0, "l_inf_dual_variable": 2.0, "l2_dual_variable": 2.0}
DONT1 = {**OPTIMAL, "primal_objective": 10.0}
DONT2 = {**OPTIMAL, "l_inf_primal_residual": 1.0, "l2_primal_residual": 1.0}
DONT3 = {**OPTIMAL, "l_inf_dual_residual": 1.0, "l2_dual_residual": 1.0}


@pytest.fixture(scope="module")
def holder():
    return oracle.problem_struct(example_qp())  # carries the CachedQuadraticProgramInfo


def test_infeasibility_predicates(holder):  # :33-69
    # eps_optimal = 0: the strict `<` tests of optimality_criteria_met never hold, and no
    # limit is reached, so the result is exactly the pair of infeasibility predicates
    crit = construct_termination_criteria(
        optimality_norm=OptimalityNorm.L2, eps_optimal_absolute=0.0, eps_optimal_relative=0.0,
        eps_primal_infeasible=1e-6, eps_dual_infeasible=1e-6, time_sec_limit=np.inf,
        iteration_limit=2 ** 31 - 1, kkt_matrix_pass_limit=np.inf)
    prm = _params(crit)
    assert oracle.check_termination(prm, holder, _stats(infeas=NO_INFEAS1)) == 0
    assert oracle.check_termination(prm, holder, _stats(infeas=NO_INFEAS2)) == 0
    assert oracle.check_termination(prm, holder, _stats(infeas=DUAL_INFEASIBLE)) == R.TERMINATION_REASON_DUAL_INFEASIBLE
    assert oracle.check_termination(prm, holder, _stats(infeas=PRIMAL_INFEASIBLE)) == R.TERMINATION_REASON_PRIMAL_INFEASIBLE


@pytest.mark.parametrize("norm", [OptimalityNorm.L_INF, OptimalityNorm.L2])
def test_optimality_and_limits(holder, norm):  # :105-191
    def criteria(**kw):
        base = dict(optimality_norm=norm, eps_optimal_absolute=1e-4, eps_optimal_relative=1e-4,
                    eps_primal_infeasible=1e-6, eps_dual_infeasible=1e-6, time_sec_limit=100.0,
                    iteration_limit=10, kkt_matrix_pass_limit=10000.0)
        base.update(kw)
        return _params(construct_termination_criteria(**base))

    full = criteria()
    for conv in (DONT1, DONT2, DONT3):
        assert oracle.check_termination(full, holder, _stats(conv)) == 0
    assert oracle.check_termination(full, holder, _stats(OPTIMAL)) == R.TERMINATION_REASON_OPTIMAL
    assert oracle.check_termination(criteria(time_sec_limit=1.0), holder, _stats(DONT1)) == R.TERMINATION_REASON_TIME_LIMIT
    assert oracle.check_termination(criteria(time_sec_limit=10.0, iteration_limit=1), holder,
                                    _stats(DONT1)) == R.TERMINATION_REASON_ITERATION_LIMIT
    assert oracle.check_termination(criteria(time_sec_limit=10.0, kkt_matrix_pass_limit=40.0), holder,
                                    _stats(DONT1)) == R.TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT
