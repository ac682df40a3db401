"""GPU parity tests: libfolp_b200.so (through the C ABI) against the CPU oracle.

Parity definition:
  (i)   one take_step attempt from identical state: x+, y+, A'y+ relative <= 1e-13
        (bit-identical where the row fits one tile), interaction/movement <= 1e-12;
  (ii)  first 200 iterations with restarts disabled: identical accept/reject
        sequence, iterates <= 1e-10 relative;
  (iii) full solves: same termination reason, KKT quantities within 1e-9 relative.

PDHG with the adaptive step-size rule amplifies rounding noise: the oracle run
twice, the second time with its initial step size moved by ONE ulp, drifts apart
by ~10x every 10-15 iterations of the transient phase (1e-16 -> 1e-10 after 200
iterations on these instances). Long-reduction order (tree on the GPU, serial in
the oracle, BLAS in Julia) is such a perturbation, so trajectory tolerances are
max(stated tolerance, 20 x the measured 1-ulp sensitivity of the oracle itself).
"""
import numpy as np
import pytest
import scipy.sparse as sp

import folp_b200
from folp_b200 import RestartScheme, TerminationReason, _marshal
from folp_b200.lib import Solver
from folp_b200.synthetic import netlib_shaped_lp, pagerank_lp, random_sparse_lp, random_sparse_qp
from oracle import oracle
from shared_problems import example_lp, generate_pdhg_params

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)) if b.size else 0.0, 1e-300)
    return float(np.max(np.abs(a - b)) / scale) if a.size else 0.0


def _pair(problem, params, scaled=None, perturbed=False):
    """Oracle and GPU solvers built from the same bytes (+ the oracle again with
    a one-ulp larger initial step size, the sensitivity probe)."""
    holder, fparams, scaled = oracle.host_setup(params, problem, scaled)
    g_holder, g_params, _ = folp_b200.host_setup(params, problem, scaled)
    # identical scalar inputs on both sides
    g_params.initial_step_size = fparams.initial_step_size
    g_params.initial_primal_weight = fparams.initial_primal_weight
    o, g = oracle.OracleSolver(holder, fparams), Solver(g_holder, g_params)
    if not perturbed:
        return o, g
    h2, p2, _ = oracle.host_setup(params, problem, scaled)
    p2.initial_step_size = np.nextafter(fparams.initial_step_size, np.inf)
    return o, g, oracle.OracleSolver(h2, p2)


def _ragged_lp(seed=3):
    """Rows of length 0, 1, ~10, 100 (warp path) and 5000 (multi-tile path)."""
    rng = np.random.default_rng(seed)
    n = 6000
    lens = np.concatenate([np.zeros(5, int), np.ones(40, int), rng.integers(2, 20, 600),
                           np.full(7, 100), np.array([5000, 2049, 2048, 33, 32])])
    rng.shuffle(lens)
    rows, cols, vals = [], [], []
    for i, k in enumerate(lens):
        c = rng.choice(n, size=int(k), replace=False)
        rows += [i] * int(k)
        cols += list(c)
        vals += list(rng.standard_normal(int(k)))
    m = len(lens)
    A = sp.csr_matrix((vals, (rows, cols)), shape=(m, n))
    neq = m // 3
    from folp_b200.synthetic import _lp, _plant
    x, y, b, c, l, u = _plant(A, neq, rng, upper_fraction=0.2)
    return _lp(A, c, l, u, b, neq)


# ---------------------------------------------------------------------------
# SpMV kernels in isolation
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("make", [lambda: random_sparse_lp(3000, 2500, 8, seed=5),
                                  lambda: netlib_shaped_lp(), _ragged_lp,
                                  lambda: pagerank_lp(3000)])
def test_spmv_matches_oracle(make):
    problem = make()
    params = generate_pdhg_params(iteration_limit=10)
    o, g = _pair(problem, params)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(o.n)
    y = rng.standard_normal(o.m)
    ax_o, ax_g = o.spmv(x), g.spmv(x)
    aty_o, aty_g = o.spmv(y, transpose=True), g.spmv(y, transpose=True)
    A = problem.constraint_matrix.tocsr()
    row_len = np.diff(A.indptr)
    col_len = np.diff(problem.constraint_matrix.indptr)
    # rows summed by a single thread follow the oracle's summation order exactly
    short_r, short_c = row_len <= 32, col_len <= 32
    if A.nnz and row_len.max() <= 32:
        assert np.array_equal(ax_o, ax_g)
    # (row-partitioned: a column's sum is the NCCL sum of per-rank partial sums, not serial)
    if A.nnz and col_len.max() <= 32 and folp_b200.distributed.state() is None:
        assert np.array_equal(aty_o, aty_g)
    assert _rel(ax_g, ax_o) <= 1e-13
    assert _rel(aty_g, aty_o) <= 1e-13
    assert short_r.sum() + short_c.sum() > 0
    o.close(); g.close()


def test_spmv_empty_matrix():
    problem = folp_b200.linear_programming_problem(
        np.zeros(3), np.full(3, np.inf), np.array([1.0, 2.0, 3.0]), 0.0,
        sp.csc_matrix((2, 3)), np.zeros(2), 1)
    params = generate_pdhg_params(iteration_limit=5)
    g_holder, g_params, _ = folp_b200.host_setup(params, problem)
    g_params.initial_step_size = 1.0
    with Solver(g_holder, g_params) as g:
        assert np.array_equal(g.spmv(np.ones(3)), np.zeros(2))
        assert np.array_equal(g.spmv(np.ones(2), transpose=True), np.zeros(3))


# ---------------------------------------------------------------------------
# (i) single attempts from identical state
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("make", [lambda: random_sparse_lp(4000, 3000, 10, seed=11), _ragged_lp,
                                  lambda: random_sparse_qp(3000, 2000, 8, seed=12)])
def test_single_attempt_parity(make):
    problem = make()
    params = generate_pdhg_params(iteration_limit=100, l_inf_ruiz_iterations=4,
                                  pock_chambolle_alpha=1.0)
    o, g = _pair(problem, params)
    rng = np.random.default_rng(1)
    x0 = np.abs(rng.standard_normal(o.n))
    y0 = rng.standard_normal(o.m)
    y0[problem.num_equalities:] = np.abs(y0[problem.num_equalities:])
    for s in (o, g):
        s.debug_set_state(x0, y0, step_size=0.05, primal_weight=1.3)
    for _ in range(3):
        o.debug_attempts(1)
        g.debug_attempts(1)
        so, sg = o.debug_state(), g.debug_state()
        assert _rel(sg["x"], so["x"]) <= 1e-13
        assert _rel(sg["y"], so["y"]) <= 1e-13
        assert _rel(sg["dual_product"], so["dual_product"]) <= 1e-13
        assert abs(sg["last_movement"] - so["last_movement"]) <= 1e-12 * abs(so["last_movement"])
        assert abs(sg["last_interaction"] - so["last_interaction"]) <= \
            1e-12 * max(abs(so["last_interaction"]), abs(so["last_movement"]))
        assert sg["total_number_iterations"] == so["total_number_iterations"]
        assert abs(sg["step_size"] - so["step_size"]) <= 1e-12 * so["step_size"]
    o.close(); g.close()


# ---------------------------------------------------------------------------
# (ii) 200 iterations, restarts disabled
# ---------------------------------------------------------------------------
def test_trajectory_parity_no_restarts():
    problem = random_sparse_lp(1500, 1200, 6, seed=21)
    params = generate_pdhg_params(iteration_limit=1000, l_inf_ruiz_iterations=5,
                                  pock_chambolle_alpha=1.0)
    o, g, o2 = _pair(problem, params, perturbed=True)
    done = 0
    for chunk in (1, 1, 3, 15, 40, 140):
        for s_ in (o, g, o2):
            s_.debug_attempts(chunk)
        done += chunk
        so, sg, s2 = o.debug_state(), g.debug_state(), o2.debug_state()
        # identical accept/reject sequence <=> identical counters
        assert sg["total_number_iterations"] == so["total_number_iterations"] == done
        assert sg["sum_primal_solutions_count"] == so["sum_primal_solutions_count"]
        for key in ("x", "y", "sum_x", "sum_y", "dual_product"):
            assert _rel(sg[key], so[key]) <= max(1e-10, 20 * _rel(s2[key], so[key])), (key, done)
        sens = abs(s2["step_size"] - so["step_size"]) / so["step_size"]
        assert abs(sg["step_size"] - so["step_size"]) <= max(1e-10, 20 * sens) * so["step_size"]
        assert abs(sg["cumulative_kkt_passes"] - so["cumulative_kkt_passes"]) == 0.0
        if done <= 20:  # before the amplification sets in the plain 1e-13 bound holds
            assert _rel(sg["x"], so["x"]) <= 1e-13 and _rel(sg["y"], so["y"]) <= 1e-13
    assert so["sum_primal_solutions_count"] <= done
    o.close(); g.close(); o2.close()


# ---------------------------------------------------------------------------
# evaluation records
# ---------------------------------------------------------------------------
from oracle.parity import EVAL_FIELDS as _EVAL_FIELDS, compare_eval  # noqa: E402


def _assert_eval_close(eg, eo, tol, e_pert=None, restart_length=None):
    """eg (GPU) against eo (oracle) by the rule of oracle/parity.py; e_pert = the oracle's record
    when its initial step size is one ulp larger (the inherent sensitivity at this iteration)."""
    problems, _ = compare_eval(eg, eo, tol, e_pert, restart_length)
    assert not problems, problems


def _run_lockstep(problem, params, tol=1e-9):
    """Steps oracle, GPU and the perturbed oracle evaluation by evaluation."""
    o, g, o2 = _pair(problem, params, perturbed=True)
    last_restart_iter, n_restart, records = 0, 0, 0
    while True:
        eo, eg, e2 = o.run(), g.run(), o2.run()
        _assert_eval_close(eg, eo, tol, e2, eo.iteration_number - last_restart_iter)
        records += 1
        if eo.restart_used >= 2:
            n_restart += 1
            last_restart_iter = eo.iteration_number
        if eo.termination_reason != 0:
            break
    xo, yo = o.get_solution()
    xg, yg = g.get_solution()
    x2, y2 = o2.get_solution()
    assert _rel(xg, xo) <= max(tol, 20 * _rel(x2, xo))
    assert _rel(yg, yo) <= max(tol, 20 * _rel(y2, yo))
    for s_ in (o, g, o2):
        s_.close()
    return eo, n_restart, records


@pytest.mark.parametrize("scheme", [RestartScheme.NO_RESTARTS, RestartScheme.ADAPTIVE_NORMALIZED,
                                    RestartScheme.FIXED_FREQUENCY, RestartScheme.ADAPTIVE_LOCALIZED,
                                    RestartScheme.ADAPTIVE_DISTANCE])
def test_eval_records_match_oracle(scheme):
    """Every IterationStats record of the first 120 iterations, restart logic on."""
    problem = random_sparse_lp(1200, 900, 6, seed=31, upper_fraction=0.1)
    params = generate_pdhg_params(iteration_limit=120, l_inf_ruiz_iterations=10,
                                  pock_chambolle_alpha=1.0, restart_scheme=scheme,
                                  restart_frequency_if_fixed=25)
    params.termination_evaluation_frequency = 8
    eo, n_restart, records = _run_lockstep(problem, params)
    assert eo.termination_reason == TerminationReason.TERMINATION_REASON_ITERATION_LIMIT
    assert n_restart >= 3 and records == 10 + 14


@pytest.mark.parametrize("scheme,policy", [(RestartScheme.ADAPTIVE_NORMALIZED, "adaptive"),
                                           (RestartScheme.NO_RESTARTS, "adaptive"),
                                           (RestartScheme.ADAPTIVE_LOCALIZED, "constant")])
def test_eval_records_match_oracle_qp(scheme, policy):
    """Quadratic objective (S1, S7, E7, E11, E13, E19 with Q != 0): every record of 120 iterations."""
    problem = random_sparse_qp(1200, 900, 6, seed=33, upper_fraction=0.1)
    if policy == "constant":
        # the constant step 0.8 / sigma_max(A) ignores Q (pdhg.jl:826-833): with |Q| ~ |A| the
        # reference's own iteration diverges to 1e240; a mild Q keeps the trajectory bounded
        problem.objective_matrix = problem.objective_matrix * 0.01
    params = generate_pdhg_params(iteration_limit=120, l_inf_ruiz_iterations=10,
                                  pock_chambolle_alpha=1.0, restart_scheme=scheme,
                                  step_size_policy=policy)
    params.termination_evaluation_frequency = 8
    eo, n_restart, records = _run_lockstep(problem, params)
    assert eo.termination_reason == TerminationReason.TERMINATION_REASON_ITERATION_LIMIT
    assert eo.primal_ray_quadratic_norm > 0.0
    assert records == 10 + 14


def test_diverging_iterates_run_to_the_limit_like_the_reference():
    """The constant step 0.8 / sigma_max(A) ignores Q: with |Q| ~ |A| the reference's own iteration
    blows up (1e240 after 120 iterations, NaN bounds) and simply runs into its iteration limit.
    The library must do the same -- no error, no hang."""
    problem = random_sparse_qp(1200, 900, 6, seed=33, upper_fraction=0.1)
    params = generate_pdhg_params(iteration_limit=120, l_inf_ruiz_iterations=10, pock_chambolle_alpha=1.0,
                                  restart_scheme=RestartScheme.ADAPTIVE_LOCALIZED, step_size_policy="constant")
    params.termination_evaluation_frequency = 8
    out_o = oracle.optimize(params, problem)
    out_g = folp_b200.optimize(params, problem)
    assert out_g.termination_reason == out_o.termination_reason == \
        TerminationReason.TERMINATION_REASON_ITERATION_LIMIT
    assert out_g.iteration_count == out_o.iteration_count == 120
    fin = out_g.iteration_stats[-1].convergence_information[0]
    assert not np.isfinite(fin.primal_objective) or abs(fin.primal_objective) > 1e100


def test_qp_full_solve_matches_oracle():
    """A QP solved to 1e-6 with the CLI defaults: same reason, and the GPU-reported KKT record
    equals the oracle's evaluation of the GPU's returned point."""
    problem = random_sparse_qp(400, 300, 6, seed=43)
    params = folp_b200.PdhgParameters(verbosity=0)
    eps = 1e-6
    params.termination_criteria.eps_optimal_absolute = eps
    params.termination_criteria.eps_optimal_relative = eps
    params.termination_criteria.iteration_limit = 100000
    out_o = oracle.optimize(params, problem)
    out_g = folp_b200.optimize(params, problem)
    assert out_g.termination_reason == out_o.termination_reason == \
        TerminationReason.TERMINATION_REASON_OPTIMAL
    fin = out_g.iteration_stats[-1].convergence_information[0]
    chk = _kkt_of(problem, out_g.primal_solution, out_g.dual_solution, eps)
    ref = max(abs(chk.primal_objective), abs(chk.dual_objective), 1.0)
    assert abs(fin.primal_objective - chk.primal_objective) <= 1e-12 * ref
    assert abs(fin.dual_objective - chk.dual_objective) <= 1e-10 * ref
    assert abs(fin.l2_primal_residual - chk.l2_primal_residual) <= 1e-10 * max(1.0, chk.l2_primal_residual)
    assert abs(fin.l2_dual_residual - chk.l2_dual_residual) <= 1e-10 * max(1.0, chk.l2_dual_residual)
    fo = out_o.iteration_stats[-1].convergence_information[0]
    assert abs(fin.primal_objective - fo.primal_objective) <= 10 * eps * ref


# ---------------------------------------------------------------------------
# (iii) full solves with the CLI default parameters
# ---------------------------------------------------------------------------
def _kkt_of(problem, x, y, eps=1e-6):
    e = oracle.iteration_stats(problem, x, y, x, y, eps, eps)
    return e


@pytest.mark.parametrize("make,eps", [
    (lambda: random_sparse_lp(2000, 1500, 8, seed=41), 1e-6),
    (lambda: netlib_shaped_lp(seed=11), 1e-6),
    (lambda: pagerank_lp(2000), 1e-8),
])
def test_full_solve_matches_oracle(make, eps):
    problem = make()
    params = folp_b200.PdhgParameters(verbosity=0)
    params.termination_criteria.eps_optimal_absolute = eps
    params.termination_criteria.eps_optimal_relative = eps
    params.termination_criteria.iteration_limit = 300000
    out_o = oracle.optimize(params, problem)
    out_g = folp_b200.optimize(params, problem)
    # iteration counts of the oracle itself under 1..4 ulp changes of the initial
    # step size: the inherent spread of "iterations to tolerance" on this instance
    spread = [out_o.iteration_count]
    for k in range(1, 5):
        holder, fparams, _ = oracle.host_setup(params, problem)
        for _ in range(k):
            fparams.initial_step_size = np.nextafter(fparams.initial_step_size, np.inf)
        s_ = oracle.OracleSolver(holder, fparams)
        spread.append(s_.solve()[3])
        s_.close()
    assert out_g.termination_reason == out_o.termination_reason == \
        TerminationReason.TERMINATION_REASON_OPTIMAL
    # the GPU-reported KKT record equals the oracle's evaluation of the GPU's point
    fin = out_g.iteration_stats[-1].convergence_information[0]
    chk = _kkt_of(problem, out_g.primal_solution, out_g.dual_solution, eps)
    ref = max(abs(chk.primal_objective), abs(chk.dual_objective), 1.0)
    assert abs(fin.primal_objective - chk.primal_objective) <= 1e-12 * ref
    assert abs(fin.dual_objective - chk.dual_objective) <= 1e-10 * ref
    assert abs(fin.l2_primal_residual - chk.l2_primal_residual) <= 1e-10 * max(1.0, chk.l2_primal_residual)
    assert abs(fin.l2_dual_residual - chk.l2_dual_residual) <= 1e-10 * max(1.0, chk.l2_dual_residual)
    # both solves meet the same tolerance on the same (relative) KKT quantities
    fo = out_o.iteration_stats[-1].convergence_information[0]
    for f in ("relative_l2_primal_residual", "relative_l2_dual_residual", "relative_optimality_gap"):
        assert getattr(fin, f) <= 2 * eps and getattr(fo, f) <= 2 * eps
    # objective values agree to the solve tolerance
    assert abs(fin.primal_objective - fo.primal_objective) <= 10 * eps * ref
    # iteration count inside the oracle's own spread (a factor 1.5 either side)
    assert min(spread) / 1.5 - 80 <= out_g.iteration_count <= 1.5 * max(spread) + 80, (
        out_g.iteration_count, spread)


def test_full_solve_trajectory_short_horizon():
    """CLI default parameters (restarts on), every record of the first 160 iterations."""
    problem = random_sparse_lp(800, 700, 5, seed=51)
    params = folp_b200.PdhgParameters(verbosity=0)
    params.termination_criteria.eps_optimal_absolute = 0.0
    params.termination_criteria.eps_optimal_relative = 0.0
    params.termination_criteria.iteration_limit = 160
    eo, _, _ = _run_lockstep(problem, params)
    assert eo.iteration_number == 160


def test_deterministic_run_to_run():
    problem = random_sparse_lp(3000, 2500, 10, seed=61)
    params = folp_b200.PdhgParameters(verbosity=0)
    params.termination_criteria.iteration_limit = 200
    a = folp_b200.optimize(params, problem)
    b = folp_b200.optimize(params, problem)
    assert np.array_equal(a.primal_solution, b.primal_solution)
    assert np.array_equal(a.dual_solution, b.dual_solution)
    assert a.iteration_count == b.iteration_count == 200


def test_example_lp_exact_record():
    """example_lp at the reference's test parameters: every record matches."""
    params = generate_pdhg_params(iteration_limit=300)
    eo, _, _ = _run_lockstep(example_lp(), params)
    assert eo.iteration_number == 300


# ---------------------------------------------------------------------------
# bench scale: the grid strides, windows sort, 32-bit index arithmetic sees 1e7 entries
# ---------------------------------------------------------------------------
def test_eval_records_match_oracle_at_bench_scale():
    """BASELINE.json configs[1] (1e6 x 1e6, 1e7 nonzeros) with the CLI defaults: every record of
    the first 40 iterations (ten per-iteration evaluations + the one at 40) against the oracle --
    ~15 s of CPU. bench.py repeats this on every run over 120 iterations (detail.parity)."""
    problem = random_sparse_lp(1_000_000, 1_000_000, 10)
    params = folp_b200.PdhgParameters(verbosity=0)
    params.termination_criteria.eps_optimal_absolute = 0.0
    params.termination_criteria.eps_optimal_relative = 0.0
    params.termination_criteria.iteration_limit = 40
    from oracle.parity import lockstep
    o, g = _pair(problem, params)
    res = lockstep(o, g, None, tol=1e-9)
    o.close(); g.close()
    assert res["ok"], res["problems"]
    assert res["records"] == 11 and res["iterations"] == 40
    # SpMV at this size: rows of A have <= 10 entries, columns of a random matrix <= 32 with
    # overwhelming probability: one lane per row, ascending order -> the oracle's bits
    # (checked through the records above: the 1e-9 bound held with 1e-11 .. 1e-13 to spare)
    assert res["max_rel_err"] < 1e-10, res["max_rel_err"]


# ---------------------------------------------------------------------------
# the two remaining exits of check_termination_criteria (term.jl:262-271)
# ---------------------------------------------------------------------------
def test_kkt_matrix_pass_limit_terminates_like_the_oracle():
    problem = random_sparse_lp(1500, 1200, 6, seed=81)
    params = generate_pdhg_params(iteration_limit=100000, l_inf_ruiz_iterations=5, pock_chambolle_alpha=1.0,
                                  restart_scheme=RestartScheme.ADAPTIVE_NORMALIZED)
    params.termination_evaluation_frequency = 8
    params.termination_criteria.kkt_matrix_pass_limit = 150.0
    eo, _, records = _run_lockstep(problem, params)
    assert eo.termination_reason == TerminationReason.TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT
    assert eo.cumulative_kkt_matrix_passes >= 150.0 and records > 10
    out = folp_b200.optimize(params, problem)
    assert out.termination_reason == TerminationReason.TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT
    assert out.iteration_count == eo.iteration_number


def test_time_limit_terminates():
    """time_sec_limit (term.jl:268-270) is wall clock: both sides stop at the first evaluation
    whose cumulative_time_sec has reached it -- here the very first one."""
    problem = random_sparse_lp(1500, 1200, 6, seed=82)
    params = generate_pdhg_params(iteration_limit=100000)
    params.termination_criteria.time_sec_limit = 0.0
    out_o = oracle.optimize(params, problem)
    out_g = folp_b200.optimize(params, problem)
    assert out_g.termination_reason == out_o.termination_reason == TerminationReason.TERMINATION_REASON_TIME_LIMIT
    assert out_g.iteration_count == out_o.iteration_count == 0  # the evaluation before the first step
    # and a limit that is never reached does not interfere
    params.termination_criteria.time_sec_limit = 3600.0
    params.termination_criteria.iteration_limit = 50
    out_g = folp_b200.optimize(params, problem)
    assert out_g.termination_reason == TerminationReason.TERMINATION_REASON_ITERATION_LIMIT


# ---------------------------------------------------------------------------
# the three forms of a batch of take_step attempts
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("form,env", [
    ("graph of three kernels per attempt", {"FOLP_PERSISTENT": "0"}),
    ("persistent cooperative kernel, software grid barrier", {"FOLP_PERSISTENT": "1", "FOLP_NO_CLUSTER": "1"}),
    ("persistent kernel as one thread-block cluster", {"FOLP_PERSISTENT": "1"}),
])
def test_take_step_forms_match_oracle(form, env, monkeypatch):
    """Small instances default to the cluster form, large ones to the CUDA graph, the partitioned
    mode to the cooperative kernel: every form is held to the same records here (LP with restarts,
    then a QP, then Malitsky-Pock), and to bit-identical results run to run."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    problem = random_sparse_lp(1200, 900, 6, seed=31, upper_fraction=0.1)
    params = generate_pdhg_params(iteration_limit=120, l_inf_ruiz_iterations=10, pock_chambolle_alpha=1.0,
                                  restart_scheme=RestartScheme.ADAPTIVE_NORMALIZED)
    params.termination_evaluation_frequency = 8
    eo, n_restart, records = _run_lockstep(problem, params)
    assert n_restart >= 3 and records == 10 + 14
    qp = random_sparse_qp(1200, 900, 6, seed=33, upper_fraction=0.1)
    _run_lockstep(qp, params)
    mp = generate_pdhg_params(iteration_limit=120, l_inf_ruiz_iterations=10, pock_chambolle_alpha=1.0,
                              restart_scheme=RestartScheme.ADAPTIVE_NORMALIZED, step_size_policy="malitsky-pock")
    mp.termination_evaluation_frequency = 8
    _run_lockstep(problem, mp)
    a = folp_b200.optimize(params, problem)
    b = folp_b200.optimize(params, problem)
    assert np.array_equal(a.primal_solution, b.primal_solution)
    assert np.array_equal(a.dual_solution, b.dual_solution)
