"""The rule by which a folp_eval record of the CUDA path is compared with the CPU oracle's.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): used by tests/test_gpu_parity.py,
tests/dist_worker.py and the cpu_baseline leg of bench.py (`detail.parity`), always as the checker.

Tolerance (north star: "matching KKT residuals to 1e-9 relative"): every floating-point field of
an IterationStats record within `tol` = 1e-9 relative, where objective-like fields are measured
against max(|primal objective|, |dual objective|, 1). PDHG with the adaptive step-size rule
amplifies rounding noise (the oracle run twice, the second time with its initial step size moved by
ONE ulp, drifts apart ~10x every 30 iterations of the transient phase), and the order of long
reductions is such a perturbation; when the caller supplies the record of the one-ulp-perturbed
oracle, the bound is max(tol * scale, 20 * |perturbed - oracle|).
"""
from __future__ import annotations

import math

EVAL_FIELDS = [
    "primal_objective", "dual_objective", "l_inf_primal_residual", "l2_primal_residual",
    "l_inf_dual_residual", "l2_dual_residual", "relative_l_inf_primal_residual",
    "relative_l2_primal_residual", "relative_l_inf_dual_residual", "relative_l2_dual_residual",
    "relative_optimality_gap", "l_inf_primal_variable", "l2_primal_variable",
    "l_inf_dual_variable", "l2_dual_variable", "max_primal_ray_infeasibility",
    "primal_ray_linear_objective", "primal_ray_quadratic_norm", "max_dual_ray_infeasibility",
    "dual_ray_objective",
    "cumulative_kkt_matrix_passes", "step_size", "primal_weight", "lagrangian_value",
    "estimated_lower_bound", "estimated_upper_bound",
]


def compare_eval(eg, eo, tol=1e-9, e_pert=None, restart_length=None):
    """eg (CUDA path) against eo (oracle). Returns (problems, max_rel_err): `problems` is a list of
    human-readable mismatches (empty = parity holds), max_rel_err the largest |a-b| / scale seen.

    e_pert: the oracle's record when its initial step size is one ulp larger (the inherent
    sensitivity at this iteration). restart_length: iterations since the last restart; with a single
    iterate in the average, average == current up to one rounding, and should_reset_to_average
    (sp.jl:530-547, a `>=` between two equal quantities) is decided by that rounding: restarting
    "to the average" (3) or "to the current iterate" (2) are both the reference's behaviour."""
    problems = []
    worst = 0.0
    if eg.iteration_number != eo.iteration_number:
        problems.append(f"iteration_number {eg.iteration_number} != {eo.iteration_number}")
    if eg.restart_used != eo.restart_used:
        if not (restart_length == 1 and {eg.restart_used, eo.restart_used} == {2, 3}):
            problems.append(f"restart_used {eg.restart_used} != {eo.restart_used} at iteration "
                            f"{eo.iteration_number}")
    if eg.termination_reason != eo.termination_reason:
        problems.append(f"termination_reason {eg.termination_reason} != {eo.termination_reason}")
    ref = max(abs(eo.primal_objective), abs(eo.dual_objective), 1.0)
    for f in EVAL_FIELDS:
        a, b = getattr(eg, f), getattr(eo, f)
        if math.isnan(b):
            if not math.isnan(a):
                problems.append(f"{f}: {a} where the oracle has NaN")
            continue
        if math.isinf(b):
            if a != b:
                problems.append(f"{f}: {a} != {b}")
            continue
        objective_like = "objective" in f or "bound" in f or "lagr" in f
        scale = max(abs(b), ref if objective_like else 0.0, 1e-12)
        bound = tol * scale
        if e_pert is not None and math.isfinite(getattr(e_pert, f)):
            bound = max(bound, 20 * abs(getattr(e_pert, f) - b))
        err = abs(a - b)
        if math.isnan(err):
            problems.append(f"{f}: {a} where the oracle has {b}")
            continue
        worst = max(worst, err / scale)
        if err > bound:
            problems.append(f"{f}: {a!r} vs {b!r} (|diff| {err:.3e} > {bound:.3e}) at iteration "
                            f"{eo.iteration_number}")
    return problems, worst


def lockstep(oracle_solver, gpu_solver, perturbed_oracle=None, tol=1e-9, until_iteration=None):
    """Steps the solvers evaluation by evaluation (folp_run / oracle_run) until termination (or
    until_iteration) and compares every record. Returns a summary dict."""
    last_restart_iter, n_restart, records = 0, 0, 0
    problems, worst = [], 0.0
    restart_equal = True
    eo = None
    while True:
        eo, eg = oracle_solver.run(), gpu_solver.run()
        e2 = perturbed_oracle.run() if perturbed_oracle is not None else None
        p, w = compare_eval(eg, eo, tol, e2, eo.iteration_number - last_restart_iter)
        problems += p
        worst = max(worst, w)
        restart_equal = restart_equal and eg.restart_used == eo.restart_used
        records += 1
        if eo.restart_used >= 2:
            n_restart += 1
            last_restart_iter = eo.iteration_number
        if eo.termination_reason != 0 or eg.termination_reason != 0:
            break
        if until_iteration is not None and eo.iteration_number >= until_iteration:
            break
    return {"records": records, "iterations": int(eo.iteration_number), "restarts": n_restart,
            "max_rel_err": worst, "tolerance": tol, "restart_choices_equal": restart_equal,
            "sensitivity_probe": perturbed_oracle is not None, "ok": not problems,
            "problems": problems[:8], "last_oracle_record": eo}
