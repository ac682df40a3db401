#!/usr/bin/env python
"""bench.py -- PDHG iterations/sec on the synthetic random sparse LP of BASELINE.json.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

A "step" is ITERS_PER_STEP PDHG iterations (outer take_step calls, evaluation /
restart blocks included, exactly the reference's iteration_count) of
optimize(PdhgParameters defaults of scripts/solve_qp.jl, lp) on one resident
problem. `value` = iterations/sec with the problem already in HBM, timed with
CUDA events on the library's own stream; `e2e` = the same metric through the
C-ABI call a host makes (folp_create from HOST arrays + folp_solve +
folp_get_solution), wall clock, uploads and downloads inside the timed region.
`--impl reference` times the CPU oracle (the single-threaded C restatement of
the reference's loop; the Julia reference itself cannot run in this image).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ITERS_PER_STEP = 400
WORKLOADS = {
    # name: (n, m, nnz_per_row)
    "c2": (1_000_000, 1_000_000, 10),      # BASELINE.json configs[1]
    "target": (10_000_000, 10_000_000, 10),  # north_star target size
    "mid": (4_000_000, 4_000_000, 10),      # between the two: where the 1-D partition starts to pay
    "small": (100_000, 100_000, 10),
    "half": (1_000_000, 500_000, 10),       # one rank's row block of c2 at 2 GPUs
    # BASELINE.json configs[2] / [3] shapes (the real files / LightGraphs' RNG are not available offline)
    "netlib": "netlib",      # block-angular, 9 902 x 230 000, ~1.4e6 nonzeros, 12 dense-ish linking rows (osa-60 class)
    "pagerank": "pagerank",  # generate_pagerank_lp.jl at 1e6 nodes: one dense row + power-law degrees, ~8e6 nonzeros
    # the median Netlib instance class (~1e3 x 2e3, ~1.5e4 nonzeros): the launch-/latency-bound regime
    "netlib_small": "netlib_small",
    # BASELINE.json configs[3]: generate_pagerank_lp.jl at 1e7 nodes (8e7 nonzeros, one dense row of 1e7 entries)
    "pagerank7": "pagerank7",
}


WORKLOAD_LABEL = {"netlib": "synthetic Netlib-shaped block-angular LP", "pagerank": "PageRank LP (Barabasi-Albert graph)",
                  "pagerank7": "PageRank LP (Barabasi-Albert graph, 1e7 nodes)",
                  "netlib_small": "synthetic Netlib-shaped block-angular LP, median Netlib size"}

KERNEL_SOURCES = ("firstorderlp.jl_b200/csrc/folp_spmv.cuh", "firstorderlp.jl_b200/csrc/folp_kernels.cu",
                  "firstorderlp.jl_b200/csrc/folp_internal.cuh")


def kernel_source_hash():
    import hashlib
    h = hashlib.sha256()
    for rel in KERNEL_SOURCES:
        with open(os.path.join(ROOT, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def ncu_traffic(workload):
    """roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum per launch of the two k_spmv
    launches of an iteration, from the committed `ncu --set full` capture. The capture is stamped
    with the hash of the kernel sources it was taken from (profiles/traffic_ncu.json, written by
    tools/traffic_stamp.py): a stale stamp reports null rather than an old number."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic_ncu.json")))
    except Exception:
        return None, "no profiles/traffic_ncu.json"
    e = d.get(workload)
    if not e:
        return None, "no ncu capture of this workload"
    if e.get("kernel_source_sha16") != kernel_source_hash():
        return None, "the kernel sources changed since the ncu capture %s" % e.get("from")
    return e["traffic_bytes"], "ncu --set full, %s" % e.get("from")



def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout. Libraries print there too (NCCL's "NCCL version ..."
# banner at the first communicator, for one): file descriptor 1 is pointed at stderr for the whole
# run and the line goes to the saved descriptor.
_REAL_STDOUT = None


def _capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def make_problem(workload):
    import folp_b200
    from folp_b200.synthetic import random_sparse_lp

    t0 = time.time()
    if workload == "netlib":
        from folp_b200.synthetic import netlib_shaped_lp
        lp = netlib_shaped_lp(num_blocks=230, block_rows=43, block_cols=1000, linking_rows=12, density=0.06)
    elif workload == "netlib_small":
        from folp_b200.synthetic import netlib_shaped_lp
        lp = netlib_shaped_lp(num_blocks=24, block_rows=40, block_cols=90, linking_rows=12, density=0.08)
    elif workload == "pagerank":
        from folp_b200.synthetic import pagerank_lp
        lp = pagerank_lp(1_000_000)
    elif workload == "pagerank7":
        from folp_b200.synthetic import pagerank_lp
        lp = pagerank_lp(10_000_000)
    else:
        n, m, k = WORKLOADS[workload]
        lp = random_sparse_lp(n, m, k)
    n, m = lp.num_variables, lp.num_constraints
    params = folp_b200.PdhgParameters(verbosity=0)  # scripts/solve_qp.jl defaults
    # fixed amount of work per step: never stop on a tolerance
    params.termination_criteria.eps_optimal_absolute = 0.0
    params.termination_criteria.eps_optimal_relative = 0.0
    params.termination_criteria.eps_primal_infeasible = 0.0
    params.termination_criteria.eps_dual_infeasible = 0.0
    # large instances: rescale_problem on the device (folp_rescale_problem, bit-identical to the host mirror)
    big = lp.constraint_matrix.nnz > 30_000_000
    holder, fparams, scaled = folp_b200.host_setup(params, lp, device_rescaling=big)
    log(f"[bench] problem {workload}: n={n} m={m} nnz={lp.constraint_matrix.nnz} "
        f"generated on host and rescaled on the {'device' if big else 'host'} in {time.time() - t0:.1f}s")
    return lp, params, holder, fparams, scaled


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.samples = []
        self.device = device
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                      "-i", str(self.device)], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=5)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            if len(s) < 9:
                continue
            try:
                sm.append(float(s[1])); mx.append(float(s[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def config_dict(args, n, m, nnz):
    """The `config` object of the JSON line: identical for both arms of one (workload, --gpus)."""
    ws = 24 * nnz + 8 * (20 * n + 12 * m)
    return {"workload": f"{WORKLOAD_LABEL.get(args.workload, 'synthetic random sparse LP')} n={n} m={m} "
                        f"nnz={nnz} fp64 ({args.workload})",
            "iterations_per_step": ITERS_PER_STEP,
            "parameters": "scripts/solve_qp.jl defaults (ruiz 10, pock-chambolle 1.0, adaptive step 0.3/0.6, "
                          "adaptive_normalized restarts, evaluation every 40 iterations), tolerances 0 so every "
                          "step does the same work",
            "parallelism": "single GPU" if args.gpus == 1 else
                           f"1-D partition over {args.gpus} GPUs (nnz-balanced row blocks + primal slices)",
            "l2_flush": ("working set 24*nnz + vectors = %.0f MB > 126 MB L2; no explicit flush" if ws > 126e6 else
                         "working set 24*nnz + vectors = %.0f MB fits the 126 MB L2 (an L2-resident instance "
                         "class by design); no flush") % (ws / 1e6)}


def algorithmic_bytes(n, m, nnz):
    """bytes one accepted iteration must move (LP, averaging on)."""
    k1 = 8 * 9 * n                                   # primal step
    k2 = 12 * nnz + 4 * (m + 1) + 8 * n + 8 * 5 * m  # A*xbar + dual step
    k3 = 12 * nnz + 4 * (n + 1) + 8 * m + 8 * 4 * n  # A'*y + interaction
    return k1, k2, k3


def run_until(solver, iteration):
    while True:
        e = solver.run()
        if e.iteration_number >= iteration or e.termination_reason != 0:
            return e


def _max_over_ranks(x, world):
    """max of a python float over all ranks (device all-reduce on the NCCL group)."""
    if world == 1:
        return x
    import torch
    import torch.distributed as td
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def _barrier(world):
    import torch
    if world > 1:
        import torch.distributed as td
        td.barrier()
    torch.cuda.synchronize()


def bench_gpu(args):
    import torch
    import folp_b200
    from folp_b200 import distributed
    from folp_b200.lib import Solver, build_info

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    devices = None
    if args.single_process and args.gpus > 1:
        # ONE process drives all N GPUs through folp_create_multi (no torchrun, no NCCL): the entry a Julia
        # optimize() would use. Everything below sees "world 1": the library's handle is the whole job.
        devices = list(range(args.gpus))
        os.environ["FOLP_DEVICES"] = ",".join(str(d) for d in devices)
        world = 1
    elif world != args.gpus:
        raise SystemExit(f"bench.py: WORLD_SIZE={world} but --gpus {args.gpus}: for N>1 launch one rank per "
                         "GPU with python -m torch.distributed.run --nproc-per-node N")
    torch.cuda.set_device(local_rank)
    distributed.init("nccl")  # no-op when world == 1
    lp, params, holder, fparams, scaled = make_problem(args.workload)  # same seed on every rank
    n, m, nnz = lp.num_variables, lp.num_constraints, lp.constraint_matrix.nnz
    total_steps = args.warmup + args.steps
    fparams.iteration_limit = ITERS_PER_STEP * (total_steps + 1) + 10_000_000

    t0 = time.time()
    solver = Solver(holder, fparams)  # row-partitioned over all ranks when world > 1
    t_create = time.time() - t0
    log(f"[bench] rank {rank}/{world} folp_create {t_create:.2f}s shard {solver.shard_info()}; {build_info()}")
    exchange = solver.shard_info()["exchange"]
    stream = torch.cuda.ExternalStream(solver.stream())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    done = 0
    for _ in range(args.warmup):
        done += ITERS_PER_STEP
        run_until(solver, done)
    c0 = solver.counters()
    _barrier(world)
    with ClockSampler(local_rank) as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            done += ITERS_PER_STEP
            e = run_until(solver, done)
        ev1.record(stream)
        _barrier(world)
    ms = _max_over_ranks(ev0.elapsed_time(ev1), world)
    c1 = solver.counters()
    iters = c1["iterations"] - c0["iterations"]
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    basic_s = c1["basic_algorithm_seconds"] - c0["basic_algorithm_seconds"]
    value = iters / (ms * 1e-3)

    b1, b2, b3 = algorithmic_bytes(n, m, nnz)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured" if "hbm_gbs" in peaks else "fallback"
    iteration = {"bytes": b1 + b2 + b3, "gbs_at_value": (b1 + b2 + b3) * value / 1e9,
                 "frac_at_value": (b1 + b2 + b3) * value / 1e9 / (peak * args.gpus),
                 "matrix_only_gbs_at_value": 24 * nnz * value / 1e9}
    if args.gpus == 1:
        # per-kernel device time of real attempts, continuing the same solve
        prof_attempts = 100
        solver.profile_attempts(10)
        kms, ran = solver.profile_attempts(prof_attempts)
        per = [k / prof_attempts for k in kms]
        spmv_ms = per[1] + per[2]
        achieved = (b2 + b3) / (spmv_ms * 1e-3) / 1e9
        roofline = {
            "bound": "hbm", "kernel": "k_spmv (A*xbar + dual step, A'*y + interaction; two launches per iteration)",
            "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
            "traffic": ncu_traffic(args.workload)[0], "traffic_source": ncu_traffic(args.workload)[1],
            "per_kernel_timing": "CUDA events around each kernel of %d real attempts launched one by one (each event "
                                 "record costs ~4 us of gap, so the three add up to more than an iteration of the "
                                 "captured CUDA graph: see in_loop)" % prof_attempts,
            "in_loop": {"iteration_us": 1e6 * basic_s / iters if iters else None,
                        "kernels_us_eager_sum": sum(per[:3]) * 1e3,
                        "iteration_gbs": (b1 + b2 + b3) * iters / basic_s / 1e9 if basic_s > 0 else None,
                        "iteration_frac": (b1 + b2 + b3) * iters / basic_s / 1e9 / peak if basic_s > 0 else None,
                        "what": "take_step alone inside the timed region (CUDA graph of 3 kernels per attempt, "
                                "events around whole batches): algorithmic bytes of an iteration / its time"},
            "per_kernel": {
                "k_primal": {"ms": per[0], "bytes": b1, "gbs": b1 / per[0] / 1e6},
                "k_spmv<EpiDual>": {"ms": per[1], "bytes": b2, "gbs": b2 / per[1] / 1e6},
                "k_spmv<EpiTrans>": {"ms": per[2], "bytes": b3, "gbs": b3 / per[2] / 1e6},
            },
            "iteration": iteration,
            "note": "random 8-byte gathers of xbar / y+ (one L2 tag lookup per nonzero) bound k_spmv on this "
                    "workload before HBM does: see DESIGN.md section 5 and tools/gather_bench.cu",
        }
    else:
        # per-phase device time of k_take_steps on rank 0 (its own phase timers: each phase ends with the grid
        # barrier and the peer exchange that follows it), continuing the same solve
        prof_attempts = 100
        solver.profile_attempts(10)
        kms, ran = solver.profile_attempts(prof_attempts)
        per = [k / prof_attempts for k in kms]
        roofline = {
            "per_kernel": {
                "primal step on the slice + xbar exchange": {"ms": per[0]},
                "A[rows,:]*xbar + dual step + y+ exchange": {"ms": per[1]},
                "A[:,slice]'*y+ + interaction + scalar exchange + step rule": {"ms": per[2]},
                "step rule kernel (kernel-per-phase form only)": {"ms": per[3]},
                "what": "rank 0, %d attempts; phases of the persistent kernel k_take_steps (phase timers) or, with "
                        "FOLP_PERSISTENT=0, CUDA events around the kernels" % prof_attempts},
            "bound": "hbm", "kernel": "whole iteration over all ranks (k_take_steps: primal / A*xbar / A'*y phases of one "
                                      "persistent cooperative kernel per rank; xbar and y+ pushed to every rank from the "
                                      "producing phase over peer memory, flags and four scalars on the grid barriers)",
            "achieved": iteration["gbs_at_value"], "peak": peak * args.gpus, "peak_kind": peak_kind, "unit": "GB/s",
            "frac": iteration["frac_at_value"], "traffic": None, "iteration": iteration,
            "nvlink_bytes_per_iteration_per_gpu": 8 * (n + m) * (args.gpus - 1) // args.gpus,
            "nvlink_gbs_in_per_gpu_at_value": 8 * (n + m) * (args.gpus - 1) / args.gpus * value / 1e9,
        }
    solver.close()

    # ---- e2e: the C-ABI call sequence with host buffers ----
    e2e = None
    if not args.skip_e2e:
        e2e_iters = args.e2e_iters
        fparams.iteration_limit = e2e_iters
        h2d = sum(a.nbytes for a in holder._keep)
        _barrier(world)
        t0 = time.perf_counter()
        s2 = Solver(holder, fparams)
        t1 = time.perf_counter()
        x, y, reason, it2, evals = s2.solve(max_evals=1)
        t2 = time.perf_counter()
        s2.close()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        t_e2e = _max_over_ranks(t3 - t0, world)
        d2h = x.nbytes + y.nbytes
        long_solve = None
        if args.e2e_long_iters > 0:  # the same call sequence on a solve of realistic length (fixed costs amortised)
            fparams.iteration_limit = args.e2e_long_iters
            _barrier(world)
            t0l = time.perf_counter()
            s3 = Solver(holder, fparams)
            t1l = time.perf_counter()
            _, _, _, it3, _ = s3.solve(max_evals=1)
            t2l = time.perf_counter()
            s3.close()
            torch.cuda.synchronize()
            t3l = time.perf_counter()
            t_long = _max_over_ranks(t3l - t0l, world)
            long_solve = {"iterations": it3, "seconds": t_long, "value": it3 / t_long,
                          "seconds_create_solve_destroy": [t1l - t0l, t2l - t1l, t3l - t2l]}
        e2e = {"value": it2 / t_e2e, "unit": "iterations/s", "h2d_bytes_per_step": h2d,
               "long_solve": long_solve,
               "d2h_bytes_per_step": d2h, "iterations": it2, "seconds": t_e2e,
               "seconds_create_solve_destroy": [t1 - t0, t2 - t1, t3 - t2],
               "what": "folp_create(host CSC arrays%s) + folp_solve + folp_get_solution, wall clock, max over ranks"
                       % (", NCCL communicator" if world > 1 else "")}

    # ---- cpu baseline + parity at bench scale: the oracle on a bounded sample of the same workload ----
    cpu = None
    parity = None
    rescale = None
    if not args.skip_cpu:
        def make_gpu_solver(step0, weight0):
            import copy
            gp = copy.copy(fparams)
            gp.iteration_limit = 1_000_000
            gp.initial_step_size = step0
            gp.initial_primal_weight = weight0
            return Solver(holder, gp)

        # N > 1: a shorter oracle run on rank 0 (the other ranks wait for it)
        sample = args.cpu_iters if args.gpus == 1 else min(args.cpu_iters, 40)
        cpu, parity = cpu_baseline(params, lp, scaled, sample_iters=sample, world=world, rank=rank,
                                   make_gpu_solver=make_gpu_solver)
        if args.gpus == 1:
            rescale = rescale_timing(params, lp)
        else:
            cpu = None  # the contract wants the CPU baseline at N = 1 only; parity is kept at every N

    # ---- the north-star target size (1e7 x 1e7, 1e8 nonzeros) on the same N GPUs: GPU arm only. LAST: its 5 GB arena
    # must not sit in the allocator when the end-to-end call is timed (cudaFree right after it took 0.9 s) ----
    target = None
    if not args.skip_target and args.workload == "c2":
        target = target_subrun(world, peak, args.gpus)

    line = {
        "metric": "PDHG iterations/sec", "value": value, "unit": "iterations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_dict(args, n, m, nnz),
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "clocks": clocks.summary(),
        "detail": {"iterations_timed": int(iters), "take_step_seconds": basic_s,
                   "pure_step_iterations_per_s": iters / basic_s if basic_s > 0 else None,
                   "final_relative_l2_primal_residual": e.relative_l2_primal_residual,
                   "final_l2_primal_residual": e.l2_primal_residual,
                   "final_l2_dual_residual": e.l2_dual_residual,
                   "folp_create_seconds": t_create, "rescale_problem": rescale, "build": build_info(),
                   "exchange": exchange, "parity": parity, "target": target,
                   "launch": ("one process driving %d GPUs (folp_create_multi)" % args.gpus) if devices else
                             ("one process per GPU (torchrun)" if world > 1 else "one process, one GPU")},
    }
    if rank == 0:
        emit(line)
    if world > 1:
        import torch.distributed as td
        td.barrier()
        td.destroy_process_group()
    if rank == 0 and parity is not None and not parity["ok"]:
        log("[bench] PARITY FAILED against the CPU oracle:", parity["problems"])
        sys.exit(3)


def target_subrun(world, peak, n_gpus):
    """BASELINE.json's north-star target (synthetic random sparse LP, 1e7 variables, 1e7 constraints,
    1e8 nonzeros) on the same GPUs as the headline line: one warm-up step and two timed steps of 400
    iterations, same parameters, same timing rule (CUDA events on the library's stream, max over
    ranks). Rescaling runs on the device (folp_rescale_problem; bit-identical to the host mirror)."""
    import torch
    import folp_b200
    from folp_b200.lib import Solver
    from folp_b200.synthetic import random_sparse_lp

    t0 = time.time()
    lp = random_sparse_lp(10_000_000, 10_000_000, 10)
    params = folp_b200.PdhgParameters(verbosity=0)
    params.termination_criteria.eps_optimal_absolute = 0.0
    params.termination_criteria.eps_optimal_relative = 0.0
    params.termination_criteria.eps_primal_infeasible = 0.0
    params.termination_criteria.eps_dual_infeasible = 0.0
    holder, fparams, _ = folp_b200.host_setup(params, lp, device_rescaling=True)
    fparams.iteration_limit = 10_000_000
    n, m, nnz = lp.num_variables, lp.num_constraints, lp.constraint_matrix.nnz
    t_host = time.time() - t0
    t0 = time.time()
    solver = Solver(holder, fparams)
    t_create = time.time() - t0
    stream = torch.cuda.ExternalStream(solver.stream())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    run_until(solver, ITERS_PER_STEP)
    c0 = solver.counters()
    _barrier(world)
    ev0.record(stream)
    run_until(solver, 3 * ITERS_PER_STEP)
    ev1.record(stream)
    _barrier(world)
    ms = _max_over_ranks(ev0.elapsed_time(ev1), world)
    c1 = solver.counters()
    solver.close()
    iters = c1["iterations"] - c0["iterations"]
    basic_s = c1["basic_algorithm_seconds"] - c0["basic_algorithm_seconds"]
    value = iters / (ms * 1e-3)
    b = sum(algorithmic_bytes(n, m, nnz))
    log(f"[bench] target sub-run: {value:.1f} it/s on {n_gpus} GPU(s), host {t_host:.1f}s, create {t_create:.2f}s")
    return {"workload": f"synthetic random sparse LP n={n} m={m} nnz={nnz} fp64 (north-star target)",
            "value": value, "unit": "iterations/s", "n_gpus": n_gpus, "iterations_timed": int(iters),
            "pure_step_iterations_per_s": iters / basic_s if basic_s > 0 else None,
            "iteration_gbs_at_value": b * value / 1e9, "iteration_frac_at_value": b * value / 1e9 / (peak * n_gpus),
            "host_generate_rescale_seconds": t_host, "folp_create_seconds": t_create}


def _collect(solver, until, records):
    """folp_run / oracle_run up to iteration `until`, keeping every record."""
    while True:
        e = solver.run()
        records.append(e)
        if e.iteration_number >= until or e.termination_reason != 0:
            return e


def cpu_baseline(params, lp, scaled, sample_iters, world=1, rank=0, make_gpu_solver=None):
    """The CPU leg: the oracle (kind "port": the Julia reference cannot run here), one thread as the
    reference's serial loop, timed from iteration 40 (after the ten per-iteration evaluations of the
    start) -- and, with the records it produced on the way, the PARITY check of the CUDA path at
    bench scale: a fresh GPU solve of the same problem from the same initial scalars is stepped
    through the same evaluations and every folp_eval record is compared by the rule of
    oracle/parity.py (1e-9 relative, widened only by the oracle's own one-ulp sensitivity, which the
    all-cores run below measures). Under torchrun rank 0 runs the oracle while the other ranks wait;
    all ranks then step the partitioned GPU solver together. Returns (cpu_baseline, parity)."""
    from oracle import oracle
    from oracle.parity import compare_eval

    horizon = 40 + sample_iters
    out, rec_o, rec_p = None, [], []
    o_params = None
    if rank == 0:
        holder, o_params, _ = oracle.host_setup(params, lp, scaled)
        o_params.iteration_limit = 1_000_000
        o = oracle.OracleSolver(holder, o_params)
        _collect(o, 40, rec_o)
        t0 = time.perf_counter()
        _collect(o, horizon, rec_o)
        dt = time.perf_counter() - t0
        o.close()
        out = {"value": sample_iters / dt, "unit": "iterations/s", "cores": 1, "kind": "port",
               "sample": f"iterations 40..{horizon} of the same problem and parameters "
                         f"({sample_iters // 40} evaluation/restart blocks included), {dt:.1f}s",
               "host_cores_available": os.cpu_count()}
        # a STRONGER baseline than the (serial) reference, labelled as such: the oracle with its two
        # sparse products on every host core (bit-identical results; vector passes and reductions stay serial).
        # Its initial step size is moved by ONE ulp: the timing does not care, and its records measure how far
        # rounding noise alone moves each field (the sensitivity the parity rule allows for).
        h2, p2, _ = oracle.host_setup(params, lp, scaled)
        p2.iteration_limit = 1_000_000
        p2.initial_step_size = float(np.nextafter(o_params.initial_step_size, np.inf))
        threads = os.cpu_count() or 1
        if oracle.openmp_enabled() and threads > 1:
            oracle.set_threads(threads)
        try:
            o2 = oracle.OracleSolver(h2, p2)
            _collect(o2, 40, rec_p)
            t0 = time.perf_counter()
            _collect(o2, horizon, rec_p)
            dt2 = time.perf_counter() - t0
            o2.close()
        finally:
            oracle.set_threads(1)
        if oracle.openmp_enabled() and threads > 1:
            out["all_cores_variant"] = {"value": sample_iters / dt2, "unit": "iterations/s", "cores": threads,
                                        "what": "oracle with A*x and A'*y on OpenMP threads (same bits), not the "
                                                "reference's behaviour: a stronger baseline"}
    parity = None
    if make_gpu_solver is not None:
        # identical scalar inputs on both sides (rank 0's, broadcast)
        init = [o_params.initial_step_size, o_params.initial_primal_weight] if rank == 0 else [0.0, 0.0]
        if world > 1:
            import torch
            import torch.distributed as td
            t = torch.tensor(init, dtype=torch.float64, device="cuda")
            td.broadcast(t, src=0)
            init = [float(v) for v in t.cpu()]
        g = make_gpu_solver(init[0], init[1])
        rec_g = []
        _collect(g, horizon, rec_g)
        g.close()
        if rank == 0:
            problems, worst, last_restart, restarts, equal = [], 0.0, 0, 0, True
            if len(rec_g) != len(rec_o):
                problems.append(f"{len(rec_g)} GPU records against {len(rec_o)} oracle records")
            for eg, eo, ep in zip(rec_g, rec_o, rec_p):
                pr, w = compare_eval(eg, eo, 1e-9, ep, eo.iteration_number - last_restart)
                problems += pr
                worst = max(worst, w)
                equal = equal and eg.restart_used == eo.restart_used
                if eo.restart_used >= 2:
                    restarts += 1
                    last_restart = eo.iteration_number
            # ("equal" up to the one tie the rule allows: a restart after ONE iteration, where average == current)
            parity = {"ok": not problems, "records": len(rec_o), "iterations": int(rec_o[-1].iteration_number),
                      "max_rel_err": worst, "tolerance": 1e-9,
                      "restart_choices_equal": not any("restart_used" in p_ for p_ in problems),
                      "restart_choices_identical": equal,
                      "restarts": restarts, "n_gpus": world, "problems": problems[:6],
                      "rule": "every floating-point field of every folp_eval record within 1e-9 relative of the "
                              "CPU oracle (objective-like fields against max(|objective|, 1)), widened only by 20x "
                              "the oracle's own sensitivity to a one-ulp change of its initial step size; identical "
                              "restart choices and termination reasons (oracle/parity.py)"}
    return out, parity


def rescale_timing(params, lp):
    """rescale_problem (src/preprocess.jl:631-687), the step before the loop: the
    device path (folp_rescale_problem, host arrays in and out) next to the oracle's C restatement
    (one thread, as the reference). Outside every timed region of the headline metric."""
    from folp_b200.lib import rescale_problem as device_rescale
    from oracle import oracle

    args_ = (params.l_inf_ruiz_iterations, params.l2_norm_rescaling, params.pock_chambolle_alpha, lp)
    device_rescale(*args_)  # warm-up (module load)
    t0 = time.perf_counter()
    g = device_rescale(*args_)
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    o = oracle.rescale_problem(*args_)
    t_cpu = time.perf_counter() - t0
    same = bool(np.array_equal(g.scaled_qp.constraint_matrix.data, o.scaled_qp.constraint_matrix.data)
                and np.array_equal(g.variable_rescaling, o.variable_rescaling)
                and np.array_equal(g.constraint_rescaling, o.constraint_rescaling))
    return {"device_seconds": t_dev, "oracle_seconds_1_thread": t_cpu, "bit_identical": same,
            "what": "ruiz %d + l2 %s + pock-chambolle %s, host arrays in / out" % (
                params.l_inf_ruiz_iterations, params.l2_norm_rescaling, params.pock_chambolle_alpha)}


def bench_reference(args):
    """--impl reference: the CPU oracle on the box's host cores (the reference's
    loop is serial: one thread). Each step = 40 iterations (one evaluation period)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lp, params, holder, fparams, scaled = make_problem(args.workload)
    from oracle import oracle

    n, m, nnz = lp.num_variables, lp.num_constraints, lp.constraint_matrix.nnz
    o_holder, o_params, _ = oracle.host_setup(params, lp, scaled)
    o_params.iteration_limit = 1_000_000
    o = oracle.OracleSolver(o_holder, o_params)
    per = 40
    done = 40
    run_until(o, done)  # the first ten iterations are evaluated one by one (pdhg.jl:894)
    for _ in range(args.warmup):
        done += per
        run_until(o, done)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        done += per
        run_until(o, done)
    dt = time.perf_counter() - t0
    o.close()
    value = per * args.steps / dt
    cpu = {"value": value, "unit": "iterations/s", "cores": 1, "kind": "port",
           "sample": f"{args.steps} steps x {per} iterations of the same problem and parameters, {dt:.1f}s",
           "host_cores_available": os.cpu_count()}
    line = {"impl": "reference", "metric": "PDHG iterations/sec", "value": value, "unit": "iterations/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config_dict(args, n, m, nnz),
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "detail": {"sample_iterations_per_step": per,
                       "note": "a step of this arm is a bounded sample (one evaluation period) of the config's step"}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="folp_b200", choices=["folp_b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-iters", type=int, default=2000)
    ap.add_argument("--e2e-long-iters", type=int, default=10000,
                    help="second end-to-end call of this many iterations (e2e.long_solve; 0 = skip)")
    ap.add_argument("--cpu-iters", type=int, default=80)
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only (ncu)")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs only (ncu)")
    ap.add_argument("--single-process", action="store_true",
                    help="with --gpus N > 1 and NO torchrun: one process drives the N GPUs (folp_create_multi)")
    ap.add_argument("--skip-target", action="store_true",
                    help="skip the 1e7 x 1e7 x 1e8 sub-run that accompanies the default workload (detail.target)")
    args = ap.parse_args()
    _capture_stdout()
    if args.warmup < 3 and args.impl != "reference":
        log("[bench] warning: fewer than 3 warm-up steps")
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_gpu(args)


if __name__ == "__main__":
    main()
