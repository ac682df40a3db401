#!/usr/bin/env python
"""Development probe: per-kernel device time of real take_step attempts for one or more
builds of libfolp_b200.so on one synthetic workload.

  python tools/probe_kernels.py [--workload c2] [--attempts 200] lib1.so [lib2.so:ENV=VALUE ...]

A library may be followed by `:NAME=VALUE[,NAME=VALUE]`: environment variables set while that
handle is created (e.g. FOLP_NO_ROW_SORT=1), so that set-up variants are compared in one process.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="*")
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--attempts", type=int, default=200)
    ap.add_argument("--iters", type=int, default=0, help="also time folp_run over this many iterations")
    args = ap.parse_args()
    import bench
    import folp_b200
    from folp_b200 import distributed, lib as L

    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        distributed.init("nccl")
    rank = int(os.environ.get("RANK", "0"))

    lp, params, holder, fparams, scaled = bench.make_problem(args.workload)
    n, m, nnz = lp.num_variables, lp.num_constraints, lp.constraint_matrix.nnz
    b1, b2, b3 = bench.algorithmic_bytes(n, m, nnz)
    fparams.iteration_limit = 10_000_000
    for spec in (args.libs or [L.LIB_PATH]):
        path, _, envs = spec.partition(":")
        env = dict(kv.split("=", 1) for kv in envs.split(",") if kv)
        L._LIB = None
        L.LIB_PATH = os.path.abspath(path)
        os.environ.update(env)
        s = L.Solver(holder, fparams)
        for k in env:
            os.environ.pop(k, None)
        bench.run_until(s, 120)
        s.profile_attempts(20)
        kms, ran = s.profile_attempts(args.attempts)
        per = [k / args.attempts for k in kms]
        if distributed.state() is not None:
            out = {"rank": rank, "env": env, "exchange": s.shard_info()["exchange"],
                   "us": {k: round(v * 1e3, 2) for k, v in zip(("primal", "dual", "trans", "finalize"), per)},
                   "iter_us": sum(per) * 1e3}
            if args.iters:
                import torch
                import torch.distributed as td
                c0 = s.counters()
                td.barrier(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                bench.run_until(s, c0["iterations"] + args.iters)
                torch.cuda.synchronize(); td.barrier()
                dt = time.perf_counter() - t0
                c1 = s.counters()
                out["run_it_per_s"] = (c1["iterations"] - c0["iterations"]) / dt
                out["pure_step_it_per_s"] = (c1["iterations"] - c0["iterations"]) / (
                    c1["basic_algorithm_seconds"] - c0["basic_algorithm_seconds"])
            s.close()
            print(json.dumps(out), flush=True)
            continue
        out = {"lib": os.path.basename(path), "env": env, "info": L.build_info(), "k_primal_us": per[0] * 1e3,
               "k_dual_us": per[1] * 1e3, "k_trans_us": per[2] * 1e3,
               "gbs": [b1 / per[0] / 1e6, b2 / per[1] / 1e6, b3 / per[2] / 1e6],
               "iter_us": sum(per) * 1e3, "iter_gbs": (b1 + b2 + b3) / sum(per) / 1e6}
        out["plain_A_us"] = s.time_spmv(False, 50) * 1e3
        out["plain_At_us"] = s.time_spmv(True, 50) * 1e3
        if args.iters:
            import torch
            c0 = s.counters()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            bench.run_until(s, c0["iterations"] + args.iters)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            c1 = s.counters()
            out["run_it_per_s"] = (c1["iterations"] - c0["iterations"]) / dt
            out["pure_step_it_per_s"] = (c1["iterations"] - c0["iterations"]) / (
                c1["basic_algorithm_seconds"] - c0["basic_algorithm_seconds"])
        s.close()
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
