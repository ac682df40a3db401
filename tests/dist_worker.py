"""Row-partitioned (multi-GPU) parity worker. Launch:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P tests/dist_worker.py          # one process per GPU (folp_dist)
  python tests/dist_worker.py --single-process N     # one process driving N GPUs (folp_create_multi)

Every rank runs the same GPU-vs-oracle checks as tests/test_gpu_parity.py, with the GPU
solver spread over all ranks (folp_b200.distributed makes every Solver join the ranks; in the
single-process form FOLP_DEVICES makes every Solver a folp_create_multi handle)."""
import os
import sys
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402


def parity_checks(check, world, rank):
    """The partitioned-mode check list (shared by both launch forms)."""
    import folp_b200
    from folp_b200 import RestartScheme
    from folp_b200.synthetic import random_sparse_lp
    import test_gpu_parity as T
    from shared_problems import generate_pdhg_params

    # the shard layout is the exported partition
    problem = random_sparse_lp(3000, 2500, 8, seed=5)
    params = generate_pdhg_params(iteration_limit=10)
    o, g = T._pair(problem, params)
    info = g.shard_info()
    rb, cb = folp_b200.lib.partition(problem.constraint_matrix, world)
    assert (info["row_begin"], info["row_end"]) == (rb[rank], rb[rank + 1]), (info, rb)
    assert (info["col_begin"], info["col_end"]) == (cb[rank], cb[rank + 1]), (info, cb)
    # peer memory; behind an NVSwitch that offers it, bound to a multicast object (FOLP_EXPECT_EXCHANGE pins one)
    assert info["exchange"] in ("peer", "multicast"), info
    assert info["exchange"] == os.environ.get("FOLP_EXPECT_EXCHANGE", info["exchange"]), info
    if rank == 0:
        print(f"[x{world}] exchange mode: {info['exchange']}", flush=True)
    o.close(); g.close()

    quick = os.environ.get("FOLP_DIST_QUICK") is not None  # a short list (expensive boxes): one check per stage
    for i, make in enumerate([lambda: random_sparse_lp(3000, 2500, 8, seed=5), T.netlib_shaped_lp,
                              T._ragged_lp, lambda: T.pagerank_lp(3000)]):
        if quick and i not in (0, 3):
            continue
        check(f"spmv[{i}]", lambda make=make: T.test_spmv_matches_oracle(make))
    if quick:
        check("single_attempt[ragged]", lambda: T.test_single_attempt_parity(T._ragged_lp))
        check("eval_records[ADAPTIVE_NORMALIZED]",
              lambda: T.test_eval_records_match_oracle(RestartScheme.ADAPTIVE_NORMALIZED))
        check("full_solve[random]", lambda: T.test_full_solve_matches_oracle(
            lambda: random_sparse_lp(2000, 1500, 8, seed=41), 1e-6))
        check("example_lp", T.test_example_lp_exact_record)
        return
    check("single_attempt[random]", lambda: T.test_single_attempt_parity(
        lambda: random_sparse_lp(4000, 3000, 10, seed=11)))
    check("single_attempt[ragged]", lambda: T.test_single_attempt_parity(T._ragged_lp))
    check("trajectory_no_restarts", T.test_trajectory_parity_no_restarts)
    for scheme in (RestartScheme.NO_RESTARTS, RestartScheme.ADAPTIVE_NORMALIZED,
                   RestartScheme.FIXED_FREQUENCY, RestartScheme.ADAPTIVE_LOCALIZED,
                   RestartScheme.ADAPTIVE_DISTANCE):
        check(f"eval_records[{scheme.name}]", lambda scheme=scheme: T.test_eval_records_match_oracle(scheme))
    check("full_solve[random]", lambda: T.test_full_solve_matches_oracle(
        lambda: random_sparse_lp(2000, 1500, 8, seed=41), 1e-6))
    check("full_solve[pagerank]", lambda: T.test_full_solve_matches_oracle(lambda: T.pagerank_lp(2000), 1e-8))
    check("short_horizon", T.test_full_solve_trajectory_short_horizon)
    check("deterministic", T.test_deterministic_run_to_run)
    check("example_lp", T.test_example_lp_exact_record)  # 3 rows on N ranks: empty shards


def main_single_process(world):
    """One process, `world` GPUs: every Solver is a folp_create_multi handle."""
    os.environ["FOLP_DEVICES"] = ",".join(str(d) for d in range(world))

    def check(name, fn):
        try:
            fn()
            print(f"[multi x{world}] {name}: ok", flush=True)
        except Exception:
            print(f"[multi x{world}] {name}: FAILED\n{traceback.format_exc()}", flush=True)
            raise

    parity_checks(check, world, 0)
    import folp_b200
    from folp_b200.synthetic import random_sparse_lp
    problem = random_sparse_lp(2500, 2000, 7, seed=71)
    params = folp_b200.PdhgParameters(verbosity=0)
    params.termination_criteria.iteration_limit = 120
    out = folp_b200.optimize(params, problem)
    assert out.iteration_count == 120
    print(f"[multi x{world}] ALL OK", flush=True)


def main():
    import torch
    import folp_b200
    from folp_b200 import distributed
    from folp_b200.synthetic import random_sparse_lp

    st = distributed.init("nccl")
    assert st is not None, "launch with torchrun, world_size >= 2"
    rank, world = st["rank"], st["world_size"]

    failures = []

    def check(name, fn):
        try:
            fn()
            if rank == 0:
                print(f"[dist x{world}] {name}: ok", flush=True)
        except Exception:  # every rank keeps the same call sequence as long as all fail alike
            failures.append(name)
            print(f"[dist x{world}] rank {rank} {name}: FAILED\n{traceback.format_exc()}", flush=True)
            raise

    parity_checks(check, world, rank)

    # every rank holds the same global records and solution
    import torch.distributed as td
    problem = random_sparse_lp(2500, 2000, 7, seed=71)
    params = folp_b200.PdhgParameters(verbosity=0)
    params.termination_criteria.iteration_limit = 120
    out = folp_b200.optimize(params, problem)
    digest = torch.tensor([float(np.sum(out.primal_solution)), float(np.sum(out.dual_solution)),
                           out.iteration_stats[-1].convergence_information[0].primal_objective],
                          dtype=torch.float64, device="cuda")
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    td.all_gather(all_d, digest)
    for d in all_d:
        assert torch.equal(d, all_d[0]), all_d
    if rank == 0:
        print(f"[dist x{world}] all ranks agree; ALL OK", flush=True)
    td.barrier()
    td.destroy_process_group()


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--single-process":
        main_single_process(int(sys.argv[2]))
    else:
        main()
