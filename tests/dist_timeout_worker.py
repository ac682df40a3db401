"""Peer-exchange time-out of the partitioned mode (the wait used to give up after a
hard-coded ~2 s). Launch with torchrun, world size 2:

  phase 1  rank 1 stalls for 3 s on the host between two batches: with the default time-out (30 s)
           the solve just continues, and both ranks return identical records;
  phase 2  FOLP_P2P_TIMEOUT_MS=400 and the same stall: rank 0's kernel gives up, the whole grid
           leaves the batch early (no attempt runs on half-delivered vectors) and folp_run returns
           an error on both ranks instead of hanging the device.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    import torch
    import torch.distributed as td
    import folp_b200
    from folp_b200 import distributed
    from folp_b200.lib import FolpError, Solver
    from folp_b200.synthetic import random_sparse_lp

    st = distributed.init("nccl")
    assert st is not None and st["world_size"] == 2
    rank = st["rank"]
    problem = random_sparse_lp(3000, 2500, 8, seed=5)
    params = folp_b200.PdhgParameters(verbosity=0)
    params.termination_criteria.iteration_limit = 400
    params.termination_criteria.eps_optimal_absolute = 0.0
    params.termination_criteria.eps_optimal_relative = 0.0
    holder, fparams, _ = folp_b200.host_setup(params, problem)

    # ---- phase 1: a 3 s stall is not an error ----
    os.environ.pop("FOLP_P2P_TIMEOUT_MS", None)
    s = Solver(holder, fparams)
    for _ in range(11):
        s.run()
    if rank == 1:
        time.sleep(3.0)
    t0 = time.time()
    e = s.run()
    waited = time.time() - t0
    digest = torch.tensor([e.primal_objective, e.l2_primal_residual, float(e.iteration_number)],
                          dtype=torch.float64, device="cuda")
    both = [torch.zeros_like(digest) for _ in range(2)]
    td.all_gather(both, digest)
    assert torch.equal(both[0], both[1]), both
    assert e.iteration_number == 80
    if rank == 0:
        assert waited > 2.0, waited  # it really waited for the stalled peer
        print("[timeout] phase 1: stall tolerated, records identical: ok", flush=True)
    s.close()
    td.barrier()

    # ---- phase 2: a short time-out surfaces as an error on both ranks ----
    os.environ["FOLP_P2P_TIMEOUT_MS"] = "400"
    s = Solver(holder, fparams)
    for _ in range(11):
        s.run()
    td.barrier()
    if rank == 1:
        time.sleep(3.0)
    t0 = time.time()
    try:
        s.run()
        raised = False
    except FolpError as err:
        raised = "timed out" in str(err)
    took = time.time() - t0
    assert raised, "folp_run returned without reporting the lost peer"
    assert took < 2.5, took  # well below the stall: the wait gave up, it did not outlast the peer
    s.close()
    print(f"[timeout] rank {rank} phase 2: error after {took:.2f} s: TIMEOUT OK", flush=True)
    td.barrier()
    td.destroy_process_group()


if __name__ == "__main__":
    main()
