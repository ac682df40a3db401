"""Host-side logic of the row-partitioned (multi-GPU) path, on CPU:

* folp_partition (the pure-host arithmetic folp_create applies) covers rows and columns,
  balances nonzeros and is what a NumPy restatement gives;
* a world_size-2 `gloo` emulation of take_step in the library's exchange pattern (DESIGN.md
  section 6) -- primal step on the rank's column slice | allgather xbar | dual step on the rank's
  row block A[rows, :] | allgather y+ (padded, rank-major) | A[:, slice]' y+ and the interaction on
  the slice (full-length rows: no partial sums cross ranks) | rank-ordered exchange of the four
  step-rule scalars -- reproduces the single-process oracle, and both ranks take bit-identical
  decisions.
"""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def _reference_partition(A, world):
    A = sp.csr_matrix(A)
    m, n = A.shape
    cost = np.concatenate([[0], np.cumsum(np.diff(A.indptr) + 2)])
    rb = [0]
    for r in range(1, world):
        target = (int(cost[m]) * r) // world
        i = int(np.searchsorted(cost, target, side="left"))
        rb.append(min(max(i, rb[-1]), m))
    rb.append(m)
    n_pad = -(-n // world)
    n_pad += n_pad & 1
    cb = [min(n, r * n_pad) for r in range(world + 1)]
    return np.array(rb), np.array(cb)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_partition_matches_restatement_and_balances(world):
    from folp_b200.lib import partition
    from folp_b200.synthetic import pagerank_lp, random_sparse_lp

    for lp in (random_sparse_lp(1000, 777, 6, seed=3), pagerank_lp(500)):
        A = lp.constraint_matrix
        rb, cb = partition(A, world)
        rb2, cb2 = _reference_partition(A, world)
        assert np.array_equal(rb, rb2) and np.array_equal(cb, cb2)
        m, n = A.shape
        assert rb[0] == 0 and rb[-1] == m and np.all(np.diff(rb) >= 0)
        assert cb[0] == 0 and cb[-1] == n and np.all(np.diff(cb) >= 0)
        assert np.all(cb[:-1] % 2 == 0)  # slices start on a 16-byte boundary (K1 moves pairs)
        # nonzero balance: no rank exceeds the ideal share by more than its heaviest row
        row_cost = np.diff(sp.csr_matrix(A).indptr) + 2
        share = row_cost.sum() / world
        for r in range(world):
            assert row_cost[rb[r]:rb[r + 1]].sum() <= share + row_cost.max() + 1


def _worker(rank, world, port, out_q):
    import torch
    import torch.distributed as td

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from folp_b200.lib import partition
        from folp_b200.synthetic import random_sparse_lp
        from oracle import oracle
        from shared_problems import generate_pdhg_params

        lp = random_sparse_lp(400, 300, 5, seed=9, upper_fraction=0.2)
        params = generate_pdhg_params(iteration_limit=50, l_inf_ruiz_iterations=3, pock_chambolle_alpha=1.0)
        holder, fparams, scaled = oracle.host_setup(params, lp)
        P = scaled.scaled_qp
        A = sp.csr_matrix(P.constraint_matrix)
        m, n = A.shape
        neq = P.num_equalities
        rb, cb = partition(P.constraint_matrix, world)
        r0, r1, c0, c1 = rb[rank], rb[rank + 1], cb[rank], cb[rank + 1]
        n_pad = int(cb[1] - cb[0]) if world > 1 else n
        m_pad = int(max(rb[q + 1] - rb[q] for q in range(world)))
        A_r = A[r0:r1, :]                                  # row block: A * xbar and the dual step
        At_s = sp.csr_matrix(sp.csc_matrix(A)[:, c0:c1].T)  # column slice: rows of full length m
        # its column indices address the padded rank-major layout of the gathered y+
        owner_off = np.zeros(m, dtype=np.int64)
        for q in range(world):
            owner_off[rb[q]:rb[q + 1]] = q * m_pad + np.arange(rb[q + 1] - rb[q])
        At_s = sp.csr_matrix((At_s.data, owner_off[At_s.indices], At_s.indptr), shape=(c1 - c0, world * m_pad))
        c, l, u, b = (np.asarray(v, dtype=np.float64) for v in (
            P.objective_vector, P.variable_lower_bound, P.variable_upper_bound, P.right_hand_side))
        rng = np.random.default_rng(1)
        x0 = np.abs(rng.standard_normal(n))
        y0 = rng.standard_normal(m)
        y0[neq:] = np.abs(y0[neq:])
        step, pw = 0.05, 1.3
        o = oracle.OracleSolver(holder, fparams)
        o.debug_set_state(x0, y0, step_size=step, primal_weight=pw)

        def gather(local, pad):
            send = torch.zeros(pad, dtype=torch.float64)
            send[: len(local)] = torch.from_numpy(np.ascontiguousarray(local))
            recv = [torch.zeros(pad, dtype=torch.float64) for _ in range(world)]
            td.all_gather(recv, send)
            return torch.cat(recv).numpy()

        # local state: primal slice + row block
        x, y = x0[c0:c1].copy(), y0[r0:r1].copy()
        aty = At_s @ gather(y, m_pad)
        total_iterations = 0
        decisions = []
        for _ in range(4):
            # K1 on the slice
            xn = np.minimum(u[c0:c1], np.maximum(l[c0:c1], x - (step / pw) * (c[c0:c1] - aty)))
            xbar_loc = xn + 1.0 * (xn - x)
            dx2 = float(np.sum((xn - x) ** 2))
            xbar = gather(xbar_loc, n_pad)[:n]           # exchange 1: allgather xbar
            # K2 on the row block
            yn = y + (pw * step) * (b[r0:r1] - A_r @ xbar)
            ineq = np.arange(r0, r1) >= neq
            yn[ineq] = np.maximum(yn[ineq], 0.0)
            dy2 = float(np.sum((yn - y) ** 2))
            y_full = gather(yn, m_pad)                   # exchange 2: allgather y+
            # K3 on the slice: full-length rows, no reduction across ranks
            atn = At_s @ y_full
            inter = float(np.sum((xn - x) * (atn - aty)))
            dp2 = float(np.sum((atn - aty) ** 2))
            # exchange 3: the four scalars, summed in rank order on every rank
            sc = [torch.zeros(4, dtype=torch.float64) for _ in range(world)]
            td.all_gather(sc, torch.tensor([dx2, dy2, inter, dp2], dtype=torch.float64))
            tot = np.zeros(4)
            for t in sc:
                tot += t.numpy()
            # the scalar rule (pdhg.jl:689-728), identical on every rank
            total_iterations += 1
            movement = 0.5 * pw * tot[0] + (0.5 / pw) * tot[1]
            interaction = abs(tot[2])
            limit = movement / interaction if interaction > 0 else np.inf
            accepted = step <= limit
            k1 = total_iterations + 1
            nxt = min((1 - k1 ** -0.3) * limit, (1 + k1 ** -0.6) * step)
            decisions.append((accepted, nxt))
            if accepted:
                x, y, aty = xn, yn, atn
            step = nxt
            # the oracle, one attempt
            o.debug_attempts(1)
            so = o.debug_state()
            x_full = gather(x, n_pad)[:n]
            assert np.max(np.abs(x_full - so["x"])) <= 1e-13 * max(1.0, np.max(np.abs(so["x"])))
            assert np.max(np.abs(y - so["y"][r0:r1])) <= 1e-13 * max(1.0, np.max(np.abs(so["y"])))
            assert np.max(np.abs(aty - so["dual_product"][c0:c1])) <= 1e-13 * max(1.0, np.max(np.abs(so["dual_product"])))
            assert abs(step - so["step_size"]) <= 1e-12 * so["step_size"]
            assert so["total_number_iterations"] == total_iterations
        o.close()
        # both ranks took bit-identical decisions
        mine = torch.tensor([float(a) for a, _ in decisions] + [s for _, s in decisions], dtype=torch.float64)
        both = [torch.zeros_like(mine) for _ in range(world)]
        td.all_gather(both, mine)
        assert all(torch.equal(b_, both[0]) for b_ in both)
        out_q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        out_q.put((rank, "FAILED: " + traceback.format_exc()))
        raise
    finally:
        td.destroy_process_group()


def test_partitioned_attempt_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sock:  # a free port: a fixed one may still be in TIME_WAIT from the last run
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
