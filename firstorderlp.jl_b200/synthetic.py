"""Synthetic LP generators for the BASELINE.json configurations.

The reference ships no random-LP generator; these define the workloads the
benchmarks and parity tests run:

* random_sparse_lp   -- "synthetic random sparse LP" (configs[1], [4] and the
  1e7 x 1e7 x 1e8 target): k i.i.d. uniform column indices per row, N(0,1)
  values, first half of the rows equalities, planted primal/dual optimal pair.
* pagerank_lp        -- restates benchmarking/generate_pagerank_lp.jl:48-73,
  114-128 without JuMP/LightGraphs: Barabasi-Albert graph, damping 0.99, one
  dense equality row sqrt(n)*sum(x) = sqrt(n) and n inequality rows. The graph
  has the reference's distribution, not its exact edges (LightGraphs' RNG
  stream is not reproducible here).
* netlib_shaped_lp   -- staircase / block-angular LP with dense-ish linking
  rows, sized like Netlib instances (the real files need network access:
  benchmarking/collect_netlib_benchmark.sh).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp

from .problem import QuadraticProgrammingProblem


def _lp(A, c, l, u, b, neq, c0=0.0) -> QuadraticProgrammingProblem:
    A = sp.csc_matrix(A, dtype=np.float64)
    A.sort_indices()
    n = A.shape[1]
    return QuadraticProgrammingProblem(
        variable_lower_bound=np.asarray(l, dtype=np.float64),
        variable_upper_bound=np.asarray(u, dtype=np.float64),
        objective_matrix=sp.csc_matrix((n, n), dtype=np.float64),
        objective_vector=np.asarray(c, dtype=np.float64),
        objective_constant=float(c0),
        constraint_matrix=A,
        right_hand_side=np.asarray(b, dtype=np.float64),
        num_equalities=int(neq),
    )


def _plant(A: sp.csr_matrix, neq: int, rng, upper_fraction=0.0):
    """Planted optimum: b and c such that (x*, y*) is primal/dual optimal."""
    m, n = A.shape
    x = rng.uniform(0.0, 1.0, n)
    x[rng.random(n) < 0.5] = 0.0
    y = rng.standard_normal(m)
    ineq = np.arange(m) >= neq
    active = rng.random(m) < 0.5
    y[ineq] = np.abs(y[ineq])
    y[ineq & ~active] = 0.0
    slack = np.zeros(m)
    idle = ineq & ~active
    slack[idle] = rng.uniform(0.1, 1.0, int(idle.sum()))
    b = A @ x - slack
    r = np.zeros(n)
    at_lower = x == 0.0
    r[at_lower] = rng.uniform(0.1, 1.0, int(at_lower.sum()))
    l = np.zeros(n)
    u = np.full(n, np.inf)
    if upper_fraction > 0.0:
        capped = rng.random(n) < upper_fraction
        u[capped] = 10.0
    c = A.T @ y + r
    return x, y, b, c, l, u


def random_sparse_lp(num_variables: int, num_constraints: int, nnz_per_row: int = 10,
                     seed: int = 20260117, upper_fraction: float = 0.0,
                     return_solution: bool = False):
    """Rows with `nnz_per_row` uniform column indices (duplicates merged), N(0,1) values."""
    rng = np.random.default_rng(seed)
    m, n, k = int(num_constraints), int(num_variables), int(nnz_per_row)
    cols = rng.integers(0, n, size=(m, k), dtype=np.int64)
    vals = rng.standard_normal((m, k))
    rows = np.repeat(np.arange(m, dtype=np.int64), k)
    A = sp.csr_matrix((vals.ravel(), (rows, cols.ravel())), shape=(m, n))
    A.sum_duplicates()
    neq = m // 2
    x, y, b, c, l, u = _plant(A, neq, rng, upper_fraction)
    lp = _lp(A, c, l, u, b, neq)
    return (lp, x, y) if return_solution else lp


def random_sparse_qp(num_variables: int, num_constraints: int, nnz_per_row: int = 10,
                     q_rank_rows: int | None = None, q_nnz_per_row: int = 3,
                     seed: int = 20260117, upper_fraction: float = 0.0):
    """random_sparse_lp plus a sparse symmetric positive semidefinite objective matrix
    Q = B'B + diag(d): B has `q_rank_rows` rows of `q_nnz_per_row` N(0,1) entries, d >= 0 is zero on
    half of the variables. The linear term is shifted so that the planted pair of the LP
    stays optimal (c <- c - Q x*): same feasible set, known optimum."""
    lp, x, _ = random_sparse_lp(num_variables, num_constraints, nnz_per_row, seed, upper_fraction,
                                return_solution=True)
    rng = np.random.default_rng(seed + 7)
    n = int(num_variables)
    r = int(q_rank_rows) if q_rank_rows is not None else max(1, n // 4)
    cols = rng.integers(0, n, size=(r, q_nnz_per_row), dtype=np.int64)
    vals = rng.standard_normal((r, q_nnz_per_row))
    rows = np.repeat(np.arange(r, dtype=np.int64), q_nnz_per_row)
    B = sp.csr_matrix((vals.ravel(), (rows, cols.ravel())), shape=(r, n))
    B.sum_duplicates()
    d = rng.uniform(0.0, 1.0, n)
    d[rng.random(n) < 0.5] = 0.0
    Q = sp.csc_matrix(B.T @ B + sp.diags(d))
    Q = sp.csc_matrix((Q + Q.T) * 0.5)  # exactly symmetric
    Q.eliminate_zeros()
    Q.sort_indices()
    lp.objective_matrix = Q
    lp.objective_vector = lp.objective_vector - Q @ x
    return lp


def barabasi_albert_edges(num_nodes: int, k: int, rng) -> np.ndarray:
    """Preferential attachment: each new node links to k distinct earlier nodes
    chosen proportionally to degree (repeated-nodes list construction)."""
    n = int(num_nodes)
    k = max(1, min(int(k), n - 1))
    targets = np.empty(2 * k * n, dtype=np.int64)  # endpoints seen so far
    filled = 0
    src = []
    dst = []
    # seed: star on the first k+1 nodes
    for v in range(1, k + 1):
        src.append(0); dst.append(v)
        targets[filled] = 0; targets[filled + 1] = v
        filled += 2
    for v in range(k + 1, n):
        chosen = set()
        while len(chosen) < k:
            cand = targets[rng.integers(0, filled, size=k - len(chosen))]
            chosen.update(int(t) for t in cand)
        for t in chosen:
            src.append(v); dst.append(t)
            targets[filled] = v; targets[filled + 1] = t
            filled += 2
    return np.stack([np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)], axis=1)


def barabasi_albert_edges_fast(num_nodes: int, k: int, rng) -> np.ndarray:
    """Vectorised batch variant for large n: nodes are attached in growing
    batches against the degree list at the start of the batch (same power-law
    degree distribution; used for the 1e7-node benchmark instance)."""
    n = int(num_nodes)
    k = max(1, min(int(k), n - 1))
    src = [np.zeros(k, dtype=np.int64)]
    dst = [np.arange(1, k + 1, dtype=np.int64)]
    endpoints = [np.zeros(k, dtype=np.int64), np.arange(1, k + 1, dtype=np.int64)]
    done = k + 1
    pool = np.concatenate(endpoints)
    while done < n:
        batch = min(max(done // 2, 1), n - done)
        new = np.arange(done, done + batch, dtype=np.int64)
        t = pool[rng.integers(0, pool.size, size=(batch, k))]
        s = np.repeat(new, k)
        t = t.ravel()
        src.append(s); dst.append(t)
        pool = np.concatenate([pool, s, t])
        done += batch
    e = np.stack([np.concatenate(src), np.concatenate(dst)], axis=1)
    # drop duplicate edges created inside a batch
    lo = np.minimum(e[:, 0], e[:, 1]); hi = np.maximum(e[:, 0], e[:, 1])
    key = np.unique(lo * n + hi)
    return np.stack([key // n, key % n], axis=1)


def pagerank_lp(num_nodes: int, approx_num_edges: int = None, damping_factor: float = 0.99,
                seed: int = 1, fast: bool = None) -> QuadraticProgrammingProblem:
    """benchmarking/generate_pagerank_lp.jl: minimise 0 subject to
    sqrt(n) * sum(x) = sqrt(n)  and  x_i - damping * sum_{j~i} x_j / deg_j >= (1-damping)/n,
    x >= 0 (:48-73); graph = barabasi_albert(n, round(approx_edges / n)) (:114-128)."""
    n = int(num_nodes)
    if approx_num_edges is None:
        approx_num_edges = 3 * n
    k = max(1, int(round(approx_num_edges / n)))
    rng = np.random.default_rng(seed)
    if fast is None:
        fast = n > 200_000
    edges = barabasi_albert_edges_fast(n, k, rng) if fast else barabasi_albert_edges(n, k, rng)
    i = np.concatenate([edges[:, 0], edges[:, 1]])
    j = np.concatenate([edges[:, 1], edges[:, 0]])
    adj = sp.csr_matrix((np.ones(i.size), (i, j)), shape=(n, n))
    adj.sum_duplicates()
    adj.data[:] = 1.0
    deg = np.asarray(adj.sum(axis=0)).ravel()
    deg[deg == 0] = 1.0
    # rows 1..n: I - damping * Adj * diag(1/deg)
    S = sp.identity(n, format="csr") - damping_factor * (adj @ sp.diags(1.0 / deg))
    dense_row = sp.csr_matrix(np.full((1, n), math.sqrt(n)))
    A = sp.vstack([dense_row, S], format="csr")
    b = np.concatenate([[math.sqrt(n)], np.full(n, (1.0 - damping_factor) / n)])
    return _lp(A, np.zeros(n), np.zeros(n), np.full(n, np.inf), b, 1)


def netlib_shaped_lp(num_blocks: int = 8, block_rows: int = 40, block_cols: int = 90,
                     linking_rows: int = 12, density: float = 0.08, seed: int = 7,
                     return_solution: bool = False):
    """Block-angular LP: `num_blocks` sparse diagonal blocks plus `linking_rows`
    dense-ish rows (30 % fill) coupling all columns; planted optimum."""
    rng = np.random.default_rng(seed)
    blocks = [sp.random(block_rows, block_cols, density=density, random_state=rng,
                        data_rvs=rng.standard_normal, format="csr") for _ in range(num_blocks)]
    D = sp.block_diag(blocks, format="csr")
    n = D.shape[1]
    L = sp.random(linking_rows, n, density=0.3, random_state=rng,
                  data_rvs=rng.standard_normal, format="csr")
    A = sp.vstack([L, D], format="csr")
    m = A.shape[0]
    neq = linking_rows + (m - linking_rows) // 3
    x, y, b, c, l, u = _plant(A, neq, rng, upper_fraction=0.1)
    lp = _lp(A, c, l, u, b, neq)
    return (lp, x, y) if return_solution else lp
