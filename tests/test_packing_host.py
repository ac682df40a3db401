"""The library's matrix packing (position-major 32-row groups, length-sorted
windows, wide rows, long-row chunks) checked on the HOST: folp_debug_host_spmv
packs exactly as folp_create does and walks the layout with the kernel's slot
arithmetic. Rows of <= 32 nonzeros must reproduce the serial ascending-column
sum bit for bit (the summation order of the reference's stdlib kernels,
SparseArrays `*`)."""
import numpy as np
import pytest
import scipy.sparse as sp

import folp_b200
from folp_b200.lib import FolpError, host_packed_spmv, host_problem_spmv
from folp_b200.synthetic import netlib_shaped_lp, pagerank_lp, random_sparse_lp
from shared_problems import generate_pdhg_params


def _chunk_nnz():
    from folp_b200.lib import build_info
    return int(build_info().split("chunk_nnz=")[1].split(";")[0])


def _serial(A, x):
    A = sp.csr_matrix(A)
    A.sort_indices()
    y = np.zeros(A.shape[0])
    for i in range(A.shape[0]):
        s = 0.0
        for k in range(A.indptr[i], A.indptr[i + 1]):
            s += A.data[k] * x[A.indices[k]]
        y[i] = s
    return y


def _ragged(seed, rows=3000, n=5000):
    rng = np.random.default_rng(seed)
    lens = np.concatenate([np.zeros(9, int), rng.poisson(10, rows), np.full(5, 100),
                           np.array([9000, 4097, 4096, 1025, 1024, 200, 129, 128, 127, 33, 32, 31])])
    lens = np.minimum(lens, n)
    rng.shuffle(lens)
    ind, ptr, val = [], [0], []
    for k in lens:
        ind += list(np.sort(rng.choice(n, size=int(k), replace=False)))
        val += list(rng.standard_normal(int(k)))
        ptr.append(len(ind))
    return sp.csr_matrix((val, ind, ptr), shape=(len(lens), n))


@pytest.mark.parametrize("make", [
    lambda: random_sparse_lp(4000, 3000, 10, seed=5).constraint_matrix.tocsr(),       # rows ~10: identity
    lambda: random_sparse_lp(4000, 3000, 10, seed=5).constraint_matrix.T.tocsr(),     # Poisson columns: sorted
    lambda: pagerank_lp(3000).constraint_matrix.T.tocsr(),                            # power-law + dense row
    lambda: pagerank_lp(3000).constraint_matrix.tocsr(),
    lambda: _ragged(1), lambda: _ragged(2, rows=257), lambda: sp.csr_matrix((5, 7)), lambda: _heavy_tailed(3, 3000, 4000),
    lambda: sp.csr_matrix(np.ones((1, 40))),
])
def test_packed_layout_reproduces_serial_row_sums(make):
    A = make()
    x = np.random.default_rng(0).standard_normal(A.shape[1])
    y, stats = host_packed_spmv(A, x)
    ref = _serial(A, x)
    row_len = np.diff(A.indptr)
    narrow = row_len <= 32
    assert np.array_equal(y[narrow], ref[narrow])          # bit-identical: same summation order
    scale = max(1.0, np.max(np.abs(ref)) if ref.size else 1.0)
    assert np.max(np.abs(y - ref), initial=0.0) <= 1e-13 * scale
    assert stats["long_rows"] == int((row_len > _chunk_nnz()).sum())


def _heavy_tailed(seed=7, rows=20000, n=20000):  # noqa: E302
    """Most rows have 2 entries, one in ten has 30: in place every group of 32 runs 30 positions."""
    rng = np.random.default_rng(seed)
    lens = np.where(rng.random(rows) < 0.1, 30, 2)
    ind = np.concatenate([np.sort(rng.choice(n, size=int(k), replace=False)) for k in lens])
    ptr = np.concatenate([[0], np.cumsum(lens)])
    return sp.csr_matrix((rng.standard_normal(ind.size), ind, ptr), shape=(rows, n))


def test_length_sorted_windows_on_heavy_tailed_rows(monkeypatch):
    M = _heavy_tailed()
    x = np.random.default_rng(1).standard_normal(M.shape[1])
    y1, s1 = host_packed_spmv(M, x)
    monkeypatch.setenv("FOLP_NO_ROW_SORT", "1")
    y0, s0 = host_packed_spmv(M, x)
    monkeypatch.delenv("FOLP_NO_ROW_SORT")
    assert np.array_equal(y0, y1) and np.array_equal(y1, _serial(M, x))
    assert s0["sorted_groups"] == 0 and s1["sorted_groups"] > 0
    assert s1["narrow_rounds"] <= 0.5 * s0["narrow_rounds"]
    # ... and the groups are dealt to the warps in rotation: the busiest warp gains as well
    # (a grid of 64 warps here, so that every warp makes several trips)
    _, r1 = host_packed_spmv(M, x, warps_total=64)
    monkeypatch.setenv("FOLP_NO_ROW_SORT", "1")
    _, r0 = host_packed_spmv(M, x, warps_total=64)
    monkeypatch.delenv("FOLP_NO_ROW_SORT")
    assert r1["busiest_warp_rounds"] <= 0.6 * r0["busiest_warp_rounds"]
    assert r1["busiest_warp_rounds"] <= 2.0 * r1["narrow_rounds"] / 64
    # mild imbalance (Poisson(10) columns of a random matrix, rows of ~10) stays in place: measured
    # on the B200 the scattered epilogue costs more than the saved rounds
    lp = random_sparse_lp(20000, 20000, 10, seed=7)
    for Mx in (lp.constraint_matrix.tocsr(), lp.constraint_matrix.T.tocsr()):
        _, s = host_packed_spmv(Mx, np.ones(Mx.shape[1]))
        assert s["sorted_groups"] == 0


@pytest.mark.parametrize("make", [lambda: random_sparse_lp(30000, 20000, 10, seed=9),
                                  lambda: pagerank_lp(5000), lambda: netlib_shaped_lp(seed=3),
                                  lambda: random_sparse_lp(300000, 150000, 4, seed=10)])
def test_host_preparation_of_a_problem(make):
    """folp_create's host half end to end (parallel two-level transposition of the caller's CSC,
    planning, packing of A and A') against the serial row sums."""
    lp = make()
    params = generate_pdhg_params(l_inf_ruiz_iterations=2)
    holder, _, scaled = folp_b200.host_setup(params, lp)
    A = scaled.scaled_qp.constraint_matrix
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal(A.shape[1]), rng.standard_normal(A.shape[0])
    ax = host_problem_spmv(holder, x)
    aty = host_problem_spmv(holder, y, transpose=True)
    Ar, Atr = A.tocsr(), A.T.tocsr()
    if Ar.shape[0] * 10 < 400000:
        ref_ax, ref_aty = _serial(Ar, x), _serial(Atr, y)
        narrow_r, narrow_c = np.diff(Ar.indptr) <= 32, np.diff(Atr.indptr) <= 32
        assert np.array_equal(ax[narrow_r], ref_ax[narrow_r])
        assert np.array_equal(aty[narrow_c], ref_aty[narrow_c])
    else:
        ref_ax, ref_aty = Ar @ x, Atr @ y
    assert np.max(np.abs(ax - ref_ax)) <= 1e-12 * max(1.0, np.max(np.abs(ref_ax)))
    assert np.max(np.abs(aty - ref_aty)) <= 1e-12 * max(1.0, np.max(np.abs(ref_aty)))


def test_host_preparation_rejects_bad_row_index():
    lp = random_sparse_lp(50, 40, 3, seed=2)
    holder, _, _ = folp_b200.host_setup(generate_pdhg_params(), lp)
    rowval = next(a for a in holder._keep if a.dtype == np.int64 and a.size == lp.constraint_matrix.nnz)
    rowval[5] = 40  # == m: out of range
    with pytest.raises(FolpError):
        host_problem_spmv(holder, np.ones(50))


def test_cost_balanced_static_order_is_a_pure_reordering(monkeypatch):
    """The cost-balanced static order (default; FOLP_NO_BALANCE_TILES=1 keeps the matrix order): the
    work items are dealt to the warps by decreasing cost and laid out for the kernel's static
    striding. Same row sums bit for bit; the busiest warp carries less."""
    for M in (random_sparse_lp(60000, 60000, 10, seed=7).constraint_matrix.T.tocsr(),
              pagerank_lp(20000).constraint_matrix.tocsr(), _ragged(5)):
        x = np.random.default_rng(2).standard_normal(M.shape[1])
        monkeypatch.setenv("FOLP_NO_BALANCE_TILES", "1")
        y0, s0 = host_packed_spmv(M, x, warps_total=256)
        monkeypatch.delenv("FOLP_NO_BALANCE_TILES")
        y1, s1 = host_packed_spmv(M, x, warps_total=256)
        assert np.array_equal(y0, y1)
        assert s1["busiest_warp_rounds"] <= s0["busiest_warp_rounds"]
        assert s1["narrow_rounds"] == s0["narrow_rounds"]
        assert s1["tiles"] % 256 == 0 if s0["tiles"] > 256 else s1["tiles"] == s0["tiles"]
