"""Property tests of the oracle (CPU, hypothesis): structural invariants of the CSR and TJDS layouts and the
agreement of the two products on arbitrary small matrices, including empty rows / columns and ragged shapes.
These are the size-independent properties the GPU tests re-use at full size."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

import util
from oracle import oracle


@st.composite
def coo_matrices(draw):
    m = draw(st.integers(1, 40))
    n = draw(st.integers(1, 40))
    nnz = draw(st.integers(0, min(m * n, 120)))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    return m, n, util.random_coo(rng, m, n, nnz), rng.uniform(-2, 2, n)


@settings(max_examples=150, deadline=None)
@given(coo_matrices())
def test_csr_layout_invariants(case):
    m, n, coo, x = case
    row_ptr, col_ind, val = oracle.csr_build(coo, m, n)
    assert row_ptr[0] == 0 and row_ptr[-1] == len(coo)
    assert np.all(np.diff(row_ptr) >= 0)
    assert np.array_equal(np.diff(row_ptr), np.bincount(coo["row"], minlength=m))
    for r in range(m):
        seg = col_ind[row_ptr[r]:row_ptr[r + 1]]
        assert np.all(np.diff(seg) > 0)  # strictly ascending columns inside a row (unique coordinates)
    # the multiset of entries is preserved
    rows = np.repeat(np.arange(m), np.diff(row_ptr))
    got = sorted(zip(rows.tolist(), col_ind.tolist(), val.tolist()))
    want = sorted(zip(coo["row"].tolist(), coo["col"].tolist(), coo["val"].tolist()))
    assert got == want
    # order of arrival does not matter
    rp2, ci2, va2 = oracle.csr_build(coo[::-1].copy(), m, n)
    assert np.array_equal(rp2, row_ptr) and np.array_equal(ci2, col_ind) and np.array_equal(va2, val)


@settings(max_examples=150, deadline=None)
@given(coo_matrices())
def test_tjds_layout_invariants_and_product(case):
    m, n, coo, x = case
    t = oracle.tjds_build(coo, m, n)
    count = np.bincount(coo["col"], minlength=n)
    assert sorted(t.perm.tolist()) == list(range(n))  # a permutation of the columns
    lens = count[t.perm]
    assert np.all(np.diff(lens) <= 0)  # columns by length, descending
    for a, b in zip(range(n - 1), range(1, n)):
        if lens[a] == lens[b]:
            assert t.perm[a] < t.perm[b]  # ties keep column order (txtable_comparator_len, main-cli.c:209-223)
    assert t.ndiag == (int(count.max()) if len(coo) else 0)
    assert t.start_pos[0] == 0 and t.start_pos[-1] == len(coo)
    diag_len = np.diff(t.start_pos)
    assert np.all(np.diff(diag_len) <= 0)  # jagged diagonals shrink
    for d in range(t.ndiag):
        assert diag_len[d] == int((count > d).sum())
    # element k of a diagonal belongs to the column at slot k, and rows ascend inside a column
    for p in range(n):
        c = t.perm[p]
        col_rows = [t.row_ind[t.start_pos[d] + p] for d in range(lens[p])]
        assert col_rows == sorted(coo["row"][coo["col"] == c].tolist())
    # the full TJDS product is the CSR product (pins the x indexing the reference gets wrong, U7)
    y_csr = oracle.csr_mult(*oracle.csr_build(coo, m, n), x)
    assert util.rel_l2(oracle.tjds_mult(t, x), y_csr) <= 1e-13
    # linearity
    x2 = x[::-1].copy()
    lhs = oracle.tjds_mult(t, 2 * x - 3 * x2)
    rhs = 2 * oracle.tjds_mult(t, x) - 3 * oracle.tjds_mult(t, x2)
    assert util.rel_l2(lhs, rhs) <= 1e-12 or np.linalg.norm(rhs) == 0
    # a diagonal limit only removes contributions of later diagonals
    if t.ndiag > 1 and np.all(coo["val"] > -10):
        pos = oracle.make_coo(coo["row"], coo["col"], np.abs(coo["val"]) + 1.0)
        tp = oracle.tjds_build(pos, m, n)
        y_lim = oracle.tjds_mult(tp, np.ones(n), diag_limit=1)
        y_all = oracle.tjds_mult(tp, np.ones(n))
        assert np.all(y_lim <= y_all + 1e-12)


def test_threaded_oracle_loop_is_bit_identical():
    """bench.py's all-cores CPU figure runs the reference loop over nnz-balanced row blocks on several threads; every row
    is still summed left to right by one thread, so the result equals the scalar loop bit for bit -- also with empty
    rows, one giant row, and more threads than rows."""
    rng = np.random.default_rng(21)
    for m, n, nnz, threads in ((5000, 4000, 60000, 7), (3, 50000, 40000, 16), (2000, 10, 5000, 3), (1, 1, 1, 4)):
        rows = rng.integers(0, m, nnz)
        cols = rng.integers(0, n, nnz)
        key = np.unique(rows.astype(np.int64) * n + cols)
        coo = oracle.make_coo(key // n, key % n, rng.uniform(-1, 1, len(key)))
        rp, ci, va = oracle.csr_build(coo, m, n)
        x = rng.uniform(-1, 1, n)
        y = oracle.csr_mult(rp, ci, va, x)
        y_mt, ms = oracle.csr_mult_timed_mt(rp, ci, va, x, 3, threads)
        assert np.array_equal(y.view(np.int64), y_mt.view(np.int64)) and len(ms) == 3 and np.all(ms >= 0)
