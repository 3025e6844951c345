"""ctypes front end of oracle/libsmvp_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module.  The product package never does (tests/test_no_oracle_in_product.py
enforces it).  See oracle/smvp_oracle.c for the cited restatement.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libsmvp_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsmvp_ref.so")

COO_DT = np.dtype([("row", "<i4"), ("col", "<i4"), ("val", "<f8")])

_lib = None


def build(ref=True):
    """Compile the C restatement (always) and oracle/_ref (when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build(ref=False)
        L = ctypes.CDLL(SO)
        vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
        L.oracle_csr_build.argtypes = [vp, i32, i32, i64, vp, vp, vp]
        L.oracle_csr_build.restype = ctypes.c_int
        L.oracle_csr_mult.argtypes = [i32, vp, vp, vp, vp, vp]
        L.oracle_csr_mult.restype = None
        L.oracle_csr_mult_timed.argtypes = [i32, vp, vp, vp, vp, vp, ctypes.c_int, vp]
        L.oracle_csr_mult_timed.restype = None
        L.oracle_tjds_build.argtypes = [vp, i32, i32, i64, vp, vp, vp, vp, vp, vp]
        L.oracle_tjds_build.restype = ctypes.c_int
        L.oracle_tjds_mult.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp, i32]
        L.oracle_tjds_mult.restype = None
        L.oracle_tjds_mult_ref_compat.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp]
        L.oracle_tjds_mult_ref_compat.restype = None
        L.oracle_tjds_mult_timed.argtypes = [i32, i32, vp, vp, vp, vp, vp, ctypes.c_int, vp]
        L.oracle_tjds_mult_timed.restype = None
        L.oracle_time_stats.argtypes = [vp, ctypes.c_int, vp]
        L.oracle_time_stats.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def make_coo(row, col, val):
    coo = np.zeros(len(row), dtype=COO_DT)
    coo["row"], coo["col"], coo["val"] = row, col, val
    return coo


def csr_build(coo, rows, cols):
    coo = np.ascontiguousarray(coo, dtype=COO_DT)
    nnz = len(coo)
    row_ptr = np.zeros(rows + 1, np.int32)
    col_ind = np.zeros(nnz, np.int32)
    val = np.zeros(nnz, np.float64)
    rc = lib().oracle_csr_build(_p(coo), rows, cols, nnz, _p(row_ptr), _p(col_ind), _p(val))
    if rc != 0:
        raise ValueError("oracle_csr_build failed (coordinates out of range?)")
    return row_ptr, col_ind, val


def csr_mult(row_ptr, col_ind, val, x):
    rows = len(row_ptr) - 1
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros(rows, np.float64)
    lib().oracle_csr_mult(rows, _p(row_ptr), _p(col_ind), _p(val), _p(x), _p(y))
    return y


def csr_mult_timed_mt(row_ptr, col_ind, val, x, iters, nthreads):
    """The reference's CSR loop over nnz-balanced row blocks on `nthreads` host threads (not what the reference does:
    it is single-threaded).  y is bit-identical to csr_mult."""
    rows = len(row_ptr) - 1
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros(rows, np.float64)
    ms = np.zeros(iters, np.float64)
    f = lib().oracle_csr_mult_timed_mt
    f.argtypes = [ctypes.c_int32] + [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    f.restype = ctypes.c_int
    if f(rows, _p(row_ptr), _p(col_ind), _p(val), _p(x), _p(y), iters, _p(ms), nthreads) != 0:
        raise RuntimeError("oracle_csr_mult_timed_mt could not start its threads")
    return y, ms


def csr_mult_timed(row_ptr, col_ind, val, x, iters):
    rows = len(row_ptr) - 1
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros(rows, np.float64)
    ms = np.zeros(iters, np.float64)
    lib().oracle_csr_mult_timed(rows, _p(row_ptr), _p(col_ind), _p(val), _p(x), _p(y), iters, _p(ms))
    return y, ms


class Tjds:
    """perm[cols], start_pos[ndiag+1], row_ind[nnz], val[nnz]; ref_limit = diagonals the shipped loop walks."""

    def __init__(self, rows, cols, perm, start_pos, ndiag, ref_limit, row_ind, val):
        self.rows, self.cols, self.perm, self.start_pos = rows, cols, perm, start_pos
        self.ndiag, self.ref_limit, self.row_ind, self.val = ndiag, ref_limit, row_ind, val


def tjds_build(coo, rows, cols):
    coo = np.ascontiguousarray(coo, dtype=COO_DT)
    nnz = len(coo)
    perm = np.zeros(cols, np.int32)
    start_pos = np.zeros(rows + 2, np.int32)
    row_ind = np.zeros(nnz, np.int32)
    val = np.zeros(nnz, np.float64)
    ndiag = ctypes.c_int32(0)
    ref_limit = ctypes.c_int32(0)
    rc = lib().oracle_tjds_build(_p(coo), rows, cols, nnz, _p(perm), _p(start_pos), ctypes.byref(ndiag),
                                 ctypes.byref(ref_limit), _p(row_ind), _p(val))
    if rc != 0:
        raise ValueError("oracle_tjds_build failed")
    return Tjds(rows, cols, perm, start_pos[: ndiag.value + 1].copy(), ndiag.value, ref_limit.value, row_ind, val)


def tjds_mult(t, x, diag_limit=0):
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros(t.rows, np.float64)
    lib().oracle_tjds_mult(t.rows, t.cols, t.ndiag, _p(t.perm), _p(t.start_pos), _p(t.row_ind), _p(t.val), _p(x),
                           _p(y), diag_limit)
    return y


def tjds_mult_ref_compat(t, x_by_row=None):
    x = np.ones(t.rows, np.float64) if x_by_row is None else np.ascontiguousarray(x_by_row, np.float64)
    y = np.zeros(t.rows, np.float64)
    lib().oracle_tjds_mult_ref_compat(t.rows, t.ndiag, t.ref_limit, _p(t.start_pos), _p(t.row_ind), _p(t.val),
                                      _p(x), _p(y))
    return y


def tjds_mult_timed(t, x, iters):
    xp = np.ascontiguousarray(np.asarray(x, np.float64)[t.perm])
    y = np.zeros(t.rows, np.float64)
    ms = np.zeros(iters, np.float64)
    lib().oracle_tjds_mult_timed(t.rows, t.ndiag, _p(t.start_pos), _p(t.row_ind), _p(t.val), _p(xp), _p(y), iters,
                                 _p(ms))
    return y, ms


def time_stats(ms_each):
    ms = np.ascontiguousarray(ms_each, np.float64)
    out = np.zeros(5, np.float64)
    lib().oracle_time_stats(_p(ms), len(ms), _p(out))
    return dict(total=out[0], avg=out[1], stdev=out[2], min=out[3], max=out[4])


# ---- the UNMODIFIED reference functions (oracle/_ref/libsmvp_ref.so), CSR only: used as the
# ---- `kind: "reference"` CPU baseline.  TJDS is not offered here: the reference's TJDS build is
# ---- O(nnz*N) (main-cli.c:894-904) and its LUT dump reads out of bounds (main-cli.c:1031-1064).
def ref_available():
    return os.path.exists(REF_SO)


def ref_csr_compute(coo, rows, iters):
    """Call the reference's own smvp_csr_compute (main-cli.c:325).  Returns (y, ms_each).

    Its always-on debug dump (main-cli.c:374-394) is sent to /dev/null; the per-iteration times are
    the reference's own clock_gettime bracket (main-cli.c:408,419) read back from struct _time_data_.
    """
    L = ctypes.CDLL(REF_SO)
    buf = np.ascontiguousarray(coo, dtype=COO_DT).copy()  # sorted in place by the callee (:340)
    tbuf = np.zeros(5 + iters, np.float64)
    f = L.smvp_csr_compute
    f.restype = ctypes.POINTER(ctypes.c_double)
    f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    libc = ctypes.CDLL(None)
    libc.fflush(None)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    try:
        yp = f(_p(buf), rows, len(buf), iters, _p(tbuf))
        libc.fflush(None)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)
    y = np.ctypeslib.as_array(yp, shape=(rows,)).copy()
    return y, tbuf[5:].copy()
