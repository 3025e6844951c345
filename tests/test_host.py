"""CPU tests of the C host side (smvp-toolkit_b200/host): Matrix Market loader and report writer, through
libsmvp_host.so, against the reference's fixtures."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import util
from oracle import oracle

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "smvp-toolkit_b200", "lib")


class TimeStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in ("time_total", "time_avg", "time_stdev", "time_min", "time_max")]


@pytest.fixture(scope="module")
def host():
    so = os.path.join(LIB, "libsmvp_host.so")
    if not os.path.exists(so):
        import __graft_entry__

        __graft_entry__.build()
    L = ctypes.CDLL(so)
    L.smvp_load_mtx.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_void_p)]
    L.smvp_load_mtx.restype = ctypes.c_int
    L.smvp_mmio_error_text.argtypes = [ctypes.c_int]
    L.smvp_mmio_error_text.restype = ctypes.c_char_p
    L.smvp_write_report.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.POINTER(TimeStats), ctypes.c_ulong, ctypes.c_char_p, ctypes.c_size_t]
    L.smvp_write_report.restype = ctypes.c_int
    return L


def load(host, path):
    code = ctypes.create_string_buffer(4)
    m, n, nnz, p = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64(), ctypes.c_void_p()
    rc = host.smvp_load_mtx(path.encode(), code, ctypes.byref(m), ctypes.byref(n), ctypes.byref(nnz), ctypes.byref(p))
    if rc != 0:
        return rc, None
    buf = (ctypes.c_char * (16 * nnz.value)).from_address(p.value)
    coo = np.frombuffer(buf, dtype=oracle.COO_DT, count=nnz.value).copy()
    ctypes.CDLL(None).free(p)
    return 0, (m.value, n.value, coo, code.raw.decode())


@pytest.mark.parametrize("name", util.SAMPLES)
def test_loader_matches_reference_semantics(host, name):
    """1-based -> 0-based, pattern => 1.0, symmetric NOT expanded (main-cli.c:1427-1441)."""
    rc, got = load(host, util.sample_path(name))
    assert rc == 0
    m, n, coo, code = got
    em, en, ecoo = util.load_sample(name)
    assert (m, n) == (em, en)
    assert np.array_equal(coo["row"], ecoo["row"]) and np.array_equal(coo["col"], ecoo["col"])
    assert np.array_equal(coo["val"], ecoo["val"])
    assert code[0] == "M" and code[1] == "C"
    if name == "pwt":
        assert code == "MCPS" and len(coo) == 181313  # stored triangle only


def test_loader_error_paths(host, tmp_path):
    rc, _ = load(host, util.sample_path("badfile"))  # 0-byte fixture of the reference
    assert rc == 12  # MM_PREMATURE_EOF
    assert b"Required parameters not present on first line" in host.smvp_mmio_error_text(rc)
    rc, _ = load(host, str(tmp_path / "missing.mtx"))
    assert rc == 101
    cases = {
        "nohdr.mtx": ("hello matrix coordinate real general\n1 1 1\n1 1 2.0\n", 14),
        "badtype.mtx": ("%%MatrixMarket matrix coordinate quaternion general\n1 1 1\n1 1 2.0\n", 15),
        "dense.mtx": ("%%MatrixMarket matrix array real general\n1 1\n2.0\n", 102),
        "complex.mtx": ("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 2.0 0.0\n", 103),
        "short.mtx": ("%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 2.0\n2 2 1.0\n", 104),
    }
    for fn, (text, want) in cases.items():
        p = tmp_path / fn
        p.write_text(text)
        rc, _ = load(host, str(p))
        assert rc == want, fn


def test_loader_comments_integer_and_blank_lines(host, tmp_path):
    p = tmp_path / "c.mtx"
    p.write_text("%%MatrixMarket MATRIX Coordinate Integer General\n% a comment\n%another\n\n3 4 2\n1 4 7\n\n3 1 -2\n")
    rc, got = load(host, str(p))
    assert rc == 0
    m, n, coo, code = got
    assert (m, n) == (3, 4) and code == "MCIG"
    assert coo["row"].tolist() == [0, 2] and coo["col"].tolist() == [3, 0] and coo["val"].tolist() == [7.0, -2.0]


def test_expand_symmetric(host, tmp_path):
    """optional mirroring of the stored triangle (the reference never does it)."""
    host.smvp_load_mtx_ex.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int),
                                      ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_void_p)]
    host.smvp_load_mtx_ex.restype = ctypes.c_int

    def load_ex(path, expand):
        code = ctypes.create_string_buffer(4)
        m, n, nnz, p = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64(), ctypes.c_void_p()
        assert host.smvp_load_mtx_ex(path.encode(), expand, code, ctypes.byref(m), ctypes.byref(n), ctypes.byref(nnz),
                                     ctypes.byref(p)) == 0
        coo = np.frombuffer((ctypes.c_char * (16 * nnz.value)).from_address(p.value), dtype=oracle.COO_DT, count=nnz.value).copy()
        ctypes.CDLL(None).free(p)
        return m.value, n.value, coo

    for kind, sign in (("symmetric", 1.0), ("skew-symmetric", -1.0)):
        p = tmp_path / (kind + ".mtx")
        diag = "" if kind == "skew-symmetric" else "2 2 5.0\n"
        p.write_text("%%MatrixMarket matrix coordinate real " + kind + "\n3 3 " + ("3" if diag else "2") + "\n2 1 4.0\n3 1 -1.5\n" + diag)
        m, n, plain = load_ex(str(p), 0)
        m, n, full = load_ex(str(p), 1)
        dense = np.zeros((3, 3))
        dense[full["row"], full["col"]] = full["val"]
        assert len(full) == len(plain) + 2
        assert np.array_equal(dense, sign * dense.T) if kind == "skew-symmetric" else np.array_equal(dense, dense.T)
        assert dense[1, 0] == 4.0 and dense[0, 1] == sign * 4.0 and dense[0, 2] == sign * -1.5
    # general files are untouched; pwt doubles its off-diagonal entries
    m, n, a = load_ex(util.sample_path("memplus"), 1)
    assert len(a) == 126150
    m, n, b = load_ex(util.sample_path("pwt"), 1)
    m, n, c = load_ex(util.sample_path("pwt"), 0)
    assert len(b) == 2 * len(c) - int((c["row"] == c["col"]).sum())


@pytest.mark.parametrize("name,alg", sorted(util.GOLDEN_REPORTS))
def test_report_writer_reproduces_golden_files(host, tmp_path, name, alg):
    """Same header, same layout, same %g formatting as generateReportText (main-cli.c:294-316): with the golden
    file's own numbers the writer must reproduce it byte for byte."""
    gpath = os.path.join(util.GOLDEN, "reports", util.GOLDEN_REPORTS[(name, alg)])
    rep = util.parse_report(gpath)
    golden = open(gpath).read()
    lines = golden.split("\n")
    unix_time = int(lines[1].split()[2])
    in_name = lines[4]
    st = TimeStats(rep["total"], rep["avg"], rep["stdev"], rep["min"], rep["max"])
    m, n, coo = util.load_sample(name)
    y = oracle.csr_mult(*oracle.csr_build(coo, m, n), np.ones(n)) if alg == "CSR" else oracle.tjds_mult_ref_compat(
        oracle.tjds_build(coo, m, n))
    y = np.ascontiguousarray(y)
    out = ctypes.create_string_buffer(4096)
    rc = host.smvp_write_report(in_name.encode(), str(tmp_path).encode(), alg.encode(), rep["nnz"], m, rep["iters"],
                                y.ctypes.data_as(ctypes.c_void_p), ctypes.byref(st), unix_time, out, 4096)
    assert rc == 0
    path = out.value.decode()
    assert os.path.basename(path) == os.path.basename(gpath)
    assert open(path).read() == golden


@pytest.mark.parametrize("fixture", sorted(f for f in os.listdir(os.path.join(util.GOLDEN, "cisr")) if f.endswith(".coe.gz")))
def test_cisr_coe_matches_reference_binary(host, tmp_path, fixture):
    """`-g`: the .coe image must equal, byte for byte, what the unmodified reference prints (tests/golden/cisr)."""
    import gzip

    name, slots = fixture[: -len(".coe.gz")].rsplit("_s", 1)
    slots = int(slots)
    want = gzip.open(os.path.join(util.GOLDEN, "cisr", fixture), "rt").read()
    m, n, coo = util.load_sample(name)
    row_ptr, col_ind, val = oracle.csr_build(coo, m, n)
    libc = ctypes.CDLL(None)
    libc.fopen.restype = ctypes.c_void_p
    libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    libc.fclose.argtypes = [ctypes.c_void_p]
    host.smvp_cisr_coe.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                   ctypes.c_int]
    host.smvp_cisr_coe.restype = ctypes.c_int
    path = str(tmp_path / "out.coe")
    f = libc.fopen(path.encode(), b"w")
    rc = host.smvp_cisr_coe(f, row_ptr.ctypes.data_as(ctypes.c_void_p), col_ind.ctypes.data_as(ctypes.c_void_p),
                            val.ctypes.data_as(ctypes.c_void_p), m, len(coo), slots)
    libc.fclose(f)
    assert rc == 0
    assert open(path).read() == want


def test_cli_usage_and_option_errors():
    cli = os.path.join(LIB, "smvp-toolkit-cli")
    if not os.path.exists(cli):
        pytest.skip("CLI not built")
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage:" in r.stderr
    r = subprocess.run([cli, "-a", "-c", "x.mtx"], capture_output=True, text=True)
    assert r.returncode == 1 and "Combining [-a|--all] with other algorithm flags is not supported." in r.stdout
    r = subprocess.run([cli, "-c", "-n", "0", "x.mtx"], capture_output=True, text=True)
    assert r.returncode == 1 and "Invalid number of algorithm iterations specified." in r.stdout
    r = subprocess.run([cli, "-c", "-n", "12x", "x.mtx"], capture_output=True, text=True)
    assert r.returncode == 1 and "non-number characters" in r.stdout
    r = subprocess.run([cli, "-c", "-d", "/nonexistent-dir", "x.mtx"], capture_output=True, text=True)
    assert r.returncode == 1 and "Report output folder not found" in r.stdout
    r = subprocess.run([cli, "-c", "/nonexistent.mtx"], capture_output=True, text=True)
    assert r.returncode == 1 and "Specified input file not found." in r.stdout
    r = subprocess.run([cli, "-c", util.sample_path("badfile")], capture_output=True, text=True)
    assert r.returncode == 1 and "Required parameters not present on first line of file." in r.stdout
    r = subprocess.run([cli, "-c", "a.mtx", "b.mtx"], capture_output=True, text=True)
    assert r.returncode == 1 and "Must specify a single input file" in r.stderr


# ------------------------------------------------------------------ chunked / threaded entry parser
def _with_threads(monkeypatch, threads, min_chunk=64):
    monkeypatch.setenv("SMVP_LOAD_THREADS", str(threads))
    monkeypatch.setenv("SMVP_LOAD_MIN_CHUNK", str(min_chunk))


def _write_mtx(path, field, m, n, lines, declared=None):
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate %s general\n%% a comment\n%d %d %d\n" % (field, m, n,
                                                                                          len(lines) if declared is None else declared))
        f.write("".join(lines))


def test_threaded_loader_is_bit_identical_to_the_sequential_one(host, tmp_path, monkeypatch):
    """The chunked parser (one line-aligned chunk per thread, exact fast path for short decimals, strtod for the rest)
    must return exactly the entries of the sequential token parser: every value format strtod accepts, every thread
    count, chunk borders anywhere."""
    rng = np.random.default_rng(5)
    m, n, nnz = 5000, 7000, 20000
    fmts = ["%.17g", "%.6f", "%.3e", "%d", "%.15g", "%.16e", "%g"]
    vals = rng.uniform(-1e3, 1e3, nnz) * 10.0 ** rng.integers(-30, 30, nnz)
    special = ["0", "-0.0", "1e-400", "1e400", "-inf", "nan", "0x1.8p3", ".5", "5.", "+7", "1E5", "00012.50", "4.9e-324",
               "123456789012345678901234567890", "0.000000000000000000001234567890123456789"]
    lines = []
    for i in range(nnz):
        tok = special[i % len(special)] if i % 97 == 0 else fmts[i % len(fmts)] % (int(vals[i]) % 10 ** 9 if fmts[i % len(fmts)] == "%d" else vals[i])
        sep = ["\n", "\r\n", "  \n", "\t\n"][i % 4]
        lines.append("%d %s%d %s%s" % (rng.integers(1, m + 1), " " * (i % 3), rng.integers(1, n + 1), tok, sep))
    path = str(tmp_path / "mixed.mtx")
    _write_mtx(path, "real", m, n, lines)
    _with_threads(monkeypatch, 1)
    rc, ref = load(host, path)
    assert rc == 0 and len(ref[2]) == nnz
    # the sequential parser is strtol/strtod: Python's float() agrees on every decimal token
    for i in (0, 1, 2, 3, 5, 6, 8, 9, 10):
        tok = lines[i].split()[2]
        assert ref[2]["val"][i] == float(tok)
    for threads in (2, 3, 8, 64):
        _with_threads(monkeypatch, threads)
        rc, got = load(host, path)
        assert rc == 0
        assert got[2].tobytes() == ref[2].tobytes(), "threads=%d" % threads
    # pattern files: two tokens per entry
    plines = ["%d %d\n" % (rng.integers(1, m + 1), rng.integers(1, n + 1)) for _ in range(5000)]
    ppath = str(tmp_path / "pattern.mtx")
    _write_mtx(ppath, "pattern", m, n, plines)
    _with_threads(monkeypatch, 1)
    rc, pref = load(host, ppath)
    _with_threads(monkeypatch, 5)
    rc2, pgot = load(host, ppath)
    assert rc == 0 and rc2 == 0 and pgot[2].tobytes() == pref[2].tobytes() and np.all(pref[2]["val"] == 1.0)


def test_threaded_loader_keeps_the_token_semantics(host, tmp_path, monkeypatch):
    """fscanf semantics (main-cli.c:1427-1441) survive the chunking: entries may be split over lines (the chunked path
    steps aside), exactly nnz entries are read and what follows is ignored, too few or malformed entries are the
    same error with any thread count."""
    m = n = 50
    entries = [(i % m + 1, (7 * i) % n + 1, 0.5 * i) for i in range(400)]
    free_form = "".join("%d\n%d\n%r\n" % e if k % 2 else "%d %d %r " % e for k, e in enumerate(entries)) + "\n"
    cases = {
        "freeform": (free_form, 400, 0),
        "trailing": ("".join("%d %d %r\n" % e for e in entries) + "9 9 9.0\nthis is ignored\n", 400, 0),
        "too_few": ("".join("%d %d %r\n" % e for e in entries[:399]), 400, 104),
        "bad_token": ("".join("%d %d %r\n" % e for e in entries[:200]) + "3 x 1.0\n" + "".join("%d %d %r\n" % e for e in entries[200:]), 400, 104),
        "fortran_d": ("".join("%d %d %r\n" % e for e in entries[:399]) + "1 1 1.5d3\n", 400, 0),
    }
    for name, (body, declared, want_rc) in cases.items():
        path = str(tmp_path / (name + ".mtx"))
        with open(path, "w") as f:
            f.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (m, n, declared))
            f.write(body)
        results = []
        for threads in (1, 4, 16):
            _with_threads(monkeypatch, threads, min_chunk=16)
            rc, got = load(host, path)
            results.append((rc, None if got is None else got[2].tobytes()))
        assert results[0][0] == want_rc, (name, results[0][0])
        assert results[1] == results[0] and results[2] == results[0], name
        if want_rc == 0 and name != "fortran_d":
            _, got = load(host, path)
            assert [(int(r) + 1, int(c) + 1, float(v)) for r, c, v in got[2]] == entries


def test_report_writer_threaded_vector_is_byte_identical(host, tmp_path, monkeypatch):
    """Vectors of 2^18 rows and more are formatted by several threads (same snprintf("%g") per row, blocks written in
    order): the file must equal, byte for byte, the one the plain fprintf loop writes -- and Python's own %g."""
    import time

    rng = np.random.default_rng(9)
    rows = (1 << 18) + 12345
    y = rng.uniform(-1, 1, rows) * 10.0 ** rng.integers(-12, 12, rows)
    y[::1000] = 0.0
    y[1::1000] = -0.0
    y[2::5000] = np.inf
    y[3::5000] = np.nan
    y[4::5000] = 1e-310  # subnormal
    y[5::5000] = np.round(y[5::5000])
    y = np.ascontiguousarray(y)
    st = TimeStats(1.5, 0.5, 0.1, 0.4, 0.6)
    texts, secs = [], []
    for threads in (1, 7):
        monkeypatch.setenv("SMVP_LOAD_THREADS", str(threads))
        d = tmp_path / ("t%d" % threads)
        d.mkdir()
        out = ctypes.create_string_buffer(4096)
        t0 = time.time()
        rc = host.smvp_write_report(b"big.mtx", str(d).encode(), b"CSR", 123, rows, 3, y.ctypes.data_as(ctypes.c_void_p),
                                    ctypes.byref(st), 1700000000, out, 4096)
        secs.append(time.time() - t0)
        assert rc == 0
        texts.append(open(out.value.decode()).read())
    assert texts[0] == texts[1]
    body = texts[1].split("[\n", 1)[1]
    assert body.endswith("\n]\n\n")
    got = body[: -len("\n]\n\n")].split("\n")
    assert len(got) == rows
    for i in list(range(0, 3000)) + list(range(rows - 50, rows)):
        assert got[i] == ("%g" % y[i]), i


def test_threaded_loader_values_are_correctly_rounded(host, tmp_path, monkeypatch):
    """200 000 random decimal tokens (1..25 significant digits, dot anywhere, exponents up to +-320, signs, leading and
    trailing zeros) through the chunked parser: every value must be the correctly rounded double, i.e. what Python's
    float() -- and strtod, which the reference's fscanf("%lg") uses -- returns for the same text."""
    import random

    rnd = random.Random(1234)
    toks = []
    for i in range(200000):
        nd = rnd.randint(1, 25) if i % 3 else rnd.randint(1, 15)
        digits = "".join(rnd.choice("0123456789") for _ in range(nd))
        dot = rnd.randint(0, nd)
        t = digits[:dot] + ("." if rnd.random() < 0.8 else "") + digits[dot:] if dot < nd else digits + rnd.choice(["", "."])
        if t.startswith("."):
            t = rnd.choice(["", "0", "00"]) + t
        r = rnd.random()
        if r < 0.5:
            t += rnd.choice("eE") + rnd.choice(["", "+", "-"]) + str(rnd.randint(0, 30 if i % 2 else 320))
        toks.append(rnd.choice(["", "-", "+"]) + t)
    m = n = 1000
    lines = ["%d %d %s\n" % (i % m + 1, (7 * i) % n + 1, t) for i, t in enumerate(toks)]
    path = str(tmp_path / "tokens.mtx")
    _write_mtx(path, "real", m, n, lines)
    want = np.array([float(t) for t in toks])
    for threads in (1, 6):
        _with_threads(monkeypatch, threads, min_chunk=4096)
        rc, got = load(host, path)
        assert rc == 0
        v = got[2]["val"]
        bad = np.nonzero(v.view(np.int64) != want.view(np.int64))[0]
        assert len(bad) == 0, (threads, [(toks[i], v[i], want[i]) for i in bad[:5]])
