#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/.

RUNS ONLY IN THE BUILD CONTAINER (needs /root/reference and oracle/_ref built by
`make -C oracle ref`).  Nothing here runs on the GPU box; the outputs are committed.

What it writes
--------------
sample-data/*.mtx      the reference's sample inputs (data, not source), copied byte for byte
reports/*.txt          the reference's shipped golden reports (output-test/, build/)
ref_y.npz              FULL-PRECISION y vectors returned by the UNMODIFIED reference functions
                       smvp_csr_compute (main-cli.c:325) and smvp_tjds_compute (main-cli.c:734),
                       called through ctypes in oracle/_ref/libsmvp_ref.so, for every sample
                       matrix and for seeded random matrices (keys "<name>/csr", "<name>/tjds")
ref_arrays.json        integer/value arrays the reference prints with its debug switches:
                       CSR row_ptr/col_ind/val (SMVP_CSR_DEBUG is 1 at HEAD, main-cli.c:10,374-394)
                       and TJDS perm/colLen/val/row_ind/start_pos/num_tjdiag (SMVP_TJDS_DEBUG
                       flipped to 1 in a scratch copy under /tmp, main-cli.c:11,870-992)
random_coo.npz         the seeded random COO inputs the ref_y entries named rand* refer to
cisr/*.coe.gz          the .coe images the reference binary prints for `-g -s <slots>` (main-cli.c:473-729)
"""
import ctypes
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
REF_SO = os.path.join(REPO, "oracle", "_ref", "libsmvp_ref.so")
SAMPLES = ["ibm32", "curtis54", "pdp08-pg4", "memplus", "pwt"]

COO_DT = np.dtype([("row", "<i4"), ("col", "<i4"), ("val", "<f8")])


def read_mtx(path):
    """Same parse as main-cli.c:1419-1441: 1-based -> 0-based, pattern => 1.0, no symmetric expansion."""
    with open(path) as f:
        banner = f.readline().lower().split()
        pattern = banner[3] == "pattern"
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        m, n, nnz = (int(t) for t in line.split())
        coo = np.zeros(nnz, dtype=COO_DT)
        for i in range(nnz):
            t = f.readline().split()
            coo["row"][i] = int(t[0]) - 1
            coo["col"][i] = int(t[1]) - 1
            coo["val"][i] = 1.0 if pattern else float(t[2])
    return m, n, coo


def _child(alg, m, n, coo_path, out_path):
    """Run one reference function in this (child) process; its stdout chatter goes to /dev/null."""
    coo = np.load(coo_path)
    lib = ctypes.CDLL(REF_SO)
    iters = 1
    # struct _time_data_ (main-cli.c:87-95): 5 doubles + flexible array
    tbuf = (ctypes.c_double * (5 + iters))()
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    try:
        buf = coo.copy()  # the reference sorts the caller's array in place (main-cli.c:340, :766)
        p = buf.ctypes.data_as(ctypes.c_void_p)
        if alg == "csr":
            f = lib.smvp_csr_compute
            f.restype = ctypes.POINTER(ctypes.c_double)
            f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
            y = f(p, m, len(buf), iters, ctypes.byref(tbuf))
        else:
            f = lib.smvp_tjds_compute
            f.restype = ctypes.POINTER(ctypes.c_double)
            f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
            y = f(p, m, n, len(buf), iters, ctypes.byref(tbuf))
        libc = ctypes.CDLL(None)
        libc.fflush(None)
    finally:
        os.dup2(saved, 1)
    np.save(out_path, np.ctypeslib.as_array(y, shape=(m,)).copy())


def run_ref(alg, m, n, coo, tmp):
    coo_path = os.path.join(tmp, "coo.npy")
    out_path = os.path.join(tmp, "y.npy")
    np.save(coo_path, coo)
    if os.path.exists(out_path):
        os.remove(out_path)
    r = subprocess.run([sys.executable, __file__, "--child", alg, str(m), str(n), coo_path, out_path],
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if r.returncode != 0 or not os.path.exists(out_path):
        return None  # the reference crashed (its LUT dump reads out of bounds on tiny inputs, U13)
    return np.load(out_path)


def random_coo(rng, m, n, nnz, empty_col0=False):
    cells = rng.choice(m * n, size=nnz, replace=False)  # unique coordinates (duplicates are UB, U9)
    coo = np.zeros(nnz, dtype=COO_DT)
    coo["row"] = cells // n
    coo["col"] = cells % n
    coo["val"] = rng.uniform(-1.0, 1.0, size=nnz)
    rng.shuffle(coo)
    return coo


def parse_list(text, conv):
    return [conv(t) for t in text.strip().strip("[]").replace("\n", " ").split(",") if t.strip()]


def debug_arrays(tmp):
    """CSR and TJDS debug dumps of the reference for the three small sample matrices."""
    out = {}
    src = open(os.path.join(REF, "main-cli.c")).read()
    assert "#define SMVP_TJDS_DEBUG 0" in src
    scratch = os.path.join(tmp, "dbg")
    os.makedirs(os.path.join(scratch, "mmio"), exist_ok=True)
    with open(os.path.join(scratch, "main-cli.c"), "w") as f:  # scratch copy under /tmp only
        f.write(src.replace("#define SMVP_TJDS_DEBUG 0", "#define SMVP_TJDS_DEBUG 1"))
    for fn in ("mmio.c", "mmio.h"):
        shutil.copy(os.path.join(REF, "mmio", fn), os.path.join(scratch, "mmio", fn))
    exe = os.path.join(scratch, "ref-dbg")
    subprocess.check_call(["gcc", "-O1", "-w", "-D_XOPEN_SOURCE=700", "-I" + os.path.join(REPO, "oracle", "stub"),
                           "-I" + scratch, "-o", exe, os.path.join(scratch, "main-cli.c"),
                           os.path.join(scratch, "mmio", "mmio.c"), "-lm"])
    for name in ("pdp08-pg4", "ibm32", "curtis54"):
        mtx = os.path.join(REF, "sample-data", name + ".mtx")
        rd = os.path.join(tmp, "rep")
        os.makedirs(rd, exist_ok=True)
        ent = {}
        r = subprocess.run(["stdbuf", "-o0", exe, "-c", "-n", "1", "-d", rd, mtx], capture_output=True, text=True)
        t = r.stdout
        ent["csr"] = {
            "row_ptr": parse_list(re.search(r"CSR JIT row_ptr:\n\t\[(.*?)\]", t, re.S).group(1), int),
            "val": parse_list(re.search(r"CSR JIT val:\n\t\[(.*?)\]", t, re.S).group(1), float),
            "col_ind": parse_list(re.search(r"CSR JIT col_ind:\n\t\[(.*?)\]", t, re.S).group(1), int),
            "y": parse_list(re.search(r"CSR JIT Vector Out:\n\t\[(.*?)\]", t, re.S).group(1), float),
        }
        r = subprocess.run(["stdbuf", "-o0", exe, "-t", "-n", "1", "-d", rd, mtx], capture_output=True, text=True,
                           errors="replace")
        t = r.stdout
        ntj = int(re.search(r"num_tjdiag \(count, not 0-index\):\t(-?\d+)", t).group(1))
        ent["tjds"] = {
            "perm": parse_list(re.search(r"origCol\t\[(.*?)\]", t, re.S).group(1), int),
            "col_len": parse_list(re.search(r"colLen\t\[(.*?)\]", t, re.S).group(1), int),
            "val": parse_list(re.search(r"\tval:\t\t\[(.*?)\]", t, re.S).group(1), float),
            "row_ind": parse_list(re.search(r"\trow_ind:\t\[(.*?)\]", t, re.S).group(1), int),
            # the dump prints num_tjdiag + 1 slots (main-cli.c:985); slots past the real
            # diagonal count are uninitialised memory and are cut by the consumer
            "start_pos_dump": parse_list(re.search(r"\tstart_pos:\t\[(.*?)\]", t, re.S).group(1), int),
            "num_tjdiag": ntj,
        }
        out[name] = ent
    return out


def cisr_fixtures():
    """.coe images printed by the UNMODIFIED reference binary for `-g -s <slots>` (smvp_cisr_coegen, main-cli.c:473-729)."""
    import gzip

    out_dir = os.path.join(HERE, "cisr")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(REPO, "oracle", "_ref", "smvp-toolkit-cli-ref")
    with tempfile.TemporaryDirectory() as tmp:
        for name, slot_list in (("pdp08-pg4", (2, 4, 16, 32)), ("ibm32", (4, 16)), ("curtis54", (4, 16)), ("memplus", (16,))):
            for slots in slot_list:
                r = subprocess.run([exe, "-g", "-s", str(slots), "-d", tmp, os.path.join(REF, "sample-data", name + ".mtx")],
                                   capture_output=True, text=True)
                assert r.returncode == 0, r.stderr
                t = r.stdout
                a = t.index("\n;*********************************************")
                b = t.index("03ffffffff;") + len("03ffffffff;\n\n")
                with gzip.open(os.path.join(out_dir, "%s_s%d.coe.gz" % (name, slots)), "wt") as f:
                    f.write(t[a:b])
                print("cisr", name, slots, len(t[a:b].splitlines()), "lines", flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--cisr-only":
        cisr_fixtures()
        return
    os.makedirs(os.path.join(HERE, "sample-data"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "reports"), exist_ok=True)
    for name in SAMPLES + ["badfile"]:
        shutil.copy(os.path.join(REF, "sample-data", name + ".mtx"), os.path.join(HERE, "sample-data", name + ".mtx"))
    for fn in sorted(os.listdir(os.path.join(REF, "output-test"))):
        shutil.copy(os.path.join(REF, "output-test", fn), os.path.join(HERE, "reports", fn))
    shutil.copy(os.path.join(REF, "build", "smvp-toolbox_report_CSR_1619162887.txt"), os.path.join(HERE, "reports"))

    ys = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name in SAMPLES:
            m, n, coo = read_mtx(os.path.join(REF, "sample-data", name + ".mtx"))
            for alg in ("csr", "tjds"):
                y = run_ref(alg, m, n, coo, tmp)
                print(name, alg, "crashed" if y is None else "ok", flush=True)
                if y is not None:
                    ys[f"{name}/{alg}"] = y
        rng = np.random.default_rng(20261018)
        rnd = {}
        shapes = [(40, 40, 200), (64, 64, 900), (97, 97, 1500), (128, 128, 400), (200, 200, 6000),
                  (33, 33, 33 * 33), (150, 150, 3000), (256, 256, 5000)]
        for k, (m, n, nnz) in enumerate(shapes):
            coo = random_coo(rng, m, n, nnz)
            rnd[f"rand{k}"] = coo
            rnd[f"rand{k}_shape"] = np.array([m, n], dtype=np.int64)
            for alg in ("csr", "tjds"):
                y = run_ref(alg, m, n, coo, tmp)
                print(f"rand{k}", alg, "crashed" if y is None else "ok", flush=True)
                if y is not None:
                    ys[f"rand{k}/{alg}"] = y
        arrays = debug_arrays(tmp)
    np.savez_compressed(os.path.join(HERE, "ref_y.npz"), **ys)
    np.savez_compressed(os.path.join(HERE, "random_coo.npz"), **rnd)
    with open(os.path.join(HERE, "ref_arrays.json"), "w") as f:
        json.dump(arrays, f)
    print("wrote", len(ys), "reference vectors")
    cisr_fixtures()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        _child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6])
    else:
        main()
