#!/usr/bin/env python3
"""Where does the vector-CSR kernel earn its place?  (GPU box)  Times VECTOR (wide 128-bit loads and lane-contiguous loads)
against MERGE on (a) the 27-point stencil 369^3 (27 nnz per row: short regular rows) and (b) a dense-band matrix with
LONG regular rows (`--band` consecutive columns per row, default 128: the `regular_long` class AUTO sends to the vector
kernel), both ~1 B nnz, and prints what AUTO picks.
    python tools/sweep_vector.py [--band 128] [--rows 8000000]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import smvp_toolkit_b200 as eng  # noqa: E402


def timeit(fn, steps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run(name, A, n):
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    eng.synth_vector(x, n, 999)
    y = torch.empty(A.rows, dtype=torch.float64, device="cuda")
    nb = A.bytes_per_mult
    A.set_x_device(x)
    ref = None
    print("%s: rows %d nnz %d (%.1f per row), AUTO -> %s" % (name, A.rows, A.nnz, A.nnz / A.rows,
                                                              {1: "vector", 2: "merge"}[A.auto_variant]), flush=True)
    for label, variant, wide in (("merge", eng.CSR_MERGE, None), ("vector, 128-bit loads", eng.CSR_VECTOR, "1"),
                                 ("vector, lane-contiguous loads", eng.CSR_VECTOR, "0"), ("vector, default", eng.CSR_VECTOR, None)):
        if wide is None:
            os.environ.pop("SMVP_VECTOR_WIDE", None)
        else:
            os.environ["SMVP_VECTOR_WIDE"] = wide
        ms = timeit(lambda: A.mult_device(None, y, variant))
        if ref is None:
            ref = y.clone()
        rel = float(torch.linalg.norm(y - ref) / torch.linalg.norm(ref))
        print("  %-30s %8.3f ms  %8.1f GB/s  (%.1f %% of 8 TB/s)  rel vs merge %.1e" % (label, ms, nb / ms / 1e6, nb / ms / 8e7, rel),
              flush=True)
    os.environ.pop("SMVP_VECTOR_WIDE", None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--band", type=int, default=128)
    ap.add_argument("--rows", type=int, default=8_000_000)
    ap.add_argument("--no-stencil", action="store_true")
    args = ap.parse_args()
    if not args.no_stencil:
        g = 369
        r, c, v = eng.synth_stencil27(g, g, g)
        A = eng.CsrMatrix.build_device(r, c, v, g ** 3, g ** 3, r.n)
        for a in (r, c, v):
            a.free()
        run("27-point stencil 369^3", A, g ** 3)
        A.free()
    m, b = args.rows, args.band
    n = m + b
    rows = torch.arange(m, dtype=torch.int32, device="cuda").repeat_interleave(b)
    cols = (torch.arange(m, dtype=torch.int32, device="cuda").view(-1, 1) + torch.arange(b, dtype=torch.int32, device="cuda")).reshape(-1)
    vals = torch.empty(m * b, dtype=torch.float64, device="cuda")
    eng.synth_vector(vals, m * b, 5)
    A = eng.CsrMatrix.build_device(rows, cols, vals, m, n, m * b)
    del rows, cols, vals
    torch.cuda.empty_cache()
    run("dense band, %d consecutive columns per row" % b, A, n)
    A.free()


if __name__ == "__main__":
    main()
