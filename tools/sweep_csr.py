#!/usr/bin/env python3
"""Tuning sweep (GPU box): time every merge-path configuration and the vector kernel on one matrix.
    python tools/sweep_csr.py --workload stencil27 --grid 369 --cfgs 0-11 [--grid-mode tiles]
Prints one line per configuration: ms, effective GB/s, rel-L2 against the vector kernel."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import smvp_toolkit_b200 as eng  # noqa: E402
from smvp_toolkit_b200 import dist as sdist  # noqa: E402


def parse_list(s):
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            out += list(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def timeit(fn, steps):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="stencil27")
    ap.add_argument("--grid", type=int, default=369)
    ap.add_argument("--scale", type=int, default=24)
    ap.add_argument("--edge-factor", type=int, default=16)
    ap.add_argument("--cfgs", default="0-11")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--grid-modes", default="persistent")
    ap.add_argument("--chunks", default="8")
    ap.add_argument("--aligned", default="auto", help="comma list of 0,1,auto")
    args = ap.parse_args()
    if args.workload == "stencil27":
        src = sdist.StencilSource(eng, args.grid, args.grid, args.grid)
    else:
        src = sdist.RmatSource(eng, args.scale, args.edge_factor << args.scale)
    op = sdist.RowBlockCsr(eng, src, 0, 1, eng.CSR_VECTOR, exchange="none", release_source=True)
    N = src.cols
    x = torch.empty(N, dtype=torch.float64, device="cuda")
    eng.synth_vector(x, N, 999)
    op.set_x(x)
    nbytes = op.global_bytes_per_mult
    print("matrix: %s rows=%d nnz=%d bytes/spmv=%d max_row=%d" % (src.desc, src.rows, op.global_nnz, nbytes, op.A.max_row_nnz), flush=True)
    ms = timeit(lambda: op.A.mult_device(x, op.y_local, eng.CSR_VECTOR), args.steps)
    y_vec = op.y_local.clone()
    print("vector            : %8.3f ms  %8.1f GB/s" % (ms, nbytes / ms / 1e6), flush=True)
    modes = [(c, a) for a in args.aligned.split(",") for c in args.chunks.split(",")]
    for chunk, al in modes:
        os.environ["SMVP_MERGE_CHUNK"] = chunk
        os.environ["SMVP_MERGE_ALIGNED"] = "" if al == "auto" else al
        mode = "chunk=%s al=%s" % (chunk, al)
        for cfg in parse_list(args.cfgs):
            os.environ["SMVP_MERGE_CFG"] = str(cfg)
            op.y_local.fill_(float("nan"))
            try:
                ms = timeit(lambda: op.A.mult_device(x, op.y_local, eng.CSR_MERGE), args.steps)
            except Exception as e:  # noqa: BLE001
                print("merge cfg %2d %-16s: FAILED %s" % (cfg, mode, e), flush=True)
                continue
            err = float(torch.linalg.norm(op.y_local - y_vec) / torch.linalg.norm(y_vec))
            print("merge cfg %2d %-16s: %8.3f ms  %8.1f GB/s  rel_l2_vs_vector=%.2e" % (cfg, mode, ms, nbytes / ms / 1e6, err), flush=True)


if __name__ == "__main__":
    main()
