#!/usr/bin/env python3
"""Tuning sweep (GPU box): TJDS multiply variants on the 27-point stencil.  python tools/sweep_tjds.py --grid 369"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import smvp_toolkit_b200 as eng  # noqa: E402
from smvp_toolkit_b200 import dist as sdist  # noqa: E402


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
    ap.add_argument("--grid", type=int, default=369)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    src = sdist.StencilSource(eng, args.grid, args.grid, args.grid)
    op = sdist.ColBlockTjds(eng, src, 0, 1, eng.TJDS_ATOMIC, exchange="none")
    x = torch.empty(src.cols, dtype=torch.float64, device="cuda")
    eng.synth_vector(x, src.cols, 999)
    op.set_x(x)
    nbytes = op.global_bytes_per_mult
    for u in ("2", "4", "8"):
        os.environ["SMVP_TJDS_UNROLL"] = u
        ms = timeit(lambda: op.T.mult_device(op.y_partial, eng.TJDS_ATOMIC), args.steps)
        print("tjds atomic unroll %s: %8.3f ms %8.1f GB/s" % (u, ms, nbytes / ms / 1e6), flush=True)
    ms = timeit(lambda: op.T.mult_device(op.y_partial, eng.TJDS_DETERMINISTIC), args.steps)
    print("tjds deterministic   : %8.3f ms %8.1f GB/s" % (ms, nbytes / ms / 1e6), flush=True)


if __name__ == "__main__":
    main()
