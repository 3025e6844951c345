#!/usr/bin/env python3
"""A/B of the row-space popularity relabelling on TJDS, R-MAT matrix (GPU box):
    python tools/sweep_rmat_tjds.py --scale 26 [--steps 5]
Builds the same TJDS twice (SMVP_TJDS_RELABEL=0 / 1) and times the atomic and the deterministic multiply;
checks the deterministic results are bit-identical and the atomic ones agree to 1e-12."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import smvp_toolkit_b200 as eng  # noqa: E402
from smvp_toolkit_b200 import dist as sdist  # noqa: E402


def timeit(fn, steps):
    for _ in range(2):
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
    ap.add_argument("--scale", type=int, default=26)
    ap.add_argument("--edge-factor", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--modes", default="0,1")
    args = ap.parse_args()
    src = sdist.RmatSource(eng, args.scale, args.edge_factor << args.scale)
    M = N = src.rows
    r, c, v = src.col_block(0, N)
    nnz = r.n
    x = torch.empty(N, dtype=torch.float64, device="cuda")
    eng.synth_vector(x, N, 999)
    y = torch.empty(M, dtype=torch.float64, device="cuda")
    keep = {}
    for mode in args.modes.split(","):
        os.environ["SMVP_TJDS_RELABEL"] = mode
        T = eng.TjdsMatrix.build_device(r, c, v, M, N, nnz)
        nbytes = T.bytes_per_mult
        T.set_x_device(x)
        for name, variant in (("atomic", eng.TJDS_ATOMIC), ("deterministic", eng.TJDS_DETERMINISTIC)):
            y.fill_(float("nan"))
            ms = timeit(lambda: T.mult_device(y, variant), args.steps)
            tag = ""
            if name in keep:
                if variant == eng.TJDS_DETERMINISTIC:
                    tag = "bit-identical" if torch.equal(y, keep[name]) else "DIFFERS"
                else:
                    tag = "rel_l2 vs plain %.2e" % float(torch.linalg.norm(y - keep[name]) / torch.linalg.norm(keep[name]))
            else:
                keep[name] = y.clone()
            print("relabel=%s state=%d ndiag=%d  %-13s: %8.3f ms  %8.1f GB/s  %s" %
                  (mode, T.y_relabel, T.ndiag, name, ms, nbytes / ms / 1e6, tag), flush=True)
        T.free()


if __name__ == "__main__":
    main()
