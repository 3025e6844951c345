#!/usr/bin/env python3
"""Tuning sweep of the pipelined host-vector entry point (GPU box): smvp_csr_mult(A, x_host, y_host, iters=1)
on the 369^3 stencil with pinned buffers, for several (tile ranges, x upload pieces) settings.
    python tools/sweep_e2e.py --settings 32x64,16x32,32x32,64x64,32x128,16x16"""
import argparse
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import smvp_toolkit_b200 as eng  # noqa: E402
from smvp_toolkit_b200 import dist as sdist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=369)
ap.add_argument("--settings", default="32x64,16x32,32x32,64x64,32x128,16x16,64x128")
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
src = sdist.StencilSource(eng, args.grid, args.grid, args.grid)
op = sdist.RowBlockCsr(eng, src, 0, 1, eng.CSR_AUTO, exchange="none", release_source=True)
N = M = src.rows
hx = torch.rand(N, dtype=torch.float64).pin_memory()
hy = torch.empty(M, dtype=torch.float64).pin_memory()
ref = None
for setting in ["off"] + args.settings.split(","):
    if setting == "off":
        os.environ["SMVP_NO_OVERLAP"] = "1"
    else:
        os.environ.pop("SMVP_NO_OVERLAP", None)
        if setting == "ramp":
            os.environ["SMVP_PIPE_PROFILE"] = "ramp"
            for k in ("SMVP_PIPE_RANGES", "SMVP_PIPE_XCHUNKS"):
                os.environ.pop(k, None)
            os.environ["SMVP_PIPE_STREAMS"] = "1"
        else:
            os.environ.pop("SMVP_PIPE_PROFILE", None)
            r, xc, ns = (setting.split("x") + ["2"])[:3]
            os.environ["SMVP_PIPE_RANGES"], os.environ["SMVP_PIPE_XCHUNKS"], os.environ["SMVP_PIPE_STREAMS"] = r, xc, ns
    ms = ctypes.c_double(0)
    for it in range(2 + args.steps):
        if it == 2:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        rc = eng.lib().smvp_csr_mult(op.A._h, ctypes.c_void_p(hx.data_ptr()), ctypes.c_void_p(hy.data_ptr()), 1,
                                     ctypes.byref(ms), eng.CSR_AUTO)
        assert rc == 0
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / args.steps
    if ref is None:
        ref = hy.clone()
    same = bool(torch.equal(ref, hy))
    print("ranges x pieces x streams %-10s: %7.3f ms per call   multiply alone %.3f ms   y identical to the plain path: %s" %
          (setting, wall, ms.value, same), flush=True)
