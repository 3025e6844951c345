#!/usr/bin/env python3
"""Probe (ONE GPU): does a copy running beside the merge-path SpMV overlap with it or serialise?  The SpMV of half of
the 369^3 stencil runs on the main stream while a high-priority side stream moves 201 MB LOCALLY (no NVLink involved)
with (a) a copy-engine transfer, (b) the 512-thread store kernel, (c) the TMA push kernel.  Prints SpMV alone, copy
alone, both together (wall, and the SpMV's own event bracket)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import smvp_toolkit_b200 as eng  # noqa: E402
from smvp_toolkit_b200 import dist as sdist  # noqa: E402


def main():
    g = int(os.environ.get("PROBE_GRID", "369"))
    src = sdist.StencilSource(eng, g, g, g)
    op = sdist.RowBlockCsr(eng, src, 0, 2, eng.CSR_AUTO, exchange="none")
    main_s = torch.cuda.current_stream()
    side = torch.cuda.Stream(priority=-1)
    x = torch.empty(src.cols, dtype=torch.float64, device="cuda")
    eng.synth_vector(x, src.cols, 1, main_s)
    op.set_x(x, main_s)
    n = op.r1 - op.r0
    a = torch.ones(n, dtype=torch.float64, device="cuda")
    b = torch.zeros(n, dtype=torch.float64, device="cuda")
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    copies = {
        "copy engine": lambda s: eng.copy_device(b, a, 8 * n, s),
        "store kernel 32 CTAs": lambda s: eng.push_device(b, a, 8 * n, 32, s),
        "tma push 128 CTAs": lambda s: eng.push_tma_device([b.data_ptr()], a, 8 * n, 128, s),
        "tma push 32 CTAs": lambda s: eng.push_tma_device([b.data_ptr()], a, 8 * n, 32, s),
    }

    def spmv(s):
        op.A.mult_device(None, y, eng.CSR_AUTO, s)

    def timeit(fn, reps=10):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_s)
        for _ in range(reps):
            fn()
        e1.record(main_s)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    t_spmv = timeit(lambda: spmv(main_s))
    print("SpMV alone: %.3f ms" % t_spmv, flush=True)
    for name, cp in copies.items():
        t_cp = timeit(lambda: cp(main_s))
        start, done = torch.cuda.Event(), torch.cuda.Event()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def both():
            start.record(main_s)
            side.wait_event(start)
            cp(side)
            done.record(side)
            k0.record(main_s)
            spmv(main_s)
            k1.record(main_s)
            main_s.wait_event(done)

        t_both = timeit(both)
        print("%-22s alone %.3f ms | together %.3f ms (sum %.3f, max %.3f), SpMV bracket %.3f ms" % (
            name, t_cp, t_both, t_spmv + t_cp, max(t_spmv, t_cp), k0.elapsed_time(k1)), flush=True)


if __name__ == "__main__":
    main()
