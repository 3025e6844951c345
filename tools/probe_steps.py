#!/usr/bin/env python3
"""Probe (torchrun, N ranks): ms per pipelined step of the 369^3 stencil CSR as a function of how the steps are driven
-- number of steps per timed region, per-step timing events, the NVML clock-sampler thread -- for one exchange scheme.
    torchrun ... tools/probe_steps.py [scheme]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import smvp_toolkit_b200 as eng  # noqa: E402
from smvp_toolkit_b200 import dist as sdist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    stream = torch.cuda.current_stream()
    src = sdist.StencilSource(eng, 369, 369, 369)
    op = sdist.RowBlockCsr(eng, src, rank, world, eng.CSR_AUTO, exchange="pipeline")
    x = torch.empty(src.cols, dtype=torch.float64, device="cuda")
    eng.synth_vector(x, src.cols, 12345, stream)
    op.set_x(x, stream)
    schemes = sys.argv[1:] or [op.scheme]
    for scheme in schemes:
        op.set_scheme(scheme)
        for steps, events, sampler in ((6, False, False), (50, False, False), (50, True, False), (50, False, True), (50, True, True),
                                       (200, False, False), (200, True, True), (6, False, False)):
            for _ in range(3):
                op.step(stream)
            op.finish(stream)
            torch.cuda.synchronize()
            dist.barrier()
            smp = bench.ClockSampler(lr) if sampler else None
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if smp:
                smp.start()
            t0 = time.perf_counter()
            e0.record(stream)
            for k in range(steps):
                op.spmv_events = ev[k] if events else None
                op.multiply(stream)
                op.exchange_y(stream)
            t_issue = time.perf_counter() - t0
            op.finish(stream)
            e1.record(stream)
            torch.cuda.synchronize()
            dist.barrier()
            op.spmv_events = None
            if smp:
                smp.stop_flag.set()
                smp.join()
            t = torch.tensor([e0.elapsed_time(e1) / steps, t_issue * 1e3 / steps], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                print("%-16s steps %3d events %-5s sampler %-5s: %.4f ms/step (host issue %.4f ms/step)" % (
                    scheme, steps, events, sampler, float(t[0]), float(t[1])), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
