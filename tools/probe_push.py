#!/usr/bin/env python3
"""Probe (torchrun, N ranks): every scheme of the pipelined exchange of y by itself -- does the data land in every
rank's y, and how long does one exchange of a 369^3-stencil row block take (no SpMV running)?
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 tools/probe_push.py [schemes]
A scheme that faults kills only this process; run doubtful ones alone (e.g. `tools/probe_push.py tma_multicast:16`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import smvp_toolkit_b200 as eng  # noqa: E402
from smvp_toolkit_b200 import dist as sdist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    stream = torch.cuda.current_stream()
    g = int(os.environ.get("PROBE_GRID", "369"))
    src = sdist.StencilSource(eng, g, g, g)
    op = sdist.RowBlockCsr(eng, src, rank, world, eng.CSR_AUTO, exchange="pipeline")
    schemes = sys.argv[1:] or op.scheme_candidates()
    nloc = op.r1 - op.r0
    for scheme in schemes:
        op.set_scheme(scheme)
        op.y_sym.fill_(float("nan"))
        pattern = torch.arange(op.r0, op.r1, dtype=torch.float64, device="cuda") * 0.5 + 1000.0 * (rank + 1)
        op._src(0).copy_(pattern)
        torch.cuda.synchronize()
        dist.barrier()
        op._push(0, stream, on_main=True)
        torch.cuda.synchronize()
        dist.barrier()
        ok = True
        for k in range(world):
            if k == rank and scheme.split(":")[0].endswith("unicast"):
                continue  # unicast schemes do not send my rows to myself
            exp = torch.arange(op.bounds[k], op.bounds[k + 1], dtype=torch.float64, device="cuda") * 0.5 + 1000.0 * (k + 1)
            ok = ok and bool(torch.equal(op.y_sym[op.bounds[k]:op.bounds[k + 1]], exp))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            op._push(0, stream, on_main=True)
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 10, 0.0 if ok else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            mb = nloc * 8 / 1e6
            print("%-16s %s  %.3f ms per exchange  (block %.1f MB; ingress per GPU %.0f GB/s)" % (
                scheme, "data ok" if float(t[1]) == 0 else "DATA WRONG", float(t[0]), mb,
                mb * (world - 1 + (1 if "multicast" in scheme else 0)) / float(t[0])), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
