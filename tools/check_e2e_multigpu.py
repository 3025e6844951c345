#!/usr/bin/env python3
"""End-to-end path on N GPUs (run under torchrun): every rank calls the pipelined host entry point on its row block.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/check_e2e_multigpu.py
Checks that the rows each rank delivers to the host equal, bit for bit, the rows its device-resident step produces,
then times the step.  --grid G chooses the stencil size (default 369)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import smvp_toolkit_b200 as eng  # noqa: E402
from smvp_toolkit_b200 import dist as sdist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=369)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
stream = torch.cuda.current_stream()
g = args.grid
src = sdist.StencilSource(eng, g, g, g)
op = sdist.RowBlockCsr(eng, src, rank, world, eng.CSR_AUTO, exchange="pipeline" if world > 1 else "none")
N = src.cols
x = torch.empty(N, dtype=torch.float64, device="cuda")
eng.synth_vector(x, N, 777, stream)
op.set_x(x, stream)
op.step(stream)
op.finish(stream)
torch.cuda.synchronize()
y_dev = op.y_local.clone()
hx = torch.empty(N, dtype=torch.float64).pin_memory()
hx.copy_(x)
hy = torch.full((op.local_rows_out,), float("nan"), dtype=torch.float64).pin_memory()
op.e2e_step(hx, hy, stream)
torch.cuda.synchronize()
same = bool(torch.equal(hy, y_dev.cpu()))
dist.barrier()
t0 = time.perf_counter()
for _ in range(args.steps):
    op.e2e_step(hx, hy, stream)
torch.cuda.synchronize()
dist.barrier()
ms = (time.perf_counter() - t0) * 1e3 / args.steps
print("rank %d/%d rows [%d,%d): host rows bit-identical to the device step: %s   e2e %.3f ms per step" % (
    rank, world, op.r0, op.r1, same, ms), flush=True)
ok = torch.tensor([1 if same else 0], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("E2E MULTIGPU CHECK %s" % ("PASSED" if int(ok[0]) == 1 else "FAILED"), flush=True)
sys.exit(0 if int(ok[0]) == 1 else 1)
