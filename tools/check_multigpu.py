#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_multigpu.py
Every rank also builds the WHOLE matrix locally and multiplies it on its own GPU; the sharded operators
(row-block CSR with NCCL allgather, row-block CSR with the fused multicast store, column-block TJDS with
NCCL reduce-scatter) must reproduce that result on every rank."""
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
    fails = 0
    for kind in ("stencil", "rmat"):
        src = sdist.StencilSource(eng, 40, 37, 29) if kind == "stencil" else sdist.RmatSource(eng, 16, 12 << 16)
        M, N = src.rows, src.cols
        x = torch.empty(N, dtype=torch.float64, device="cuda")
        eng.synth_vector(x, N, 4242, stream)
        whole = sdist.RowBlockCsr(eng, src, 0, 1, eng.CSR_VECTOR, exchange="none")
        whole.set_x(x)
        whole.step(stream)
        torch.cuda.synchronize()
        y_ref = whole.y_full.clone()
        nrm = float(torch.linalg.norm(y_ref))
        # the sharded R-MAT operators run with the popularity relabelling forced on (small matrices never choose it)
        os.environ["SMVP_CSR_RELABEL"] = "1" if kind == "rmat" else "0"
        for exch in ("nccl", "multicast", "p2p", "copy", "pipeline"):
            for variant in (eng.CSR_VECTOR, eng.CSR_MERGE):
                try:
                    op = sdist.RowBlockCsr(eng, src, rank, world, variant, exchange=exch)
                except RuntimeError as e:
                    if rank == 0:
                        print("SKIP csr %s: %s" % (exch, e), flush=True)
                    break
                op.set_x(x)
                # the pipelined push: every scheme this box offers (copy engines / SM kernel, unicast / multicast)
                schemes = op.scheme_candidates() if (exch == "pipeline" and variant == eng.CSR_MERGE) else [op.scheme]
                for scheme in schemes:
                    if scheme:
                        op.set_scheme(scheme)
                    for _ in range(3):
                        op.y_full.fill_(float("nan")) if exch == "nccl" else None
                        op.step(stream)
                    op.finish(stream)
                    torch.cuda.synchronize()
                    dist.barrier()
                    err = float(torch.linalg.norm(op.last_y() - y_ref)) / nrm
                    ok = err <= 1e-12
                    fails += 0 if ok else 1
                    print("rank %d %-7s csr %-9s %-16s variant %d rows [%d,%d) x_relabel %d: rel_l2 %.2e %s" % (
                        rank, kind, exch, scheme or "", variant, op.r0, op.r1, op.A.x_relabel, err, "ok" if ok else "FAIL"),
                        flush=True)
                if exch == "pipeline" and variant == eng.CSR_MERGE:
                    tuned = op.tune_pipeline(stream, steps=3)
                    op.step(stream)
                    op.finish(stream)
                    torch.cuda.synchronize()
                    dist.barrier()
                    err = float(torch.linalg.norm(op.last_y() - y_ref)) / nrm
                    fails += 0 if err <= 1e-12 else 1
                    if rank == 0:
                        print("tuned pipeline: %s -> rel_l2 %.2e" % (tuned, err), flush=True)
                op.free()
                del op
        if True:
            for variant in (eng.TJDS_ATOMIC, eng.TJDS_DETERMINISTIC, eng.TJDS_DETERMINISTIC_FAST):
                op = sdist.ColBlockTjds(eng, src, rank, world, variant, exchange="nccl")
                op.set_x(x, stream)
                op.step(stream)
                torch.cuda.synchronize()
                lo = rank * (op.Mp // world)
                hi = min(lo + op.Mp // world, M)
                err = float(torch.linalg.norm(op.y_owned[: hi - lo] - y_ref[lo:hi])) / nrm
                ok = err <= 1e-12
                tag = ""
                if variant != eng.TJDS_ATOMIC:
                    # reproducible run to run INCLUDING the exchange: the N-way sum is taken in rank order
                    keep = op.y_owned.clone()
                    for _ in range(3):
                        op.step(stream)
                        torch.cuda.synchronize()
                        ok = ok and bool(torch.equal(op.y_owned, keep))
                    tag = " ordered=%s run-to-run %s" % (op.ordered, "bit-identical" if ok else "DIFFERS")
                fails += 0 if ok else 1
                print("rank %d %-7s tjds nccl     variant %d cols [%d,%d): rel_l2 %.2e %s%s" % (
                    rank, kind, variant, op.c0, op.c1, err, "ok" if ok else "FAIL", tag), flush=True)
                op.free()
        whole.free()
    t = torch.tensor([fails], device="cuda")
    dist.all_reduce(t)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIGPU CHECK %s" % ("PASSED" if int(t[0]) == 0 else "FAILED"), flush=True)
    sys.exit(0 if int(t[0]) == 0 else 1)


if __name__ == "__main__":
    main()
