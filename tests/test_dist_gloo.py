"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: nnz-balanced partitioning, the row-block
all-gather and the column-block reduce-scatter.  The local multiply is injected (the oracle's loops, test
infrastructure); the partition arithmetic and the collective wiring are the product's (smvp-toolkit_b200/dist.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from oracle import oracle

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from smvp_toolkit_b200 import dist as sdist

        m, n, coo = util.load_sample(name)
        x = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, n))
        y_ref = oracle.csr_mult(*oracle.csr_build(coo, m, n), x.numpy())

        # ---- CSR: row blocks balanced by nnz, x replicated, y all-gathered
        rb = sdist.bounds_from_counts(np.bincount(coo["row"], minlength=m), world)
        r0, r1 = rb[rank], rb[rank + 1]
        blk = coo[(coo["row"] >= r0) & (coo["row"] < r1)].copy()
        blk["row"] -= r0
        arrays = oracle.csr_build(blk, r1 - r0, n)
        y_full = torch.full((m,), float("nan"), dtype=torch.float64)
        sdist.row_block_spmv(dist, lambda xx: torch.from_numpy(oracle.csr_mult(*arrays, xx.numpy())), rb, rank, x, y_full)
        err_csr = util.rel_l2(y_full.numpy(), y_ref)

        # ---- TJDS: column blocks balanced by nnz, x sliced, partial y reduce-scattered
        cb = sdist.bounds_from_counts(np.bincount(coo["col"], minlength=n), world)
        c0, c1 = cb[rank], cb[rank + 1]
        cblk = coo[(coo["col"] >= c0) & (coo["col"] < c1)].copy()
        cblk["col"] -= c0
        t = oracle.tjds_build(cblk, m, c1 - c0)
        owned = sdist.col_block_spmv(dist, lambda xs: torch.from_numpy(oracle.tjds_mult(t, xs.numpy())), cb, rank, world, x, m)
        per = -(-m // world)
        lo, hi = rank * per, min((rank + 1) * per, m)
        err_tjds = float(np.linalg.norm(owned.numpy()[: hi - lo] - y_ref[lo:hi]) / np.linalg.norm(y_ref))
        # the rank-ordered combine (what the deterministic variant uses at N > 1): same blocks, fixed summation order
        owned2 = sdist.col_block_spmv(dist, lambda xs: torch.from_numpy(oracle.tjds_mult(t, xs.numpy())), cb, rank, world, x, m,
                                      ordered=True)
        owned3 = sdist.col_block_spmv(dist, lambda xs: torch.from_numpy(oracle.tjds_mult(t, xs.numpy())), cb, rank, world, x, m,
                                      ordered=True)
        assert torch.equal(owned2, owned3), "rank-ordered combine must be bit-identical run to run"
        err_tjds = max(err_tjds, float(np.linalg.norm(owned2.numpy()[: hi - lo] - y_ref[lo:hi]) / np.linalg.norm(y_ref)))
        # the NCCL-baseline wiring: ONE all-gather on equal padded slots + compaction into the contiguous y
        per = max(rb[g + 1] - rb[g] for g in range(world))
        y_pad = torch.full((per * world,), float("nan"), dtype=torch.float64)
        y_pad[rank * per:rank * per + (r1 - r0)] = torch.from_numpy(oracle.csr_mult(*arrays, x.numpy()))
        y_full2 = torch.full((m,), float("nan"), dtype=torch.float64)
        sdist.allgather_padded(dist, y_pad, y_full2, rb, rank, per)
        assert torch.equal(y_full2, y_full), "padded all-gather + compaction must reproduce the block-wise all-gather"
        nnz_share = len(blk) / max(len(coo), 1)
        out[rank] = (err_csr, err_tjds, nnz_share, rb, cb)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["memplus", "curtis54"])
def test_row_and_column_partition_world2(name):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), name, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        err_csr, err_tjds, share, rb, cb = out[rank]
        assert err_csr <= 1e-12, "every rank must hold the full y after the all-gather"
        assert err_tjds <= 1e-12
        assert 0.35 <= share <= 0.65, "row blocks are balanced by nnz"
        assert rb == out[0][3] and cb == out[0][4], "all ranks agree on the partition"


def test_balanced_bounds_properties():
    sys.path.insert(0, REPO)
    from smvp_toolkit_b200 import dist as sdist

    rng = np.random.default_rng(0)
    for parts in (1, 2, 3, 8):
        counts = rng.integers(0, 50, size=1000)
        counts[10] = 5000  # one hub row
        b = sdist.bounds_from_counts(counts, parts)
        assert b[0] == 0 and b[-1] == len(counts) and len(b) == parts + 1
        assert all(b[i] <= b[i + 1] for i in range(parts))
        csum = np.concatenate([[0], np.cumsum(counts)])
        total = csum[-1]
        for g in range(1, parts):
            # b_g = lower_bound(prefix, g * total / parts)  
            target = total * g // parts
            assert csum[b[g]] >= target and (b[g] == 0 or csum[b[g] - 1] < target)
    # empty matrix
    assert sdist.bounds_from_counts(np.zeros(7, int), 2) == [0, 0, 7]
