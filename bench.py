#!/usr/bin/env python3
"""bench.py -- headline benchmark of the CSR / TJDS SpMV path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload stencil27|rmat] [--format csr|tjds] [--variant ...]

Default (N=1): BASELINE.json configs[2], the HBM-roofline run the metric is quoted on --
27-point stencil on a 369^3 grid (50 243 409 rows, 1 349 232 625 nnz, fp64), CSR, one SpMV per step.
N>1: the same matrix cut into nnz-balanced row blocks, x replicated, y all-gathered (strong scaling).

One JSON line on stdout (rank 0).  `value` = effective GB/s = algorithmic bytes of one SpMV
(12 nnz + 4 (M+1) + 8 N + 8 M for CSR; 12 nnz + 4 (ndiag+1) + 8 N + 8 M for TJDS) / step time, inputs
resident in HBM.  `e2e` = the same metric through the host-buffer C-ABI call (x from pinned host
memory, y back to pinned host memory, copies inside the timed region).
`--impl reference` times the reference's own CPU loop (oracle/_ref when it was built, else the
oracle port) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

METRIC = "spmv_effective_bandwidth"
UNIT = "GB/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="stencil27", choices=["stencil27", "rmat"])
    p.add_argument("--grid", type=int, default=369, help="stencil27: grid edge (369 -> 50.2M rows, 1.35B nnz)")
    p.add_argument("--scale", type=int, default=26, help="rmat: log2(rows)")
    p.add_argument("--edge-factor", type=int, default=16)
    p.add_argument("--format", default="csr", choices=["csr", "tjds"])
    p.add_argument("--variant", default="auto", choices=["auto", "vector", "merge", "atomic", "deterministic"])
    p.add_argument("--exchange", default="auto", choices=["auto", "pipeline", "copy", "multicast", "p2p", "nccl", "none"],
                   help="N>1: collective after the multiply (allgather of y for CSR, reduce-scatter for TJDS)")
    p.add_argument("--sub-blocks", type=int, default=4, help="exchange=copy: sub-blocks per rank")
    p.add_argument("--cpu-grid", type=int, default=100, help="grid edge of the bounded CPU sample (stencil27)")
    p.add_argument("--cpu-scale", type=int, default=20, help="scale of the bounded CPU sample (rmat)")
    p.add_argument("--cpu-iters", type=int, default=10)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-all-cores", action="store_true", help="skip the extra all-cores CPU figure")
    p.add_argument("--no-e2e", action="store_true")
    return p.parse_args()


# --------------------------------------------------------------------------------------- helpers
def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def csr_bytes(rows, cols, nnz):
    return 12 * nnz + 4 * (rows + 1) + 8 * cols + 8 * rows


def tjds_bytes(rows, cols, nnz, ndiag):
    return 12 * nnz + 4 * (ndiag + 1) + 8 * cols + 8 * rows


# --------------------------------------------------------------------------------------- CPU arm
def stencil_coo_numpy(g):
    """27-point stencil COO on a g^3 grid, (row,col)-sorted, {26,-1} values -- numpy only (no engine)."""
    from oracle import oracle

    idx = np.arange(g ** 3, dtype=np.int64)
    ix, iy, iz = idx % g, (idx // g) % g, idx // (g * g)
    rows, cols = [], []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ok = ((ix + dx >= 0) & (ix + dx < g) & (iy + dy >= 0) & (iy + dy < g) & (iz + dz >= 0) & (iz + dz < g))
                rows.append(idx[ok])
                cols.append(idx[ok] + dx + g * (dy + g * dz))
    row = np.concatenate(rows)
    col = np.concatenate(cols)
    order = np.lexsort((col, row))
    row, col = row[order], col[order]
    return oracle.make_coo(row, col, np.where(row == col, 26.0, -1.0))


def rmat_coo_numpy(scale, edge_factor):
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import synth_ref

    return synth_ref.rmat(scale, edge_factor << scale, seed=42)


def cpu_reference_run(args, iters, drop=0):
    """The reference's CPU implementation of the path on a bounded sample of the workload, 1 thread
    (the reference is single-threaded: main-cli.c has no threads)."""
    from oracle import oracle

    if args.workload == "stencil27":
        g = args.cpu_grid
        coo = stencil_coo_numpy(g)
        m = n = g ** 3
        sample = "27-point stencil %d^3 (%d rows, %d nnz), %d iterations, x = ones" % (g, m, len(coo), iters)
    else:
        coo = rmat_coo_numpy(args.cpu_scale, args.edge_factor)
        m = n = 1 << args.cpu_scale
        sample = "R-MAT scale %d, edge factor %d (%d rows, %d nnz after dedupe), %d iterations" % (
            args.cpu_scale, args.edge_factor, m, len(coo), iters)
    nnz = len(coo)
    if args.format == "csr":
        nbytes = csr_bytes(m, n, nnz)
        # the verbatim reference leaves row_ptr slots unwritten for empty rows (U3): only usable on
        # matrices without empty rows, i.e. the stencil
        if oracle.ref_available() and args.workload == "stencil27":
            kind = "reference"
            _, ms = oracle.ref_csr_compute(coo, m, iters)
            what = "oracle/_ref/libsmvp_ref.so: the unmodified smvp_csr_compute (main-cli.c:325), its own clock_gettime bracket"
        else:
            kind = "port"
            rp, ci, va = oracle.csr_build(coo, m, n)
            _, ms = oracle.csr_mult_timed(rp, ci, va, np.ones(n), iters)
            what = "oracle/smvp_oracle.c: oracle_csr_mult_timed (restatement of main-cli.c:402-420)"
    else:
        kind = "port"  # the reference's TJDS build is O(nnz*N) (main-cli.c:894-904): not runnable at this size
        t = oracle.tjds_build(coo, m, n)
        nbytes = tjds_bytes(m, n, nnz, t.ndiag)
        _, ms = oracle.tjds_mult_timed(t, np.ones(n), iters)
        what = "oracle/smvp_oracle.c: oracle_tjds_mult_timed (restatement of main-cli.c:1004-1024, all diagonals)"
    ms = np.asarray(ms)[drop:]
    avg_ms = float(ms.mean())
    res = {"value": nbytes / (avg_ms * 1e-3) / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
           "sample": sample + "; " + what, "gflops": 2 * nnz / (avg_ms * 1e-3) / 1e9, "ms_per_step": avg_ms,
           "min_ms": float(ms.min())}
    if args.format == "csr" and not args.no_all_cores:
        # NOT the reference (it is single-threaded): its loop over nnz-balanced row blocks on every core of the box,
        # same sample, so the 1-thread figure can be put in proportion.  Reported beside it, never instead of it.
        cores = os.cpu_count() or 1
        rp, ci, va = oracle.csr_build(coo, m, n)
        _, ms_mt = oracle.csr_mult_timed_mt(rp, ci, va, np.ones(n), max(iters, 3) + 2, cores)
        mt = float(np.asarray(ms_mt)[2:].mean())
        res["all_cores"] = {"value": nbytes / (mt * 1e-3) / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                            "note": "oracle/smvp_oracle.c oracle_csr_mult_timed_mt: the reference loop (main-cli.c:410-416) "
                                    "row-parallel on all host cores; the reference itself is single-threaded"}
    return res


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    iters = max(1, args.steps)
    # warm-up iterations are part of the same call; the first `warmup` timings are dropped
    res = cpu_reference_run(args, iters + args.warmup, drop=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, None),
        "gflops": res["gflops"],
        "cpu_baseline": dict({"value": res["value"], "unit": UNIT, "cores": 1, "kind": res["kind"], "sample": res["sample"]},
                             **({"all_cores": res["all_cores"]} if "all_cores" in res else {})),
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print_json(line)


def workload_config(args, extra):
    if args.workload == "stencil27":
        g = args.grid
        w = "27-point stencil %d^3 (BASELINE.json configs[2]): %d rows, %d nnz, fp64, %s" % (
            g, g ** 3, (3 * g - 2) ** 3, args.format.upper())
    else:
        w = "R-MAT scale %d edge factor %d (BASELINE.json configs[3]/[4]), fp64, %s" % (args.scale, args.edge_factor,
                                                                                         args.format.upper())
    cfg = {"workload": w, "format": args.format, "variant": args.variant,
           "l2_hygiene": "matrix streams (>= 12 B/nnz) are far larger than the 126 MB L2; no flush needed"}
    if extra:
        cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import smvp_toolkit_b200 as eng

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun (one rank per GPU)" % args.gpus)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from smvp_toolkit_b200 import dist as sdist

    variant_map = {"auto": eng.CSR_AUTO, "vector": eng.CSR_VECTOR, "merge": eng.CSR_MERGE}
    tj_map = {"auto": eng.TJDS_ATOMIC, "atomic": eng.TJDS_ATOMIC, "deterministic": eng.TJDS_DETERMINISTIC}

    # ---------------- build the shard of this rank
    t_build0 = time.time()
    if args.workload == "stencil27":
        g = args.grid
        M = N = g ** 3
        gen = sdist.StencilSource(eng, g, g, g)
    else:
        M = N = 1 << args.scale
        gen = sdist.RmatSource(eng, args.scale, args.edge_factor << args.scale)
    exch = "none" if world == 1 else args.exchange
    if args.format == "csr":
        # "auto", in the order measured on B200 (profiles/r01_multigpu.md): copy engines pushing finished
        # sub-blocks over NVLink while the next sub-block multiplies; in-kernel stores to the NVSwitch multicast
        # address; in-kernel unicast fan-out; SpMV followed by an NCCL allgather as the fallback of last resort
        candidates = ["pipeline", "copy", "multicast", "p2p", "nccl"] if exch == "auto" else [exch]
        op, err = None, None
        for cand in candidates:
            try:
                op = sdist.RowBlockCsr(eng, gen, rank, world, variant_map.get(args.variant, eng.CSR_AUTO), exchange=cand,
                                       release_source=(cand == candidates[-1]), sub_blocks=args.sub_blocks)
            except Exception as e:  # noqa: BLE001  (e.g. no NVSwitch multicast on this box)
                op, err = None, e
            if world > 1:
                ok = torch.tensor([1 if op is not None else 0], device="cuda")
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok[0]) == 0 and op is not None:
                    op.free()
                    op = None
            if op is not None:
                exch = cand
                break
        if op is None:
            raise err if err is not None else RuntimeError("no exchange could be set up on every rank")
    else:
        op = sdist.ColBlockTjds(eng, gen, rank, world, tj_map.get(args.variant, eng.TJDS_ATOMIC),
                                exchange="nccl" if exch in ("pipeline", "copy", "multicast", "p2p", "nccl") else "none", release_source=True)
    torch.cuda.synchronize()
    build_s = time.time() - t_build0
    nnz_total = op.global_nnz
    nbytes = op.global_bytes_per_mult
    stream = torch.cuda.current_stream()

    x = torch.empty(N, dtype=torch.float64, device="cuda")
    eng.synth_vector(x, N, 12345, stream)
    op.set_x(x, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- warm-up, then the timed region: exactly K steps
    for _ in range(max(args.warmup, 3)):
        op.step(stream)
    op.finish(stream)
    barrier()
    sampler = ClockSampler(local_rank)
    launches0 = eng.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e_begin.record(stream)
    for k in range(args.steps):
        ev[k][0].record(stream)
        op.multiply(stream)      # the SpMV kernel(s) of this rank
        ev[k][1].record(stream)
        op.exchange_y(stream)    # N>1: allgather / reduce-scatter
    op.finish(stream)            # drains a pipelined exchange: still inside the timed region
    e_end.record(stream)
    barrier()
    sampler.stop_flag.set()
    sampler.join()
    launches = eng.launch_count() - launches0
    total_ms = e_begin.elapsed_time(e_end)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    t = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms_max = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    value = nbytes / (ms_per_step * 1e-3) / 1e9

    # ---------------- roofline of the dominant kernel (this rank's SpMV launch)
    peak, peak_src = load_peaks()
    local_bytes = op.local_bytes_per_mult
    achieved = local_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                # ncu capture of THIS configuration only (N=1, stencil 369^3): profiles/traffic.json
                "traffic": (op.measured_traffic_bytes() if (world == 1 and args.workload == "stencil27" and args.grid == 369)
                            else None),
                "kernel": op.kernel_name, "kernel_ms": kern_ms,
                "algorithmic_bytes_per_launch": local_bytes, "peak_source": peak_src,
                "frac_of_nominal_8000": achieved / 8000.0}

    # ---------------- end to end: host x -> device -> SpMV (-> exchange) -> host y, every step
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 20))
        hx = torch.empty(N, dtype=torch.float64).pin_memory()
        hx.copy_(x)
        hy = torch.empty(op.local_rows_out, dtype=torch.float64).pin_memory()
        for _ in range(2):
            op.e2e_step(hx, hy, stream)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            op.e2e_step(hx, hy, stream)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0]) / e2e_steps
        # whole-job bytes per step: x crosses PCIe once (each rank uploads its slice), y comes back once
        e2e = {"value": nbytes / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(N * 8),
               "d2h_bytes_per_step": int(M * 8), "ms_per_step": e2e_ms, "steps": e2e_steps, "api": op.e2e_api}
        del hx, hy

    clocks = sampler.result()
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, {"rows": M, "cols": N, "nnz": nnz_total, "bytes_per_spmv": nbytes,
                                             "partition": op.partition_desc, "kernel_variant": op.variant_name,
                                             "build_s": build_s, "multiply_plan": op.plan_desc()}),
            "gflops": 2 * nnz_total / (ms_per_step * 1e-3) / 1e9,
            "pct_of_hbm_peak_8000": 100.0 * value / 8000.0 / world,
            "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
        }
        if e2e:
            line["e2e"] = e2e
    # the operator is released before the CPU leg so that leg has the host to itself
    op.free()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            res = cpu_reference_run(args, args.cpu_iters)
            line["cpu_baseline"] = {"value": res["value"], "unit": UNIT, "cores": 1, "kind": res["kind"],
                                    "sample": res["sample"], "gflops": res["gflops"]}
            if "all_cores" in res:
                line["cpu_baseline"]["all_cores"] = res["all_cores"]
        print_json(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: libraries that chat on fd 1 (NCCL prints its version there) are
    # sent to stderr, and the line is written to the real stdout at the end
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global print_json

    def print_json(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
