#!/usr/bin/env python3
"""Probe (torchrun, N ranks): cost of the symmetric-memory barrier and achievable store bandwidth to the
NVSwitch multicast address / to a unicast peer mapping, with fully coalesced 8-byte stores."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.distributed._symmetric_memory as symm_mem  # noqa: E402

import smvp_toolkit_b200 as eng  # noqa: E402


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    n = 50_000_000
    buf = symm_mem.empty(n, dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD.group_name)
    per = n // world
    s = torch.cuda.current_stream()
    res = {}
    res["barrier_ms"] = timed(lambda: hdl.barrier(channel=0))
    res["local_fill_ms"] = timed(lambda: eng.synth_vector(buf.data_ptr() + 8 * per * rank, per, 7, s))
    mc = int(hdl.multicast_ptr)
    if mc:
        res["mc_fill_ms"] = timed(lambda: eng.synth_vector(mc + 8 * per * rank, per, 7, s))
        res["mc_fill_plus_barrier_ms"] = timed(lambda: (eng.synth_vector(mc + 8 * per * rank, per, 7, s), hdl.barrier(channel=0)))
    peer = (rank + 1) % world
    pp = int(hdl.buffer_ptrs[peer])
    res["peer_fill_ms"] = timed(lambda: eng.synth_vector(pp + 8 * per * rank, per, 7, s))
    res["peer_fill_plus_barrier_ms"] = timed(lambda: (eng.synth_vector(pp + 8 * per * rank, per, 7, s), hdl.barrier(channel=0)))
    local = torch.empty(per, dtype=torch.float64, device="cuda")
    remote = hdl.get_buffer(peer, (per,), torch.float64, per * rank)
    res["peer_copy_ms"] = timed(lambda: remote.copy_(local))
    # copy-engine write to the MULTICAST address (one copy, the switch replicates it): allowed?
    try:
        import ctypes
        import glob

        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
        rt = ctypes.CDLL(cands[0] if cands else "libcudart.so.12")
        rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        rt.cudaMemcpyAsync.restype = ctypes.c_int
        src = torch.full((per,), float(rank + 1), dtype=torch.float64, device="cuda")
        if mc:
            def ce_mc():
                rc = rt.cudaMemcpyAsync(ctypes.c_void_p(mc + 8 * per * rank), ctypes.c_void_p(src.data_ptr()), per * 8, 3,
                                        ctypes.c_void_p(s.cuda_stream))
                if rc != 0:
                    raise RuntimeError("cudaMemcpyAsync to the multicast address failed: %d" % rc)
            buf.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            ce_mc()
            torch.cuda.synchronize()
            hdl.barrier(channel=0)
            torch.cuda.synchronize()
            ok = all(bool((buf[per * k:per * (k + 1)] == float(k + 1)).all()) for k in range(world))
            res["ce_multicast_copy_ms"] = timed(ce_mc)
            res["ce_multicast_copy_ok"] = 1.0 if ok else 0.0
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print("CE multicast probe failed:", e)
    mb = per * 8 / 1e6
    if rank == 0:
        print("block = %.1f MB per rank, world %d" % (mb, world))
        for k, v in res.items():
            print("%-28s %8.3f ms  %8.1f GB/s" % (k, v, mb / v))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
