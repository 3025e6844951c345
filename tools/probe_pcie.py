#!/usr/bin/env python3
"""PCIe floor of the host-vector entry point (GPU box): time a pinned H2D and a pinned D2H transfer of one
50.2M-entry fp64 vector (402 MB) alone and both at once.  The pipelined smvp_csr_mult cannot beat the last number."""
import torch

n = 50243409
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.zeros(n, dtype=torch.float64, device="cuda")
s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s_up.wait_event(e0)
        s_down.wait_event(e0)
        if up:
            with torch.cuda.stream(s_up):
                d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s_down):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_up)
        torch.cuda.current_stream().wait_stream(s_down)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


mb = n * 8 / 1e6
for name, up, down in (("H2D alone", True, False), ("D2H alone", False, True), ("H2D + D2H at once", True, True)):
    ms = run(up, down)
    print("%-18s %7.3f ms  (%.1f GB/s per direction)" % (name, ms, mb / ms))
