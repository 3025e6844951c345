import os, sys
sys.path.insert(0, os.getcwd())
import torch
import smvp_toolkit_b200 as eng
from smvp_toolkit_b200 import dist as sdist
def timeit(fn, steps=8):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / steps
src = sdist.RmatSource(eng, 26, 16 << 26)
x = torch.empty(src.cols, dtype=torch.float64, device="cuda"); eng.synth_vector(x, src.cols, 12345)
for world in (8, 4, 2):
    for mode in ("auto", "0"):
        if mode == "auto": os.environ.pop("SMVP_CSR_RELABEL", None)
        else: os.environ["SMVP_CSR_RELABEL"] = mode
        op = sdist.RowBlockCsr(eng, src, world // 2, world, eng.CSR_AUTO, exchange="none")
        op.set_x(x)
        ms = timeit(lambda: op.multiply())
        print("world %d rank %d relabel=%-4s x_relabel %2d: %.3f ms (local nnz %d)" % (world, world // 2, mode, op.A.x_relabel, ms, op.local_nnz), flush=True)
        op.free()
