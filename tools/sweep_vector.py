import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, os.getcwd())
import torch
import smvp_toolkit_b200 as eng
from smvp_toolkit_b200 import dist as sdist
src = sdist.StencilSource(eng, 369, 369, 369)
op = sdist.RowBlockCsr(eng, src, 0, 1, eng.CSR_VECTOR, exchange="none", release_source=True)
x = torch.empty(src.cols, dtype=torch.float64, device="cuda"); eng.synth_vector(x, src.cols, 999); op.set_x(x)
nb = op.global_bytes_per_mult
def timeit(fn, steps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / steps
ref = None
for wide in ("1", "0"):
    os.environ["SMVP_VECTOR_WIDE"] = wide
    ms = timeit(lambda: op.A.mult_device(x, op.y_local, eng.CSR_VECTOR))
    y = op.y_local.clone()
    if ref is None: ref = y
    print("vector wide=%s: %.3f ms %.1f GB/s rel=%.2e" % (wide, ms, nb / ms / 1e6, float(torch.linalg.norm(y - ref) / torch.linalg.norm(ref))), flush=True)
