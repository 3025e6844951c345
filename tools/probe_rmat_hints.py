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
r, c, v = src.row_block(0, src.rows)
x = torch.empty(src.cols, dtype=torch.float64, device="cuda"); eng.synth_vector(x, src.cols, 12345)
y = torch.empty(src.rows, dtype=torch.float64, device="cuda")
A = eng.CsrMatrix.build_device(r, c, v, src.rows, src.cols, r.n)
A.set_x_device(x)
cfgs = [("hints on, persist 0", "1", "0"), ("hints on, persist 16", "1", "16"), ("hints on, persist 32", "1", "32"),
        ("hints off, persist 0", "0", "0"), ("hints off, persist 64", "0", "64")]
res = {k[0]: [] for k in cfgs}
for rep in range(4):
    for name, h, p in cfgs:
        os.environ["SMVP_RANKED_HINTS"], os.environ["SMVP_L2_PERSIST_MB"] = h, p
        res[name].append(timeit(lambda: A.mult_device(None, y, eng.CSR_MERGE)))
for name, ts in res.items():
    print("%-24s %s  median %.3f ms" % (name, " ".join("%.3f" % t for t in ts), sorted(ts)[len(ts) // 2]), flush=True)
