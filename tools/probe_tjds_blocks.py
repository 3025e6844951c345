import os, sys
sys.path.insert(0, os.getcwd())
import torch
import smvp_toolkit_b200 as eng
from smvp_toolkit_b200 import dist as sdist
def timeit(fn, steps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / steps
src = sdist.RmatSource(eng, 26, 16 << 26)
x = torch.empty(src.cols, dtype=torch.float64, device="cuda"); eng.synth_vector(x, src.cols, 12345)
y = torch.empty(src.rows, dtype=torch.float64, device="cuda")
for world in (1, 8):
    b = sdist.balanced_bounds(src.col_prefix, src.cols, world)
    c0, c1 = b[world // 2], b[world // 2 + 1]
    r, c, v = src.col_block(c0, c1)
    for mode in ("auto", "0", "1"):
        if mode == "auto": os.environ.pop("SMVP_TJDS_RELABEL", None)
        else: os.environ["SMVP_TJDS_RELABEL"] = mode
        T = eng.TjdsMatrix.build_device(r, c, v, src.rows, c1 - c0, r.n)
        T.set_x_device(x[c0:c1])
        ms = timeit(lambda: T.mult_device(y, eng.TJDS_ATOMIC))
        print("1/%d column block, relabel=%-4s y_relabel %2d: atomic %.3f ms" % (world, mode, T.y_relabel, ms), flush=True)
        T.free()
    for a in (r, c, v): a.free()
