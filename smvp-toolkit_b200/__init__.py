"""smvp-toolkit_b200 -- B200-native engine for the CSR / TJDS SpMV path of circletile/smvp-toolkit.

Only what the hot path needs:
    csrc/      CUDA kernels (sm_100a) + the C ABI (include/smvp_cuda.h, include/smvp_synth.h)
    host/      the C host side: Matrix Market loader, report writer, smvp-toolkit-cli
    engine.py  ctypes mirror of the reference's interface for the path
    dist.py    multi-GPU partitioning (row blocks / column blocks) over torch.distributed

The directory name contains a hyphen; import it as `smvp_toolkit_b200` (shim at the repo root).
"""
from .engine import *  # noqa: F401,F403
from .engine import lib, LIB_PATH, SIGNATURES  # noqa: F401
