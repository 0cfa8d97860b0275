"""FP64 issue rate against instruction-level parallelism (chains per thread) and resident warps per SM:
   python tools/probe_ilp.py        (on the GPU box)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from py_vollib_vectorized_b200 import _engine as E  # noqa: E402

torch.cuda.set_device(0)
peak = 148 * 64 * 1.965e9
print("T DP-instr/s (fraction of 148 SM x 64 lanes x 1.965 GHz); rows: warps per SM, columns: chains per thread")
print("warps " + " ".join(f"{c:>14d}" for c in (1, 2, 4, 8)))
for w in (8, 16, 24, 32, 48, 64):
    row = []
    for c in (1, 2, 4, 8):
        r = E.fp64_ilp_probe(c, w)
        row.append(f"{r / 1e12:6.2f} ({r / peak * 100:3.0f}%)")
    print(f"{w:5d} " + " ".join(f"{x:>14s}" for x in row))
