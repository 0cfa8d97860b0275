"""IV -> greeks (K4): the single fused kernel (PVV_IVG_FUSED=1) against the two-launch route pvv_iv_greeks_dev takes by
default (K3, then K2 at the solved volatility with its zero divisions ORed into the status bytes), and against the two
public calls made by hand.  Same cfg2 chain as bench.py's iv_greeks workload, device-resident, CUDA events; also checks
that the routes agree bit for bit.   python tools/probe_k4_split.py [n]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from py_vollib_vectorized_b200 import _engine as E  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
model = "black_scholes"
dev = torch.device("cuda:0")
c = bench.make_chain("iv_greeks", n, seed=1000)
d = {k: torch.from_numpy(v).to(dev) for k, v in c.items() if isinstance(v, np.ndarray) and v.ndim == 1}
price = torch.empty(n, dtype=torch.float64, device=dev)
E.price_dev(model, n, d["flag"], d["S"], d["K"], d["t"], d["r"], d["q"], d["sigma"], price)
new = lambda: torch.empty(n, dtype=torch.float64, device=dev)
o4 = {k: new() for k in E.GREEK_NAMES}
o2 = {k: new() for k in E.GREEK_NAMES}
iv4, iv3 = new(), new()
st4 = torch.empty(n, dtype=torch.uint8, device=dev)
st3 = torch.empty(n, dtype=torch.uint8, device=dev)


o5 = {k: new() for k in E.GREEK_NAMES}
iv5 = new()
st5 = torch.empty(n, dtype=torch.uint8, device=dev)


def fused():
    os.environ["PVV_IVG_FUSED"] = "1"
    E.iv_greeks_dev(model, n, d["flag"], price, d["S"], d["K"], d["t"], d["r"], d["q"], iv4, st4, o4)
    os.environ["PVV_IVG_FUSED"] = "0"


def routed():
    E.iv_greeks_dev(model, n, d["flag"], price, d["S"], d["K"], d["t"], d["r"], d["q"], iv5, st5, o5)


def k3():
    E.iv_dev(model, n, d["flag"], price, d["S"], d["K"], d["t"], d["r"], d["q"], iv3, st3)


def k2():
    E.greeks_dev(model, n, d["flag"], d["S"], d["K"], d["t"], d["r"], d["q"], iv3, o2)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


t4, t5, t3, t2 = timed(fused), timed(routed), timed(k3), timed(k2)
print(f"n={n}: fused K4 {t4:.3f} ms = {n / t4 / 1e6:.3f} G/s | pvv_iv_greeks_dev {t5:.3f} ms = {n / t5 / 1e6:.3f} G/s | K3 {t3:.3f} ms ({n / t3 / 1e6:.2f} G/s) + K2 {t2:.3f} ms ({n / t2 / 1e6:.2f} G/s)"
      f" = {t3 + t2:.3f} ms = {n / (t3 + t2) / 1e6:.3f} G/s")
same = lambda a, b: int((a.view(torch.int64) != b.view(torch.int64)).sum())
print("bit mismatches fused vs by hand: iv", same(iv4, iv3), {k: same(o4[k], o2[k]) for k in E.GREEK_NAMES}, "status:", int((st4 != st3).sum()))
print("bit mismatches fused vs routed:  iv", same(iv4, iv5), {k: same(o4[k], o5[k]) for k in E.GREEK_NAMES}, "status:", int((st4 != st5).sum()))
