"""GPU-box experiment: how much do rare asymptotic-region items (x < -10 s) cost the tiled kernels?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from py_vollib_vectorized_b200 import _engine as E
from py_vollib_vectorized_b200.util.chains import chain_cfg2

n = 4_000_000
c = chain_cfg2(n, 0)
col = lambda a: (np.ascontiguousarray(a), 1)
price = E.price("black_scholes", n, col(c["flag"]), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), None, col(c["sigma"]))
F = c["S"] * np.exp(c["r"] * c["t"])
h = -np.abs(np.log(F / c["K"])) / (c["sigma"] * np.sqrt(c["t"]))
for label, keep in (("all rows", np.ones(n, bool)), ("rows with |h| < 8 only (no asymptotic items)", h > -8)):
    idx = np.flatnonzero(keep)
    idx = np.resize(idx, n)  # same size, rows repeated
    d = {k: torch.from_numpy(np.ascontiguousarray(c[k][idx])).cuda() for k in ("flag", "S", "K", "t", "r", "sigma")}
    d["price"] = torch.from_numpy(np.ascontiguousarray(price[idx])).cuda()
    out = torch.empty(n, dtype=torch.float64, device="cuda"); st = torch.empty(n, dtype=torch.uint8, device="cuda")
    outs = {k: torch.empty(n, dtype=torch.float64, device="cuda") for k in ("price",) + E.GREEK_NAMES}
    for name, f in (("iv", lambda: E.iv_dev("black_scholes", n, d["flag"], d["price"], d["S"], d["K"], d["t"], d["r"], None, out, st)),
                    ("greeks", lambda: E.greeks_dev("black_scholes", n, d["flag"], d["S"], d["K"], d["t"], d["r"], None, d["sigma"], outs)),
                    ("price", lambda: E.price_dev("black_scholes", n, d["flag"], d["S"], d["K"], d["t"], d["r"], None, d["sigma"], outs["price"]))):
        for _ in range(3): f()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): f()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        print(f"{label:48s} {name:7s} {n / dt / 1e9:7.3f} G/s")
