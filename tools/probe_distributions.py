"""GPU-box experiment: how the device-resident throughput of K2 (price + FD greeks) and K3 (IV) depends on the
input distribution: the seeded random chain of the benchmark (cfg2), a strike-sorted surface grid (cfg5: every
warp sees neighbouring strikes of one expiry), and the stress chain that populates all four regions / branches."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from py_vollib_vectorized_b200 import _engine as E
from py_vollib_vectorized_b200.util.chains import chain_cfg2, chain_cfg5, chain_wide

n = 4_000_000
chains = {"cfg2 random": chain_cfg2(n, 0), "cfg5 sorted grid": chain_cfg5(2000, 20, 100), "wide stress": chain_wide(n, 7)}
for name, c in chains.items():
    m = c["S"].size
    d = {k: torch.from_numpy(np.ascontiguousarray(c[k])).cuda() for k in ("flag", "S", "K", "t", "r", "sigma", "q")}
    outs = {k: torch.empty(m, dtype=torch.float64, device="cuda") for k in ("price",) + E.GREEK_NAMES}
    price = torch.empty(m, dtype=torch.float64, device="cuda")
    E.price_dev("black_scholes", m, d["flag"], d["S"], d["K"], d["t"], d["r"], None, d["sigma"], price)
    iv = torch.empty(m, dtype=torch.float64, device="cuda")
    st = torch.empty(m, dtype=torch.uint8, device="cuda")
    runs = {
        "K1 price": lambda: E.price_dev("black_scholes", m, d["flag"], d["S"], d["K"], d["t"], d["r"], None, d["sigma"], price),
        "K2 price+greeks": lambda: E.greeks_dev("black_scholes", m, d["flag"], d["S"], d["K"], d["t"], d["r"], None, d["sigma"], outs),
        "K3 IV": lambda: E.iv_dev("black_scholes", m, d["flag"], price, d["S"], d["K"], d["t"], d["r"], None, iv, st),
    }
    for kname, f in runs.items():
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name:18s} {kname:16s} n={m:8d} {ms:8.3f} ms  {m / ms / 1e6:8.3f} G/s", flush=True)
