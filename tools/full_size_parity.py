"""GPU box: BASELINE.json's full-size configurations 4 and 5, bit-compared with the CPU oracle (all host threads).

    python tools/full_size_parity.py [--n 100000000] > profiles/r02_full_size_parity.log

cfg4: one 100 M-contract mixed call/put chain (the strong-scaling chain of bench.py: seeded 12.5 M blocks), IV then the
      greeks at the solved volatility through the HOST entry point (pinned streaming pipeline), block by block;
cfg5: the 2000 strikes x 500 expiries x 100 underlyings Black-76 surface (100 M rows) from a pandas frame with a string
      flag column through price_dataframe(model="black", sigma_col=...) in one call.
Size-independent properties on top: price -> IV -> price round trip, status byte census."""
import argparse
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (make_chain / BLOCK: the same seeded blocks as the benchmark)
from py_vollib_vectorized_b200 import _engine as E  # noqa: E402
from oracle import get_oracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100_000_000)
ap.add_argument("--skip-surface", action="store_true")
a = ap.parse_args()
O = get_oracle()
th = O.max_threads
col = lambda v: (np.ascontiguousarray(v), 1)  # noqa: E731


def same(x, y):
    nx, ny = np.isnan(x), np.isnan(y)
    return (nx == ny) & (nx | (x == y))


print(f"host threads for the oracle: {th}; contracts: {a.n}", flush=True)
# ---- cfg4 ---------------------------------------------------------------------------------------------
bad_total, rows, t_gpu, t_cpu = 0, 0, 0.0, 0.0
census = np.zeros(256, dtype=np.int64)
rt_max = 0.0
for b in range((a.n + bench.BLOCK - 1) // bench.BLOCK):
    m = min(bench.BLOCK, a.n - b * bench.BLOCK)
    c = bench.make_chain("iv_greeks", m, seed=7000 + b)
    price = E.price("black_scholes", m, col(c["flag"]), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), None, col(c["sigma"]))
    t0 = time.perf_counter()
    outs, status, zde = E.iv_greeks("black_scholes", m, col(c["flag"]), col(price), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), None)
    t_gpu += time.perf_counter() - t0
    t0 = time.perf_counter()
    ref, st, _ = O.iv_greeks("black_scholes", price, c["S"], c["K"], c["t"], c["r"], c["flag"], threads=th)
    t_cpu += time.perf_counter() - t0
    bad = {k: int((~same(ref[k], outs[k])).sum()) for k in ref}
    bad["status"] = int((st != status).sum())
    bad_total += sum(bad.values())
    rows += m
    census += np.bincount(status, minlength=256)
    ok = np.isfinite(outs["iv"]) & (outs["iv"] > 0)
    back = E.price("black_scholes", m, col(c["flag"]), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), None, col(np.where(ok, outs["iv"], 0.2)))
    rel = np.abs(back - price)[ok] / np.maximum(price[ok], 1e-300)
    rt_max = max(rt_max, float(np.quantile(rel, 0.999)))
    print(f"cfg4 block {b}: {m} contracts, mismatches {bad}, zero-division index {zde}", flush=True)
print(f"cfg4 TOTAL: {rows} contracts, {bad_total} mismatching values (iv, 5 greeks, status) against the oracle; "
      f"engine (host entry point, pageable numpy in, pinned out) {rows / t_gpu / 1e6:.0f} M contracts/s, oracle on {th} threads {rows / t_cpu / 1e6:.1f} M contracts/s")
print(f"cfg4 status census: ok {census[0]}, below intrinsic {census[1]}, above maximum {census[2]}, other {int(census.sum() - census[:3].sum())}; "
      f"price -> IV -> price round trip, 99.9th percentile relative error {rt_max:.2e}", flush=True)

# ---- cfg5 ---------------------------------------------------------------------------------------------
if not a.skip_surface:
    import pandas as pd
    import py_vollib_vectorized_b200 as pvv
    ne = max(1, a.n // 200000)
    c = bench._chains_module().chain_cfg5(2000, ne, 100)
    n = c["S"].size
    frame = pd.DataFrame({"Flag": np.where(c["flag"] > 0, "c", "p"), "F": c["S"], "K": c["K"], "T": c["t"], "R": c["r"], "IV": c["sigma"]})
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = pvv.price_dataframe(frame, flag_col="Flag", underlying_price_col="F", strike_col="K", annualized_tte_col="T", riskfree_rate_col="R",
                                  sigma_col="IV", model="black", inplace=False)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    p, _ = O.price("black", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], threads=th)
    g, _ = O.greeks("black", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], threads=th)
    dc = time.perf_counter() - t0
    bad = {"Price": int((~same(p, res["Price"].to_numpy())).sum())}
    for k in ("delta", "gamma", "theta", "rho", "vega"):
        bad[k] = int((~same(g[k], res[k].to_numpy())).sum())
    print(f"cfg5: {n} rows (2000 x {ne} x 100), price_dataframe from pandas {dt:.2f} s = {n / dt / 1e6:.0f} M rows/s; oracle {dc:.1f} s; mismatches {bad}")
    print("cfg5 TOTAL MISMATCHES", sum(bad.values()))
print("FULL SIZE TOTAL MISMATCHES", bad_total)
