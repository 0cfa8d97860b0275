"""GPU-box experiment: wall-clock latency of the public API for small calls (the drop-in must not be slower than the
reference's numba loops where those are fast: a few contracts per call)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import py_vollib_vectorized_b200 as pvv
from py_vollib_vectorized_b200.util.chains import chain_cfg2

for n in (1, 10, 100, 1000, 10_000, 100_000):
    c = chain_cfg2(n, 0)
    fl = np.where(c["flag"] > 0, "c", "p")
    price = pvv.vectorized_black_scholes(fl, c["S"], c["K"], c["t"], c["r"], c["sigma"], return_as="numpy")
    fs = {
        "price": lambda: pvv.vectorized_black_scholes(fl, c["S"], c["K"], c["t"], c["r"], c["sigma"], return_as="numpy"),
        "greeks": lambda: pvv.get_all_greeks(fl, c["S"], c["K"], c["t"], c["r"], c["sigma"], model="black_scholes", return_as="dict"),
        "iv": lambda: pvv.vectorized_implied_volatility(price, c["S"], c["K"], c["t"], c["r"], fl, on_error="ignore", return_as="numpy"),
    }
    for name, f in fs.items():
        for _ in range(5):
            f()
        reps = 200 if n <= 10_000 else 50
        t0 = time.perf_counter()
        for _ in range(reps):
            f()
        dt = (time.perf_counter() - t0) / reps
        print(f"n={n:7d} {name:7s} {dt*1e6:9.1f} us/call  {n/dt/1e6:9.3f} M contracts/s", flush=True)
