"""GPU-box experiment: end-to-end (pinned host buffers) throughput of pvv_greeks_host / pvv_iv_host
for several chunk sizes and stream counts.   python tools/tune_e2e.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from py_vollib_vectorized_b200 import _engine as E  # noqa: E402
from py_vollib_vectorized_b200.util.chains import chain_cfg2  # noqa: E402

pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
for n in [int(x) for x in os.environ.get("TUNE_N", "1000000,10000000").split(",")]:
    c = chain_cfg2(n, 0)
    h = {k: pin(c[k]) for k in ("flag", "S", "K", "t", "r", "sigma", "q")}
    out = {k: pin(np.empty(n)) for k in ("price", "delta", "gamma", "theta", "rho", "vega")}
    col = lambda k: (h[k], 1)  # noqa: E731
    for chunk in (65536, 131072, 262144, 524288, 1048576, 4194304):
        if chunk > n:
            continue
        for ns in (2, 3, 4, 6):
            ctx = E.get_ctx(0, chunk=chunk, nstreams=ns)
            f = lambda: E.greeks("black_scholes", n, col("flag"), col("S"), col("K"), col("t"), col("r"), None, col("sigma"),  # noqa: E731
                                 with_price=True, ctx=ctx, outs=out)
            for _ in range(3):
                f()
            reps = 20 if n <= 1_000_000 else 5
            t0 = time.perf_counter()
            for _ in range(reps):
                f()
            dt = (time.perf_counter() - t0) / reps
            print(f"greeks n={n} chunk={chunk} streams={ns}: {dt * 1e3:.3f} ms  {n / dt / 1e9:.3f} G/s  {n * 89 / dt / 1e9:.1f} GB/s", flush=True)
            ctx.close()
            E._ctxs.pop((0, chunk, ns), None)
