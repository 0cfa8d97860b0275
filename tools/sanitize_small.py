"""Small run of every kernel (a few CTAs each) for compute-sanitizer:
   compute-sanitizer --tool racecheck python tools/sanitize_small.py
   compute-sanitizer --tool memcheck  python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from py_vollib_vectorized_b200 import _engine as E  # noqa: E402
from py_vollib_vectorized_b200.util.chains import chain_cfg2, chain_wide  # noqa: E402

col = lambda a: (np.ascontiguousarray(a), 1)  # noqa: E731
for name, c in (("cfg2", chain_cfg2(3001, 0)), ("wide", chain_wide(2500, 7))):
    n = c["S"].size
    for model in ("black_scholes", "black_scholes_merton"):
        q = col(c["q"]) if model == "black_scholes_merton" else None
        p = E.price(model, n, col(c["flag"]), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), q, col(c["sigma"]))
        g = E.greeks(model, n, col(c["flag"]), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), q, col(c["sigma"]), with_price=True)
        iv, st, z = E.iv(model, n, col(c["flag"]), col(p), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), q)
        o, st2, z2 = E.iv_greeks(model, n, col(c["flag"]), col(p), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), q)
        a = E.analytic_greeks(model, n, col(c["flag"]), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), q, col(c["sigma"]), with_price=True)
        print(name, model, "ok", float(np.nansum(p)), float(np.nansum(iv)), int(st.sum()), z, z2)
print("done")
