"""GPU-box experiment: wall-time breakdown of price_dataframe on the cfg5 surface grid (pandas in/out)."""
import os
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import py_vollib_vectorized_b200 as pvv  # noqa: E402
from py_vollib_vectorized_b200 import _engine as E  # noqa: E402
from py_vollib_vectorized_b200.util import data_format as D  # noqa: E402
from py_vollib_vectorized_b200.util.chains import chain_cfg5  # noqa: E402


def T(label, f, reps=3):
    f()
    t = time.perf_counter()
    for _ in range(reps):
        r = f()
    print("%-46s %8.2f ms" % (label, (time.perf_counter() - t) / reps * 1e3), flush=True)
    return r


c = chain_cfg5(2000, 50, 100)
n = c["S"].size
df = pd.DataFrame({"Flag": np.where(c["flag"] > 0, "c", "p"), "F": c["S"], "K": c["K"], "T": c["t"], "R": c["r"], "IV": c["sigma"]})
print("rows", n, "flag dtype", df["Flag"].dtype)
kw = dict(flag_col="Flag", underlying_price_col="F", strike_col="K", annualized_tte_col="T", riskfree_rate_col="R")
T("price_dataframe(model=black, sigma_col)", lambda: pvv.price_dataframe(df, **kw, sigma_col="IV", model="black"))
T("price_dataframe(model=black_scholes, sigma_col)", lambda: pvv.price_dataframe(df, **kw, sigma_col="IV", model="black_scholes"))
flag = T("  flag encode", lambda: D._preprocess_flags(df["Flag"]))
cols = T("  broadcast", lambda: D.maybe_format_data_and_broadcast(df["F"], df["K"], df["T"], df["R"], flag, df["IV"]))
pc = T("  pin_columns (6 cols)", lambda: E.pin_columns(cols.cols))
S, K, t, r, f, sg = pc
outs = T("  engine.greeks with_price (pinned in)", lambda: E.greeks("black_scholes", n, f, S, K, t, r, None, sg, with_price=True))
T("  engine.price black (pinned in)", lambda: E.price("black", n, f, S, K, t, None, None, sg))
T("  DataFrame assembly", lambda: pd.DataFrame({k: outs[k] for k in outs}, index=df.index, copy=False))
df["Px"] = outs["price"]
import warnings
warnings.simplefilter("ignore")
T("price_dataframe(model=black_scholes, price_col)", lambda: pvv.price_dataframe(df, **kw, price_col="Px", model="black_scholes"))
