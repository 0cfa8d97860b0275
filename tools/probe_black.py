import os, sys, time
import numpy as np, pandas as pd
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import py_vollib_vectorized_b200 as pvv
from py_vollib_vectorized_b200 import _engine as E
from py_vollib_vectorized_b200.util import data_format as D
from py_vollib_vectorized_b200.util.chains import chain_cfg5
c = chain_cfg5(2000, 50, 100); n = c["S"].size
df = pd.DataFrame({"Flag": np.where(c["flag"] > 0, "c", "p"), "F": c["S"], "K": c["K"], "T": c["t"], "R": c["r"], "IV": c["sigma"]})
flag = D._preprocess_flags(df["Flag"])
cols = D.maybe_format_data_and_broadcast(df["F"], df["K"], df["T"], df["R"], flag, df["IV"])
for rep in range(4):
    t0 = time.perf_counter(); pc = E.pin_columns(cols.cols); S, K, t, r, f, sg = pc
    t1 = time.perf_counter(); p = E.price("black", n, f, S, K, t, None, None, sg)
    t2 = time.perf_counter(); g = E.greeks("black", n, f, S, K, t, r, None, sg)
    t3 = time.perf_counter(); g2 = E.greeks("black_scholes", n, f, S, K, t, r, None, sg, with_price=True)
    t4 = time.perf_counter()
    print("pin %.1f  price(black) %.1f  greeks(black) %.1f  greeks(bs,+price) %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3), flush=True)
kw = dict(flag_col="Flag", underlying_price_col="F", strike_col="K", annualized_tte_col="T", riskfree_rate_col="R")
for rep in range(4):
    t0 = time.perf_counter(); r1 = pvv.price_dataframe(df, **kw, sigma_col="IV", model="black"); t1 = time.perf_counter()
    r2 = pvv.price_dataframe(df, **kw, sigma_col="IV", model="black_scholes"); t2 = time.perf_counter()
    print("price_dataframe black %.1f  bs %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3), flush=True)
    del r1, r2
