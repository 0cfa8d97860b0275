import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from py_vollib_vectorized_b200 import _engine as E
from py_vollib_vectorized_b200.util.chains import chain_cfg2
from oracle import get_oracle
O = get_oracle()
n = 1_000_000
c = chain_cfg2(n, seed=5)
col = lambda a: (np.ascontiguousarray(a), 1)
out = E.analytic_greeks("black_scholes", n, col(c["flag"]), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), None, col(c["sigma"]), with_price=True)
ref = O.analytic_greeks("black_scholes", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], threads=O.max_threads)
for k in out:
    a = np.abs(out[k] - ref[k]); rel = a / np.maximum(np.abs(ref[k]), 1e-300)
    i = int(np.argmax(rel))
    print(k, "max abs %.2e  max rel %.2e  p99 rel %.2e  p50 rel %.2e | worst row: ref %.6e got %.6e S %.2f K %.2f t %.4f sig %.3f flag %d" % (a.max(), rel.max(), np.quantile(rel, .99), np.quantile(rel, .5), ref[k][i], out[k][i], c["S"][i], c["K"][i], c["t"][i], c["sigma"][i], c["flag"][i]))
