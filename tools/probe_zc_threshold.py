import os, sys, time, subprocess
import numpy as np
if len(sys.argv) == 1:
    for zc in ("0", "100000000"):
        subprocess.run([sys.executable, __file__, zc], env=dict(os.environ, PVV_ZERO_COPY_MAX=zc, PVV_ZERO_COPY_MAX_IVG=zc))
    sys.exit(0)
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from py_vollib_vectorized_b200 import _engine as E
from py_vollib_vectorized_b200.util.chains import chain_cfg2
for n in (100_000, 500_000, 1_000_000, 2_000_000, 4_000_000, 8_000_000):
    c = chain_cfg2(n, 0)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h = {k: pin(c[k]) for k in ("flag", "S", "K", "t", "r", "sigma")}
    on = {k: pin(np.empty(n)) for k in ("price",) + E.GREEK_NAMES}
    col = lambda k: (h[k], 1)
    price = E.price("black_scholes", n, col("flag"), col("S"), col("K"), col("t"), col("r"), None, col("sigma"))
    ivo = pin(np.empty(n)); st = pin(np.empty(n, dtype=np.uint8)); hp = pin(price)
    fs = {"greeks": lambda: E.greeks("black_scholes", n, col("flag"), col("S"), col("K"), col("t"), col("r"), None, col("sigma"), with_price=True, outs=on),
          "iv_greeks": lambda: E.iv_greeks("black_scholes", n, col("flag"), (hp, 1), col("S"), col("K"), col("t"), col("r"), None, status=st,
                                           outs={k: (ivo if k == "iv" else on[k]) for k in ("iv",) + E.GREEK_NAMES}),
          "iv": lambda: E.iv("black_scholes", n, col("flag"), (hp, 1), col("S"), col("K"), col("t"), col("r"), None, out=ivo, status=st)}
    for name, f in fs.items():
        for _ in range(3): f()
        reps = 10; t0 = time.perf_counter()
        for _ in range(reps): f()
        dt = (time.perf_counter() - t0) / reps
        print(f"PVV_ZERO_COPY_MAX={os.environ['PVV_ZERO_COPY_MAX']:>10s} n={n:8d} {name:6s} {dt*1e3:8.3f} ms {n/dt/1e9:.3f} G/s", flush=True)
