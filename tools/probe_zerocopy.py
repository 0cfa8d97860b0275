"""GPU-box experiment: kernels reading/writing PINNED HOST memory directly (UVA zero-copy) vs the chunked
copy pipeline, end to end."""
import ctypes, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from py_vollib_vectorized_b200 import _engine as E
from py_vollib_vectorized_b200.util.chains import chain_cfg2
L = E.lib()
vp = ctypes.c_void_p
for n in (1_000_000, 10_000_000):
    c = chain_cfg2(n, 0)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h = {k: pin(c[k]) for k in ("flag", "S", "K", "t", "r", "sigma")}
    outs = {k: pin(np.empty(n)) for k in ("price",) + E.GREEK_NAMES}
    stride = np.ones(7, dtype=np.int64)
    P = lambda t: vp(t.data_ptr())
    torch.cuda.synchronize()
    def zero_copy():
        rc = L.pvv_greeks_dev(1, n, P(h["flag"]), P(h["S"]), P(h["K"]), P(h["t"]), P(h["r"]), None, P(h["sigma"]), E._hp(stride),
                              P(outs["price"]), P(outs["delta"]), P(outs["gamma"]), P(outs["theta"]), P(outs["rho"]), P(outs["vega"]), None, None)
        assert rc == 0, rc
        torch.cuda.synchronize()
    hn = {k: v.numpy() for k, v in h.items()}; on = {k: v.numpy() for k, v in outs.items()}
    col = lambda k: (hn[k], 1)
    def pipeline():
        E.greeks("black_scholes", n, col("flag"), col("S"), col("K"), col("t"), col("r"), None, col("sigma"), with_price=True, outs=on)
    for name, f in (("zero-copy single launch", zero_copy), ("chunked copy pipeline", pipeline)):
        for _ in range(3): f()
        t0 = time.perf_counter(); reps = 10
        for _ in range(reps): f()
        dt = (time.perf_counter() - t0) / reps
        print(f"n={n} greeks {name:26s} {dt*1e3:8.3f} ms  {n/dt/1e9:.3f} G/s  {n*89/dt/1e9:.1f} GB/s", flush=True)
    ref = on["delta"].copy(); zero_copy(); print("same result:", np.array_equal(ref, outs["delta"].numpy(), equal_nan=True))
