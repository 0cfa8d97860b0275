"""GPU-box experiment: where the wall time of the PUBLIC API goes for pageable numpy inputs."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import py_vollib_vectorized_b200 as pvv  # noqa: E402
from py_vollib_vectorized_b200 import _engine as E  # noqa: E402
from py_vollib_vectorized_b200.util.chains import chain_cfg2  # noqa: E402
from py_vollib_vectorized_b200.util.data_format import _preprocess_flags  # noqa: E402


def timeit(f, reps=5):
    f()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    return (time.perf_counter() - t0) / reps


cudart = ctypes.CDLL("libcudart.so.12")
for n in (1_000_000, 10_000_000):
    c = chain_cfg2(n, 0)
    fl_str = np.where(c["flag"] > 0, "c", "p").astype(object)
    col = lambda k: (c[k], 1)  # noqa: E731
    print(f"--- n = {n}")
    print("flag encode (object str array): %.2f ms" % (1e3 * timeit(lambda: _preprocess_flags(fl_str))))
    print("public get_all_greeks, str flags, pageable: %.2f ms" % (1e3 * timeit(lambda: pvv.get_all_greeks(fl_str, c["S"], c["K"], c["t"], c["r"], c["sigma"], return_as="dict"))))
    print("public get_all_greeks, int8 flags, pageable: %.2f ms" % (1e3 * timeit(lambda: pvv.get_all_greeks(c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], return_as="dict"))))
    print("engine.greeks pageable in/out: %.2f ms" % (1e3 * timeit(lambda: E.greeks("black_scholes", n, col("flag"), col("S"), col("K"), col("t"), col("r"), None, col("sigma")))))
    outs = {k: np.empty(n) for k in E.GREEK_NAMES}
    print("engine.greeks pageable in, preallocated pageable out: %.2f ms" % (1e3 * timeit(lambda: E.greeks("black_scholes", n, col("flag"), col("S"), col("K"), col("t"), col("r"), None, col("sigma"), outs=outs))))
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    h = {k: pin(c[k]) for k in ("flag", "S", "K", "t", "r", "sigma")}
    pouts = {k: pin(np.empty(n)) for k in E.GREEK_NAMES}
    pc = lambda k: (h[k], 1)  # noqa: E731
    print("engine.greeks pinned in/out: %.2f ms" % (1e3 * timeit(lambda: E.greeks("black_scholes", n, pc("flag"), pc("S"), pc("K"), pc("t"), pc("r"), None, pc("sigma"), outs=pouts))))
    print("engine.greeks pinned in, pageable out: %.2f ms" % (1e3 * timeit(lambda: E.greeks("black_scholes", n, pc("flag"), pc("S"), pc("K"), pc("t"), pc("r"), None, pc("sigma"), outs=outs))))
    print("engine.greeks pageable in, pinned out: %.2f ms" % (1e3 * timeit(lambda: E.greeks("black_scholes", n, col("flag"), col("S"), col("K"), col("t"), col("r"), None, col("sigma"), outs=pouts))))
    a = c["S"]

    def reg():
        assert cudart.cudaHostRegister(ctypes.c_void_p(a.ctypes.data), ctypes.c_size_t(a.nbytes), 0) == 0
        assert cudart.cudaHostUnregister(ctypes.c_void_p(a.ctypes.data)) == 0
    print("cudaHostRegister+Unregister of one %d MB column: %.2f ms" % (a.nbytes >> 20, 1e3 * timeit(reg)))
    print("np.empty(n) x5 + first touch: %.2f ms" % (1e3 * timeit(lambda: [np.empty(n).fill(0) for _ in range(5)])))
    print("memcpy pageable->pinned one column: %.2f ms" % (1e3 * timeit(lambda: np.copyto(h["S"], c["S"]))))
    print("torch pin_memory alloc of one column: %.2f ms" % (1e3 * timeit(lambda: torch.empty(n, dtype=torch.float64, pin_memory=True))))
