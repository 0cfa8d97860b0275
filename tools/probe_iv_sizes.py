"""K3 (implied volatility) device-resident throughput against the column length: how the tile plan (iv_tile_plan:
even tiles over whole waves; PVV_IV_EVEN_TILES=0 = full tiles + remainder) and the shape threshold (PVV_IV_BIG_MIN)
behave between a thousand and ten million contracts.  Both switches are read once per process:
    PVV_IV_EVEN_TILES=0 python tools/probe_iv_sizes.py ; python tools/probe_iv_sizes.py
Prints one line per size and a checksum of the outputs (must not depend on the plan)."""
import os
import sys
import zlib

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from py_vollib_vectorized_b200 import _engine as E  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [1000, 8192, 32768, 131072, 262144, 524288, 1_000_000, 2_000_000, 3_000_000, 4_000_000, 10_000_000]
model = "black_scholes_merton"
dev = torch.device("cuda:0")
col = lambda a: (np.ascontiguousarray(a), 1)


def pricer(f, S, K, t, r, sg, q):
    return E.price(model, S.size, col(f), col(S), col(K), col(t), col(r), col(q), col(sg), raise_zde=False)


big = bench.make_chain("iv", max(sizes), seed=1, pricer=pricer)
tag = f"even={os.environ.get('PVV_IV_EVEN_TILES', '1')} big_min={os.environ.get('PVV_IV_BIG_MIN', 'default')}"
for n in sizes:
    d = {k: torch.from_numpy(np.ascontiguousarray(v[:n])).to(dev) for k, v in big.items() if isinstance(v, np.ndarray) and v.ndim == 1}
    iv = torch.empty(n, dtype=torch.float64, device=dev)
    st = torch.empty(n, dtype=torch.uint8, device=dev)
    call = lambda: E.iv_dev(model, n, d["flag"], d["price"], d["S"], d["K"], d["t"], d["r"], d["q"], iv, st)
    for _ in range(5):
        call()
    torch.cuda.synchronize()
    reps = max(10, min(200, int(2e8 // n)))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        call()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    crc = zlib.crc32(iv.cpu().numpy().tobytes()) ^ zlib.crc32(st.cpu().numpy().tobytes())
    print(f"{tag} n={n:9d}: {ms * 1e3:8.1f} us  {n / ms / 1e6:6.3f} G/s  crc {crc:08x}", flush=True)
