"""GPU-box experiment: which mix of DMA copies and zero-copy (UVA) accesses moves a host-resident call fastest?
   A  zero-copy in + zero-copy out, one launch        (what pvv_*_host does for short calls)
   B  DMA in (chunked, S streams) + kernel writes its outputs straight to pinned host memory
   C  kernel reads pinned host memory + outputs to device, DMA out (chunked)
   D  DMA in + DMA out (the chunked pipeline of pvv_*_host)"""
import ctypes, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from py_vollib_vectorized_b200 import _engine as E
from py_vollib_vectorized_b200.util.chains import chain_cfg2
L = E.lib()
vp = ctypes.c_void_p
IN = ("flag", "S", "K", "t", "r", "sigma")
OUT = ("price",) + E.GREEK_NAMES
stride = np.ones(7, dtype=np.int64)


def launch(n, ins, outs, stream):
    P = lambda t_: vp(t_.data_ptr())
    rc = L.pvv_greeks_dev(1, n, P(ins["flag"]), P(ins["S"]), P(ins["K"]), P(ins["t"]), P(ins["r"]), None, P(ins["sigma"]), E._hp(stride),
                          *[P(outs[k]) for k in OUT], None, vp(stream.cuda_stream))
    assert rc == 0, rc


for n in (1_000_000, 4_000_000, 16_000_000):
    c = chain_cfg2(n, 0)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h = {k: pin(c[k]) for k in IN}
    ho = {k: pin(np.empty(n)) for k in OUT}
    res = {}
    for S_, chunk in ((1, n), (2, n // 2), (3, n // 4), (3, n // 8), (4, n // 16)):
        if chunk < 100_000:
            continue
        streams = [torch.cuda.Stream() for _ in range(S_)]
        din = [{k: torch.empty(chunk, dtype=h[k].dtype, device="cuda") for k in IN} for _ in range(S_)]
        dout = [{k: torch.empty(chunk, dtype=torch.float64, device="cuda") for k in OUT} for _ in range(S_)]

        def run(mode):
            for ci, a in enumerate(range(0, n, chunk)):
                m = min(chunk, n - a)
                s = streams[ci % S_]
                with torch.cuda.stream(s):
                    if mode in "BD":
                        ins = {k: din[ci % S_][k][:m] for k in IN}
                        for k in IN:
                            ins[k].copy_(h[k][a:a + m], non_blocking=True)
                    else:
                        ins = {k: h[k][a:a + m] for k in IN}
                    if mode in "CD":
                        outs = {k: dout[ci % S_][k][:m] for k in OUT}
                    else:
                        outs = {k: ho[k][a:a + m] for k in OUT}
                    launch(m, ins, outs, s)
                    if mode in "CD":
                        for k in OUT:
                            ho[k][a:a + m].copy_(outs[k], non_blocking=True)
            torch.cuda.synchronize()

        for mode in "ABCD":
            if mode == "A" and S_ > 2:
                continue
            for _ in range(2):
                run(mode)
            reps = 8
            t0 = time.perf_counter()
            for _ in range(reps):
                run(mode)
            dt = (time.perf_counter() - t0) / reps
            print(f"n={n:9d} streams={S_} chunk={chunk:9d} mode {mode}: {dt*1e3:8.3f} ms {n/dt/1e9:.3f} G/s {n*89/dt/1e9:5.1f} GB/s", flush=True)
