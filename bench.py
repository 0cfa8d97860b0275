#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path on B200 (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload greeks|iv|iv_greeks|price]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU path (oracle port, all host threads)

One "step" = one pass of the hot path over one batch of synthetic contracts.  Default workload =
BASELINE.json configs[1]: 1 M-contract seeded Black-Scholes chain, price + all five FD greeks
(kernel K2).  Every rank owns its own chain (weak scaling, no collective on the data path).

  value   device-resident throughput (inputs already in HBM), CUDA events, max over ranks
  e2e     the same metric through the C ABI with HOST buffers (pvv_*_host, pinned memory), H2D and
          D2H inside the timed region
  roofline / fp64   achieved algorithmic GB/s against the measured HBM peak, and the measured FP64
          issue rate (the kernels are FP64-pipe bound)
  cpu_baseline      the oracle port timed on this box's host cores (rank 0, N=1 only)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# algorithmic bytes per contract: int8 flag + fp64 columns in, fp64 columns out
WORKLOADS = {
    # name: (metric, model, n_per_rank, bytes_in, bytes_out, description)
    "greeks": ("price+greeks contracts/s", "black_scholes", 1_000_000, 41, 48,
               "cfg2: 1M-contract synthetic Black-Scholes chain, price + get_all_greeks (kernel K2)"),
    "price": ("prices/s", "black_scholes", 1_000_000, 41, 8, "cfg2: 1M-contract Black-Scholes chain, price only (kernel K1)"),
    "iv": ("IV solves/s", "black_scholes_merton", 10_000_000, 49, 9,
           "cfg3: 10M-contract BSM implied volatility with deep-OTM / short-expiry / below-intrinsic strata (kernel K3)"),
    "iv_greeks": ("IV+greeks contracts/s", "black_scholes", 12_500_000, 41, 49,
                  "cfg4: mixed call/put chain, IV then greeks at the solved vol (kernel K4), 12.5M contracts per GPU"),
    "analytic": ("price+analytical-greeks contracts/s", "black_scholes", 16_000_000, 41, 48,
                 "opt-in closed-form greeks (SURVEY 8(f).3): 16M-contract Black-Scholes chain, price + 5 analytical greeks, one HBM-bound pass"),
    "surface": ("price_dataframe rows/s", "black", 10_000_000, 41, 48,
                "cfg5 (scaled 10x down in expiries): Black-76 surface grid 2000 strikes x 50 expiries x 100 underlyings via "
                "price_dataframe(model='black', sigma_col=...), e2e from a pandas frame with a string flag column"),
}
N_SETS = 4  # rotating input/output sets so that the working set exceeds the 126 MB L2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="greeks", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="contracts per rank (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def make_surface(n):
    from py_vollib_vectorized_b200.util.chains import chain_cfg5
    ne = max(1, n // (2000 * 100))
    return chain_cfg5(2000, ne, 100)


def make_chain(workload, n, seed, pricer=None):
    """cfg2 chain for price / greeks / iv_greeks; cfg3 (strata: deep OTM, short expiry, below intrinsic,
    above maximum, t < 0) for the IV workload.  `pricer` supplies cfg3's clean BSM prices."""
    from py_vollib_vectorized_b200.util.chains import chain_cfg2, chain_cfg3
    if workload == "iv" and pricer is not None:
        return chain_cfg3(n, seed=seed, pricer=pricer)
    if workload == "surface":
        return make_surface(n)
    return chain_cfg2(n, seed=seed)


# DP instructions (DFMA+DMUL+DADD+DSETP) per contract from the ncu captures under profiles/r01j_ncu_full_summary.txt:
#   algorithmic = executed on ACTIVE lanes (the work the reference's formulas need; invariant under regrouping),
#   slots       = warp-level issues x 32 / contracts (what occupies the FP64 pipe, inactive lanes included).
DP_INST_PER_CONTRACT = {"greeks": 1861.0, "price": 268.0, "iv": 1336.0, "iv_greeks": 3175.0, "surface": 1861.0 + 268.0, "analytic": 174.0}
DP_SLOTS_PER_CONTRACT = {"greeks": 2145.0, "price": 314.0, "iv": 1479.0, "iv_greeks": 3555.0, "surface": 2145.0 + 314.0, "analytic": 223.0}


# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per contract and launch, from the same ncu --set full
# captures.  K1/K2 stay below the algorithmic bytes (the outputs of a launch are still in the 126 MB L2 when the
# kernel ends); K3/K4 exceed them by their register-spill traffic and the second read of t (HBM is 5 % busy).
TRAFFIC_BYTES_PER_CONTRACT = {"greeks": 46.1, "price": 41.1, "iv": 74.3, "iv_greeks": 137.4, "analytic": 76.4, "surface": 46.1 + 41.1}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines = []
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, windows):
        """windows: [(label, t0, t1), ...] in order of preference; the first one that contains samples wins
        (a timed region shorter than the 50 ms sampling period has none: the next window is under load too)."""
        if not self.p:
            return None
        time.sleep(0.06)
        self.p.terminate()
        rows, label = [], None
        for label, t0, t1 in windows:
            rows = [ln.split(", ") for ts, ln in self.lines if t0 <= ts <= t1 + 0.06]
            if rows:
                break
        if not rows:
            rows, label = [ln.split(", ") for ts, ln in self.lines[-3:]], "last samples"
        sm, mx, reasons = [], 0.0, set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "window": label}


def cpu_baseline(workload, model, n_target_s=2.0):
    """The oracle port on this box's host cores, all threads, on a bounded sample of the workload."""
    from oracle import get_oracle
    O = get_oracle()
    threads = O.max_threads
    n = 1_000_000
    if workload == "surface":
        n = 2_000_000
    c = make_chain(workload, n, seed=0, pricer=lambda f, S, K, t, r, sg, q: O.price("black_scholes_merton", f, S, K, t, r, sg, q, threads=threads)[0])
    if workload == "iv":
        price = c["price"]
    elif workload == "iv_greeks":
        price, _ = O.price(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)

    def one():
        if workload == "surface":
            O.price("black", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], threads=threads)
            O.greeks("black", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], threads=threads)
        elif workload == "analytic":
            O.analytic_greeks(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)
        elif workload == "greeks":
            O.greeks(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)
        elif workload == "price":
            O.price(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)
        elif workload == "iv":
            O.iv(model, price, c["S"], c["K"], c["t"], c["r"], c["flag"], c["q"], threads=threads)
        else:
            O.iv_greeks(model, price, c["S"], c["K"], c["t"], c["r"], c["flag"], c["q"], threads=threads)

    one()  # warm (page faults, OpenMP pool)
    reps, t0 = 0, time.perf_counter()
    while True:
        one()
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= n_target_s and reps >= 3:
            break
    return {"value": n * reps / dt, "unit": "contracts/s", "cores": threads, "kind": "port",
            "sample": f"{reps} passes over a 1M-contract slice of the same seeded chain ({dt:.1f} s wall, OpenMP over {threads} threads; "
                      "C restatement of the reference's numba path, which itself is single-threaded)"}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the numba original cannot travel)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    metric, model, n, bi, bo, desc = WORKLOADS[args.workload]
    from oracle import get_oracle
    O = get_oracle()
    threads = O.max_threads
    n = 1_000_000  # bounded sample per step
    if args.workload == "surface":
        n = 2_000_000
    c = make_chain(args.workload, n, seed=0, pricer=lambda f, S, K, t, r, sg, q: O.price("black_scholes_merton", f, S, K, t, r, sg, q, threads=threads)[0])
    price = None
    if args.workload == "iv":
        price = c["price"]
    elif args.workload == "iv_greeks":
        price, _ = O.price(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)

    def step():
        if args.workload == "surface":
            O.price("black", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], threads=threads)
            O.greeks("black", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], threads=threads)
        elif args.workload == "analytic":
            O.analytic_greeks(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)
        elif args.workload == "greeks":
            O.greeks(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)
        elif args.workload == "price":
            O.price(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)
        elif args.workload == "iv":
            O.iv(model, price, c["S"], c["K"], c["t"], c["r"], c["flag"], c["q"], threads=threads)
        else:
            O.iv_greeks(model, price, c["S"], c["K"], c["t"], c["r"], c["flag"], c["q"], threads=threads)

    steps = min(args.steps, 40)  # keep the whole run within a few minutes on any host
    warm = min(args.warmup, 3)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    v = n * steps / dt
    sample = f"each step = one pass over a 1M-contract slice of the workload's seeded chain; {steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": v, "unit": "contracts/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": desc, "contracts_per_step": n, "model": model},
        "cpu_baseline": {"value": v, "unit": "contracts/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "contracts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from py_vollib_vectorized_b200 import _engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    metric, model, n, bytes_in, bytes_out, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    if args.workload == "surface":
        n = max(1, n // 200000) * 200000  # whole expiries of the 2000 x 100 grid
    wl = args.workload
    bsm = model == "black_scholes_merton"
    if bsm and wl != "iv":
        bytes_in += 8

    # ---- inputs: N_SETS independent seeded chains per rank, resident in HBM ------------------------
    def gpu_pricer(f, S_, K_, t_, r_, sg_, q_):
        col = lambda a: (np.ascontiguousarray(a), 1)  # noqa: E731
        return E.price("black_scholes_merton", f.size, col(f), col(S_), col(K_), col(t_), col(r_), col(q_), col(sg_), ctx=E.get_ctx(local))

    sets = []
    for k in range(N_SETS):
        c = make_chain(wl, n, seed=1000 * rank + k, pricer=gpu_pricer)
        d = {key: torch.from_numpy(np.ascontiguousarray(c[key])).to(dev) for key in ("flag", "S", "K", "t", "r", "sigma", "q")}
        d["host"] = c
        outs = {key: torch.empty(n, dtype=torch.float64, device=dev) for key in ("price", "delta", "gamma", "theta", "rho", "vega", "iv")}
        d["out"] = outs
        d["status"] = torch.empty(n, dtype=torch.uint8, device=dev)
        if wl == "iv":
            d["mprice"] = torch.from_numpy(np.ascontiguousarray(c["price"])).to(dev)
        elif wl == "iv_greeks":
            d["mprice"] = torch.empty(n, dtype=torch.float64, device=dev)
            E.price_dev(model, n, d["flag"], d["S"], d["K"], d["t"], d["r"], d["q"] if bsm else None, d["sigma"], d["mprice"])
        sets.append(d)
    torch.cuda.synchronize()
    q_of = (lambda d: d["q"]) if bsm else (lambda d: None)

    def step_dev(d):
        o = d["out"]
        if wl == "surface":
            E.price_dev("black", n, d["flag"], d["S"], d["K"], d["t"], None, None, d["sigma"], o["price"])
            E.greeks_dev("black", n, d["flag"], d["S"], d["K"], d["t"], d["r"], None, d["sigma"],
                         {k: o[k] for k in ("delta", "gamma", "theta", "rho", "vega")})
        elif wl == "analytic":
            E.analytic_greeks_dev(model, n, d["flag"], d["S"], d["K"], d["t"], d["r"], q_of(d), d["sigma"],
                                  {k: o[k] for k in ("price", "delta", "gamma", "theta", "rho", "vega")})
        elif wl == "greeks":
            E.greeks_dev(model, n, d["flag"], d["S"], d["K"], d["t"], d["r"], q_of(d), d["sigma"],
                         {k: o[k] for k in ("price", "delta", "gamma", "theta", "rho", "vega")})
        elif wl == "price":
            E.price_dev(model, n, d["flag"], d["S"], d["K"], d["t"], d["r"], q_of(d), d["sigma"], o["price"])
        elif wl == "iv":
            E.iv_dev(model, n, d["flag"], d["mprice"], d["S"], d["K"], d["t"], d["r"], q_of(d), o["iv"], d["status"])
        else:
            E.iv_greeks_dev(model, n, d["flag"], d["mprice"], d["S"], d["K"], d["t"], d["r"], q_of(d), o["iv"], d["status"],
                            {k: o[k] for k in ("delta", "gamma", "theta", "rho", "vega")})

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- measured FP64 issue rate (no FP64 entry in MEASURED_PEAKS.json) ---------------------------
    fp64_peak = E.fp64_probe(use_fma=False)
    fp64_peak_fma = E.fp64_probe(use_fma=True)

    # ---- device-resident: W warm-up + exactly K timed steps ----------------------------------------
    for i in range(max(args.warmup, 3)):
        step_dev(sets[i % N_SETS])
    barrier()
    sampler = ClockSampler(local)
    launches0 = E.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        step_dev(sets[i % N_SETS])
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = E.launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tm = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    value = world * n * args.steps / (ms * 1e-3)
    kernel_ms = ms / args.steps  # one kernel launch per step (two for the surface workload)

    # ---- end to end: host (pinned) buffers through pvv_*_host, H2D + D2H inside the timed region ----
    h = sets[0]["host"]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    hin = {k: pin(h[k]) for k in ("flag", "S", "K", "t", "r", "sigma", "q")}
    if wl in ("iv", "iv_greeks"):
        hin["price"] = pin(sets[0]["mprice"].cpu().numpy())
    hout = {k: pin(np.empty(n)) for k in ("price", "delta", "gamma", "theta", "rho", "vega", "iv")}
    hstatus = pin(np.empty(n, dtype=np.uint8))
    ctx = E.get_ctx(local)  # library defaults: 512 Ki-contract chunks, 3 streams
    col = lambda k: (hin[k], 1)  # noqa: E731
    qh = col("q") if bsm else None

    if wl == "surface":
        import pandas as pd
        import py_vollib_vectorized_b200 as pvv
        frame = pd.DataFrame({"Flag": np.where(h["flag"] > 0, "c", "p"), "F": h["S"], "K": h["K"], "T": h["t"], "R": h["r"], "IV": h["sigma"]})

    def step_host():
        if wl == "surface":
            res = pvv.price_dataframe(frame, flag_col="Flag", underlying_price_col="F", strike_col="K", annualized_tte_col="T",
                                      riskfree_rate_col="R", sigma_col="IV", model="black", inplace=False)
            assert res.shape == (n, 6)
        elif wl == "analytic":
            E.analytic_greeks(model, n, col("flag"), col("S"), col("K"), col("t"), col("r"), qh, col("sigma"), with_price=True, ctx=ctx,
                              outs={k: hout[k] for k in ("price", "delta", "gamma", "theta", "rho", "vega")})
        elif wl == "greeks":
            E.greeks(model, n, col("flag"), col("S"), col("K"), col("t"), col("r"), qh, col("sigma"), with_price=True, ctx=ctx,
                     outs={k: hout[k] for k in ("price", "delta", "gamma", "theta", "rho", "vega")})
        elif wl == "price":
            E.price(model, n, col("flag"), col("S"), col("K"), col("t"), col("r"), qh, col("sigma"), ctx=ctx, out=hout["price"])
        elif wl == "iv":
            E.iv(model, n, col("flag"), col("price"), col("S"), col("K"), col("t"), col("r"), qh, ctx=ctx, out=hout["iv"], status=hstatus)
        else:
            E.iv_greeks(model, n, col("flag"), col("price"), col("S"), col("K"), col("t"), col("r"), qh, ctx=ctx, status=hstatus,
                        outs={k: hout[k] for k in ("iv", "delta", "gamma", "theta", "rho", "vega")})

    e2e_steps = max(3, min(args.steps, 50 if wl != "surface" else 5))
    for _ in range(3):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e2e_s = t1 - t0
    clocks = sampler.stop([("device-resident timed region", t_wall0, t_wall1), ("end-to-end timed region", t0, t1)])
    if world > 1:
        tm = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_s = float(tm.item())
    e2e_value = world * n * e2e_steps / e2e_s
    h2d = n * (bytes_in)
    d2h = n * (bytes_out)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = n * (bytes_in + bytes_out) / (kernel_ms * 1e-3) / 1e9
        out = {
            "metric": metric, "value": value, "unit": "contracts/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc, "contracts_per_gpu_per_step": n, "model": model,
                       "l2": f"{N_SETS} rotating input/output sets per GPU ({N_SETS * n * (bytes_in + bytes_out) / 1e6:.0f} MB > 126 MB L2)",
                       "parallelism": f"contracts sharded over {world} GPU(s), no collective on the data path"},
            "e2e": {"value": e2e_value, "unit": "contracts/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "how": ("price_dataframe() on a pandas frame (pageable memory, string flag column): flag encode + pinned staging + pipeline + frame assembly, wall clock" if wl == "surface" else "pvv_*_host C-ABI call on pinned host buffers, wall clock: up to 1.5 Mi contracts one launch that reads and writes the host memory over PCIe (zero-copy), above that the chunked 3-stream H2D/kernel/D2H pipeline")},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": TRAFFIC_BYTES_PER_CONTRACT[wl] * n, "traffic_unit": "bytes per launch (ncu dram read+write, scaled by contracts)",
                         "algorithmic_bytes": n * (bytes_in + bytes_out), "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)",
                         "note": ("closed-form kernel: HBM is the binding roofline" if wl == "analytic" else "kernel is FP64-pipe bound, not HBM bound (SURVEY.md 8(d)); see fp64")},
            "fp64": {"achieved_dp_inst_per_s": value / world * DP_INST_PER_CONTRACT[wl], "peak_dp_inst_per_s_nofma": fp64_peak,
                     "peak_dp_inst_per_s_fma": fp64_peak_fma, "frac": value / world * DP_INST_PER_CONTRACT[wl] / fp64_peak,
                     "pipe_frac": value / world * DP_SLOTS_PER_CONTRACT[wl] / fp64_peak,
                     "dp_inst_per_contract": DP_INST_PER_CONTRACT[wl], "dp_slots_per_contract": DP_SLOTS_PER_CONTRACT[wl],
                     "how": "peak: dependent DMUL+DADD (and DFMA) chains, 8 per thread, measured in this run; frac: contracts/s x DP "
                            "instructions per contract executed on active lanes (ncu, profiles/) / peak; pipe_frac: the same with "
                            "warp-level issues x 32 (inactive lanes occupy the pipe too; compare sm__pipe_fp64_cycles_active)"},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(wl, model)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
