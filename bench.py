#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path on B200 (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W]                      # the driver's call
    python bench.py --workload greeks|price|iv|iv_greeks|analytic|surface    # one kernel only
    python bench.py --workload iv_greeks --scaling strong [--n 100000000]    # cfg4: ONE chain split over the ranks
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                                     # the reference's own numba CPU path

One "step" = one pass of the hot path over one batch of synthetic contracts.  BASELINE.json's metric is
"IV solves/s & price+greeks contracts/s": the default line's headline (`metric`, `value`, `e2e`, `roofline`) is
BASELINE.json configs[1] — 1 M-contract seeded Black-Scholes chain, price + all five FD greeks (kernel K2) — and
its `kernels` block carries the other kernels of the path measured in the same run: `iv` (cfg3, 10 M-contract
BSM implied volatility, K3), `iv_greeks` (cfg4's per-GPU share, K4), `price` (K1) and `analytic` (K5, opt-in).

  value     device-resident throughput (inputs already in HBM), CUDA events on the launching stream, max over ranks
  e2e       the same metric through the C ABI with HOST buffers (pvv_*_host, pinned memory), H2D and D2H inside the
            timed region; `e2e.ceiling` = plain cudaMemcpyAsync of the same bytes (the link's limit for this traffic)
  roofline  the kernels are FP64-pipe bound: achieved DP instructions/s (DP instructions per
            contract from the ncu capture named in profiles/dp_per_contract.json x contracts/s) against the FP64
            issue rate measured in this run; `roofline.hbm` is the (non-binding) HBM view with the measured copy peak
  cpu_baseline  the reference's own numba path (baseline/_ref, 1 core as shipped) and the C port on all host threads

Weak scaling by default (every rank owns its own chain, no collective on the data path); --scaling strong splits ONE
chain of --n contracts into contiguous shards.  With N > 1 every rank bit-compares a sample of its results with the
CPU oracle after the timed region and the output column is gathered to rank 0 over NCCL (`parity_checked`, `gather_ms`).
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

# algorithmic bytes per contract: int8 flag + fp64 columns in, fp64 columns (+ status byte) out
WORKLOADS = {
    # name: (metric, model, n_per_rank, bytes_in, bytes_out, description)
    "greeks": ("price+greeks contracts/s", "black_scholes", 1_000_000, 41, 48,
               "cfg2: 1M-contract synthetic Black-Scholes chain, price + get_all_greeks (kernel K2)"),
    "price": ("prices/s", "black_scholes", 1_000_000, 41, 8, "cfg2: 1M-contract Black-Scholes chain, price only (kernel K1)"),
    "iv": ("IV solves/s", "black_scholes_merton", 10_000_000, 49, 9,
           "cfg3: 10M-contract BSM implied volatility with deep-OTM / short-expiry / below-intrinsic strata (kernel K3)"),
    "iv_greeks": ("IV+greeks contracts/s", "black_scholes", 12_500_000, 41, 49,
                  "cfg4: mixed call/put chain, IV then greeks at the solved vol (K4 route: K3 then K2 on one stream), 12.5M contracts per GPU"),
    "analytic": ("price+analytical-greeks contracts/s", "black_scholes", 16_000_000, 41, 48,
                 "opt-in closed-form greeks (SURVEY 8(f).3): 16M-contract Black-Scholes chain, price + 5 analytical greeks, one HBM-bound pass"),
    "surface": ("price_dataframe rows/s", "black", 10_000_000, 41, 48,
                "cfg5: Black-76 surface grid 2000 strikes x E expiries x 100 underlyings via price_dataframe(model='black', "
                "sigma_col=...), e2e from a pandas frame with a string flag column (E = rows / 200000; 500 = the full grid)"),
}
STRONG_TOTAL = {"iv_greeks": 100_000_000, "iv": 100_000_000, "greeks": 100_000_000, "price": 100_000_000, "analytic": 100_000_000}
BLOCK = 12_500_000  # a strong-scaling chain is the concatenation of seeded blocks of this size (rank-count independent)
SUB_KERNELS = ("iv", "iv_greeks", "price", "analytic")
REF_SAMPLE = {"greeks": 200_000, "price": 500_000, "iv": 200_000, "iv_greeks": 100_000, "surface": 200_000}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all"] + sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="contracts per rank (weak) or in total (strong); default: the workload's")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ceiling", action="store_true", help="(kept for compatibility: the copy-only ceiling of the e2e traffic is always measured)")
    return ap.parse_args()


def config_of(wl, n, world, scaling):
    """The `config` object: identical in both arms (b200 and reference) for the same command line."""
    metric, model, _, bi, bo, desc = WORKLOADS[wl]
    if model == "black_scholes_merton" and wl != "iv":
        bi += 8
    per = "contracts_per_gpu_per_step" if scaling == "weak" else "contracts_per_step_total"
    sets = n_sets(wl, n)
    return {"workload": desc, per: n, "model": model,
            "l2": f"{sets} rotating input/output set(s) per GPU ({sets * (n if scaling == 'weak' else n // world) * (bi + bo) / 1e6:.0f} MB per GPU > 126 MB L2)",
            "parallelism": f"contracts sharded over {world} GPU(s), contiguous shards, no collective on the data path"}


def n_sets(wl, n):
    """Rotating input/output sets so that consecutive steps never find their data in the 126 MB L2."""
    bi, bo = WORKLOADS[wl][3], WORKLOADS[wl][4]
    return max(1, min(4, -(-300_000_000 // (n * (bi + bo)))))


def load_dp_table():
    """DP instructions / DRAM bytes per contract, written next to the ncu captures by tools/ncu_dp_counts.py."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "dp_per_contract.json")))
    except Exception:
        return {"source": None, "kernels": {}}


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


# ---- chains -------------------------------------------------------------------------------------------------
def _chains_module():
    import importlib.util  # by path: the reference arm must not import the engine package (it would map libpvv.so)
    spec = importlib.util.spec_from_file_location("_pvv_chains", os.path.join(ROOT, "py_vollib_vectorized_b200", "util", "chains.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def make_chain(wl, n, seed, pricer=None):
    """cfg2 chain for price / greeks / iv_greeks / analytic; cfg3 (strata: deep OTM, short expiry, below intrinsic,
    above maximum, t < 0) for the IV workload (`pricer` supplies its clean BSM prices); cfg5 grid for surface."""
    m = _chains_module()
    if wl == "iv":
        return m.chain_cfg3(n, seed=seed, pricer=pricer)
    if wl == "surface":
        return m.chain_cfg5(2000, max(1, n // (2000 * 100)), 100)
    return m.chain_cfg2(n, seed=seed)


def make_shard(wl, n_total, world, rank, pricer):
    """Strong scaling: the chain is the concatenation of BLOCK-sized seeded blocks; a rank builds only the blocks
    its contiguous shard [lo, hi) touches."""
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    parts = []
    for b in range(lo // BLOCK, (hi - 1) // BLOCK + 1):
        size = min(BLOCK, n_total - b * BLOCK)
        c = make_chain(wl, size, seed=7000 + b, pricer=pricer)
        a, z = max(lo, b * BLOCK) - b * BLOCK, min(hi, b * BLOCK + size) - b * BLOCK
        parts.append({k: v[a:z] for k, v in c.items() if isinstance(v, np.ndarray) and v.ndim == 1})
    return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}


# ---- CPU legs -----------------------------------------------------------------------------------------------
def port_baseline(wl, model, n_target_s=2.0):
    """The oracle port (C restatement, OpenMP) on this box's host cores, all threads, on a bounded sample."""
    from oracle import get_oracle
    O = get_oracle()
    threads = O.max_threads
    n = 2_000_000 if wl == "surface" else 1_000_000
    c = make_chain(wl, n, seed=0, pricer=lambda f, S, K, t, r, sg, q: O.price("black_scholes_merton", f, S, K, t, r, sg, q, threads=threads)[0])
    n = c["S"].size
    price = None
    if wl == "iv":
        price = c["price"]
    elif wl == "iv_greeks":
        price, _ = O.price(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)

    def one():
        if wl == "surface":
            O.price("black", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], threads=threads)
            O.greeks("black", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], threads=threads)
        elif wl == "analytic":
            O.analytic_greeks(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)
        elif wl == "greeks":
            O.greeks(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)
        elif wl == "price":
            O.price(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)
        elif wl == "iv":
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
            "sample": f"{reps} passes over a {n // 1000}k-contract slice of the same seeded chain ({dt:.1f} s wall, OpenMP over {threads} threads; "
                      "C restatement of the reference's numba path, which itself is single-threaded)"}


def numba_reference(wl, n, steps, warmup, procs=0):
    """The UNMODIFIED reference (baseline/_ref on oracle/shim) in a process of its own: oracle/run_reference.py.
    Returns its JSON (public / loop_only contracts/s, 1 core) or {"unavailable": reason}."""
    if wl not in REF_SAMPLE:
        return {"unavailable": f"the reference has no path for workload {wl!r}"}
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "run_reference.py"), "--workload", wl, "--n", str(n), "--steps", str(steps),
           "--warmup", str(warmup)] + (["--procs", str(procs)] if procs > 1 else [])
    env = {k: v for k, v in os.environ.items() if k not in ("PYTHONPATH", "PVV_LIB")}
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=env, cwd="/tmp")
        line = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else ""
        return json.loads(line) if line.startswith("{") else {"unavailable": (res.stderr.strip().splitlines() or ["no output"])[-1][:300]}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


def cpu_baseline(wl, model):
    """cpu_baseline of the b200 arm: the reference's numba path (1 core, as shipped) on a bounded sample, with the
    C port on all host threads beside it.  Falls back to the port alone, saying why."""
    port = port_baseline(wl, model)
    ref = numba_reference(wl, REF_SAMPLE.get(wl, 0) // 2, steps=3, warmup=1, procs=os.cpu_count() or 1)
    if "public" not in ref:
        port["reference_unavailable"] = ref.get("unavailable")
        return port
    out = {"value": ref["public"], "unit": "contracts/s", "cores": 1, "kind": "reference",
           "sample": f"{ref['steps']} passes of the unmodified reference's public API (numba {ref.get('numba')}, JIT warmed on a 1k-contract call) "
                     f"over a {ref['n'] // 1000}k-contract slice of the same seeded chain; the numba loops are single-threaded as shipped",
           "loop_only": ref.get("loop_only"), "port_all_threads": port}
    if "all_cores" in ref:
        out["all_cores_multiprocessing"] = ref["all_cores"]
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (numba, public API, stock code path) on
    this box's host cores; each step = a bounded sample of the workload.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = "greeks" if args.workload == "all" else args.workload
    metric, model, n, bi, bo, desc = WORKLOADS[wl]
    world = args.gpus
    n_cfg = args.n or (STRONG_TOTAL.get(wl, n) if args.scaling == "strong" else n)
    base = {"impl": "reference", "metric": metric, "unit": "contracts/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(wl, n_cfg, world, args.scaling)}
    ref = numba_reference(wl, REF_SAMPLE.get(wl, 0), steps=args.steps, warmup=args.warmup)
    if "public" in ref:
        v = ref["public"]
        base.update({"value": v, "ms_per_step": ref["public_s_per_step"] * 1e3,
                     "cpu_baseline": {"value": v, "unit": "contracts/s", "cores": 1, "kind": "reference", "loop_only": ref.get("loop_only"),
                                      "sample": f"each step = the unmodified reference's public API (numba {ref.get('numba')}, single-threaded as shipped, JIT "
                                                f"warmed) over a {ref['n'] // 1000}k-contract slice of the workload's seeded chain; host has {ref.get('host_cpus')} CPUs"},
                     "e2e": {"value": v, "unit": "contracts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    else:  # fall back to the C port (all host threads), stating why
        from oracle import get_oracle
        O = get_oracle()
        threads = O.max_threads
        nn = 2_000_000 if wl == "surface" else 1_000_000
        c = make_chain(wl, nn, seed=0, pricer=lambda f, S, K, t, r, sg, q: O.price("black_scholes_merton", f, S, K, t, r, sg, q, threads=threads)[0])
        nn = c["S"].size
        price = c.get("price") if wl == "iv" else None
        if wl == "iv_greeks":
            price, _ = O.price(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)

        def step():
            if wl == "surface":
                O.price("black", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], threads=threads)
                O.greeks("black", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], threads=threads)
            elif wl == "analytic":
                O.analytic_greeks(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)
            elif wl == "greeks":
                O.greeks(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)
            elif wl == "price":
                O.price(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=threads)
            elif wl == "iv":
                O.iv(model, price, c["S"], c["K"], c["t"], c["r"], c["flag"], c["q"], threads=threads)
            else:
                O.iv_greeks(model, price, c["S"], c["K"], c["t"], c["r"], c["flag"], c["q"], threads=threads)

        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
        v = nn * args.steps / dt
        base.update({"value": v, "ms_per_step": dt / args.steps * 1e3,
                     "cpu_baseline": {"value": v, "unit": "contracts/s", "cores": threads, "kind": "port",
                                      "sample": f"each step = one pass of the C port (OpenMP, {threads} threads) over a {nn // 1000}k-contract slice",
                                      "reference_unavailable": ref.get("unavailable")},
                     "e2e": {"value": v, "unit": "contracts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


# ---- GPU legs -----------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from py_vollib_vectorized_b200 import _engine as E
        self.torch, self.dist, self.E, self.args = torch, dist, E, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU path")
        if self.world > 1 and os.environ.get("PVV_BENCH_PIN", "1") != "0":
            # one slice of the host cores per rank: the ranks' pinned staging and copy submission stop sharing cores
            try:
                cpus = sorted(os.sched_getaffinity(0))
                per = max(1, len(cpus) // self.world)
                os.sched_setaffinity(0, set(cpus[self.local * per:(self.local + 1) * per]) or set(cpus))
            except (AttributeError, OSError):
                pass
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = E.get_ctx(self.local)  # library defaults: 512 Ki-contract chunks, 3 streams
        self.dp = load_dp_table()
        self.fp64_peak = E.fp64_probe(use_fma=False)
        self.fp64_peak_fma = E.fp64_probe(use_fma=True)
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            self.peaks = {}

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gpu_pricer(self, f, S_, K_, t_, r_, sg_, q_):
        col = lambda a: (np.ascontiguousarray(a), 1)  # noqa: E731
        return self.E.price("black_scholes_merton", f.size, col(f), col(S_), col(K_), col(t_), col(r_), col(q_), col(sg_), ctx=self.ctx)

    def build_sets(self, wl, n, strong_total=0):
        torch, E, dev = self.torch, self.E, self.dev
        model = WORKLOADS[wl][1]
        bsm = model == "black_scholes_merton"
        sets = []
        for k in range(1 if strong_total else n_sets(wl, n)):
            if strong_total:
                c = make_shard(wl, strong_total, self.world, self.rank, self.gpu_pricer)
            else:
                c = make_chain(wl, n, seed=1000 * self.rank + k, pricer=self.gpu_pricer)
            d = {key: torch.from_numpy(np.ascontiguousarray(c[key])).to(dev) for key in ("flag", "S", "K", "t", "r", "sigma", "q")}
            d["host"] = c
            d["out"] = {key: torch.empty(n, dtype=torch.float64, device=dev) for key in ("price", "delta", "gamma", "theta", "rho", "vega", "iv")}
            d["status"] = torch.empty(n, dtype=torch.uint8, device=dev)
            if wl == "iv":
                d["mprice"] = torch.from_numpy(np.ascontiguousarray(c["price"])).to(dev)
            elif wl == "iv_greeks":
                d["mprice"] = torch.empty(n, dtype=torch.float64, device=dev)
                E.price_dev(model, n, d["flag"], d["S"], d["K"], d["t"], d["r"], d["q"] if bsm else None, d["sigma"], d["mprice"])
            sets.append(d)
        torch.cuda.synchronize()
        return sets

    def step_dev(self, wl, d, n):
        E = self.E
        model = WORKLOADS[wl][1]
        q = d["q"] if model == "black_scholes_merton" else None
        o = d["out"]
        if wl == "surface":
            E.price_dev("black", n, d["flag"], d["S"], d["K"], d["t"], None, None, d["sigma"], o["price"])
            E.greeks_dev("black", n, d["flag"], d["S"], d["K"], d["t"], d["r"], None, d["sigma"], {k: o[k] for k in ("delta", "gamma", "theta", "rho", "vega")})
        elif wl == "analytic":
            E.analytic_greeks_dev(model, n, d["flag"], d["S"], d["K"], d["t"], d["r"], q, d["sigma"],
                                  {k: o[k] for k in ("price", "delta", "gamma", "theta", "rho", "vega")})
        elif wl == "greeks":
            E.greeks_dev(model, n, d["flag"], d["S"], d["K"], d["t"], d["r"], q, d["sigma"], {k: o[k] for k in ("price", "delta", "gamma", "theta", "rho", "vega")})
        elif wl == "price":
            E.price_dev(model, n, d["flag"], d["S"], d["K"], d["t"], d["r"], q, d["sigma"], o["price"])
        elif wl == "iv":
            E.iv_dev(model, n, d["flag"], d["mprice"], d["S"], d["K"], d["t"], d["r"], q, o["iv"], d["status"])
        else:
            E.iv_greeks_dev(model, n, d["flag"], d["mprice"], d["S"], d["K"], d["t"], d["r"], q, o["iv"], d["status"],
                            {k: o[k] for k in ("delta", "gamma", "theta", "rho", "vega")})

    def time_device(self, wl, sets, n, steps, warmup, sampler_windows=None):
        """W warm-up + exactly K timed steps, CUDA events on the launching (current torch) stream."""
        torch = self.torch
        for i in range(max(warmup, 3)):
            self.step_dev(wl, sets[i % len(sets)], n)
        self.barrier()
        launches0 = self.E.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        for i in range(steps):
            self.step_dev(wl, sets[i % len(sets)], n)
        ev1.record()
        self.barrier()
        t1 = time.perf_counter()
        if sampler_windows is not None:
            sampler_windows.append(("device-resident timed region", t0, t1))
        return self.max_over_ranks(ev0.elapsed_time(ev1)), self.E.launch_count() - launches0

    def host_buffers(self, wl, d, n):
        torch = self.torch
        h = d["host"]
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
        hin = {k: pin(h[k]) for k in ("flag", "S", "K", "t", "r", "sigma", "q")}
        if wl in ("iv", "iv_greeks"):
            hin["price"] = pin(d["mprice"].cpu().numpy())
        hout = {k: pin(np.empty(n)) for k in ("price", "delta", "gamma", "theta", "rho", "vega", "iv")}
        return hin, hout, pin(np.empty(n, dtype=np.uint8))

    def time_e2e(self, wl, d, n, steps, sampler_windows=None):
        """The same metric through pvv_*_host on pinned host buffers, wall clock, H2D + D2H inside the timed region."""
        E, ctx = self.E, self.ctx
        model = WORKLOADS[wl][1]
        bsm = model == "black_scholes_merton"
        h = d["host"]
        if wl == "surface":
            import pandas as pd
            import py_vollib_vectorized_b200 as pvv
            frame = pd.DataFrame({"Flag": np.where(h["flag"] > 0, "c", "p"), "F": h["S"], "K": h["K"], "T": h["t"], "R": h["r"], "IV": h["sigma"]})

            def step():
                res = pvv.price_dataframe(frame, flag_col="Flag", underlying_price_col="F", strike_col="K", annualized_tte_col="T",
                                          riskfree_rate_col="R", sigma_col="IV", model="black", inplace=False)
                assert res.shape == (n, 6)
        else:
            hin, hout, hstatus = self.host_buffers(wl, d, n)
            col = lambda k: (hin[k], 1)  # noqa: E731
            qh = col("q") if bsm else None
            six = ("price", "delta", "gamma", "theta", "rho", "vega")

            def step():
                if wl == "analytic":
                    E.analytic_greeks(model, n, col("flag"), col("S"), col("K"), col("t"), col("r"), qh, col("sigma"), with_price=True, ctx=ctx,
                                      outs={k: hout[k] for k in six})
                elif wl == "greeks":
                    E.greeks(model, n, col("flag"), col("S"), col("K"), col("t"), col("r"), qh, col("sigma"), with_price=True, ctx=ctx,
                             outs={k: hout[k] for k in six})
                elif wl == "price":
                    E.price(model, n, col("flag"), col("S"), col("K"), col("t"), col("r"), qh, col("sigma"), ctx=ctx, out=hout["price"])
                elif wl == "iv":
                    E.iv(model, n, col("flag"), col("price"), col("S"), col("K"), col("t"), col("r"), qh, ctx=ctx, out=hout["iv"], status=hstatus)
                else:
                    E.iv_greeks(model, n, col("flag"), col("price"), col("S"), col("K"), col("t"), col("r"), qh, ctx=ctx, status=hstatus,
                                outs={k: hout[k] for k in ("iv", "delta", "gamma", "theta", "rho", "vega")})
        for _ in range(3):
            step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        self.torch.cuda.synchronize()
        t1 = time.perf_counter()
        if sampler_windows is not None:
            sampler_windows.append(("end-to-end timed region", t0, t1))
        return self.max_over_ranks(t1 - t0) / steps

    def copy_ceiling(self, n, bytes_in, bytes_out, steps=10):
        """Plain cudaMemcpyAsync of one step's traffic (H2D on one stream, D2H on another, pinned memory), all ranks at
        once: what the host link gives this byte mix with no kernel in the way.  Returns seconds per step."""
        torch = self.torch
        hin = torch.empty(n * bytes_in, dtype=torch.uint8).pin_memory()
        hout = torch.empty(n * bytes_out, dtype=torch.uint8).pin_memory()
        din = torch.empty(n * bytes_in, dtype=torch.uint8, device=self.dev)
        dout = torch.empty(n * bytes_out, dtype=torch.uint8, device=self.dev)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def step():
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
        for _ in range(2):
            step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        self.torch.cuda.synchronize()
        sec = self.max_over_ranks(time.perf_counter() - t0) / steps
        del hin, hout, din, dout
        return sec

    def parity_sample(self, wl, d, n, m=65536):
        """Bit comparison of the first m contracts of this rank's results with the CPU oracle (the checker)."""
        from oracle import get_oracle
        O = get_oracle()
        m = min(m, n)
        model = WORKLOADS[wl][1]
        h = {k: np.ascontiguousarray(v[:m]) for k, v in d["host"].items() if isinstance(v, np.ndarray) and v.ndim == 1}
        got = {k: v[:m].cpu().numpy() for k, v in d["out"].items()}
        same = lambda a, b: bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())  # noqa: E731
        th = max(1, O.max_threads // max(1, self.world))
        if wl == "greeks":
            ref, _ = O.greeks(model, h["flag"], h["S"], h["K"], h["t"], h["r"], h["sigma"], h["q"], threads=th)
            return all(same(ref[k], got[k]) for k in ref)
        if wl == "price":
            return same(O.price(model, h["flag"], h["S"], h["K"], h["t"], h["r"], h["sigma"], h["q"], threads=th)[0], got["price"])
        price = d["mprice"][:m].cpu().numpy()
        if wl == "iv":
            iv, st, _ = O.iv(model, price, h["S"], h["K"], h["t"], h["r"], h["flag"], h["q"], threads=th)
            return same(iv, got["iv"]) and bool((st == d["status"][:m].cpu().numpy()).all())
        if wl == "iv_greeks":
            ref, st, _ = O.iv_greeks(model, price, h["S"], h["K"], h["t"], h["r"], h["flag"], h["q"], threads=th)
            return all(same(ref[k], got[k]) for k in ref) and bool((st == d["status"][:m].cpu().numpy()).all())
        return None

    def gather_check(self, col, n_total):
        """distributed.gather_column over NCCL to rank 0, timed on the device; the gathered column must carry the same
        64-bit checksum (sum of the bit patterns, wrapping) as the shards."""
        torch, dist = self.torch, self.dist
        from py_vollib_vectorized_b200 import distributed as D
        local_sum = col.view(torch.int64).sum()
        dist.all_reduce(local_sum, op=dist.ReduceOp.SUM)
        D.gather_column(col, n_total)  # warm (NCCL channels)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        full = D.gather_column(col, n_total)
        e1.record()
        torch.cuda.synchronize()
        ok = True
        if self.rank == 0:
            ok = bool(full.numel() == n_total and int(full.view(torch.int64).sum().item()) == int(local_sum.item()))
        return self.max_over_ranks(e0.elapsed_time(e1)), ok

    # ---- one workload -------------------------------------------------------------------------------------
    def measure(self, wl, steps, warmup, headline=False, e2e=True):
        args = self.args
        metric, model, n, bytes_in, bytes_out, desc = WORKLOADS[wl]
        strong = args.scaling == "strong" and headline
        if strong:
            total = args.n or STRONG_TOTAL[wl]
            n = total * (self.rank + 1) // self.world - total * self.rank // self.world
        else:
            total = 0
            if args.n and headline:
                n = args.n
            if wl == "surface":
                n = max(1, n // 200000) * 200000  # whole expiries of the 2000 x 100 grid
        if model == "black_scholes_merton" and wl != "iv":
            bytes_in += 8
        sets = self.build_sets(wl, n, strong_total=total)
        windows = [] if headline else None
        sampler = ClockSampler(self.local) if headline else None
        ms, launches = self.time_device(wl, sets, n, steps, warmup, windows)
        done = total if strong else self.world * n
        value = done * steps / (ms * 1e-3)
        kernel_ms = ms / steps
        res = {"metric": metric, "value": value, "unit": "contracts/s", "ms_per_step": kernel_ms, "contracts_per_gpu_per_step": n, "model": model,
               "gpu_launches": int(launches)}
        if e2e and not args.no_e2e:
            e2e_steps = max(3, min(steps, 50 if wl != "surface" else 5, max(3, int(2e8 // n))))
            sec = self.time_e2e(wl, sets[0], n, e2e_steps, windows)
            res["e2e"] = {"value": done / sec, "unit": "contracts/s", "h2d_bytes_per_step": n * bytes_in, "d2h_bytes_per_step": n * bytes_out,
                          "steps": e2e_steps,
                          "how": ("price_dataframe() on a pandas frame (pageable memory, string flag column): flag encode + pinned staging + pipeline + "
                                  "frame assembly, wall clock" if wl == "surface" else
                                  "pvv_*_host C-ABI call on PRE-PINNED host buffers, wall clock: short calls (<= 1.5 Mi contracts; K4: 128 Ki) are one launch "
                                  "that reads and writes the host memory over PCIe (zero-copy), longer ones the chunked 3-stream H2D/kernel/D2H pipeline; "
                                  "from pageable numpy / pandas the public API adds one staging copy (DESIGN.md section 6: 0.36 G contracts/s at 10 M rows)")}
            if wl != "surface":
                csec = self.copy_ceiling(n, bytes_in, bytes_out)
                res["e2e"]["ceiling"] = {"value": done / csec, "unit": "contracts/s", "gbs_both_directions": self.world * n * (bytes_in + bytes_out) / csec / 1e9,
                                         "frac": (done / sec) / (done / csec),
                                         "how": "cudaMemcpyAsync of the same bytes per step (H2D and D2H on two streams, pinned memory), all ranks at once, no kernel"}
        if headline:
            res["clocks"] = sampler.stop(windows)
        # rooflines
        dpk = self.dp.get("kernels", {}).get(wl)
        per_gpu = value / self.world
        hbm_peak = float(self.peaks.get("hbm_gbs", 6650.0))
        achieved_gbs = n * (bytes_in + bytes_out) / (kernel_ms * 1e-3) / 1e9
        hbm = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
               "algorithmic_bytes": n * (bytes_in + bytes_out), "traffic": (dpk["traffic_bytes"] * n if dpk and dpk.get("traffic_bytes") else None),
               "peak_source": "MEASURED_PEAKS.json (burst: the kernel is timed alone)" if self.peaks else "fallback (B200_PROFILING.md)"}
        if wl == "analytic" or not dpk:
            res["roofline"] = hbm
            if not dpk:
                res["roofline"]["note"] = "no DP-instruction count for this workload in profiles/dp_per_contract.json"
        else:
            a = per_gpu * dpk["dp_inst"]
            res["roofline"] = {"bound": "fp64", "achieved": a / 1e12, "peak": self.fp64_peak / 1e12, "unit": "TFLOP/s", "frac": a / self.fp64_peak,
                               "traffic": hbm["traffic"], "pipe_frac": per_gpu * dpk["dp_slots"] / self.fp64_peak,
                               "dp_inst_per_contract": dpk["dp_inst"], "dp_slots_per_contract": dpk["dp_slots"], "peak_fma_tflops": 2 * self.fp64_peak_fma / 1e12,
                               "source": self.dp.get("source"),
                               "how": "the engine is compiled without FMA contraction (parity), so one DP instruction is one flop: achieved = contracts/s x DP "
                                      "instructions per contract executed on active lanes (ncu capture named in `source`); peak = dependent DMUL+DADD chains "
                                      "measured in this run; pipe_frac counts warp-level issues x 32 (inactive lanes occupy the pipe too; compare ncu's "
                                      "sm__pipe_fp64_cycles_active); traffic = ncu dram read+write bytes per launch"
                                      + (" (IV -> greeks: a step is two launches, K3 then K2; counts, traffic and the time are those of the pair)" if wl == "iv_greeks" else ""),
                               "hbm": {k: hbm[k] for k in ("achieved", "peak", "unit", "frac", "algorithmic_bytes", "peak_source")}}
        if self.world > 1 and wl in ("greeks", "price", "iv", "iv_greeks"):
            ok = self.parity_sample(wl, sets[0], n)
            t = self.torch.tensor([1 if ok else 0], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
            res["parity_checked"] = {"ok": bool(t.item()), "what": "every rank: first 64 Ki contracts of its shard, all outputs, bit-compared with the CPU oracle"}
            if headline:
                gms, gok = self.gather_check(sets[0]["out"]["iv" if wl in ("iv", "iv_greeks") else "price"], total if strong else n * self.world)
                res["gather_ms"] = gms
                res["gather_ok"] = gok
        del sets
        self.torch.cuda.empty_cache()
        return res, n, (total if strong else n)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    B = Bench(args)
    wl = "greeks" if args.workload == "all" else args.workload
    head, n, n_cfg = B.measure(wl, args.steps, args.warmup, headline=True)
    kernels = {}
    if args.workload == "all" and args.scaling == "weak":
        for k in SUB_KERNELS:
            r, _, _ = B.measure(k, min(args.steps, 20), 3, headline=False)
            kernels[k] = r
    if B.rank == 0:
        out = {"metric": head["metric"], "value": head["value"], "unit": "contracts/s", "n_gpus": B.world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": config_of(wl, n_cfg, B.world, args.scaling)}
        for k in ("e2e", "gpu_launches", "clocks", "roofline", "parity_checked", "gather_ms", "gather_ok"):
            if k in head:
                out[k] = head[k]
        if kernels:
            out["kernels"] = kernels
        if B.world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(wl, WORKLOADS[wl][1])
        print(json.dumps(out))
    if B.world > 1:
        B.dist.barrier()
        B.dist.destroy_process_group()


if __name__ == "__main__":
    main()
