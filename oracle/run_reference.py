"""Times the UNMODIFIED reference (baseline/_ref, installed by oracle/install_reference.py) on the host cores.

BENCHMARK INFRASTRUCTURE: run by bench.py (--impl reference and the cpu_baseline block) in a process of its own,
so that nothing of the engine (libpvv.so, the alias package) is mapped while the reference is on the clock.

    python oracle/run_reference.py --workload greeks|price|iv|iv_greeks|surface --n 200000 --steps 5 --warmup 1
                                   [--procs P]

The reference's numba loops are single-threaded (no parallel=True, GIL held): `public` and `loop_only` are 1-core
figures, as shipped.  --procs P adds an all-cores figure the way BASELINE.md section 3 plans it: P processes, one
contiguous chunk each, results pickled back (that is outside the reference's own code path and reported separately).
Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFDIR = os.path.join(ROOT, "baseline", "_ref")
os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join("/tmp", "pvv_numba_cache"))
# the reference first, then its dependency shim; the repo root (it holds the alias package of the same name) is removed
sys.path = [p for p in sys.path if os.path.abspath(p or ".") not in (HERE, ROOT)]
sys.path[:0] = [REFDIR, os.path.join(HERE, "shim")]
sys.dont_write_bytecode = True

import numpy as np  # noqa: E402


def load_chain(workload, n):
    import importlib.util  # the seeded generators, loaded by path: the engine package itself is never imported here
    spec = importlib.util.spec_from_file_location("_pvv_chains", os.path.join(ROOT, "py_vollib_vectorized_b200", "util", "chains.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    chain_cfg2, chain_cfg3, chain_cfg5 = m.chain_cfg2, m.chain_cfg3, m.chain_cfg5
    if workload == "surface":
        ne = max(1, n // (2000 * 100))
        return chain_cfg5(2000, ne, 100)
    if workload == "iv":
        import py_vollib_vectorized as ref

        def pricer(f, S, K, t, r, sg, q):
            return ref.vectorized_black_scholes_merton(f, S, K, t, r, sg, q, return_as="numpy")
        return chain_cfg3(n, seed=0, pricer=pricer)
    return chain_cfg2(n, seed=0)


def make_steps(workload, c):
    """(public, loop_only): callables running one pass through the reference's public API / its numba loops only."""
    import pandas as pd
    import py_vollib_vectorized as ref
    from py_vollib_vectorized import _iv_models, _model_calls, _numerical_greeks as ng
    n = c["S"].size
    fl = np.where(c["flag"] > 0, "c", "p")
    ff = c["flag"].astype(np.float64)
    S, K, t, r, sg, q = (np.ascontiguousarray(c[k]) for k in ("S", "K", "t", "r", "sigma", "q"))
    if workload == "price":
        return (lambda: ref.vectorized_black_scholes(fl, S, K, t, r, sg, return_as="numpy"),
                lambda: np.ascontiguousarray(_model_calls._black_scholes_vectorized_call(ff, S, K, t, r, sg)))
    if workload == "greeks":
        def public():
            ref.vectorized_black_scholes(fl, S, K, t, r, sg, return_as="numpy")
            ref.get_all_greeks(fl, S, K, t, r, sg, model="black_scholes", return_as="dict")

        def loops():
            _model_calls._black_scholes_vectorized_call(ff, S, K, t, r, sg)
            for g in ("delta", "gamma", "theta", "rho", "vega"):
                getattr(ng, f"numerical_{g}_black_scholes")(ff, S, K, t, r, sg, r)
        return public, loops
    if workload == "iv":
        price = np.ascontiguousarray(c["price"])

        def public():
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref.vectorized_implied_volatility(price, S, K, t, r, fl, q, model="black_scholes_merton", on_error="warn", return_as="numpy")
        D = np.exp(-r * t)
        F, und = S * np.exp((r - q) * t), price / D
        return public, (lambda: _iv_models.implied_volatility_from_a_transformed_rational_guess(und, F, K, t, ff))
    if workload == "iv_greeks":
        price = ref.vectorized_black_scholes(fl, S, K, t, r, sg, return_as="numpy")
        df = pd.DataFrame({"Flag": fl, "S": S, "K": K, "T": t, "R": r, "Px": price})

        def public():
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref.price_dataframe(df, flag_col="Flag", underlying_price_col="S", strike_col="K", annualized_tte_col="T",
                                    riskfree_rate_col="R", price_col="Px", model="black_scholes", inplace=False)
        return public, None
    if workload == "surface":
        df = pd.DataFrame({"Flag": fl, "F": S, "K": K, "T": t, "R": r, "IV": sg})
        return (lambda: ref.price_dataframe(df, flag_col="Flag", underlying_price_col="F", strike_col="K", annualized_tte_col="T",
                                            riskfree_rate_col="R", sigma_col="IV", model="black", inplace=False)), None
    raise SystemExit(f"no reference path for workload {workload!r}")


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return (time.perf_counter() - t0) / steps


_CHAIN = None  # inherited by the forked workers of --procs


def _worker(args):
    workload, lo, hi = args
    public, _ = make_steps(workload, {k: v[lo:hi] for k, v in _CHAIN.items()})
    public()
    return hi - lo


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="greeks")
    ap.add_argument("--n", type=int, default=200_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--procs", type=int, default=0)
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(REFDIR, "py_vollib_vectorized")):
        print(json.dumps({"unavailable": "baseline/_ref is not installed (python oracle/install_reference.py in the build container)"}))
        return
    t_import = time.perf_counter()
    c = load_chain(a.workload, a.n)
    n = c["S"].size
    public, loops = make_steps(a.workload, c)
    c = {k: v for k, v in c.items() if isinstance(v, np.ndarray) and v.shape == c["S"].shape}
    small = {k: v[:1000] for k, v in c.items()}
    ps, ls = make_steps(a.workload, small)
    ps()  # JIT compile on a 1 k-contract call, excluded from every rate
    if ls:
        ls()
    jit_s = time.perf_counter() - t_import
    out = {"workload": a.workload, "n": n, "steps": a.steps, "warmup": a.warmup, "cores": 1, "jit_and_setup_s": jit_s}
    sec = timed(public, a.steps, a.warmup)
    out["public_s_per_step"] = sec
    out["public"] = n / sec
    if loops:
        sec = timed(loops, max(1, a.steps // 2), 1)
        out["loop_only"] = n / sec
    if a.procs > 1:
        import multiprocessing as mp
        global _CHAIN
        _CHAIN = c
        bounds = [(a.workload, n * i // a.procs, n * (i + 1) // a.procs) for i in range(a.procs)]
        with mp.get_context("fork").Pool(a.procs) as pool:
            pool.map(_worker, bounds)  # warm every worker (JIT from the on-disk cache, page faults)
            t0 = time.perf_counter()
            pool.map(_worker, bounds)
            wall = time.perf_counter() - t0
        out["all_cores"] = {"procs": a.procs, "value": n / wall,
                            "note": "one contiguous chunk per forked process, public call, wall clock; outside the reference's own code path"}
    import numba
    out["numba"] = numba.__version__
    out["host_cpus"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
