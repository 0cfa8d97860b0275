"""Pin the C oracle to the reference: bitwise comparison on dense seeded inputs.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference and numba):

    python oracle/validate_oracle.py [--n 200000] [--native-simd]

It imports the UNMODIFIED reference on top of oracle/shim and compares, bit for bit (NaNs by
position), every array entry point of oracle/libpvv_oracle.so with the reference's numba loops and
public API.  numpy's own np.exp (used by the reference's IV pre-processing,
implied_volatility.py:48,71 / _iv_models.py:15) is an AVX-512 SIMD routine on hosts that have it
and glibc's scalar exp elsewhere, so the reference's public IV results depend on the host CPU at
the 1-ulp level.  The oracle restates the glibc flavour; by default this script therefore re-execs
itself with NPY_DISABLE_CPU_FEATURES set so that numpy dispatches to glibc, and demands exact
equality.  --native-simd keeps numpy's SIMD exp and reports the (small) differences instead.

The last run's output is committed as oracle/VALIDATION.txt.
"""
import argparse
import os
import sys
import warnings

_AVX512 = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR"

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=200000)
ap.add_argument("--native-simd", action="store_true")
args = ap.parse_args()

if not args.native_simd and os.environ.get("NPY_DISABLE_CPU_FEATURES") is None:
    os.environ["NPY_DISABLE_CPU_FEATURES"] = _AVX512
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
    os.execv(sys.executable, [sys.executable] + sys.argv)

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "shim"), "/root/reference", os.path.dirname(HERE)]
sys.dont_write_bytecode = True

import numpy as np  # noqa: E402

import py_vollib_vectorized as ref  # noqa: E402
from py_vollib_vectorized import _iv_models, _model_calls, _numerical_greeks as ng  # noqa: E402

from oracle import get_oracle  # noqa: E402
from oracle.chains import chain_cfg2, chain_cfg3, chain_wide, chain_edge  # noqa: E402

O = get_oracle()
FAIL = 0


def same(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    return (na == nb) & (na | (a == b))


def report(name, a, b):
    global FAIL
    ok = same(a, b)
    bad = int((~ok).sum())
    msg = f"{name:58s} n={ok.size:8d} bit-mismatches={bad}"
    if bad:
        FAIL += 1
        i = np.flatnonzero(~ok)[:3]
        msg += f"  first at {i.tolist()}: ref={np.asarray(a)[i].tolist()} oracle={np.asarray(b)[i].tolist()}"
    print(msg, flush=True)


def ref_call(fn, *a):
    """numba loops raise ZeroDivisionError for the whole call; return (array | None, raised)."""
    try:
        return np.array(fn(*a), dtype=np.float64), False
    except ZeroDivisionError:
        return None, True


def check_prices(tag, c):
    f = c["flag"].astype(np.float64)
    a, _ = ref_call(_model_calls._black_scholes_vectorized_call, f, c["S"], c["K"], c["t"], c["r"], c["sigma"])
    b, z = O.price("black_scholes", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"])
    assert z == -1, z
    report(f"{tag} price black_scholes", a, b)
    a, _ = ref_call(_model_calls._black_scholes_merton_vectorized_call, f, c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"])
    b, z = O.price("black_scholes_merton", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"])
    report(f"{tag} price black_scholes_merton", a, b)
    a, _ = ref_call(_model_calls._black_vectorized_call, c["S"], c["K"], c["sigma"], c["t"], f)
    b, z = O.price("black", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"])
    report(f"{tag} price black (undiscounted)", a, b)


def check_greeks(tag, c):
    f = c["flag"].astype(np.float64)
    g, z = O.greeks("black_scholes", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"])
    for k in ("delta", "gamma", "theta", "rho", "vega"):
        fn = getattr(ng, f"numerical_{k}_black_scholes")
        a, _ = ref_call(fn, f, c["S"], c["K"], c["t"], c["r"], c["sigma"], c["r"])
        report(f"{tag} {k} black_scholes", a, g[k])
    g, z = O.greeks("black_scholes_merton", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"])
    bcol = c["r"] - c["q"]
    for k in ("delta", "gamma", "theta", "rho", "vega"):
        fn = getattr(ng, f"numerical_{k}_black_scholes_merton")
        a, _ = ref_call(fn, f, c["S"], c["K"], c["t"], c["r"], c["sigma"], bcol)
        report(f"{tag} {k} black_scholes_merton", a, g[k])


def check_iv_loop(tag, c, price):
    """numba loop only: undiscounted price and forward handed over directly."""
    f = c["flag"].astype(np.float64)
    a, raised = ref_call(_iv_models.implied_volatility_from_a_transformed_rational_guess, price, c["S"], c["K"], c["t"], f)
    b, st, z = O.iv("black", price, c["S"], c["K"], c["t"], 0.0, c["flag"], mask_errors=False)
    if raised:
        print(f"{tag} IV numba loop: reference raised ZeroDivisionError; oracle zde index = {z}")
        assert z >= 0
    else:
        assert z == -1, z
        report(f"{tag} IV numba loop (_iv_models.py:18-44)", a, b)


def check_iv_public(tag, c, price, model):
    for on_error in ("warn", "ignore"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            kw = dict(q=c["q"]) if model == "black_scholes_merton" else {}
            with np.errstate(all="ignore"):
                a = ref.vectorized_implied_volatility(price, c["S"], c["K"], c["t"], c["r"], c["flag"].astype(np.int64), model=model,
                                                      on_error=on_error, return_as="numpy", **kw)
        b, st, z = O.iv(model, price, c["S"], c["K"], c["t"], c["r"], c["flag"], c["q"] if kw else None,
                        mask_errors=(on_error != "ignore"))
        assert z == -1
        if args.native_simd:
            ok = same(a, b)
            d = np.abs(a - b)[~ok & ~np.isnan(a) & ~np.isnan(b) & np.isfinite(a) & np.isfinite(b)]
            print(f"{tag} IV public {model} on_error={on_error}: {(~ok).sum()} of {ok.size} rows differ from the AVX-512 numpy run; "
                  f"NaN positions equal: {bool((np.isnan(a) == np.isnan(b)).all())}; max |diff| = {d.max() if d.size else 0:.3e}; "
                  f"rows with |diff|>1e-10: {(d > 1e-10).sum()}")
        else:
            report(f"{tag} IV public {model} on_error={on_error}", a, b)


n = args.n
print(f"numpy dispatch: {'native SIMD' if args.native_simd else 'glibc exp/log (AVX512 disabled)'}; n={n}")
c2 = chain_cfg2(n, seed=0)
cw = chain_wide(n, seed=7)
ce = chain_edge()
for tag, c in (("cfg2", c2), ("wide", cw), ("edge", ce)):
    if not args.native_simd:
        check_prices(tag, c)
        check_greeks(tag, c)
# IV: prices produced by the oracle's own pricer (then perturbed strata in cfg3)
for tag, c in (("cfg2", c2), ("wide", cw)):
    p, _ = O.price("black", c["flag"], c["S"], c["K"], c["t"], 0.0, c["sigma"])
    if not args.native_simd:
        check_iv_loop(tag, c, p)
c3 = chain_cfg3(n, seed=1, oracle=O)
check_iv_public("cfg3", c3, c3["price"], "black_scholes_merton")
check_iv_public("cfg3", c3, c3["price"], "black_scholes")
check_iv_public("cfg3", c3, c3["price"], "black")
if not args.native_simd:
    keep = (ce["t"] != 0) & (ce["S"] != 0)
    cek = {k: v[keep] for k, v in ce.items()}
    check_iv_loop("edge", cek, cek["price"])
    for tag, c, p in (("cfg3", c3, c3["price"]),):
        _, _, _, br = O.iv("black_scholes_merton", p, c["S"], c["K"], c["t"], c["r"], c["flag"], c["q"], want_branch=True)
        print(f"{tag} solver branch mix (0 beta<=0, 1 lowest, 2 lower-mid, 3 upper-mid, 4 highest):", np.bincount(br.astype(np.int64) + 0, minlength=5).tolist())
    p, _ = O.price("black", cw["flag"], cw["S"], cw["K"], cw["t"], 0.0, cw["sigma"])
    _, _, _, br = O.iv("black", p, cw["S"], cw["K"], cw["t"], 0.0, cw["flag"], want_branch=True)
    print("wide solver branch mix:", np.bincount(br.astype(np.int64), minlength=5).tolist())
    # ZeroDivisionError rows
    for name, kw in (("t==0", dict(t=0.0)), ("K==0", dict(K=0.0)), ("S==0", dict(S=0.0))):
        base = dict(price=5.0, S=100.0, K=100.0, t=0.5, r=0.01)
        base.update(kw)
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                with np.errstate(all="ignore"):
                    ref.vectorized_implied_volatility(base["price"], base["S"], base["K"], base["t"], base["r"], ["c"], on_error="ignore", return_as="numpy")
            raised = False
        except ZeroDivisionError:
            raised = True
        _, st, z = O.iv("black_scholes", base["price"], base["S"], base["K"], base["t"], base["r"], [1], mask_errors=False)
        print(f"IV {name}: reference raised={raised}; oracle zde={z} status={st.tolist()}")
        if raised != (z >= 0):
            FAIL += 1
    try:
        _model_calls._black_scholes_vectorized_call(np.array([1.0]), np.array([100.0]), np.array([0.0]), np.array([0.5]), np.array([0.01]), np.array([0.2]))
        raised = False
    except ZeroDivisionError:
        raised = True
    _, z = O.price("black_scholes", [1], 100.0, 0.0, 0.5, 0.01, 0.2)
    print(f"price K==0: reference raised={raised}; oracle zde={z}")
    if raised != (z >= 0):
        FAIL += 1
print("RESULT:", "FAIL" if FAIL else "PASS")
sys.exit(1 if FAIL else 0)
