"""Generate tests/golden/*.npz by running the UNMODIFIED reference (on oracle/shim) in this container.

TEST INFRASTRUCTURE.  Needs /root/reference and numba, so it only runs in the build container;
the fixtures it writes are committed and are what travels to the GPU box.

    python oracle/gen_golden.py            # writes tests/golden/{cfg1_json,fake_data,cfg2,cfg3,wide,edge,kat}.npz
    python oracle/gen_golden.py composed   # only (re)writes the price_dataframe fixtures into cfg3.npz / cfg2.npz

Two flavours of the reference's public IV API are recorded (see oracle/validate_oracle.py):
  *_glibc  numpy dispatching np.exp to glibc (NPY_DISABLE_CPU_FEATURES=AVX512*): CPU-independent,
           the flavour the oracle and the CUDA engine reproduce bit for bit;
  *_simd   numpy's AVX-512 np.exp as it runs on this host: 1-ulp different forwards / deflaters.
Everything that goes through numba only (prices, greeks, the IV numba loop) has a single flavour.
"""
import json
import os
import subprocess
import sys
import warnings

_AVX512 = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
N = 4096

if len(sys.argv) == 1 or sys.argv[1] == "composed":
    # driver: run the two numpy flavours in fresh interpreters (CPU features are fixed at import)
    env = dict(os.environ, NUMBA_CACHE_DIR=os.environ.get("NUMBA_CACHE_DIR", "/tmp/nbcache"))
    pre = "composed_" if len(sys.argv) > 1 else ""
    subprocess.run([sys.executable, __file__, pre + "simd"], check=True, env=env)
    subprocess.run([sys.executable, __file__, pre + "glibc"], check=True, env=dict(env, NPY_DISABLE_CPU_FEATURES=_AVX512))
    sys.exit(0)

flavour = sys.argv[1]
composed_only = flavour.startswith("composed_")
flavour = flavour.replace("composed_", "")
sys.path[:0] = [os.path.join(HERE, "shim"), "/root/reference", ROOT]
sys.dont_write_bytecode = True

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402

import py_vollib_vectorized as ref  # noqa: E402
from py_vollib_vectorized import _iv_models, _model_calls, _numerical_greeks as ng  # noqa: E402

from oracle.chains import chain_cfg2, chain_cfg3, chain_edge, chain_wide  # noqa: E402

os.makedirs(OUT, exist_ok=True)
REF = "/root/reference"


def loop(fn, *a):
    return np.array(fn(*a), dtype=np.float64)


def ref_pricer(flag, S, K, t, r, sigma, q):
    return loop(_model_calls._black_scholes_merton_vectorized_call, flag.astype(np.float64), S, K, t, r, sigma, q)


def public_iv(c, price, model, on_error):
    kw = dict(q=c["q"]) if model == "black_scholes_merton" else {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            return ref.vectorized_implied_volatility(price, c["S"], c["K"], c["t"], c["r"], c["flag"].astype(np.int64), model=model,
                                                     on_error=on_error, return_as="numpy", **kw)


def numba_outputs(c, with_iv_loop=True):
    """Everything that only touches numba: 3 price models, 10 greeks, IV loop on the Black price."""
    f = c["flag"].astype(np.float64)
    o = {}
    o["price_bs"] = loop(_model_calls._black_scholes_vectorized_call, f, c["S"], c["K"], c["t"], c["r"], c["sigma"])
    o["price_bsm"] = loop(_model_calls._black_scholes_merton_vectorized_call, f, c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"])
    o["price_black"] = loop(_model_calls._black_vectorized_call, c["S"], c["K"], c["sigma"], c["t"], f)
    b = c["r"] - c["q"]
    for k in ("delta", "gamma", "theta", "rho", "vega"):
        o[f"{k}_bs"] = loop(getattr(ng, f"numerical_{k}_black_scholes"), f, c["S"], c["K"], c["t"], c["r"], c["sigma"], c["r"])
        o[f"{k}_bsm"] = loop(getattr(ng, f"numerical_{k}_black_scholes_merton"), f, c["S"], c["K"], c["t"], c["r"], c["sigma"], b)
    if with_iv_loop:
        o["ivloop_black"] = loop(_iv_models.implied_volatility_from_a_transformed_rational_guess, o["price_black"], c["S"], c["K"], c["t"], f)
    return o


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    if os.path.exists(path):
        old = dict(np.load(path))
        old.update(arrays)
        arrays = old
    np.savez_compressed(path, **arrays)
    print("wrote", path, sorted(arrays))


inputs = ("flag", "S", "K", "t", "r", "sigma", "q")


def composed():
    """The composed public paths (api.py:19-181): price_dataframe(price_col=...) = public IV, then get_all_greeks at that
    IV (NaN-sigma rows of masked contracts included), for the three models on the cfg3 rows; and
    price_dataframe(sigma_col=...) for black_scholes_merton and black on the cfg2 rows.  The IV leg goes through numpy's
    np.exp, so the price_col fixtures exist in both flavours; the sigma_col ones are numba only."""
    c3 = dict(np.load(os.path.join(OUT, "cfg3.npz")))
    df = pd.DataFrame({"Flag": np.where(c3["flag"] > 0, "c", "p"), "S": c3["S"], "K": c3["K"], "T": c3["t"], "R": c3["r"], "Q": c3["q"], "Px": c3["price"]})
    kw = dict(flag_col="Flag", underlying_price_col="S", strike_col="K", annualized_tte_col="T", riskfree_rate_col="R")
    for model, tag in (("black_scholes_merton", "bsm"), ("black_scholes", "bs"), ("black", "black")):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with np.errstate(all="ignore"):
                res = ref.price_dataframe(df, **kw, price_col="Px", model=model, inplace=False,
                                          **(dict(dividend_col="Q") if model == "black_scholes_merton" else {}))
        assert list(res.columns) == ["IV", "delta", "gamma", "theta", "rho", "vega"], list(res.columns)
        save("cfg3", **{f"pdf_{tag}_{col}_{flavour}": res[col].values.astype(np.float64) for col in res.columns})
    if flavour == "glibc":
        c2 = dict(np.load(os.path.join(OUT, "cfg2.npz")))
        df2 = pd.DataFrame({"Flag": np.where(c2["flag"] > 0, "c", "p"), "S": c2["S"], "K": c2["K"], "T": c2["t"], "R": c2["r"], "Q": c2["q"], "IV": c2["sigma"]})
        for model, tag in (("black_scholes_merton", "bsm"), ("black", "black"), ("black_scholes", "bs")):
            res = ref.price_dataframe(df2, **kw, sigma_col="IV", model=model, inplace=False,
                                      **(dict(dividend_col="Q") if model == "black_scholes_merton" else {}))
            assert list(res.columns) == ["Price", "delta", "gamma", "theta", "rho", "vega"], list(res.columns)
            save("cfg2", **{f"pdf_{tag}_{col}": res[col].values.astype(np.float64) for col in res.columns})


if composed_only:
    composed()
    sys.exit(0)

if flavour == "simd":
    for f in ("cfg1_json", "fake_data", "cfg2", "cfg3", "wide", "edge", "kat"):
        p = os.path.join(OUT, f + ".npz")
        if os.path.exists(p):
            os.remove(p)
    # ---- cfg 1: the reference's own regression table -------------------------------------------
    with open(os.path.join(REF, "tests", "test_data_py_vollib.json"), "rb") as fh:
        d = json.load(fh)
    df = pd.DataFrame(d["data"], index=d["index"], columns=d["columns"])
    out = {col: df[col].values.astype(np.float64) for col in df.columns}
    for fl, tag in ((1, "c"), (-1, "p")):
        c = dict(flag=np.full(len(df), fl, dtype=np.int8), S=out["S"], K=out["K"], t=out["t"], r=out["R"], sigma=out["v"], q=np.zeros(len(df)))
        o = numba_outputs(c, with_iv_loop=False)
        for k, v in o.items():
            out[f"ref_{tag}_{k}"] = v
    save("cfg1_json", **out)
    # ---- fake_data.csv (test_validity_greeks) ---------------------------------------------------
    fd = pd.read_csv(os.path.join(REF, "tests", "fake_data.csv"))
    c = dict(flag=np.where(fd["Flag"].values == "c", 1, -1).astype(np.int8), S=fd["Px"].values.astype(np.float64),
             K=fd["Strike"].values.astype(np.float64), t=fd["Annualized Time To Expiration"].values.astype(np.float64),
             r=fd["Interest Free Rate"].values.astype(np.float64), sigma=np.zeros(len(fd)), q=np.zeros(len(fd)))
    mid = fd["MidPx"].values.astype(np.float64)
    save("fake_data", MidPx=mid, **{k: c[k] for k in inputs})
    # ---- seeded chains: numba-only outputs -------------------------------------------------------
    c2, cw, ce = chain_cfg2(N, 0), chain_wide(N, 7), chain_edge()
    save("cfg2", **{k: c2[k] for k in inputs}, **numba_outputs(c2))
    save("wide", **{k: cw[k] for k in inputs}, **numba_outputs(cw))
    save("edge", price=ce["price"], **{k: ce[k] for k in inputs}, **numba_outputs(ce, with_iv_loop=False))
    keep = (ce["t"] != 0) & (ce["S"] != 0)
    save("edge", ivloop_keep=keep, ivloop_black=loop(_iv_models.implied_volatility_from_a_transformed_rational_guess,
                                                     ce["price"][keep], ce["S"][keep], ce["K"][keep], ce["t"][keep],
                                                     ce["flag"][keep].astype(np.float64)))
    c3 = chain_cfg3(N, 1, pricer=ref_pricer)
    save("cfg3", price=c3["price"], **{k: c3[k] for k in inputs})

# ---- both flavours: public IV API -------------------------------------------------------------------
fd = dict(np.load(os.path.join(OUT, "fake_data.npz")))
iv = public_iv(fd, fd["MidPx"], "black_scholes", "warn")
save("fake_data", **{f"iv_bs_warn_{flavour}": iv})
if flavour == "glibc":
    c = dict(fd, sigma=iv)
    o = numba_outputs(c, with_iv_loop=False)
    save("fake_data", sigma=iv, **{k: v for k, v in o.items() if k.endswith("_bs")})
c3 = dict(np.load(os.path.join(OUT, "cfg3.npz")))
for model, tag in (("black_scholes_merton", "bsm"), ("black_scholes", "bs"), ("black", "black")):
    for on_error in ("warn", "ignore"):
        save("cfg3", **{f"iv_{tag}_{on_error}_{flavour}": public_iv(c3, c3["price"], model, on_error)})
ce = dict(np.load(os.path.join(OUT, "edge.npz")))
keep = ce["ivloop_keep"]
cek = {k: ce[k][keep] for k in inputs + ("price",)}
for on_error in ("warn", "ignore"):
    save("edge", **{f"iv_bsm_{on_error}_{flavour}": public_iv(cek, cek["price"], "black_scholes_merton", on_error)})
c1 = dict(np.load(os.path.join(OUT, "cfg1_json.npz")))
for tag, fl in (("c", 1), ("p", -1)):
    c = dict(flag=np.full(c1["S"].size, fl, dtype=np.int8), S=c1["S"], K=c1["K"], t=c1["t"], r=c1["R"], q=np.zeros(c1["S"].size))
    save("cfg1_json", **{f"ref_{tag}_ivroundtrip_{flavour}": public_iv(c, c1[f"ref_{tag}_price_bs"], "black_scholes", "warn")})

composed()

if flavour == "glibc":
    # ---- docstring known answers and SURVEY Appendix D, recomputed from the reference ---------------
    kat = {}
    kat["black_doc"] = ref.vectorized_black(["c", "p"], 95, [100, 90], .2, .2, .2, return_as="numpy")
    kat["bs_doc"] = ref.vectorized_black_scholes(["c", "p"], 95, [100, 90], .2, .2, .2, return_as="numpy")
    kat["bsm_doc"] = ref.vectorized_black_scholes_merton(["c", "p"], [95, 99], [100, 90], .2, .2, .2, 0, return_as="numpy")
    kat["iv_bsm_doc"] = ref.vectorized_implied_volatility(.10, 95, [100, 90], .2, .2, ["c", "p"], q=0, model="black_scholes_merton", return_as="numpy")
    kat["iv_black_doc"] = ref.vectorized_implied_volatility_black(.10, 95, [100, 90], .2, .2, ["c", "p"], return_as="numpy")
    kat["iv_bs_5"] = ref.vectorized_implied_volatility(5, 100, 100, .5, .01, ["c", "p"], return_as="numpy")
    g = ref.get_all_greeks(["c", "p"], 95, [100, 90], .2, .2, .2, return_as="dict")
    for k, v in g.items():
        kat[f"greeks_bs_{k}"] = np.array(v)
    g = ref.get_all_greeks(["c", "p"], 100, [100, 90], .5, .05, .2, q=.02, model="black_scholes_merton", return_as="dict")
    for k, v in g.items():
        kat[f"greeks_bsm_{k}"] = np.array(v)
    g = ref.get_all_greeks(["c", "p"], 95, [100, 90], .2, .2, .2, model="black", return_as="dict")
    for k, v in g.items():
        kat[f"greeks_black_{k}"] = np.array(v)
    dfq = pd.DataFrame({"Flag": ["c", "p"], "S": 95, "K": [100, 90], "T": 0.2, "R": 0.2, "IV": 0.2})
    res = ref.price_dataframe(dfq, flag_col="Flag", underlying_price_col="S", strike_col="K", annualized_tte_col="T",
                              riskfree_rate_col="R", sigma_col="IV", model="black_scholes", inplace=False)
    for col in res.columns:
        kat[f"price_dataframe_{col}"] = res[col].values
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kat["quickstart_iv"] = ref.vectorized_implied_volatility([0.5, 0.05], 95, [100, 90], .2, .2, ["c", "p"], return_as="numpy")
    kat["repr_black"] = np.array(repr(__import__("py_vollib.black", fromlist=["black"]).black))
    save("kat", **kat)
