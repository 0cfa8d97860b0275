"""GPU tier: the public Python surface, written the way the reference's own tests are
(/root/reference/tests/test_entrypoints.py, test_monkeypatch.py) — same calls, same assertions —
plus the docstring known answers, on_error semantics and broadcasting."""
import json
import warnings

import numpy as np
import pandas as pd
import pytest
from numpy.testing import assert_array_almost_equal

from tests.util import assert_bit_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pvv():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import py_vollib_vectorized_b200 as m
    return m


@pytest.fixture(scope="module")
def frames(golden):
    g = golden("cfg1_json")
    cols = ["S", "K", "R", "t", "v", "bs_call", "bs_put", "CD", "CG", "CT", "CV", "CR", "PD", "PG", "PT", "PV", "PR", "call_vals"]
    df = pd.DataFrame({c: g[c] for c in cols})
    calls, puts = df.copy(), df.copy()
    calls["flag"], calls["q"] = "c", 0
    puts["flag"], puts["q"] = "p", 0
    return df, calls, puts


def test_implied_volatility_vectorized(pvv, frames):
    """test_entrypoints.py:24-68"""
    _, calls, puts = frames
    for d, as_values in ((calls, False), (puts, True)):
        price = pvv.vectorized_black_scholes(d["flag"], d["S"], d["K"], d["t"], d["R"], d["v"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ivs = pvv.vectorized_implied_volatility(price=price.values if as_values else price, S=d["S"].values, K=d["K"].values,
                                                    t=d["t"].values, r=d["R"].values, flag=d["flag"].values)
        true_sigmas = d["v"].values.ravel()
        pred = ivs.values.ravel()
        ok = np.isclose(true_sigmas, pred, atol=1e-2)
        ok[~ok] = pred[~ok] == 0
        assert all(ok)


def test_black_scholes_and_merton_vectorized(pvv, frames):
    """test_entrypoints.py:70-122"""
    _, calls, puts = frames
    for d, col in ((calls, "bs_call"), (puts, "bs_put")):
        p = pvv.vectorized_black_scholes(d["flag"], d["S"], d["K"], d["t"], d["R"], d["v"])
        assert_array_almost_equal(p.values.ravel(), d[col].values)
        p = pvv.vectorized_black_scholes_merton(d["flag"], d["S"], d["K"], d["t"], d["R"], d["v"], d["q"])
        assert_array_almost_equal(p.values.ravel(), d[col].values)


def test_api_get_all_greeks(pvv, frames):
    """test_entrypoints.py:125-152"""
    _, calls, puts = frames
    for d, D, G, V in ((calls, "CD", "CG", "CV"), (puts, "PD", "PG", "PV")):
        g = pvv.get_all_greeks(d["flag"], d["S"], d["K"], d["t"], d["R"], d["v"], model="black_scholes", return_as="dataframe")
        assert_array_almost_equal(g["delta"].values, d[D].values)
        assert_array_almost_equal(g["gamma"].values, d[G].values)
        assert_array_almost_equal(g["vega"].values, d[V].values * .01, 3)
        assert list(g.columns) == ["delta", "gamma", "theta", "rho", "vega"]


def test_api_price_dataframe(pvv, frames):
    """test_entrypoints.py:154-177: 3 models x {sigma, price, both}"""
    _, calls, puts = frames
    for model in ("black", "black_scholes", "black_scholes_merton"):
        for kw, expect in ((dict(sigma_col="v"), {"Price"}), (dict(price_col="bs_put"), {"IV"}), (dict(sigma_col="v", price_col="bs_put"), set())):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res = pvv.price_dataframe(puts, flag_col="flag", underlying_price_col="S", strike_col="K", annualized_tte_col="t",
                                          riskfree_rate_col="R", dividend_col="q" if model == "black_scholes_merton" else None,
                                          model=model, inplace=False, **kw)
            assert set(res.columns) == expect | {"delta", "gamma", "theta", "vega", "rho"}
            assert len(res) == len(puts) and (res.index == puts.index).all()
    d = puts.copy()
    assert pvv.price_dataframe(d, flag_col="flag", underlying_price_col="S", strike_col="K", annualized_tte_col="t", riskfree_rate_col="R",
                               sigma_col="v", inplace=True) is None
    assert {"Price", "delta", "gamma", "theta", "rho", "vega"} <= set(d.columns)


def test_validity_greeks(pvv, golden):
    """test_entrypoints.py:179-297: vector call == per-row scalar call through the patched py_vollib,
    on fake_data.csv — and both equal the reference's numbers."""
    import py_vollib.black_scholes.greeks.numerical as num
    g = golden("fake_data")
    fl = np.where(g["flag"] > 0, "c", "p")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        iv = pvv.vectorized_implied_volatility(g["MidPx"], g["S"], g["K"], g["t"], g["r"], fl, return_as="numpy")
    assert np.nanmax(np.abs(iv - g["iv_bs_warn_glibc"])) <= 1e-10
    sigma = g["sigma"]
    for name in ("delta", "gamma", "theta", "rho", "vega"):
        vec = getattr(pvv, "vectorized_" + name)(fl, g["S"], g["K"], g["t"], g["r"], sigma, model="black_scholes", return_as="numpy")
        assert_bit_equal(g[f"{name}_bs"], vec, name)
        rows = [getattr(num, name)(fl[i], float(g["S"][i]), float(g["K"][i]), float(g["t"][i]), float(g["r"][i]), float(sigma[i]),
                                   return_as="numpy")[0] for i in range(0, len(fl), 13)]
        assert_array_almost_equal(vec[::13], rows)


def test_broadcasting(pvv):
    """test_entrypoints.py:299-323"""
    args = dict(price=0.10, S=95, K=100, t=.2, r=.2, flag="c")
    for key in args:
        a = dict(args)
        a[key] = [args[key], args[key]]
        ivs = pvv.vectorized_implied_volatility(**a, return_as="numpy")
        assert ivs.shape == (2,) and ivs[0] == ivs[1]
    gargs = dict(flag="c", S=95, K=100, t=.2, r=.2, sigma=.2)
    for key in gargs:
        a = dict(gargs)
        a[key] = [gargs[key], gargs[key]]
        g = pvv.get_all_greeks(**a, return_as="numpy")
        assert all(v.shape == (2,) and v[0] == v[1] for v in g.values())


def test_monkeypatched_entry_points_compute(pvv, golden):
    """test_monkeypatch.py + the docstring known answers of models.py / implied_volatility.py / greeks.py / api.py"""
    import py_vollib.black
    import py_vollib.black.implied_volatility
    import py_vollib.black_scholes
    import py_vollib.black_scholes.greeks.numerical
    import py_vollib.black_scholes_merton
    import py_vollib.black_scholes_merton.implied_volatility
    k = golden("kat")
    flag = ['c', 'p']
    assert_bit_equal(k["black_doc"], py_vollib.black.black(flag, 95, [100, 90], .2, .2, .2, return_as='numpy'))
    assert_bit_equal(k["bs_doc"], py_vollib.black_scholes.black_scholes(flag, 95, [100, 90], .2, .2, .2, return_as='numpy'))
    assert_bit_equal(k["bsm_doc"], py_vollib.black_scholes_merton.black_scholes_merton(flag, [95, 99], [100, 90], .2, .2, .2, 0, return_as='numpy'))
    np.testing.assert_allclose(py_vollib.black_scholes.black_scholes(flag, 95, [100, 90], .2, .2, .2, return_as='numpy'), [2.89558836, 0.61109351], atol=5e-9)
    iv = py_vollib.black_scholes_merton.implied_volatility.implied_volatility(0.10, 95, [100, 90], .2, .2, flag, q=0, return_as='numpy')
    np.testing.assert_allclose(iv, k["iv_bsm_doc"], atol=1e-10, rtol=0)
    np.testing.assert_allclose(iv, [0.02621257, 0.12585767], atol=5e-9)
    iv = py_vollib.black.implied_volatility.implied_volatility(0.10, 95, [100, 90], .2, .2, flag, return_as='numpy')
    np.testing.assert_allclose(iv, k["iv_black_doc"], atol=1e-10, rtol=0)
    d = py_vollib.black_scholes.greeks.numerical.delta(flag, 95, [100, 90], .2, .2, .2, return_as='numpy')
    assert_bit_equal(k["greeks_bs_delta"], d)
    for model, tag, a in (("black_scholes", "bs", (flag, 95, [100, 90], .2, .2, .2)), ("black", "black", (flag, 95, [100, 90], .2, .2, .2)),
                          ("black_scholes_merton", "bsm", (flag, 100, [100, 90], .5, .05, .2, .02))):
        g = pvv.get_all_greeks(*a, model=model, return_as="dict")
        for name, v in g.items():
            assert_bit_equal(k[f"greeks_{tag}_{name}"], v, f"{model} {name}")
    df = pd.DataFrame({"Flag": ['c', 'p'], "S": 95, "K": [100, 90], "T": 0.2, "R": 0.2, "IV": 0.2})
    res = pvv.price_dataframe(df, flag_col='Flag', underlying_price_col='S', strike_col='K', annualized_tte_col='T', riskfree_rate_col='R',
                              sigma_col='IV', model='black_scholes', inplace=False)
    assert list(res.columns) == ["Price", "delta", "gamma", "theta", "rho", "vega"]
    for c in res.columns:
        assert_bit_equal(k[f"price_dataframe_{c}"], res[c].values, c)
    js = json.loads(pvv.get_all_greeks(flag, 95, [100, 90], .2, .2, .2, return_as="json"))
    assert js["delta"] == k["greeks_bs_delta"].tolist()


def test_return_containers(pvv):
    s = pvv.vectorized_black_scholes(['c'], 100, 100, .5, .01, .2, return_as="series")
    assert isinstance(s, pd.Series) and s.name == "Price"
    d = pvv.vectorized_black_scholes(['c'], 100, 100, .5, .01, .2)
    assert isinstance(d, pd.DataFrame) and list(d.columns) == ["Price"]
    a = pvv.vectorized_black_scholes(['c'], 100, 100, .5, .01, .2, return_as="array")
    assert isinstance(a, np.ndarray) and a.dtype == np.float64
    iv = pvv.vectorized_implied_volatility(5, 100, 100, .5, .01, ['c'])
    assert list(iv.columns) == ["IV"]
    assert pvv.vectorized_delta(['c'], 100, 100, .5, .01, .2, return_as="series").name == "delta"
    assert pvv.vectorized_black_scholes(['c'], np.float32(100.0).item(), 100, .5, .01, .2, dtype=np.float32, return_as="numpy").dtype == np.float64


def test_on_error_semantics(pvv):
    """SURVEY.md Appendix B, IV column."""
    from py_vollib_vectorized_b200.util.data_format import PriceIsAboveMaximum, PriceIsBelowIntrinsic
    # rows: ok, below intrinsic (ITM call priced under intrinsic), above maximum (call), above maximum (put)
    price = [5.0, 1.0, 150.0, 100.0]
    S = [100.0, 110.0, 100.0, 100.0]
    flag = ['c', 'c', 'c', 'p']
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        iv = pvv.vectorized_implied_volatility(price, S, 100.0, .5, .01, flag, on_error="warn", return_as="numpy")
    msgs = [str(x.message) for x in w]
    assert msgs == ["Found Below Intrinsic contracts at index [1]", "Found Above Maximum Price contracts at index [2, 3]"]
    assert abs(iv[0] - 0.16877851941995567) < 1e-12 and np.isnan(iv[1:]).all()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        iv = pvv.vectorized_implied_volatility(price, S, 100.0, .5, .01, flag, on_error="ignore", return_as="numpy")
    assert not w and iv[1] == 0.0 and not np.isfinite(iv[2]) and not np.isfinite(iv[3])
    with pytest.raises(PriceIsBelowIntrinsic):
        pvv.vectorized_implied_volatility(price, S, 100.0, .5, .01, flag, on_error="raise")
    with pytest.raises(PriceIsAboveMaximum):
        pvv.vectorized_implied_volatility(price[2:], S[2:], 100.0, .5, .01, flag[2:], on_error="raise")
    for kw in (dict(t=0.0), dict(K=0.0), dict(S=0.0)):
        a = dict(price=5.0, S=100.0, K=100.0, t=.5, r=.01, flag=['c'])
        a.update(kw)
        with pytest.raises(ZeroDivisionError), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pvv.vectorized_implied_volatility(**a, on_error="ignore")
    with pytest.raises(PriceIsBelowIntrinsic), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pvv.vectorized_implied_volatility(5.0, 100.0, 0.0, .5, .01, ['c'], on_error="raise")   # intrinsic check fires first
    with pytest.raises(ZeroDivisionError):
        pvv.vectorized_black_scholes(['c'], 100.0, 0.0, .5, .01, .2)
    # t < 0 and NaN price: NaN without a warning; price == 0 out of the money: 0.0
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        iv = pvv.vectorized_implied_volatility([5.0, np.nan, 0.0], [100.0, 100.0, 90.0], 100.0, [-0.5, 0.5, 0.5], .01, ['c', 'c', 'c'], return_as="numpy")
    assert not w and np.isnan(iv[0]) and np.isnan(iv[1]) and iv[2] == 0.0
    # exp(-r t) == 0
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        iv = pvv.vectorized_implied_volatility(5.0, 100.0, 100.0, 1.0, 3000.0, ['c'], return_as="numpy")
    assert any("Interest Free Rate" in str(x.message) for x in w) and np.isnan(iv[0])
    with pytest.raises(ValueError):
        pvv.vectorized_implied_volatility(5.0, 100.0, 100.0, 1.0, 3000.0, ['c'], on_error="raise")


def test_pricing_edge_cases(pvv):
    """SURVEY.md Appendix B, pricing and greeks rows."""
    p = pvv.vectorized_black_scholes(['c', 'c', 'c', 'p'], [110., 90., 100., 90.], 100., 0.0, .05, .2, return_as="numpy")
    assert p.tolist() == [10.0, 0.0, 0.0, 10.0]
    assert np.isnan(pvv.vectorized_black_scholes(['c'], 100., 100., -0.5, .05, .2, return_as="numpy")[0])
    assert abs(pvv.vectorized_black_scholes(['c'], 110., 100., .5, .05, 0.0, return_as="numpy")[0] - 12.469008797154906) < 1e-9
    g = pvv.get_all_greeks(['c', 'p', 'c', 'p'], [100., 100., 110., 110.], 100., 0.0, .05, .2, return_as="dict")
    assert g["delta"].tolist() == [0.5, -0.5, 1.0, 0.0] and np.isinf(g["gamma"][:2]).all() and g["gamma"][2:].tolist() == [0.0, 0.0]
    assert g["rho"].tolist() == [0.0] * 4 and g["vega"].tolist() == [0.0] * 4


def test_scalar_q_broadcast_and_reference_bug_fixes(pvv, oracle):
    """Scalar q with vector inputs: rho raises like the reference (greeks.py:240-242), the other four greeks broadcast
    it.  fix_reference_bugs=True: rho broadcasts too, and model='black' differentiates the discounted Black-76 price
    (= the Black-Scholes-Merton stencil at zero carry), here checked against the oracle's BSM greeks with q = r."""
    fl, S, K, t, r, sg = ['c', 'p', 'c'], 95.0, [100.0, 90.0, 95.0], .2, .05, [.2, .3, .25]
    n = 3
    f8 = np.array([1, -1, 1], dtype=np.int8)
    full = lambda v: np.full(n, v) if np.isscalar(v) else np.asarray(v, dtype=np.float64)  # noqa: E731
    ref, _ = oracle.greeks("black_scholes_merton", f8, full(S), full(K), full(t), full(r), full(sg), full(0.01))
    fns = {"delta": pvv.vectorized_delta, "theta": pvv.vectorized_theta, "vega": pvv.vectorized_vega, "gamma": pvv.vectorized_gamma}
    for name, fn in fns.items():
        got = fn(fl, S, K, t, r, sg, q=0.01, model="black_scholes_merton", return_as="numpy")
        assert_bit_equal(ref[name], got, name)
    with pytest.raises(ValueError, match="same number of elements"):
        pvv.vectorized_rho(fl, S, K, t, r, sg, q=0.01, model="black_scholes_merton")
    got = pvv.vectorized_rho(fl, S, K, t, r, sg, q=0.01, model="black_scholes_merton", return_as="numpy", fix_reference_bugs=True)
    assert_bit_equal(ref["rho"], got, "rho (fixed)")
    # Black-76, fixed: zero carry
    ref0, _ = oracle.greeks("black_scholes_merton", f8, full(S), full(K), full(t), full(r), full(sg), full(r))
    allg = pvv.get_all_greeks(fl, S, K, t, r, sg, model="black", return_as="dict", fix_reference_bugs=True)
    for name in ("delta", "gamma", "theta", "rho", "vega"):
        assert_bit_equal(ref0[name], allg[name], "black " + name)
        one = getattr(pvv, "vectorized_" + name)(fl, S, K, t, r, sg, model="black", return_as="numpy", fix_reference_bugs=True)
        assert_bit_equal(ref0[name], one, "black " + name)
    # sanity: the discounted Black-76 rho is -t * price (py_vollib.black.greeks.analytical.rho, per point: / 100)
    price = np.exp(-r * t) * pvv.vectorized_black(fl, S, K, t, r, sg, return_as="numpy")
    assert np.allclose(allg["rho"], -t * price * 0.01, rtol=1e-4)
    with pytest.raises(TypeError):
        pvv.vectorized_delta(fl, S, K, t, r, sg, model="black")
