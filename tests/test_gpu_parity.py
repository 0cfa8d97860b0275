"""GPU tier: the CUDA engine (through the C ABI, host-buffer and device-pointer entry points) against
the CPU oracle and the committed golden vectors generated from the reference.

Tolerances (BASELINE.json north_star): prices and greeks 1e-12 relative, IV 1e-10 absolute, NaN /
error positions identical.  Prices and greeks are in fact required to be BIT-IDENTICAL here (that is
what makes the finite-difference greeks meet 1e-12); so is IV, including the lowest solver branch
whose pow(x, 1/3) is glibc's algorithm too (pow_g).
"""
import warnings

import numpy as np
import pytest

from tests.util import assert_bit_equal, bit_equal

pytestmark = pytest.mark.gpu

GREEKS = ("delta", "gamma", "theta", "rho", "vega")


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from py_vollib_vectorized_b200 import _engine
    return _engine


def col(a):
    a = np.ascontiguousarray(a)
    return (a, 1)


def run_price(eng, model, g):
    n = g["S"].size
    return eng.price(model, n, col(g["flag"]), col(g["S"]), col(g["K"]), col(g["t"]), col(g["r"]),
                     col(g["q"]) if model == "black_scholes_merton" else None, col(g["sigma"]))


def run_greeks(eng, model, g, **kw):
    n = g["S"].size
    return eng.greeks(model, n, col(g["flag"]), col(g["S"]), col(g["K"]), col(g["t"]), col(g["r"]),
                      col(g["q"]) if model == "black_scholes_merton" else None, col(g["sigma"]), **kw)


def run_iv(eng, model, g, price, mask=True):
    n = g["S"].size
    return eng.iv(model, n, col(g["flag"]), col(price), col(g["S"]), col(g["K"]), col(g["t"]), col(g["r"]),
                  col(g["q"]) if model == "black_scholes_merton" else None, mask_errors=mask)


def test_device_exp_log_are_the_host_libm(eng, oracle):
    """exp_g / log_g must return glibc's bits (the reference's numba loops call the host libm)."""
    import torch
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.uniform(-745, 709, 200000), rng.uniform(-2, 2, 200000), rng.normal(0, 1e-3, 50000),
                        rng.uniform(-1100, 1100, 20000), [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-300, 709.79, -745.2, -708.4]])
    y = eng.unary_dev("exp", torch.from_numpy(x).cuda()).cpu().numpy()
    ref = np.array([oracle.lib.pvv_oracle_exp(v) for v in x])
    assert_bit_equal(ref, y, "exp")
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 200000)), rng.uniform(0.9, 1.1, 200000), rng.uniform(0, 3, 50000),
                        [0.0, -1.0, np.inf, np.nan, 5e-324, 1e-310, 1.0, 2.2250738585072014e-308]])
    y = eng.unary_dev("log", torch.from_numpy(x).cuda()).cpu().numpy()
    ref = np.array([oracle.lib.pvv_oracle_log(v) for v in x])
    assert_bit_equal(ref, y, "log")


def test_device_pow_is_the_host_libm(eng):
    import math
    import torch
    rng = np.random.default_rng(13)
    x = np.concatenate([np.exp(rng.uniform(-700, 10, 200000)), rng.uniform(0, 1, 200000), rng.uniform(0.9, 1.1, 50000),
                        1 - np.exp(rng.uniform(-40, -1, 50000)), np.exp(rng.uniform(-745, -708, 20000)), [1.0, 0.5, 1e-300, 2.0, 5e-324, 1e-310]])
    y = eng.unary_dev("cbrt_pow", torch.from_numpy(x).cuda()).cpu().numpy()
    ref = np.array([math.pow(v, 1. / 3.) for v in x])
    assert_bit_equal(ref, y, "pow(x, 1/3)")


def test_guard_free_division_and_sqrt_are_ieee_inside_their_bounds(eng):
    """div_bounded / sqrt_bounded are the fast paths of nvcc's own fp64 divider and square root without the operand
    guards (csrc/pvv_math.cuh): inside the documented bounds they must equal the correctly rounded IEEE results —
    numpy's `/` and `sqrt` on the host — on every operand, including the edges of the bounds and hard-to-round cases."""
    import torch
    rng = np.random.default_rng(21)
    m = 2_000_000
    mag = np.exp2(rng.uniform(-300, 300, m))
    a = np.concatenate([mag * rng.uniform(1, 2, m), rng.uniform(0.02, 200.0, m), 1.0 + rng.integers(0, 1 << 20, m) * 2.0 ** -52,
                        [2.0 ** -300, 2.0 ** 300, 0.02, 1.0, 4.0, 2.0, 3.0, 1e-13, 2844.236833439171]])
    a = a[: a.size // 2 * 2]
    y = eng.unary_dev("sqrt_bounded", torch.from_numpy(a).cuda()).cpu().numpy()
    assert_bit_equal(np.sqrt(a), y, "sqrt_bounded")
    # pairs (numerator, denominator) and (denominator, numerator): quotients between 2^-600 and 2^600
    signs = np.where(rng.random(a.size) < 0.5, -1.0, 1.0)
    x = a * signs
    q = eng.unary_dev("div_bounded_pairs", torch.from_numpy(x).cuda()).cpu().numpy()
    partner = x.reshape(-1, 2)[:, ::-1].reshape(-1)
    assert_bit_equal(x / partner, q, "div_bounded")


@pytest.mark.parametrize("name", ["erfcx", "erfc", "norm_cdf", "inverse_norm_cdf"])
def test_device_special_functions(eng, oracle, name):
    import torch
    rng = np.random.default_rng(12)
    if name == "inverse_norm_cdf":
        x = np.concatenate([rng.uniform(0, 1, 100000), np.exp(rng.uniform(-700, 0, 50000)), [0.0, 1.0, 0.5, -1.0, 2.0]])
    else:
        x = np.concatenate([rng.uniform(-30, 30, 150000), rng.uniform(-1, 1, 50000), rng.uniform(-1e5, 1e9, 1000), [0.0, np.nan, np.inf, -np.inf]])
    y = eng.unary_dev(name, torch.from_numpy(x).cuda()).cpu().numpy()
    fn = getattr(oracle.lib, "pvv_oracle_" + name)
    ref = np.array([fn(v) for v in x])
    assert_bit_equal(ref, y, name)


@pytest.mark.parametrize("chain", ["cfg2", "wide", "edge"])
def test_prices_golden(eng, golden, chain):
    g = golden(chain)
    for model, tag in (("black_scholes", "bs"), ("black_scholes_merton", "bsm"), ("black", "black")):
        assert_bit_equal(g[f"price_{tag}"], run_price(eng, model, g), f"{chain} {model}")


@pytest.mark.parametrize("chain", ["cfg2", "wide", "edge"])
def test_fd_greeks_golden(eng, golden, chain):
    g = golden(chain)
    for model, tag in (("black_scholes", "bs"), ("black_scholes_merton", "bsm")):
        o = run_greeks(eng, model, g, with_price=True)
        assert_bit_equal(g[f"price_{tag}"], o["price"], f"{chain} fused price {model}")
        for k in GREEKS:
            assert_bit_equal(g[f"{k}_{tag}"], o[k], f"{chain} {k} {model}")
            rel = np.abs(o[k] - g[f"{k}_{tag}"]) / np.maximum(np.abs(g[f"{k}_{tag}"]), 1e-300)
            assert np.nanmax(np.where(np.isfinite(rel), rel, 0)) <= 1e-12
        # each greek on its own evaluates only its stencil points and must give the same bits
        for k in GREEKS:
            assert_bit_equal(o[k], run_greeks(eng, model, g, want=(k,))[k], f"{chain} single {k}")


@pytest.mark.parametrize("model,tag", [("black_scholes_merton", "bsm"), ("black_scholes", "bs"), ("black", "black")])
@pytest.mark.parametrize("on_error", ["warn", "ignore"])
def test_iv_golden_cfg3(eng, oracle, golden, model, tag, on_error):
    g = golden("cfg3")
    iv, status, zde = run_iv(eng, model, g, g["price"], mask=(on_error != "ignore"))
    ref = g[f"iv_{tag}_{on_error}_glibc"]
    assert zde == -1
    assert (np.isnan(iv) == np.isnan(ref)).all(), "NaN positions differ"
    inf = np.isinf(ref)
    assert (np.isinf(iv) == inf).all() and (iv[inf] == ref[inf]).all()
    fin = np.isfinite(ref)
    assert np.abs(iv[fin] - ref[fin]).max() <= 1e-10
    _, st_ref, _, br = oracle.iv(model, g["price"], g["S"], g["K"], g["t"], g["r"], g["flag"], g["q"], mask_errors=(on_error != "ignore"),
                                 want_branch=True)
    assert (status == st_ref).all(), "status bytes differ from the oracle"
    assert_bit_equal(ref, iv, "IV (all four solver branches)")
    assert set(np.unique(br)) >= {1, 2, 3}


def test_iv_loop_and_edges(eng, golden):
    for chain in ("cfg2", "wide"):
        g = golden(chain)
        z = dict(g, r=np.zeros_like(g["r"]))
        iv, status, zde = run_iv(eng, "black", z, g["price_black"], mask=False)
        ref = g["ivloop_black"]
        assert zde == -1
        assert (np.isnan(iv) == np.isnan(ref)).all()
        fin = np.isfinite(ref) & np.isfinite(iv)
        assert (np.isfinite(ref) == np.isfinite(iv)).all()
        assert np.abs(iv[fin] - ref[fin]).max() <= 1e-10, chain
        assert_bit_equal(ref, iv, f"{chain} IV numba loop")
    g = golden("edge")
    k = g["ivloop_keep"]
    gk = {n: g[n][k] for n in ("flag", "S", "K", "t", "r", "q", "sigma", "price")}
    for on_error in ("warn", "ignore"):
        iv, status, zde = run_iv(eng, "black_scholes_merton", gk, gk["price"], mask=(on_error != "ignore"))
        ref = g[f"iv_bsm_{on_error}_glibc"]
        assert (np.isnan(iv) == np.isnan(ref)).all()
        fin = np.isfinite(ref)
        assert (np.isfinite(iv) == fin).all()
        assert np.abs(iv[fin] - ref[fin]).max() <= 1e-10
        assert_bit_equal(ref, iv, f"edge public IV {on_error}")


def test_reference_regression_table(eng, golden):
    """The reference's own test_entrypoints.py:24-152 assertions on its 101-row JSON table."""
    from numpy.testing import assert_array_almost_equal
    g = golden("cfg1_json")
    n = g["S"].size
    for fl, tag, colname, D, G, V in ((1, "c", "bs_call", "CD", "CG", "CV"), (-1, "p", "bs_put", "PD", "PG", "PV")):
        c = dict(flag=np.full(n, fl, dtype=np.int8), S=g["S"], K=g["K"], t=g["t"], r=g["R"], sigma=g["v"], q=np.zeros(n))
        p = run_price(eng, "black_scholes", c)
        assert_array_almost_equal(p, g[colname])
        assert_array_almost_equal(run_price(eng, "black_scholes_merton", c), g[colname])
        o = run_greeks(eng, "black_scholes", c)
        assert_array_almost_equal(o["delta"], g[D])
        assert_array_almost_equal(o["gamma"], g[G])
        assert_array_almost_equal(o["vega"], g[V] * .01, 3)
        assert_bit_equal(g[f"ref_{tag}_price_bs"], p)
        for k in GREEKS:
            assert_bit_equal(g[f"ref_{tag}_{k}_bs"], o[k], k)
        iv, status, _ = run_iv(eng, "black_scholes", c, p)
        ref = g[f"ref_{tag}_ivroundtrip_glibc"]
        assert (np.isnan(iv) == np.isnan(ref)).all()
        assert np.nanmax(np.abs(iv - ref)) <= 1e-10
        assert_bit_equal(ref, iv, "json IV round trip")
        ok = np.isclose(g["v"], iv, atol=1e-2)
        ok[~ok] = iv[~ok] == 0
        assert ok.all()


def test_zero_division_is_reported(eng):
    one = lambda v: (np.array([v], dtype=np.float64), 1)  # noqa: E731
    f = (np.array([1], dtype=np.int8), 1)
    with pytest.raises(ZeroDivisionError):
        eng.price("black_scholes", 1, f, one(100.0), one(0.0), one(0.5), one(0.01), None, one(0.2))
    for kw in (dict(t=0.0), dict(K=0.0), dict(S=0.0)):
        b = dict(price=5.0, S=100.0, K=100.0, t=0.5, r=0.01)
        b.update(kw)
        iv, status, zde = eng.iv("black_scholes", 1, f, one(b["price"]), one(b["S"]), one(b["K"]), one(b["t"]), one(b["r"]), None)
        assert zde == 0 and status[0] & 4
    # the index reported is the first offending contract, also across pipeline chunks
    n = 50000
    K = np.full(n, 100.0)
    K[[31337, 40000]] = 0.0
    ctx = eng.get_ctx(chunk=8192, nstreams=3)
    iv, status, zde = eng.iv("black_scholes", n, (np.ones(n, dtype=np.int8), 1), (np.full(n, 5.0), 1), (np.full(n, 100.0), 1), (K, 1),
                             (np.full(n, 0.5), 1), (np.full(n, 0.01), 1), None, ctx=ctx)
    assert zde == 31337
    assert np.flatnonzero(status & 4).tolist() == [31337, 40000]


def test_broadcast_scalars_and_chunking(eng, oracle, golden):
    """stride-0 columns and a multi-chunk pipeline give the same bits as one dense launch."""
    g = golden("cfg2")
    n = g["S"].size
    ctx = eng.get_ctx(chunk=1000, nstreams=3)  # 5 chunks, ragged tail
    r0, s0 = 0.03, 0.25
    out = eng.greeks("black_scholes", n, col(g["flag"]), col(g["S"]), col(g["K"]), col(g["t"]), (np.array([r0]), 0), None,
                     (np.array([s0]), 0), with_price=True, ctx=ctx)
    ref, _ = oracle.greeks("black_scholes", g["flag"], g["S"], g["K"], g["t"], r0, s0)
    for k in ref:
        assert_bit_equal(ref[k], out[k], k)
    # empty call
    e = np.empty(0)
    assert eng.price("black_scholes", 0, (np.empty(0, dtype=np.int8), 1), (e, 1), (e, 1), (e, 1), (e, 1), None, (e, 1)).size == 0


def test_device_pointer_entry_points(eng, oracle, golden):
    import torch
    g = golden("cfg3")
    n = g["S"].size
    d = {k: torch.from_numpy(np.ascontiguousarray(g[k])).cuda() for k in ("flag", "S", "K", "t", "r", "sigma", "q", "price")}
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    eng.price_dev("black_scholes_merton", n, d["flag"], d["S"], d["K"], d["t"], d["r"], d["q"], d["sigma"], out)
    ref, _ = oracle.price("black_scholes_merton", g["flag"], g["S"], g["K"], g["t"], g["r"], g["sigma"], g["q"])
    assert_bit_equal(ref, out.cpu().numpy())
    # K4: IV then greeks at the solved vol; greeks are checked against the oracle evaluated at the
    # GPU's own IV (bit-exact), IV against the oracle's (1e-10)
    iv = torch.empty(n, dtype=torch.float64, device="cuda")
    st = torch.empty(n, dtype=torch.uint8, device="cuda")
    outs = {k: torch.empty(n, dtype=torch.float64, device="cuda") for k in GREEKS}
    zde = eng.new_zde_word("cuda")
    eng.iv_greeks_dev("black_scholes_merton", n, d["flag"], d["price"], d["S"], d["K"], d["t"], d["r"], d["q"], iv, st, outs, first_zde=zde)
    torch.cuda.synchronize()
    assert zde.item() == np.iinfo(np.int64).max
    iv_h = iv.cpu().numpy()
    ref_o, st_ref, _ = oracle.iv_greeks("black_scholes_merton", g["price"], g["S"], g["K"], g["t"], g["r"], g["flag"], g["q"])
    assert (st.cpu().numpy() == st_ref).all()
    assert_bit_equal(ref_o["iv"], iv_h, "fused IV")
    for k in GREEKS:
        assert_bit_equal(ref_o[k], outs[k].cpu().numpy(), f"fused {k}")


def test_iv_greeks_routes_agree(eng, oracle, monkeypatch):
    """pvv_iv_greeks_dev's default route (K3, then K2 at the solved volatility, which ORs its own zero divisions into the
    status bytes K3 wrote) against the single fused kernel (PVV_IVG_FUSED=1) and the oracle: values, status and the
    first-ZeroDivisionError index, on a chain with rows whose ONLY zero division sits in the greeks pass
    (exp(-(r + 0.01) t) underflows to zero although exp(-r t) does not), for every subset size the stencil prunes by."""
    import torch
    rng = np.random.default_rng(77)
    n = 70_001  # ragged against every tile size, odd so the status word of the last byte is shared
    S = rng.uniform(50, 150, n)
    K = S * rng.uniform(0.6, 1.5, n)
    t = rng.uniform(0.02, 2.0, n)
    r = rng.uniform(0.0, 0.08, n)
    sig = rng.uniform(0.05, 0.9, n)
    flag = np.where(rng.random(n) < 0.5, 1, -1).astype(np.int8)
    price, _ = oracle.price("black_scholes", flag, S, K, t, r, sig)
    price[rng.random(n) < 0.01] *= 0.2   # below intrinsic: NaN sigma goes into the greeks pass
    price[rng.random(n) < 0.01] *= 50.0  # above maximum
    late = np.array([5, 4097, 4098, 4099, 33_333, n - 1])  # neighbours share a status word
    t[late], r[late] = 100.0, 7.445       # exp(-744.5) is a subnormal, exp(-745.5) is zero
    d = {k: torch.from_numpy(v).cuda() for k, v in dict(flag=flag, price=price, S=S, K=K, t=t, r=r).items()}
    ref, st_ref, zde_ref = oracle.iv_greeks("black_scholes", price, S, K, t, r, flag)
    assert (st_ref[late] & 4).all() and zde_ref == 5
    st_iv = oracle.iv("black_scholes", price, S, K, t, r, flag)[1]
    assert not (st_iv[late] & 4).any()  # the solve alone does not divide by zero on those rows

    def run(want):
        iv = torch.empty(n, dtype=torch.float64, device="cuda")
        st = torch.empty(n, dtype=torch.uint8, device="cuda")
        outs = {k: torch.empty(n, dtype=torch.float64, device="cuda") for k in want}
        zde = eng.new_zde_word("cuda")
        eng.iv_greeks_dev("black_scholes", n, d["flag"], d["price"], d["S"], d["K"], d["t"], d["r"], None, iv, st, outs, first_zde=zde)
        torch.cuda.synchronize()
        return iv.cpu().numpy(), st.cpu().numpy(), {k: v.cpu().numpy() for k, v in outs.items()}, zde.item()

    for want in (GREEKS, ("rho",), ("delta", "vega")):
        monkeypatch.setenv("PVV_IVG_FUSED", "0")
        before = eng.launch_count()
        iv_a, st_a, o_a, z_a = run(want)
        assert eng.launch_count() - before == 2  # K3 + K2
        monkeypatch.setenv("PVV_IVG_FUSED", "1")
        before = eng.launch_count()
        iv_b, st_b, o_b, z_b = run(want)
        assert eng.launch_count() - before == 1
        assert_bit_equal(iv_a, iv_b, "iv")
        assert (st_a == st_b).all() and z_a == z_b
        for k in want:
            assert_bit_equal(o_a[k], o_b[k], k)
        if "rho" in want:  # the r +- 0.01 stencil points are the ones that divide by zero
            assert z_a == zde_ref and (st_a == st_ref).all()
            ok = (st_ref & 4) == 0
            assert_bit_equal(ref["iv"][ok], iv_a[ok], "iv vs oracle")
            for k in want:
                assert_bit_equal(ref[k][ok], o_a[k][ok], f"{k} vs oracle")
    monkeypatch.setenv("PVV_IVG_FUSED", "0")
