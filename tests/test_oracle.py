"""CPU tier: the C oracle against the committed golden vectors (generated from the reference by
oracle/gen_golden.py) and the reference's own regression table / docstring known answers."""
import numpy as np
import pytest
from numpy.testing import assert_array_almost_equal

from tests.util import assert_bit_equal

INPUTS = ("flag", "S", "K", "t", "r", "sigma", "q")
GREEKS = ("delta", "gamma", "theta", "rho", "vega")


@pytest.mark.parametrize("chain", ["cfg2", "wide", "edge"])
def test_prices_bit_identical_to_reference(oracle, golden, chain):
    g = golden(chain)
    for model, tag in (("black_scholes", "bs"), ("black_scholes_merton", "bsm"), ("black", "black")):
        p, zde = oracle.price(model, g["flag"], g["S"], g["K"], g["t"], g["r"], g["sigma"], g["q"])
        assert zde == -1
        assert_bit_equal(g[f"price_{tag}"], p, f"{chain} price {model}")


@pytest.mark.parametrize("chain", ["cfg2", "wide", "edge"])
def test_fd_greeks_bit_identical_to_reference(oracle, golden, chain):
    g = golden(chain)
    for model, tag in (("black_scholes", "bs"), ("black_scholes_merton", "bsm")):
        o, zde = oracle.greeks(model, g["flag"], g["S"], g["K"], g["t"], g["r"], g["sigma"], g["q"])
        assert zde == -1
        for k in GREEKS:
            assert_bit_equal(g[f"{k}_{tag}"], o[k], f"{chain} {k} {model}")
        assert_bit_equal(g[f"price_{tag}"], o["price"], f"{chain} fused price {model}")


@pytest.mark.parametrize("chain", ["cfg2", "wide"])
def test_iv_numba_loop_bit_identical(oracle, golden, chain):
    g = golden(chain)
    iv, st, zde = oracle.iv("black", g["price_black"], g["S"], g["K"], g["t"], 0.0, g["flag"], mask_errors=False)
    assert zde == -1
    assert_bit_equal(g["ivloop_black"], iv, f"{chain} IV loop")


def test_iv_edge_rows(oracle, golden):
    g = golden("edge")
    k = g["ivloop_keep"]
    iv, st, zde = oracle.iv("black", g["price"][k], g["S"][k], g["K"][k], g["t"][k], 0.0, g["flag"][k], mask_errors=False)
    assert zde == -1
    assert_bit_equal(g["ivloop_black"], iv, "edge IV loop")
    for on_error in ("warn", "ignore"):
        iv, st, zde = oracle.iv("black_scholes_merton", g["price"][k], g["S"][k], g["K"][k], g["t"][k], g["r"][k], g["flag"][k],
                                g["q"][k], mask_errors=(on_error != "ignore"))
        assert_bit_equal(g[f"iv_bsm_{on_error}_glibc"], iv, f"edge public IV {on_error}")


@pytest.mark.parametrize("model,tag", [("black_scholes_merton", "bsm"), ("black_scholes", "bs"), ("black", "black")])
@pytest.mark.parametrize("on_error", ["warn", "ignore"])
def test_public_iv_cfg3(oracle, golden, model, tag, on_error):
    g = golden("cfg3")
    iv, st, zde = oracle.iv(model, g["price"], g["S"], g["K"], g["t"], g["r"], g["flag"], g["q"], mask_errors=(on_error != "ignore"))
    assert zde == -1
    # the CPU-independent flavour of the reference (numpy -> glibc exp): exact
    assert_bit_equal(g[f"iv_{tag}_{on_error}_glibc"], iv, f"cfg3 IV {model} {on_error}")
    # the AVX-512 numpy flavour of the same reference run: 1-ulp different forwards; away from the
    # ill-conditioned rows (price - intrinsic cancellation) results agree to 1e-10 absolute
    simd = g[f"iv_{tag}_{on_error}_simd"]
    both = np.isfinite(simd) & np.isfinite(iv)
    d = np.abs(simd - iv)[both]
    assert (d > 1e-10).mean() < 5e-3


def test_status_bits(oracle):
    #            below    above(call)  above(put)  ok     t==0   K==0
    price = [1.0, 150.0, 100.0, 5.0, 5.0, 5.0]
    S = [110.0, 100.0, 100.0, 100.0, 100.0, 100.0]
    K = [100.0, 100.0, 100.0, 100.0, 100.0, 0.0]
    t = [0.5, 0.5, 0.5, 0.5, 0.0, 0.5]
    flag = [1, 1, -1, 1, 1, 1]
    iv, st, zde = oracle.iv("black_scholes", price, S, K, t, 0.01, flag)
    assert st.tolist()[:4] == [1, 2, 2, 0]
    assert st[4] & 4 and st[5] & 4 and zde == 4
    assert np.isnan(iv[:3]).all() and abs(iv[3] - 0.16877851941995567) < 1e-12
    iv2, st2, _ = oracle.iv("black_scholes", price[:4], S[:4], K[:4], t[:4], 0.01, flag[:4], mask_errors=False)
    assert iv2[0] == 0.0 and not np.isfinite(iv2[1]) and not np.isfinite(iv2[2])  # SURVEY.md Appendix B, on_error="ignore"


def test_reference_regression_table(oracle, golden):
    """tests/test_entrypoints.py:70-152 of the reference, on the oracle."""
    g = golden("cfg1_json")
    n = g["S"].size
    assert n == 101
    for fl, tag, col, D, G, V in ((1, "c", "bs_call", "CD", "CG", "CV"), (-1, "p", "bs_put", "PD", "PG", "PV")):
        flag = np.full(n, fl, dtype=np.int8)
        p, _ = oracle.price("black_scholes", flag, g["S"], g["K"], g["t"], g["R"], g["v"])
        assert_array_almost_equal(p, g[col])
        pm, _ = oracle.price("black_scholes_merton", flag, g["S"], g["K"], g["t"], g["R"], g["v"], np.zeros(n))
        assert_array_almost_equal(pm, g[col])
        o, _ = oracle.greeks("black_scholes", flag, g["S"], g["K"], g["t"], g["R"], g["v"])
        assert_array_almost_equal(o["delta"], g[D])
        assert_array_almost_equal(o["gamma"], g[G])
        assert_array_almost_equal(o["vega"], g[V] * .01, 3)
        assert_bit_equal(g[f"ref_{tag}_price_bs"], p, "json price")
        for k in GREEKS:
            assert_bit_equal(g[f"ref_{tag}_{k}_bs"], o[k], f"json {k}")
        iv, st, _ = oracle.iv("black_scholes", p, g["S"], g["K"], g["t"], g["R"], flag)
        assert_bit_equal(g[f"ref_{tag}_ivroundtrip_glibc"], iv, "json IV round trip")
        ok = np.isclose(g["v"], iv, atol=1e-2)
        ok[~ok] = iv[~ok] == 0
        assert ok.all()


def test_docstring_known_answers(oracle, golden):
    k = golden("kat")
    p, _ = oracle.price("black", [1, -1], 95, [100, 90], .2, .2, .2)
    assert_bit_equal(k["black_doc"], p)
    np.testing.assert_allclose(p, [1.53408169, 1.38409245], rtol=0, atol=5e-9)
    p, _ = oracle.price("black_scholes", [1, -1], 95, [100, 90], .2, .2, .2)
    assert_bit_equal(k["bs_doc"], p)
    np.testing.assert_allclose(p, [2.89558836, 0.61109351], rtol=0, atol=5e-9)
    p, _ = oracle.price("black_scholes_merton", [1, -1], [95, 99], [100, 90], .2, .2, .2, 0.0)
    np.testing.assert_allclose(p, [2.89558836, 0.23536284], rtol=0, atol=5e-9)
    iv, _, _ = oracle.iv("black_scholes_merton", .10, 95, [100, 90], .2, .2, [1, -1], 0.0)
    assert_bit_equal(k["iv_bsm_doc"], iv)
    np.testing.assert_allclose(iv, [0.02621257, 0.12585767], rtol=0, atol=5e-9)
    iv, _, _ = oracle.iv("black", .10, 95, [100, 90], .2, .2, [1, -1])
    assert_bit_equal(k["iv_black_doc"], iv)
    for model, tag, args in (("black_scholes", "bs", ([1, -1], 95, [100, 90], .2, .2, .2)),
                             ("black_scholes_merton", "bsm", ([1, -1], 100, [100, 90], .5, .05, .2, .02)),
                             ("black", "black", ([1, -1], 95, [100, 90], .2, .2, .2))):
        o, _ = oracle.greeks(model, *args)
        for g in GREEKS:
            assert_bit_equal(k[f"greeks_{tag}_{g}"], o[g], f"{model} {g}")


def test_region_and_branch_coverage(oracle):
    regions = {oracle.normalised_black_call(x, s)[1] for x, s in ((-0.5, 0.02), (-0.05, 0.3), (0.5, 3.0), (-0.5, 1.0), (-1.0, 0.0))}
    assert regions == {0, 1, 2, 3, 4}
    for x, s, br in ((-0.5, 0.15, 1), (-0.5, 0.6, 2), (-0.5, 1.5, 3), (-0.5, 4.0, 4)):
        b, _ = oracle.normalised_black_call(x, s)
        s_out, branch = oracle.normalised_iv(b, x)
        assert branch == br and abs(s_out - s) < 1e-14 * max(1, s)


def test_threads_do_not_change_bits(oracle, golden):
    g = golden("cfg2")
    a, _ = oracle.greeks("black_scholes_merton", g["flag"], g["S"], g["K"], g["t"], g["r"], g["sigma"], g["q"], threads=1)
    b, _ = oracle.greeks("black_scholes_merton", g["flag"], g["S"], g["K"], g["t"], g["r"], g["sigma"], g["q"], threads=4)
    for k in a:
        assert_bit_equal(a[k], b[k])


def test_closed_form_greeks_oracle_is_pinned_by_the_reference_table(oracle, golden):
    """The checker of the opt-in analytical mode: py_vollib greeks.analytical conventions, pinned by the
    analytical columns of the reference's own regression table (6 significant digits)."""
    g = golden("cfg1_json")
    n = g["S"].size
    for fl, D, G, T, V, R, P in ((1, "CD", "CG", "CT", "CV", "CR", "bs_call"), (-1, "PD", "PG", "PT", "PV", "PR", "bs_put")):
        a = oracle.analytic_greeks("black_scholes", np.full(n, fl, np.int8), g["S"], g["K"], g["t"], g["R"], g["v"])
        assert_array_almost_equal(a["delta"], g[D])
        assert_array_almost_equal(a["gamma"], g[G])
        assert_array_almost_equal(a["theta"] * 365, g[T], 4)
        assert_array_almost_equal(a["vega"], g[V] * .01, 5)
        assert_array_almost_equal(a["rho"], g[R] * .01, 5)
        assert_array_almost_equal(a["price"], g[P], 5)
        b = oracle.analytic_greeks("black_scholes_merton", np.full(n, fl, np.int8), g["S"], g["K"], g["t"], g["R"], g["v"], np.zeros(n))
        for k in a:
            assert np.array_equal(a[k], b[k])


def test_closed_form_greeks_oracle_with_dividends_is_pinned_by_mpmath(oracle):
    """The reference's regression table has q = 0 only: the Black-Scholes-Merton closed forms with a dividend yield are
    pinned here by an independent 40-digit evaluation (mpmath) in py_vollib's greeks.analytical conventions
    (theta per calendar day, vega and rho per point)."""
    import mpmath as mp
    from py_vollib_vectorized_b200.util.chains import chain_cfg2
    mp.mp.dps = 40
    c = chain_cfg2(300, seed=9)
    c["q"] = np.where(np.arange(300) % 3 == 0, 0.0, c["q"] + 0.01 * (np.arange(300) % 7))  # q from 0 to 0.1
    got = oracle.analytic_greeks("black_scholes_merton", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"])
    N = lambda x: mp.ncdf(x)  # noqa: E731
    phi = lambda x: mp.npdf(x)  # noqa: E731
    for i in range(300):
        f, S, K, t, r, sg, q = (mp.mpf(float(c[k][i])) for k in ("flag", "S", "K", "t", "r", "sigma", "q"))
        sq = mp.sqrt(t)
        d1 = (mp.log(S / K) + (r - q + sg * sg / 2) * t) / (sg * sq)
        d2 = d1 - sg * sq
        Dq, Dr = mp.exp(-q * t), mp.exp(-r * t)
        ref = {"price": f * (S * Dq * N(f * d1) - K * Dr * N(f * d2)), "delta": f * Dq * N(f * d1), "gamma": Dq * phi(d1) / (S * sg * sq),
               "vega": S * Dq * phi(d1) * sq / 100,
               "theta": (-S * Dq * phi(d1) * sg / (2 * sq) - f * r * K * Dr * N(f * d2) + f * q * S * Dq * N(f * d1)) / 365,
               "rho": f * t * K * Dr * N(f * d2) / 100}
        # price and theta subtract nearly equal terms: 1e-12 relative to the size of those terms, else to the value
        scale = {"price": S * Dq * N(f * d1) + K * Dr * N(f * d2), "theta": (S * Dq * phi(d1) * sg / (2 * sq) + r * K * Dr + q * S * Dq) / 365}
        for k, v in ref.items():
            tol = 1e-12 * float(scale.get(k, abs(v))) + 1e-300
            assert abs(float(got[k][i]) - float(v)) <= tol, (i, k, float(got[k][i]), float(v))
