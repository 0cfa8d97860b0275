"""GPU tier: the opt-in closed-form greeks kernel against the closed-form
oracle (1e-12 relative) and the analytical columns of the reference's regression table."""
import numpy as np
import pytest
from numpy.testing import assert_array_almost_equal

pytestmark = pytest.mark.gpu
NAMES = ("price", "delta", "gamma", "theta", "rho", "vega")


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from py_vollib_vectorized_b200 import _engine
    return _engine


def col(a):
    return (np.ascontiguousarray(a), 1)


@pytest.mark.parametrize("model", ["black_scholes", "black_scholes_merton"])
def test_analytic_greeks_vs_closed_form_oracle(eng, oracle, model):
    from py_vollib_vectorized_b200.util.chains import chain_cfg2
    n = 1_000_000
    c = chain_cfg2(n, seed=5)
    q = col(c["q"]) if model == "black_scholes_merton" else None
    out = eng.analytic_greeks(model, n, col(c["flag"]), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), q, col(c["sigma"]), with_price=True)
    ref = oracle.analytic_greeks(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=oracle.max_threads)
    # tolerance of the analytical mode: 1e-12 relative — to the value, or where the closed form itself
    # subtracts two nearly equal terms (price, theta), to the size of those terms
    from scipy.special import ndtr
    sgn = np.where(c["flag"] > 0, 1.0, -1.0)
    qq = c["q"] if model == "black_scholes_merton" else 0.0
    sq = np.sqrt(c["t"])
    d1 = (np.log(c["S"] / c["K"]) + (c["r"] - qq + 0.5 * c["sigma"] ** 2) * c["t"]) / (c["sigma"] * sq)
    term_p = c["S"] * np.exp(-qq * c["t"]) * ndtr(sgn * d1)
    term_t = (c["S"] * np.exp(-0.5 * d1 * d1) * c["sigma"] / sq + c["r"] * term_p + np.abs(qq) * term_p) / 365.0
    terms = {"price": term_p, "theta": term_t}
    for k in NAMES:
        tol = 1e-12 * (np.abs(ref[k]) + terms.get(k, 0.0)) + 1e-300
        err = np.abs(out[k] - ref[k])
        assert (err <= tol).all(), (k, float((err / tol).max()))
        rel = err / np.maximum(np.abs(ref[k]), 1e-300)
        assert np.quantile(rel, 0.5) <= 1e-15, (k, np.quantile(rel, 0.5))


def test_analytic_greeks_regression_table(golden):
    """The JSON table's CD/CG/CT/CV/CR, PD/PG/PT/PV/PR columns ARE analytical greeks (raw theta per year,
    raw vega / rho): py_vollib's conventions divide theta by 365 and scale vega, rho by 0.01."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import py_vollib_vectorized_b200 as pvv
    g = golden("cfg1_json")
    n = g["S"].size
    for fl, D, G, T, V, R in (("c", "CD", "CG", "CT", "CV", "CR"), ("p", "PD", "PG", "PT", "PV", "PR")):
        a = pvv.get_all_greeks([fl] * n, g["S"], g["K"], g["t"], g["R"], g["v"], model="black_scholes", method="analytical", return_as="dict")
        assert_array_almost_equal(a["delta"], g[D])
        assert_array_almost_equal(a["gamma"], g[G])
        assert_array_almost_equal(a["theta"] * 365.0, g[T], 4)
        assert_array_almost_equal(a["vega"], g[V] * .01, 5)
        assert_array_almost_equal(a["rho"], g[R] * .01, 5)
        # and the finite-difference greeks of the reference agree with them to FD accuracy
        f = pvv.get_all_greeks([fl] * n, g["S"], g["K"], g["t"], g["R"], g["v"], model="black_scholes", return_as="dict")
        assert np.abs(f["delta"] - a["delta"]).max() < 1e-6
    with pytest.raises(ValueError):
        pvv.get_all_greeks(['c'], 100, 100, .5, .01, .2, model="black", method="analytical")
    with pytest.raises(ValueError):
        pvv.get_all_greeks(['c'], 100, 100, .5, .01, .2, method="exact")


@pytest.mark.parametrize("model", ["black_scholes", "black_scholes_merton"])
def test_analytic_streaming_and_general_forms_agree(eng, model):
    """The dense, 16-byte-aligned, all-outputs call runs the two-contracts-per-thread streaming kernel (128-bit
    accesses); an odd length leaves one contract to the general kernel, and misaligned or broadcast columns or a
    missing output take the general kernel for the whole call.  Same arithmetic: identical bits."""
    import torch
    from py_vollib_vectorized_b200.util.chains import chain_cfg2
    n = 200_001  # odd: streaming form + one-contract tail
    c = chain_cfg2(n + 1, seed=9)
    d = {k: torch.from_numpy(np.ascontiguousarray(c[k])).cuda() for k in ("flag", "S", "K", "t", "r", "q", "sigma")}
    bsm = model == "black_scholes_merton"

    def run(sl, names=NAMES, align_out=True):
        m = sl.stop - sl.start
        outs = {k: torch.empty(m + 1, dtype=torch.float64, device="cuda")[(0 if align_out else 1):][:m] for k in names}
        eng.analytic_greeks_dev(model, m, d["flag"][sl], d["S"][sl], d["K"][sl], d["t"][sl], d["r"][sl], d["q"][sl] if bsm else None,
                                d["sigma"][sl], outs)
        torch.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in outs.items()}

    fast = run(slice(0, n))                        # aligned inputs and outputs: streaming + tail
    shifted = run(slice(1, n))                     # inputs start 8 bytes (flag: 1 byte) off alignment: general form
    no_price = run(slice(0, n), names=NAMES[1:])   # one output not wanted: general form
    odd_out = run(slice(0, n), align_out=False)    # outputs off alignment: general form
    for k in NAMES:
        assert np.array_equal(fast[k][1:], shifted[k], equal_nan=True), k
        assert np.array_equal(fast[k], odd_out[k], equal_nan=True), k
        if k != "price":
            assert np.array_equal(fast[k], no_price[k], equal_nan=True), k
