"""CPU tier: the PRODUCT's device math header (csrc/pvv_math.cuh) compiled for the host by
tests/native/host_selfcheck.cpp, against the golden vectors generated from the reference and the
host libm.  This is the same source the kernels compile, so it pins the CUDA arithmetic (operation
order, constants, glibc-exact exp/log) before a GPU is involved."""
import ctypes
import math

import numpy as np
import pytest

from tests.util import assert_bit_equal


@pytest.fixture(scope="module")
def H():
    import __graft_entry__ as g
    lib = ctypes.CDLL(g.build_hostcheck())
    for n in ("pvvh_exp", "pvvh_log", "pvvh_const"):
        getattr(lib, n).restype = ctypes.c_double
    lib.pvvh_exp.argtypes = [ctypes.c_double]
    lib.pvvh_log.argtypes = [ctypes.c_double]
    return lib


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_constants(H):
    import sys
    eps = sys.float_info.epsilon
    r16 = math.sqrt(math.sqrt(math.sqrt(math.sqrt(eps))))
    want = [math.sqrt(eps), math.sqrt(math.sqrt(eps)), r16, 2 * r16, math.sqrt(sys.float_info.min), math.sqrt(sys.float_info.max),
            -(1 - math.sqrt(eps)), 2 / (eps * eps), -1 / math.sqrt(eps), 0.01 ** 2.]
    for i, w in enumerate(want):
        assert H.pvvh_const(i) == w, i


def test_exp_log_bit_identical_to_host_libm(H):
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-745, 710, 100000), rng.uniform(-2, 2, 100000), rng.normal(0, 1e-3, 30000),
                        rng.uniform(-1100, 1100, 20000), [0.0, -0.0, np.inf, -np.inf, 1e-300, -1e-300, 709.782712893384, 709.79, -745.2, -745.13, -708.4]])
    y = np.empty_like(x)
    H.pvvh_exp_array(x.size, P(x), P(y))
    ref = np.array([math.exp(v) if v < 709.7827128933841 else math.inf for v in x])
    assert_bit_equal(ref, y, "exp")
    assert math.isnan(H.pvvh_exp(math.nan))
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 100000)), rng.uniform(0.9, 1.1, 100000), rng.uniform(1e-3, 3, 30000),
                        [5e-324, 1e-310, 1.0, 2.2250738585072014e-308, np.inf]])
    y = np.empty_like(x)
    H.pvvh_log_array(x.size, P(x), P(y))
    ref = np.array([math.log(v) if math.isfinite(v) else v for v in x])
    assert_bit_equal(ref, y, "log")
    assert H.pvvh_log(0.0) == -math.inf and math.isnan(H.pvvh_log(-1.0)) and math.isnan(H.pvvh_log(math.nan))


@pytest.mark.parametrize("chain", ["cfg2", "wide", "edge"])
def test_prices_and_greeks_bit_identical_to_reference(H, golden, chain):
    g = golden(chain)
    n = g["S"].size
    c = {k: np.ascontiguousarray(g[k]) for k in ("flag", "S", "K", "t", "r", "q", "sigma")}
    st = np.empty(n, dtype=np.uint8)
    for m, tag in ((0, "black"), (1, "bs"), (2, "bsm")):
        out = np.empty(n)
        H.pvvh_price(m, n, P(c["flag"]), P(c["S"]), P(c["K"]), P(c["t"]), P(c["r"]), P(c["q"]), P(c["sigma"]), P(out), P(st))
        assert_bit_equal(g[f"price_{tag}"], out, f"{chain} price {tag}")
        assert not st.any()
    for m, tag in ((1, "bs"), (2, "bsm")):
        out = np.empty((n, 6))
        H.pvvh_greeks(m, n, P(c["flag"]), P(c["S"]), P(c["K"]), P(c["t"]), P(c["r"]), P(c["q"]), P(c["sigma"]), P(out), P(st))
        for i, k in enumerate(("price", "delta", "gamma", "theta", "rho", "vega")):
            assert_bit_equal(g[f"{k}_{tag}"], out[:, i], f"{chain} {k} {tag}")


@pytest.mark.parametrize("on_error", ["warn", "ignore"])
def test_iv_bit_identical_to_reference(H, golden, on_error):
    g = golden("cfg3")
    n = g["S"].size
    c = {k: np.ascontiguousarray(g[k]) for k in ("flag", "S", "K", "t", "r", "q", "price")}
    for m, tag in ((0, "black"), (1, "bs"), (2, "bsm")):
        iv, st, br = np.empty(n), np.empty(n, dtype=np.uint8), np.empty(n, dtype=np.int8)
        H.pvvh_iv(m, n, P(c["flag"]), P(c["price"]), P(c["S"]), P(c["K"]), P(c["t"]), P(c["r"]), P(c["q"]), int(on_error != "ignore"),
                  P(iv), P(st), P(br))
        # on the host pow() is glibc's too, so even the lowest branch is bit-identical
        assert_bit_equal(g[f"iv_{tag}_{on_error}_glibc"], iv, f"cfg3 IV {tag} {on_error}")
        assert set(np.unique(br)) >= {1, 2, 3}


@pytest.mark.parametrize("c", [6.0, 2844.23683343917062])
def test_zero_safe_constant_division_is_the_ieee_quotient(H, c):
    """div_by_const_zn (Householder's / 6, Cody's small-argument quotient) must return the correctly rounded x / c
    for every x: Markstein's three operations in the ordinary range, the zero itself for +-0, the divider elsewhere."""
    rng = np.random.default_rng(17)
    n = 4_000_000
    mant = rng.uniform(1.0, 2.0, n)
    x = np.concatenate([mant * 2.0 ** rng.integers(-1074, 1023, n), mant[:n // 4] * 2.0 ** rng.integers(-60, 10, n // 4),
                        # quotients next to a rounding boundary: c * (k + 1/2 ulp) patterns
                        c * (rng.integers(1, 2 ** 52, n // 4).astype(np.float64) * 2.0 ** -52 + 1.0) * (1 + rng.integers(-2, 3, n // 4) * 2.0 ** -53),
                        [0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 2.0 ** -900, 2.0 ** 900, 1.7976931348623157e308]])
    x = np.ascontiguousarray(np.concatenate([x, -x]))
    y = np.empty_like(x)
    H.pvvh_div_const_array.argtypes = [ctypes.c_long, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
    H.pvvh_div_const_array(x.size, P(x), c, P(y))
    with np.errstate(all="ignore"):
        ref = x / c
    assert_bit_equal(ref, y, f"x / {c}")
    assert np.array_equal(np.signbit(ref), np.signbit(y))  # -0 / c = -0


def test_sort_key_region_is_the_exact_classification(H):
    """The tiled kernels trust key_region(nbc_sort_key(x, s)) to be the reference's region (_model_calls.py:18-42):
    the branch-free key must classify exactly like the chain of comparisons, also on the region boundaries."""
    rng = np.random.default_rng(23)
    n = 2_000_000
    s = np.concatenate([rng.uniform(0, 0.5, n // 2), rng.uniform(0, 12.0, n // 2)])
    x = -np.abs(rng.normal(0, 1, n)) * s * np.where(rng.random(n) < 0.2, 12.0, 1.0)
    # points ON the thresholds: x = -10 s, s/2 = tau, x + s^2/2 = 0.85 s, and their neighbours by one ulp
    tau = H.pvvh_const(3)
    sb = rng.uniform(0.01, 8.0, 200000)
    xb = np.concatenate([-10.0 * sb[:50000], 0.85 * sb[50000:100000] - 0.5 * sb[50000:100000] ** 2, sb[100000:150000] * (tau - 10.0) - 0.5 * sb[100000:150000] ** 2,
                         -rng.uniform(0, 3, 50000)])
    sb[150000:] = 2 * tau
    xb = np.where(rng.random(200000) < 0.5, np.nextafter(xb, 0), np.nextafter(xb, -np.inf))
    x = np.concatenate([x, xb, [0.0, -0.0, 1.0, -np.inf, np.inf, np.nan, -1.0, -1.0, -1.0, 5.0]])
    s = np.concatenate([s, sb, [1.0, 0.0, 1.0, 1.0, 1.0, 1.0, 0.0, -1.0, np.nan, 0.0]])
    x, s = np.ascontiguousarray(x), np.ascontiguousarray(s)
    H.pvvh_sort_key_mismatches.restype = ctypes.c_long
    assert H.pvvh_sort_key_mismatches(x.size, P(x), P(s)) == 0
