"""CPU tier: the host-side mirror of the reference interface (flags, broadcasting, error / warning
semantics, monkey-patches).  No kernel is launched here."""
import warnings

import numpy as np
import pandas as pd
import pytest

import py_vollib_vectorized_b200 as pvv
from py_vollib_vectorized_b200.util import data_format as df_


def test_flag_encoding_matches_reference_dict():
    enc = df_._preprocess_flags
    assert enc(['c', 'p', 'c']).tolist() == [1, -1, 1]
    assert enc(('c', 'p')).tolist() == [1, -1]
    assert enc('c').tolist() == [1]
    assert enc('cp').tolist() == [1, -1]          # the reference iterates a bare string
    assert enc([1, -1]).tolist() == [1, -1]
    assert enc(np.array([1.0, -1.0])).tolist() == [1, -1]
    assert enc(['1', '-1']).tolist() == [1, -1]
    assert enc(pd.Series(['c', 'p'], dtype=object)).tolist() == [1, -1]
    assert enc(pd.Series(['c', 1, '-1', -1], dtype=object)).tolist() == [1, 1, -1, -1]
    assert enc(np.array(['p', 'c'])).tolist() == [-1, 1]
    assert enc(np.array([1, -1], dtype=np.int8)).dtype == np.int8
    with pytest.raises(KeyError):
        enc(['c', 'x'])
    with pytest.raises(KeyError):
        enc([1, 0])
    with pytest.raises(KeyError):
        enc(pd.Series(['c', 'call'], dtype=object))
    with pytest.raises(TypeError):
        enc(1)                                     # a bare int is not iterable
    big = np.where(np.arange(200000) % 3 == 0, 'c', 'p').astype(object)
    out = enc(pd.Series(big))
    assert out.dtype == np.int8 and out[:4].tolist() == [1, -1, -1, 1]


def test_flag_encoding_fast_paths_agree_with_the_dict():
    """Every array layout has its own encoder (C passes over 4- / 8-byte cells, the Arrow data buffer, PyObject*
    identity): all of them must give what the reference's per-element dict lookup gives, threaded or not."""
    enc = df_._preprocess_flags
    rng = np.random.default_rng(0)
    for n in (1, 7, 70_001, 2_300_000):  # the last one crosses the threading threshold of the C helpers
        base = rng.choice(np.array([1, -1], dtype=np.int8), n)
        chars = np.where(base > 0, 'c', 'p')
        cases = {
            "<U1": chars,
            "<U1 strided": np.repeat(chars, 2)[::2],
            "<U2": np.where(base > 0, '1', '-1'),
            "object": chars.astype(object),
            "int64": base.astype(np.int64), "int32": base.astype(np.int32), "int16": base.astype(np.int16),
            "float64": base.astype(np.float64), "float32": base.astype(np.float32),
            "arrow": pd.Series(chars, dtype="string"),
            "arrow mixed width": pd.Series(np.where(base > 0, 'c', '-1'), dtype="string"),
            "list": chars.tolist() if n < 100_000 else None,
        }
        for name, flags in cases.items():
            if flags is None:
                continue
            out = enc(flags)
            assert out.dtype == np.int8 and np.array_equal(out, base), (name, n)
    assert enc(np.array([1, 1], dtype=np.uint32)).tolist() == [1, 1]
    for bad in (np.array(['c', 'x', 'p']), np.array([1, 2, -1]), np.array([1.0, 0.5]), np.array([1, 3], dtype=np.uint32),
                np.array([1, -1, 0], dtype=np.int32), pd.Series(['c', 'q'], dtype="string"), np.array([np.nan, 1.0])):
        with pytest.raises(KeyError):
            enc(bad)


def test_broadcast_keeps_scalars_as_stride_zero():
    c = df_.maybe_format_data_and_broadcast(95, [100, 90], .2, pd.Series([.1, .2]), np.array([1, -1], dtype=np.int8), dtype=np.float64)
    assert c.n == 2
    assert [s for _, s in c.cols] == [0, 1, 0, 1, 1]
    assert c.cols[0][0].tolist() == [95.0] and c.dense(0).tolist() == [95.0, 95.0]
    assert c.cols[4][0].dtype == np.int8
    one = df_.maybe_format_data_and_broadcast(1.0, 2, dtype=np.float64)
    assert one.n == 1 and [s for _, s in one.cols] == [1, 1]
    with pytest.raises(ValueError):
        df_.maybe_format_data_and_broadcast([1, 2, 3], [1, 2], dtype=np.float64)   # numpy's shape mismatch
    with pytest.raises(ValueError, match="unsupported"):
        df_.maybe_format_data_and_broadcast(np.int64(3), dtype=np.float64)         # the reference rejects numpy scalars too
    with pytest.raises(AssertionError):
        df_.maybe_format_data_and_broadcast(pd.DataFrame({"a": [1], "b": [2]}), dtype=np.float64)
    two_d = df_.maybe_format_data_and_broadcast(np.ones((2, 3)), dtype=np.float64)
    assert two_d.n == 6                                                             # 2-D is silently flattened
    f32 = df_.maybe_format_data_and_broadcast([0.1], dtype=np.float32)
    assert f32.cols[0][0].dtype == np.float32


def test_iv_status_reporting_order_and_texts():
    st = np.array([0, 1, 2, 0, 1], dtype=np.uint8)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        df_.report_iv_status(st, "warn")
    assert [str(x.message) for x in w] == ["Found Below Intrinsic contracts at index [1, 4]",
                                           "Found Above Maximum Price contracts at index [2]"]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        df_.report_iv_status(st, "ignore")
    assert not w
    with pytest.raises(df_.PriceIsBelowIntrinsic):
        df_.report_iv_status(st, "raise")                     # below-intrinsic wins when both are present
    with pytest.raises(df_.PriceIsAboveMaximum):
        df_.report_iv_status(np.array([0, 2], dtype=np.uint8), "raise")
    with pytest.raises(ValueError, match="Interest Free Rate"):
        df_.report_iv_status(np.array([8], dtype=np.uint8), "raise")
    with pytest.raises(ZeroDivisionError):
        df_.report_iv_status(np.array([0, 4], dtype=np.uint8), "ignore")
    with pytest.raises(df_.PriceIsBelowIntrinsic):           # K == 0: the intrinsic check fires before numba would
        df_.report_iv_status(np.array([5], dtype=np.uint8), "raise")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        with pytest.raises(ZeroDivisionError):
            df_.report_iv_status(np.array([8 | 4], dtype=np.uint8), "warn")
    assert "Unexpected value encountered in Interest Free Rate (r)" in str(w[0].message)


def test_monkeypatches_like_the_reference():
    import py_vollib.black
    import py_vollib.black.greeks.numerical
    import py_vollib.black.implied_volatility
    import py_vollib.black_scholes
    import py_vollib.black_scholes.greeks.numerical
    import py_vollib.black_scholes.implied_volatility
    import py_vollib.black_scholes_merton
    import py_vollib.black_scholes_merton.greeks.numerical
    import py_vollib.black_scholes_merton.implied_volatility
    pkg = pvv.__name__
    for m in (py_vollib.black.implied_volatility, py_vollib.black_scholes.implied_volatility, py_vollib.black_scholes_merton.implied_volatility):
        assert m.implied_volatility.__module__ == pkg + ".implied_volatility"
    assert py_vollib.black.black.__module__ == pkg + ".models"
    assert py_vollib.black_scholes.black_scholes.__module__ == pkg + ".models"
    assert py_vollib.black_scholes_merton.black_scholes_merton.__module__ == pkg + ".models"
    for m in (py_vollib.black.greeks.numerical, py_vollib.black_scholes.greeks.numerical, py_vollib.black_scholes_merton.greeks.numerical):
        for g in ("delta", "gamma", "rho", "theta", "vega"):
            assert getattr(m, g).__module__ == pkg + ".greeks"
    assert repr(py_vollib.black.black) == "Vectorized implementation of <black(flag, F, K, t, r, sigma)>"
    assert repr(py_vollib.black_scholes.greeks.numerical.delta) == \
        "Vectorized implementation of <delta(flag, S, K, t, r, sigma, q, model=black_scholes)>"
    assert py_vollib.black_scholes_merton.implied_volatility.implied_volatility.keywords == {"model": "black_scholes_merton"}


def test_argument_validation_happens_before_any_kernel():
    with pytest.raises(AssertionError):
        pvv.vectorized_implied_volatility(1, 100, 100, .5, .01, ['c'], on_error="nope")
    with pytest.raises(ValueError, match="Model must be one of"):
        pvv.vectorized_implied_volatility(1, 100, 100, .5, .01, ['c'], model="bachelier")
    with pytest.raises(ValueError, match="Must pass a `q`"):
        pvv.vectorized_implied_volatility(1, 100, 100, .5, .01, ['c'], model="black_scholes_merton")
    with pytest.raises(ValueError, match="Must pass a `q`"):
        pvv.get_all_greeks(['c'], 100, 100, .5, .01, .2, model="black_scholes_merton")
    with pytest.raises(ValueError, match="Model must be one of"):
        pvv.get_all_greeks(['c'], 100, 100, .5, .01, .2, model="x")
    with pytest.raises(TypeError):
        pvv.vectorized_delta(['c'], 100, 100, .5, .01, .2, model="black")      # broken in the reference too (Q3)
    df = pd.DataFrame({"Flag": ['c'], "S": [100.], "K": [100.], "T": [.5], "R": [.01], "IV": [.2]})
    kw = dict(flag_col="Flag", underlying_price_col="S", strike_col="K", annualized_tte_col="T", riskfree_rate_col="R")
    with pytest.raises(AssertionError):
        pvv.price_dataframe(df, underlying_price_col="S")
    with pytest.raises(ValueError, match="not found"):
        pvv.price_dataframe(df, **dict(kw, strike_col="nope"), sigma_col="IV")
    with pytest.raises(ValueError, match="either `sigma_col`, `price_col`"):
        pvv.price_dataframe(df, **kw)
    with pytest.raises(ValueError, match="dividend_col"):
        pvv.price_dataframe(df, **kw, sigma_col="IV", model="black_scholes_merton")
    with pytest.raises(ValueError, match="must be a string"):
        pvv.price_dataframe(df, **dict(kw, strike_col=3), sigma_col="IV")


def test_scalar_q_with_vector_inputs_like_the_reference():
    """greeks.py:54-56, 117, 178, 301 re-broadcast q for delta / theta / vega / gamma; only rho (greeks.py:240-242)
    checks the length of q against r and raises.  Without a GPU the four working greeks get as far as the kernel call."""
    from py_vollib_vectorized_b200 import _engine
    args = (['c', 'p'], 95, [100, 90], .2, .2, .2)
    with pytest.raises(ValueError, match="same number of elements"):
        pvv.vectorized_rho(*args, q=0.0, model="black_scholes_merton")
    for fn in (pvv.vectorized_delta, pvv.vectorized_theta, pvv.vectorized_vega, pvv.vectorized_gamma):
        try:
            out = fn(*args, q=0.0, model="black_scholes_merton", return_as="numpy")
            assert out.shape == (2,)
        except _engine.PvvError:
            pass  # no CUDA device in the CPU tier: the arguments were accepted
    try:  # the opt-in fix makes rho broadcast q too
        assert pvv.vectorized_rho(*args, q=0.0, model="black_scholes_merton", return_as="numpy", fix_reference_bugs=True).shape == (2,)
    except _engine.PvvError:
        pass
    with pytest.raises(TypeError):
        pvv.vectorized_delta(*args, model="black")
    try:
        assert pvv.vectorized_delta(*args, model="black", return_as="numpy", fix_reference_bugs=True).shape == (2,)
    except _engine.PvvError:
        pass


def test_iv_tile_plan_invariants():
    """K3's launch plan (pvv_iv_tile_plan, pure host logic): a column is cut into waves x slots equal tiles.  Whatever n:
    the tiles cover the column, fit the tile capacity, start on 32-contract boundaries, use the minimal number of waves,
    and no tile is more than one 32-contract step larger than an even split."""
    import ctypes
    from py_vollib_vectorized_b200 import _engine
    lib = _engine.lib()
    lib.pvv_iv_tile_plan.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int]
    lib.pvv_iv_tile_plan.restype = ctypes.c_int
    rng = np.random.default_rng(5)
    sizes = [1, 31, 32, 33, 1000, 65535, 65536, 131072, 454655, 454656, 454657, 1_000_000, 10_000_000, 100_000_000, 2**40]
    sizes += [int(x) for x in 10 ** rng.uniform(0, 9, 300)]
    for capacity, slots in ((3072, 148), (1536, 296), (3072, 132), (1536, 7)):
        for n in sizes:
            cpt = lib.pvv_iv_tile_plan(n, capacity, slots)
            assert 32 <= cpt <= capacity and cpt % 32 == 0, (n, capacity, slots, cpt)
            grid = -(-n // cpt)
            waves_min = -(-n // (slots * capacity))
            assert grid * cpt >= n
            assert -(-grid // slots) == waves_min, (n, capacity, slots, cpt, grid)      # never an extra wave
            even = -(-n // (waves_min * slots))
            assert cpt < even + 32 or cpt == capacity, (n, capacity, slots, cpt, even)  # as even as the alignment allows
    # full tiles once the column is long: the plan converges to the capacity
    assert lib.pvv_iv_tile_plan(148 * 3072 * 50, 3072, 148) == 3072
    for bad in ((0, 3072, 148), (10, 100, 148), (10, 3072, 0)):
        assert lib.pvv_iv_tile_plan(*bad) < 0
