"""GPU tier at BASELINE.json's sizes: cfg2 (1 M Black-Scholes chain, price + greeks) and cfg3 (10 M
BSM implied volatility with strata) against the multi-threaded oracle bit for bit, plus the
size-independent properties the domain offers (price -> IV -> price round trip, put-call parity,
launch-shape invariance: one launch == chunked pipeline == device-pointer path)."""
import numpy as np
import pytest

from tests.util import assert_bit_equal

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
    return (np.ascontiguousarray(a), 1)


def test_cfg2_one_million_price_and_greeks(eng, oracle):
    from py_vollib_vectorized_b200.util.chains import chain_cfg2
    n = 1_000_000
    c = chain_cfg2(n, seed=0)
    th = oracle.max_threads
    for model in ("black_scholes", "black_scholes_merton"):
        q = col(c["q"]) if model == "black_scholes_merton" else None
        out = eng.greeks(model, n, col(c["flag"]), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), q, col(c["sigma"]), with_price=True)
        ref, zde = oracle.greeks(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=th)
        assert zde == -1
        for k in ("price",) + GREEKS:
            assert_bit_equal(ref[k], out[k], f"{model} {k}")
    p = eng.price("black", n, col(c["flag"]), col(c["S"]), col(c["K"]), col(c["t"]), None, None, col(c["sigma"]))
    assert_bit_equal(oracle.price("black", c["flag"], c["S"], c["K"], c["t"], 0.0, c["sigma"], threads=th)[0], p, "black")
    # put-call parity of the undiscounted Black price: C - P = F - K (to rounding of the larger leg)
    call = eng.price("black", n, (np.ones(1, dtype=np.int8), 0), col(c["S"]), col(c["K"]), col(c["t"]), None, None, col(c["sigma"]))
    put = eng.price("black", n, (-np.ones(1, dtype=np.int8), 0), col(c["S"]), col(c["K"]), col(c["t"]), None, None, col(c["sigma"]))
    assert np.abs((call - put) - (c["S"] - c["K"])).max() <= 1e-12 * np.maximum(c["S"], c["K"]).max()


def test_cfg3_ten_million_iv_with_strata(eng, oracle):
    import torch
    from py_vollib_vectorized_b200.util.chains import chain_cfg3
    n = 10_000_000
    th = oracle.max_threads

    def gpu_pricer(f, S, K, t, r, sg, q):
        return eng.price("black_scholes_merton", f.size, col(f), col(S), col(K), col(t), col(r), col(q), col(sg))

    c = chain_cfg3(n, seed=1, pricer=gpu_pricer)
    args = (col(c["flag"]), col(c["price"]), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), col(c["q"]))
    for mask in (True, False):
        iv, status, zde = eng.iv("black_scholes_merton", n, *args, mask_errors=mask)
        ref, st_ref, zde_ref = oracle.iv("black_scholes_merton", c["price"], c["S"], c["K"], c["t"], c["r"], c["flag"], c["q"],
                                         mask_errors=mask, threads=th)
        assert zde == zde_ref == -1
        assert (status == st_ref).all()
        assert_bit_equal(ref, iv, f"10M IV mask={mask}")
    st = c["strata"]
    iv_w, status, _ = eng.iv("black_scholes_merton", n, *args, mask_errors=True)
    assert (status[st["below"]] & 1).all() and np.isnan(iv_w[st["below"]]).all()
    assert (status[st["above"]] & 2).all() and np.isnan(iv_w[st["above"]]).all()
    assert np.isnan(iv_w[st["tneg"]]).all()  # t < 0: sqrt(t) is NaN (no ZeroDivisionError, no warning of its own)
    # round trip: repricing at the solved volatility gives the price back
    ok = np.isfinite(iv_w) & (iv_w > 0) & ~st["deep"]
    idx = np.flatnonzero(ok)[::7]
    rp = eng.price("black_scholes_merton", idx.size, col(c["flag"][idx]), col(c["S"][idx]), col(c["K"][idx]), col(c["t"][idx]),
                   col(c["r"][idx]), col(c["q"][idx]), col(iv_w[idx]))
    rel = np.abs(rp - c["price"][idx]) / c["price"][idx]
    assert np.quantile(rel, 0.999) < 1e-9
    # launch-shape invariance: device-pointer path (one launch) == host pipeline (chunks)
    d = {k: torch.from_numpy(np.ascontiguousarray(c[k])).cuda() for k in ("flag", "price", "S", "K", "t", "r", "q")}
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    stt = torch.empty(n, dtype=torch.uint8, device="cuda")
    eng.iv_dev("black_scholes_merton", n, d["flag"], d["price"], d["S"], d["K"], d["t"], d["r"], d["q"], out, stt)
    assert_bit_equal(iv_w, out.cpu().numpy(), "dev vs host pipeline")
    assert (stt.cpu().numpy() == status).all()


def test_iv_big_and_small_tile_shapes_agree_on_ragged_sizes(eng, oracle):
    """K3 has two launch shapes (4096-contract tiles from 2 424 832 contracts on, 2048-contract tiles below):
    a ragged column just above the switch through the device path (big tiles, partial last tile) must equal the
    host pipeline (chunks -> small tiles) and the oracle bit for bit, for all three models."""
    import torch
    from py_vollib_vectorized_b200.util.chains import chain_cfg2
    n = 4 * 148 * 4096 + 1234
    c = chain_cfg2(n, seed=5)
    th = oracle.max_threads
    for model in ("black", "black_scholes", "black_scholes_merton"):
        bsm = model == "black_scholes_merton"
        q = col(c["q"]) if bsm else None
        r = None if model == "black" else col(c["r"])
        price = eng.price(model, n, col(c["flag"]), col(c["S"]), col(c["K"]), col(c["t"]), r, q, col(c["sigma"]))
        price[::97] *= 0.5  # some rows below intrinsic / far from the clean price
        iv_h, st_h, _ = eng.iv(model, n, col(c["flag"]), col(price), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), q)
        d = {k: torch.from_numpy(np.ascontiguousarray(c[k])).cuda() for k in ("flag", "S", "K", "t", "r", "q")}
        out = torch.empty(n, dtype=torch.float64, device="cuda")
        stt = torch.empty(n, dtype=torch.uint8, device="cuda")
        eng.iv_dev(model, n, d["flag"], torch.from_numpy(price).cuda(), d["S"], d["K"], d["t"], d["r"], d["q"] if bsm else None, out, stt)
        assert_bit_equal(iv_h, out.cpu().numpy(), f"{model}: big tiles vs small tiles")
        assert (stt.cpu().numpy() == st_h).all()
        ref, st_ref, _ = oracle.iv(model, price, c["S"], c["K"], c["t"], c["r"], c["flag"], c["q"] if bsm else None, threads=th)
        assert_bit_equal(ref, iv_h, f"{model}: vs oracle")
        assert (st_ref == st_h).all()


def test_cfg4_fused_iv_greeks_two_million(eng, oracle):
    from py_vollib_vectorized_b200.util.chains import chain_cfg2
    n = 2_000_000
    c = chain_cfg2(n, seed=2)
    th = oracle.max_threads
    price = eng.price("black_scholes", n, col(c["flag"]), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), None, col(c["sigma"]))
    outs, status, zde = eng.iv_greeks("black_scholes", n, col(c["flag"]), col(price), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), None)
    ref, st_ref, _ = oracle.iv_greeks("black_scholes", price, c["S"], c["K"], c["t"], c["r"], c["flag"], threads=th)
    assert zde == -1 and (status == st_ref).all()
    for k in ("iv",) + GREEKS:
        assert_bit_equal(ref[k], outs[k], f"fused {k}")
    good = np.isfinite(outs["iv"]) & (outs["iv"] > 0)
    assert np.quantile(np.abs(outs["iv"][good] - c["sigma"][good]), 0.99) < 1e-6


def test_cfg5_surface_price_dataframe(eng, oracle):
    """Black-76 surface grid through the public price_dataframe (string flag column, pandas in / out)."""
    import pandas as pd
    import py_vollib_vectorized_b200 as pvv
    from py_vollib_vectorized_b200.util.chains import chain_cfg5
    c = chain_cfg5(2000, 10, 100)  # 2 M rows, F-major, K fastest
    n = c["S"].size
    df = pd.DataFrame({"Flag": np.where(c["flag"] > 0, "c", "p"), "F": c["S"], "K": c["K"], "T": c["t"], "R": c["r"], "IV": c["sigma"]})
    df.index = np.arange(n)[::-1]  # a non-default index must not disturb positional assignment
    res = pvv.price_dataframe(df, flag_col="Flag", underlying_price_col="F", strike_col="K", annualized_tte_col="T",
                              riskfree_rate_col="R", sigma_col="IV", model="black", inplace=False)
    assert list(res.columns) == ["Price", "delta", "gamma", "theta", "rho", "vega"] and (res.index == df.index).all()
    th = oracle.max_threads
    p, _ = oracle.price("black", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], threads=th)
    g, _ = oracle.greeks("black", c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], threads=th)
    assert_bit_equal(p, res["Price"].values, "surface price")
    for k in GREEKS:
        assert_bit_equal(g[k], res[k].values, f"surface {k}")
    # and back: implied volatility + greeks from the prices (K4 through the public API)
    df["Px"] = res["Price"].values
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        back = pvv.price_dataframe(df, flag_col="Flag", underlying_price_col="F", strike_col="K", annualized_tte_col="T",
                                   riskfree_rate_col="R", price_col="Px", model="black", inplace=False)
    ref, st, _ = oracle.iv_greeks("black", res["Price"].values, c["S"], c["K"], c["t"], c["r"], c["flag"], threads=th)
    assert_bit_equal(ref["iv"], back["IV"].values, "surface IV")
    for k in GREEKS:
        assert_bit_equal(ref[k], back[k].values, f"surface IV->{k}")


def test_single_process_multi_gpu_sharding(eng, oracle, monkeypatch):
    """PVV_DEVICES shards one host call over several devices (drop-in mode, no collective).  With one
    visible GPU the same device is listed twice: the sharding, offsets and zero-division index logic
    are exercised all the same."""
    import torch
    from py_vollib_vectorized_b200.util.chains import chain_cfg2
    nd = torch.cuda.device_count()
    monkeypatch.setenv("PVV_DEVICES", ",".join(str(i % nd) for i in range(max(2, nd))))
    n = 1_000_003
    c = chain_cfg2(n, seed=9)
    out = eng.greeks("black_scholes", n, col(c["flag"]), col(c["S"]), col(c["K"]), col(c["t"]), (np.array([0.03]), 0), None, col(c["sigma"]),
                     with_price=True)
    ref, _ = oracle.greeks("black_scholes", c["flag"], c["S"], c["K"], c["t"], 0.03, c["sigma"], threads=oracle.max_threads)
    for k in ref:
        assert_bit_equal(ref[k], out[k], f"sharded {k}")
    K = c["K"].copy()
    K[[700_001, 900_000]] = 0.0
    price = ref["price"]
    iv, status, zde = eng.iv("black_scholes", n, col(c["flag"]), col(price), col(c["S"]), col(K), col(c["t"]), (np.array([0.03]), 0), None)
    assert zde == 700_001 and np.flatnonzero(status & 4).tolist() == [700_001, 900_000]


def test_repeated_ragged_launches_are_deterministic(eng, oracle):
    """compute-sanitizer is not available on the pool, so shared-memory hazards in the tiled kernels (the
    sorts reuse their counters and arrays with a minimum of barriers) are hunted the blunt way: many
    launches of ragged sizes — partial tiles, single contracts, tile-size multiples — every output
    compared bit for bit with the oracle."""
    from py_vollib_vectorized_b200.util.chains import chain_cfg2, chain_wide
    rng = np.random.default_rng(2024)
    th = oracle.max_threads
    sizes = [1, 2, 31, 127, 128, 129, 511, 512, 513, 1023, 1024, 1025, 4097] + [int(x) for x in rng.integers(1000, 300000, 12)]
    for rep, n in enumerate(sizes):
        c = (chain_cfg2 if rep % 2 == 0 else chain_wide)(n, seed=100 + rep)
        for model in ("black_scholes", "black_scholes_merton"):
            q = col(c["q"]) if model == "black_scholes_merton" else None
            out = eng.greeks(model, n, col(c["flag"]), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), q, col(c["sigma"]), with_price=True)
            ref, _ = oracle.greeks(model, c["flag"], c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], threads=th)
            for k in ref:
                assert_bit_equal(ref[k], out[k], f"n={n} {model} {k}")
            p = eng.price(model, n, col(c["flag"]), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), q, col(c["sigma"]))
            assert_bit_equal(ref["price"], p, f"n={n} {model} price kernel")
            outs, status, zde = eng.iv_greeks(model, n, col(c["flag"]), col(p), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), q)
            ref2, st2, _ = oracle.iv_greeks(model, p, c["S"], c["K"], c["t"], c["r"], c["flag"], c["q"], threads=th)
            assert (status == st2).all() and zde == -1, n
            for k in ref2:
                assert_bit_equal(ref2[k], outs[k], f"n={n} {model} fused {k}")
            iv, status, _ = eng.iv(model, n, col(c["flag"]), col(p), col(c["S"]), col(c["K"]), col(c["t"]), col(c["r"]), q, mask_errors=False)
            ref3, st3, _ = oracle.iv(model, p, c["S"], c["K"], c["t"], c["r"], c["flag"], c["q"], mask_errors=False, threads=th)
            assert_bit_equal(ref3, iv, f"n={n} {model} iv")
