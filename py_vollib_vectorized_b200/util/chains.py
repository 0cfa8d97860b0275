"""Seeded synthetic option chains: the workloads of BASELINE.json (SURVEY.md §8(d) cfg 2-5).

numpy only.  Used by bench.py, the parity tests and the oracle validation scripts so that every
leg sees identical inputs.  Columns are float64, flags int8 (+1 call / -1 put).
"""
import numpy as np


def chain_cfg2(n, seed=0):
    """cfg 2: broad random Black-Scholes chain (the generator of the survey's probes)."""
    rng = np.random.default_rng(seed)
    S = rng.uniform(50.0, 150.0, n)
    K = S * np.exp(rng.normal(0.0, 0.25, n))
    t = rng.uniform(1.0 / 365.0, 2.0, n)
    r = rng.uniform(0.0, 0.08, n)
    sigma = rng.uniform(0.05, 0.9, n)
    flag = np.where(rng.random(n) < 0.5, 1, -1).astype(np.int8)
    q = rng.uniform(0.0, 0.04, n)
    return dict(flag=flag, S=S, K=K, t=t, r=r, sigma=sigma, q=q)


def chain_cfg3(n, seed=1, pricer=None):
    """cfg 3: BSM implied-vol chain with deep-OTM, short-expiry, below-intrinsic, above-maximum
    and t<0 strata.  `pricer(flag,S,K,t,r,sigma,q) -> price` (BSM) supplies the clean prices."""
    c = chain_cfg2(n, seed)
    rng = np.random.default_rng(seed + 1000)
    u = rng.random(n)
    deep = u < 0.02
    short = (u >= 0.02) & (u < 0.04)
    below = (u >= 0.04) & (u < 0.05)
    above = (u >= 0.05) & (u < 0.055)
    tneg = (u >= 0.055) & (u < 0.056)
    S, K, t, r, sigma, q, flag = c["S"], c["K"], c["t"], c["r"], c["sigma"], c["q"], c["flag"]
    t[short] = rng.uniform(1e-4, 1.0 / 365.0, int(short.sum()))
    # deep OTM: |ln(F/K)| / (sigma sqrt t) in [6, 12], on the OTM side of each flag
    m = rng.uniform(6.0, 12.0, int(deep.sum())) * sigma[deep] * np.sqrt(t[deep])
    F = S[deep] * np.exp((r[deep] - q[deep]) * t[deep])
    K[deep] = F * np.exp(np.where(flag[deep] > 0, m, -m))
    if pricer is None:
        raise ValueError("chain_cfg3 needs a BSM pricer callable")
    price = np.array(pricer(flag, S, K, t, r, sigma, q), dtype=np.float64)
    F = S * np.exp((r - q) * t)
    D = np.exp(-r * t)
    intrinsic = np.maximum(np.where(flag > 0, F - K, K - F), 0.0) * D
    itm = intrinsic > 0
    bi = below & itm
    price[bi] = intrinsic[bi] * (1.0 - rng.uniform(1e-6, 1e-2, int(bi.sum())))
    mx = np.where(flag > 0, F, K) * D
    price[above] = mx[above] * (1.0 + rng.uniform(0.0, 0.5, int(above.sum())))
    t[tneg] = -t[tneg]
    c["price"] = price
    c["strata"] = dict(deep=deep, short=short, below=bi, above=above, tneg=tneg)
    return c


def chain_wide(n, seed=7):
    """Stress chain: log-uniform moneyness / vol / expiry far beyond market ranges, so that all four
    normalised-Black regions and all four solver branches are populated."""
    rng = np.random.default_rng(seed)
    S = np.exp(rng.uniform(np.log(1e-2), np.log(1e4), n))
    sigma = np.exp(rng.uniform(np.log(1e-3), np.log(6.0), n))
    t = np.exp(rng.uniform(np.log(1e-5), np.log(30.0), n))
    # moneyness in units of total vol, mixed with absolute log-moneyness
    z = rng.uniform(-14.0, 14.0, n) * sigma * np.sqrt(t)
    z = np.where(rng.random(n) < 0.3, rng.uniform(-3.0, 3.0, n), z)
    K = S * np.exp(z)
    r = rng.uniform(-0.05, 0.3, n)
    q = rng.uniform(-0.02, 0.2, n)
    flag = np.where(rng.random(n) < 0.5, 1, -1).astype(np.int8)
    return dict(flag=flag, S=S, K=K, t=t, r=r, sigma=sigma, q=q)


def chain_edge():
    """Hand-made rows: NaN inputs, t<0, t==0 at/in/out of the money, sigma<=0, S==0, tiny and huge
    prices.  No row divides by zero in the reference (K==0 and, for IV, t==0 are tested apart)."""
    nan = float("nan")
    rows = [
        # flag, S, K, t, r, sigma, q, price
        (1, 100.0, 100.0, 0.0, 0.05, 0.2, 0.0, 5.0),
        (-1, 100.0, 100.0, 0.0, 0.05, 0.2, 0.0, 5.0),
        (1, 110.0, 100.0, 0.0, 0.05, 0.2, 0.01, 12.0),
        (-1, 110.0, 100.0, 0.0, 0.05, 0.2, 0.01, 1.0),
        (1, 90.0, 100.0, 0.0, 0.05, 0.2, 0.01, 1.0),
        (-1, 90.0, 100.0, 0.0, 0.05, 0.2, 0.01, 12.0),
        (1, 100.0, 100.0, -0.5, 0.05, 0.2, 0.0, 5.0),
        (-1, 100.0, 90.0, -0.1, 0.05, 0.2, 0.0, 5.0),
        (1, 110.0, 100.0, 0.5, 0.05, 0.0, 0.0, 12.5),
        (-1, 110.0, 100.0, 0.5, 0.05, 0.0, 0.0, 0.0),
        (1, 110.0, 100.0, 0.5, 0.05, -0.3, 0.0, 12.4690088),
        (1, 0.0, 100.0, 0.5, 0.05, 0.2, 0.0, 1.0),
        (1, nan, 100.0, 0.5, 0.05, 0.2, 0.0, 5.0),
        (1, 100.0, nan, 0.5, 0.05, 0.2, 0.0, 5.0),
        (1, 100.0, 100.0, nan, 0.05, 0.2, 0.0, 5.0),
        (1, 100.0, 100.0, 0.5, nan, 0.2, 0.0, 5.0),
        (1, 100.0, 100.0, 0.5, 0.05, nan, 0.0, nan),
        (-1, 100.0, 100.0, 0.5, 0.05, 0.2, nan, 5.0),
        (1, 50.0, 100.0, 0.1, 0.0, 0.2, 0.0, 1e-300),
        (1, 50.0, 100.0, 0.1, 0.0, 0.2, 0.0, 0.0),
        (1, 100.0, 100.0, 0.5, 0.01, 0.2, 0.0, -1.0),
        (1, 100.0, 100.0, 0.5, 0.01, 0.2, 0.0, 100.0),
        (1, 100.0, 100.0, 0.5, 0.01, 0.2, 0.0, 250.0),
        (-1, 100.0, 100.0, 0.5, 0.01, 0.2, 0.0, 100.0),
        (-1, 100.0, 100.0, 0.5, 0.01, 0.2, 0.0, 99.9999),
        (1, 100.0, 100.0, 0.5, 0.01, 0.2, 0.0, 99.9999),
        (1, 100.0, 100.0, 1.0 / 365.0, 0.01, 0.2, 0.0, 0.4),
        (-1, 100.0, 100.0, 0.5 / 365.0, 0.01, 0.2, 0.0, 0.3),
        (1, 100.0, 100.0, 1e-5, 0.01, 0.2, 0.0, 0.02),
        (1, 1e-5, 1e5, 1.0, 0.0, 0.5, 0.0, 1e-200),
        (-1, 1e5, 1e-5, 1.0, 0.0, 0.5, 0.0, 1e-200),
        (1, 1e5, 1e-5, 1.0, 0.0, 0.5, 0.0, 1e5 - 1e-5 + 1e-9),
        (1, 100.0, 100.0, 0.5, 0.01, 8.0, 0.0, 99.0),
        (1, 100.0, 100.0, 0.5, 0.01, 20.0, 0.0, 99.99999),
        (1, 100.0, 100.0, 40.0, 0.01, 3.0, 0.0, 66.9),
        (1, 100.0, 100.0, 0.5, 30.0, 0.2, 0.0, 5.0),
        (1, 1e300, 1e-300, 1.0, 0.0, 0.5, 0.0, 1e299),
        (-1, 1e-300, 1e300, 1.0, 0.0, 0.5, 0.0, 1e299),
        (1, 100.0, 100.0 * (1 + 2 ** -40), 0.5, 0.0, 0.2, 0.0, 5.6),
        (1, 100.0, 100.0, 0.5, 0.0, 1e-12, 0.0, 1e-11),
    ]
    a = np.array(rows, dtype=np.float64)
    return dict(flag=a[:, 0].astype(np.int8), S=a[:, 1].copy(), K=a[:, 2].copy(), t=a[:, 3].copy(), r=a[:, 4].copy(),
                sigma=a[:, 5].copy(), q=a[:, 6].copy(), price=a[:, 7].copy())


def chain_cfg5(n_strikes=2000, n_expiries=500, n_underlyings=100):
    """cfg 5: Black-76 futures-option surface grid (F-major, K fastest), alternating c/p flags."""
    K = np.linspace(50.0, 150.0, n_strikes)
    t = np.linspace(1.0 / 365.0, 2.0, n_expiries)
    F = np.linspace(80.0, 120.0, n_underlyings)
    Fg, tg, Kg = np.meshgrid(F, t, K, indexing="ij")
    Fg, tg, Kg = Fg.ravel(), tg.ravel(), Kg.ravel()
    n = Fg.size
    flag = np.where(np.arange(n) % 2 == 0, 1, -1).astype(np.int8)
    sigma = 0.2 + 0.1 * np.abs(np.log(Fg / Kg))
    return dict(flag=flag, S=Fg, K=Kg, t=tg, r=np.full(n, 0.03), sigma=sigma, q=np.zeros(n))
