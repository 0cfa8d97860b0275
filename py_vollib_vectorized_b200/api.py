"""`get_all_greeks` and `price_dataframe` on the B200 engine.

Public surface of /root/reference/py_vollib_vectorized/api.py (:19-181, :184-256).  The reference
runs five numba passes for the greeks (11 price evaluations per contract, 8 distinct) and re-runs
flag pre-processing and broadcasting three to four times per `price_dataframe` call; here one
flag encode + one broadcast feed kernel K2 (all five greeks from the 8 distinct stencil points in
one pass) or, when only prices are given, the K4 route (one call: implied volatility, then the
greeks at the solved volatility; sigma never leaves the device).

Deliberate differences (values are identical):
  * `get_all_greeks(return_as != "dataframe")` returns a dict of numpy arrays, where the reference
    returns Python lists (api.py:256; its README promises arrays); `return_as="json"` works instead
    of raising AttributeError (api.py:254-255);
  * `price_dataframe` assigns results by POSITION.  The reference assigns RangeIndex-ed frames into
    `index=df.index` (api.py:125,140,169,178), which silently yields NaN for any non-default index.
"""
import json
import warnings

import numpy as np
import pandas as pd

from . import _engine
from .implied_volatility import _MODEL_ERROR
from .util.data_format import (_preprocess_flags, _validate_df_col, maybe_format_data_and_broadcast,
                               report_iv_status)

_Q_ERROR = "Must pass a `q` to black scholes merton model (annualized continuous dividend yield)."
_GREEKS = ("delta", "gamma", "theta", "rho", "vega")


def get_all_greeks(flag, S, K, t, r, sigma, q=None, *, model="black_scholes", return_as="dataframe", dtype=np.float64,
                   method="numerical", fix_reference_bugs=False):
    """
    All five greeks of each contract, as specified by `model`
    ("black" uses the Black-Scholes stencil, exactly as the reference does).

    `method="numerical"` (default) reproduces the reference's finite differences bit for bit.
    `method="analytical"` (an extension; models black_scholes / black_scholes_merton) returns the
    closed-form greeks in py_vollib's `greeks.analytical` conventions — theta per calendar day, vega
    and rho per point — from a single HBM-bound pass.

    >>> get_all_greeks(['c', 'p'], 95, [100, 90], .2, .2, .2, model='black_scholes', return_as='dict')
    {'delta': array([ 0.46750566, -0.1364465 ]),
     'gamma': array([0.0467948, 0.0257394]),
     'theta': array([-0.04589963, -0.00533543]),
     'rho': array([ 0.0830349 , -0.02715114]),
     'vega': array([0.16892575, 0.0928379 ])}
    """
    if method not in ("numerical", "analytical"):
        raise ValueError("`method` must be 'numerical' or 'analytical'")
    compute = _engine.greeks
    if method == "analytical":
        if model == "black" and not fix_reference_bugs:
            raise ValueError("Analytical greeks are available for `black_scholes` and `black_scholes_merton`")
        compute = _engine.analytic_greeks
    flag = _preprocess_flags(flag, dtype=dtype)
    c = maybe_format_data_and_broadcast(S, K, t, r, sigma, flag, dtype=dtype)
    if model == "black" and fix_reference_bugs:
        S, K, t, r, sigma, flag = c
        greeks = compute("black_scholes_merton", c.n, flag, S, K, t, r, r, sigma)  # zero carry: q = r
    elif model in ("black", "black_scholes"):
        S, K, t, r, sigma, flag = c
        greeks = compute(model, c.n, flag, S, K, t, r, None, sigma)
    elif model == "black_scholes_merton":
        if q is None:
            raise ValueError(_Q_ERROR)
        c = maybe_format_data_and_broadcast(*c.cols, q, dtype=dtype)
        S, K, t, r, sigma, flag, q = c
        greeks = compute(model, c.n, flag, S, K, t, r, q, sigma)
    else:
        raise ValueError(_MODEL_ERROR)
    greeks = {k: greeks[k] for k in _GREEKS}
    if return_as == "dataframe":
        return pd.DataFrame.from_dict(greeks)
    elif return_as == "json":
        return json.dumps({k: v.tolist() for k, v in greeks.items()})
    return greeks


def price_dataframe(df, *, flag_col=None, underlying_price_col=None, strike_col=None, annualized_tte_col=None,
                    riskfree_rate_col=None,
                    sigma_col=None, price_col=None, dividend_col=None,
                    model="black_scholes", inplace=False, dtype=np.float64):
    """
    Price a DataFrame of option contracts by naming the column of each input.

    `sigma_col` only  -> option prices and greeks.
    `price_col` only  -> implied volatilities and greeks (at the implied volatility).
    both              -> greeks only (and a warning).

    :return: None if `inplace`, else a DataFrame (index = df.index) with the columns among
             Price, IV, delta, gamma, theta, rho, vega.

    >>> df = pd.DataFrame({'Flag': ['c', 'p'], 'S': 95, 'K': [100, 90], 'T': 0.2, 'R': 0.2, 'IV': 0.2})
    >>> price_dataframe(df, flag_col='Flag', underlying_price_col='S', strike_col='K', annualized_tte_col='T', riskfree_rate_col='R', sigma_col='IV', model='black_scholes', inplace=False)
          Price     delta     gamma     theta       rho      vega
    0  2.895588  0.467506  0.046795 -0.045900  0.083035  0.168926
    1  0.611094 -0.136447  0.025739 -0.005335 -0.027151  0.092838
    """
    assert flag_col is not None, "You must specify a `flag_col` argument!"
    assert underlying_price_col is not None, "You must specify a `underlying_price_col` argument!"
    assert strike_col is not None, "You must specify a `strike_col` argument!"
    assert annualized_tte_col is not None, "You must specify a `annualized_tte_col` argument!"
    assert riskfree_rate_col is not None, "You must specify a `riskfree_rate_col` argument!"
    for col in [flag_col, underlying_price_col, strike_col, annualized_tte_col, riskfree_rate_col]:
        _validate_df_col(col, df)

    flag = _preprocess_flags(df[flag_col], dtype=dtype)
    inputs = [df[underlying_price_col], df[strike_col], df[annualized_tte_col], df[riskfree_rate_col], flag]

    _px_calc, _iv_calc = False, False
    price = sigma = None
    if price_col is not None and sigma_col is not None:
        warnings.warn(
            "Both `price_col` and `sigma_col` were specified. If these two calculations were not calculated using the same "
            "library, this may lead to inconsistencies.",
            stacklevel=2)
        _validate_df_col(price_col, df)
        _validate_df_col(sigma_col, df)
        sigma = df[sigma_col]
    elif price_col is not None:
        _validate_df_col(price_col, df)
        price = df[price_col]
        _iv_calc = True
    elif sigma_col is not None:
        _validate_df_col(sigma_col, df)
        sigma = df[sigma_col]
        _px_calc = True
    else:
        raise ValueError("You must specify either `sigma_col`, `price_col`, or both!")

    q = None
    if model == "black_scholes_merton":
        if dividend_col is not None:
            _validate_df_col(dividend_col, df)
            q = df[dividend_col]
        else:
            raise ValueError("You must specify a `dividend_col` value with `black_scholes_merton` model.")
    elif model not in ("black", "black_scholes"):
        raise ValueError(_MODEL_ERROR)

    # one broadcast for everything the kernels need
    extra = [x for x in (price, sigma, q) if x is not None]
    c = maybe_format_data_and_broadcast(*inputs, *extra, dtype=dtype)
    cols = _engine.pin_columns(c.cols)  # one pinned staging copy serves every kernel call below
    S, K, t, r, flag = cols[:5]
    rest = list(cols[5:])
    price_c = rest.pop(0) if price is not None else None
    sigma_c = rest.pop(0) if sigma is not None else None
    q_c = rest.pop(0) if q is not None else None

    results = {}
    if _iv_calc:
        # K4: implied volatility and the greeks at that volatility in one call (api.py:142-178)
        outs, status, _ = _engine.iv_greeks(model, c.n, flag, price_c, S, K, t, r, q_c, mask_errors=True)
        report_iv_status(status, "warn")
        results["IV"] = outs["iv"]
        for k in _GREEKS:
            results[k] = outs[k]
    else:
        if _px_calc:
            if model == "black":
                # the reference prices Black-76 undiscounted (models.py:38) but takes the greeks from the
                # Black-Scholes stencil (api.py:217-225): two different price functions
                results["Price"] = _engine.price("black", c.n, flag, S, K, t, None, None, sigma_c)
                outs = _engine.greeks(model, c.n, flag, S, K, t, r, q_c, sigma_c)
            else:
                outs = _engine.greeks(model, c.n, flag, S, K, t, r, q_c, sigma_c, with_price=True)
                results["Price"] = outs["price"]
        else:
            outs = _engine.greeks(model, c.n, flag, S, K, t, r, q_c, sigma_c)
        for k in _GREEKS:
            results[k] = outs[k]

    if inplace:
        for k, v in results.items():
            df[k] = v
        return None
    return pd.DataFrame(results, index=df.index, copy=False)
