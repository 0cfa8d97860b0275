"""Vectorised Black / Black-Scholes / Black-Scholes-Merton pricing on the B200 engine.

Public surface of /root/reference/py_vollib_vectorized/models.py (:8-45, :48-86, :89-129): same
names, argument order, keyword-only `return_as` / `dtype`, broadcasting and return containers.  The
numba loops `_black_vectorized_call`, `_black_scholes_vectorized_call` and
`_black_scholes_merton_vectorized_call` (_model_calls.py:81-102) are replaced by kernel K1
(csrc/pvv_kernels.cu) reached through `pvv_price_host` (include/pvv.h).
"""
import numpy as np
import pandas as pd

from . import _engine
from .util.data_format import _preprocess_flags, maybe_format_data_and_broadcast


def _wrap(values, name, return_as):
    if return_as == "series":
        return pd.Series(values, name=name)
    elif return_as == "dataframe":
        return pd.DataFrame(values, columns=[name])
    return values


def vectorized_black(flag, F, K, t, r, sigma, *, return_as="dataframe", dtype=np.float64):
    """
    Price a Future option using the Black model.  Broadcasting is applied on the inputs.

    As in the reference (models.py:38) `r` is accepted but unused: the result is the UNDISCOUNTED
    Black-76 price.

    :param flag: 'c' / 'p' (or 1 / -1) per contract.
    :param F: The price of the underlying asset (forward).
    :param K: The strike price.
    :param t: The annualized time to expiration.
    :param r: Ignored (kept for signature compatibility).
    :param sigma: The volatility (as a decimal).
    :param return_as: "series" -> pd.Series, "dataframe" -> pd.DataFrame, anything else -> numpy array.
    :param dtype: Input data type (computation and result are float64).

    >>> vectorized_black(['c', 'p'], 95, [100, 90], .2, .2, .2, return_as='numpy')
    array([1.53408169, 1.38409245])
    """
    flag = _preprocess_flags(flag, dtype=dtype)
    c = maybe_format_data_and_broadcast(F, K, sigma, t, flag, dtype=dtype)
    F, K, sigma, t, flag = c
    prices = _engine.price("black", c.n, flag, F, K, t, None, None, sigma)
    return _wrap(prices, "Price", return_as)


def vectorized_black_scholes(flag, S, K, t, r, sigma, *, return_as="dataframe", dtype=np.float64):
    """
    Price an option using the Black-Scholes model.  Broadcasting is applied on the inputs.

    :param flag: 'c' / 'p' (or 1 / -1) per contract.
    :param S: The underlying asset price.
    :param K: The strike price.
    :param t: The annualized time to expiration.
    :param r: The interest free rate.
    :param sigma: The volatility (as a decimal).
    :param return_as: "series" -> pd.Series, "dataframe" -> pd.DataFrame, anything else -> numpy array.
    :param dtype: Input data type (computation and result are float64).

    >>> vectorized_black_scholes(['c', 'p'], 95, [100, 90], .2, .2, .2, return_as='numpy')
    array([2.89558836, 0.61109351])
    """
    flag = _preprocess_flags(flag, dtype=dtype)
    c = maybe_format_data_and_broadcast(S, K, sigma, t, r, flag, dtype=dtype)
    S, K, sigma, t, r, flag = c
    prices = _engine.price("black_scholes", c.n, flag, S, K, t, r, None, sigma)
    return _wrap(prices, "Price", return_as)


def vectorized_black_scholes_merton(flag, S, K, t, r, sigma, q, *, return_as="dataframe", dtype=np.float64):
    """
    Price an option using the Black-Scholes-Merton model.  Broadcasting is applied on the inputs.

    :param q: The annualized continuous dividend yield.  Other parameters as `vectorized_black_scholes`.

    >>> vectorized_black_scholes_merton(['c', 'p'], [95, 99], [100, 90], .2, .2, .2, 0, return_as='numpy')
    array([2.89558836, 0.23536284])
    """
    flag = _preprocess_flags(flag, dtype=dtype)
    c = maybe_format_data_and_broadcast(flag, S, K, t, r, sigma, q, dtype=dtype)
    flag, S, K, t, r, sigma, q = c
    prices = _engine.price("black_scholes_merton", c.n, flag, S, K, t, r, q, sigma)
    return _wrap(prices, "Price", return_as)
