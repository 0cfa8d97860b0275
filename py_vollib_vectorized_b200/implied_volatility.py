"""Vectorised "Let's be rational" implied volatility on the B200 engine.

Public surface of /root/reference/py_vollib_vectorized/implied_volatility.py (:8-92, :95-129).
Everything the reference does around its numba loop — deflater, undiscounted price, forward,
below-intrinsic / above-maximum masks, NaN writes (implied_volatility.py:48-86,
util/data_format.py:42-94) — is fused into kernel K3 together with the solver itself
(_iv_models.py:18-276); the kernel hands back one status byte per contract from which the
reference's warnings and exceptions are reproduced here.
"""
import numpy as np

from . import _engine
from .models import _wrap
from .util.data_format import _preprocess_flags, maybe_format_data_and_broadcast, report_iv_status

_MODEL_ERROR = "Model must be one of: `black`, `black_scholes`, `black_scholes_merton`"
_Q_ERROR = "Must pass a `q` to black scholes merton model (annualized continuous dividend yield)."


def _iv_columns(price, S, K, t, r, flag, q, model, dtype):
    flag = _preprocess_flags(flag, dtype)
    if model == "black_scholes_merton":
        if q is None:
            # the reference broadcasts the six mandatory columns before it looks at q
            maybe_format_data_and_broadcast(price, S, K, t, r, flag, dtype=dtype)
            raise ValueError(_Q_ERROR)
        c = maybe_format_data_and_broadcast(price, S, K, t, r, flag, q, dtype=dtype)
    else:
        c = maybe_format_data_and_broadcast(price, S, K, t, r, flag, dtype=dtype)
    return c


def vectorized_implied_volatility(price, S, K, t, r, flag, q=None, *, on_error="warn",
                                  model="black_scholes", return_as="dataframe",
                                  dtype=np.float64, **kwargs):
    """
    Implied volatility of option contracts (Jaeckel's "Let's be rational", two Householder steps).
    Inputs can be lists, tuples, floats, `pd.Series` or numpy arrays; broadcasting is applied.

    :param price: The price of the option.
    :param S: The price of the underlying asset.
    :param K: The strike price.
    :param t: The annualized time to expiration.
    :param r: The interest free rate.
    :param flag: 'c' / 'p' (or 1 / -1) per contract.
    :param q: The annualized continuous dividend yield (black_scholes_merton only).
    :param on_error: "raise", "warn" or "ignore" — prices below intrinsic / above the maximum price.
    :param model: "black", "black_scholes" or "black_scholes_merton".
    :param return_as: "series" -> pd.Series, "dataframe" -> pd.DataFrame, anything else -> numpy array.
    :param dtype: Input data type (computation and result are float64).
    :param kwargs: Other keyword arguments are ignored.

    >>> vectorized_implied_volatility(0.10, 95, [100, 90], .2, .2, ['c', 'p'], q=0, model='black_scholes_merton', return_as='numpy')
    array([0.02621257, 0.12585767])
    """
    assert on_error in ["warn", "ignore", "raise"], "`on_error` must be 'warn', 'ignore' or 'raise'."
    if model not in ("black", "black_scholes", "black_scholes_merton"):
        flag = _preprocess_flags(flag, dtype)
        maybe_format_data_and_broadcast(price, S, K, t, r, flag, dtype=dtype)
        raise ValueError(_MODEL_ERROR)
    c = _iv_columns(price, S, K, t, r, flag, q, model, dtype)
    price, S, K, t, r, flag = c.cols[:6]
    q = c.cols[6] if model == "black_scholes_merton" else None
    sigma_calc, status, _ = _engine.iv(model, c.n, flag, price, S, K, t, r, q, mask_errors=(on_error != "ignore"))
    report_iv_status(status, on_error)
    return _wrap(sigma_calc, "IV", return_as)


def vectorized_implied_volatility_black(price, F, K, r, t, flag, *, on_error="warn", return_as="dataframe",
                                        dtype=np.float64, **kwargs):
    """
    Implied volatility in the Black (futures) model.  NOTE the argument order `r` before `t`,
    kept consistent with py_vollib (implied_volatility.py:95).

    >>> vectorized_implied_volatility_black(0.10, 95, [100, 90], .2, .2, ['c', 'p'], return_as='numpy')
    array([0.07758129, 0.08177732])
    """
    return vectorized_implied_volatility(price, F, K, t, r, flag, model="black", on_error=on_error, return_as=return_as,
                                         dtype=dtype)
