"""Vectorised finite-difference greeks on the B200 engine.

Public surface of /root/reference/py_vollib_vectorized/greeks.py (:12-316): `delta`, `theta`,
`vega`, `rho`, `gamma` (exported by the package as `vectorized_*`).  The reference's greeks are
NUMERICAL (central differences of the price function, _numerical_greeks.py:87-253) and so are
these: kernel K2 evaluates exactly the stencil points the requested greek needs.

Reference behaviours kept on purpose (each can be switched off with `fix_reference_bugs=True`):
  * model="black" raises TypeError: the reference passes 7 arguments to 6-parameter numba functions
    (greeks.py:46-47 vs _numerical_greeks.py:11) — `get_all_greeks(model="black")` is the working
    route (it uses the Black-Scholes stencil, api.py:217-225);
  * rho for black_scholes_merton does not re-broadcast `q` against the other columns
    (greeks.py:240-242): a scalar q with vector inputs raises ValueError there, as here.  delta,
    theta, vega and gamma DO re-broadcast it (greeks.py:54-56, 117, 178, 301).

`fix_reference_bugs=True` (opt-in, not part of the reference's surface):
  * model="black" works: finite differences of the DISCOUNTED Black-76 price exp(-r t) black(F, K, sigma, t),
    the function py_vollib.black.greeks.numerical differentiates and the reference's broken
    `numerical_*_black` (_numerical_greeks.py:10-83) set out to.  That is the Black-Scholes-Merton
    stencil at zero carry (q = r: the forward S exp((r - q) t) is F exactly), so it runs on the same kernel;
  * rho for black_scholes_merton broadcasts a scalar `q` like the other four greeks.
"""
import numpy as np

from . import _engine
from .models import _wrap
from .util.data_format import _preprocess_flags, _validate_data, maybe_format_data_and_broadcast

_MODEL_ERROR = "Model must be one of: `black`, `black_scholes`, `black_scholes_merton`"
_Q_ERROR = "Must pass a `q` to black scholes merton model (annualized continuous dividend yield)."


def _one_greek(name, flag, S, K, t, r, sigma, q, model, return_as, dtype, rebroadcast_q, fix_reference_bugs=False):
    flag = _preprocess_flags(flag, dtype=dtype)
    c = maybe_format_data_and_broadcast(S, K, t, r, sigma, flag, dtype=dtype)
    if model == "black":
        if not fix_reference_bugs:
            raise TypeError("too many arguments: expected 6, got 7")
        S, K, t, r, sigma, flag = c
        out = _engine.greeks("black_scholes_merton", c.n, flag, S, K, t, r, r, sigma, want=(name,))  # zero carry: q = r
    elif model == "black_scholes":
        S, K, t, r, sigma, flag = c
        out = _engine.greeks("black_scholes", c.n, flag, S, K, t, r, None, sigma, want=(name,))
    elif model == "black_scholes_merton":
        if q is None:
            raise ValueError(_Q_ERROR)
        if rebroadcast_q or fix_reference_bugs:
            c2 = maybe_format_data_and_broadcast(*c.cols, q, dtype=dtype)
        else:
            qa = maybe_format_data_and_broadcast(q, dtype=dtype)
            _validate_data(c.dense(3), qa.dense(0))  # the reference's length check of r against q
            c2 = maybe_format_data_and_broadcast(*c.cols, qa.cols[0], dtype=dtype)
        S, K, t, r, sigma, flag, q = c2
        out = _engine.greeks("black_scholes_merton", c2.n, flag, S, K, t, r, q, sigma, want=(name,))
    else:
        raise ValueError(_MODEL_ERROR)
    return _wrap(out[name], name, return_as)


def delta(flag, S, K, t, r, sigma, q=None, *, model="black_scholes", return_as="dataframe", dtype=np.float64, fix_reference_bugs=False):
    """
    Delta of each contract (central difference in S, dS = 0.01), as specified by `model`.

    >>> delta(['c', 'p'], 95, [100, 90], .2, .2, .2, model='black_scholes', return_as='numpy')
    array([ 0.46750566, -0.1364465 ])
    """
    return _one_greek("delta", flag, S, K, t, r, sigma, q, model, return_as, dtype, rebroadcast_q=True,
                      fix_reference_bugs=fix_reference_bugs)


def theta(flag, S, K, t, r, sigma, q=None, *, model="black_scholes", return_as="dataframe", dtype=np.float64, fix_reference_bugs=False):
    """
    Theta of each contract: one calendar day of decay, p(t - 1/365) - p(t) (not annualised).

    >>> theta(['c', 'p'], 95, [100, 90], .2, .2, .2, model='black_scholes', return_as='numpy')
    array([-0.04589963, -0.00533543])
    """
    return _one_greek("theta", flag, S, K, t, r, sigma, q, model, return_as, dtype, rebroadcast_q=True,
                      fix_reference_bugs=fix_reference_bugs)


def vega(flag, S, K, t, r, sigma, q=None, *, model="black_scholes", return_as="dataframe", dtype=np.float64, fix_reference_bugs=False):
    """
    Vega of each contract per volatility point: (p(sigma + 0.01) - p(sigma - 0.01)) / 2.

    >>> vega(['c', 'p'], 95, [100, 90], .2, .2, .2, model='black_scholes', return_as='numpy')
    array([0.16892575, 0.0928379 ])
    """
    return _one_greek("vega", flag, S, K, t, r, sigma, q, model, return_as, dtype, rebroadcast_q=True,
                      fix_reference_bugs=fix_reference_bugs)


def rho(flag, S, K, t, r, sigma, q=None, *, model="black_scholes", return_as="dataframe", dtype=np.float64, fix_reference_bugs=False):
    """
    Rho of each contract per rate point: (p(r + 0.01) - p(r - 0.01)) / 2 (for black_scholes_merton
    the carry r - q is held fixed, _numerical_greeks.py:229-232).

    >>> rho(['c', 'p'], 95, [100, 90], .2, .2, .2, model='black_scholes', return_as='numpy')
    array([ 0.0830349 , -0.02715114])
    """
    return _one_greek("rho", flag, S, K, t, r, sigma, q, model, return_as, dtype, rebroadcast_q=False,
                      fix_reference_bugs=fix_reference_bugs)


def gamma(flag, S, K, t, r, sigma, q=None, *, model="black_scholes", return_as="dataframe", dtype=np.float64, fix_reference_bugs=False):
    """
    Gamma of each contract (second central difference in S, dS = 0.01).

    >>> gamma(['c', 'p'], 95, [100, 90], .2, .2, .2, model='black_scholes', return_as='numpy')
    array([0.0467948, 0.0257394])
    """
    return _one_greek("gamma", flag, S, K, t, r, sigma, q, model, return_as, dtype, rebroadcast_q=True,
                      fix_reference_bugs=fix_reference_bugs)
