"""The pieces of Jaeckel's "Let's be rational" that the reference reaches through this module.

`normalised_black_call` is bound to the reference's own vendored four-region function
(`py_vollib_vectorized/_model_calls.py:18-42`; upstream's is the same algorithm), so import
`py_vollib_vectorized` BEFORE this module (the reference's own import order does that).
"""
from math import exp, fabs

from py_lets_be_rational.constants import *  # noqa: F401,F403
from py_lets_be_rational.constants import (DBL_EPSILON, DBL_MAX, DBL_MIN, DENORMALIZATION_CUTOFF,
                                           ONE_OVER_SQRT_TWO_PI, PI_OVER_SIX, SQRT_DBL_MAX, SQRT_DBL_MIN,
                                           SQRT_ONE_OVER_THREE, SQRT_PI_OVER_TWO, SQRT_THREE, TWO_PI,
                                           TWO_PI_OVER_SQRT_TWENTY_SEVEN)
from py_lets_be_rational.normaldistribution import inverse_norm_cdf, norm_cdf, norm_pdf
from py_lets_be_rational.numba_helper import maybe_jit
from py_lets_be_rational.rationalcubic import (  # noqa: F401
    convex_rational_cubic_control_parameter_to_fit_second_derivative_at_left_side,
    convex_rational_cubic_control_parameter_to_fit_second_derivative_at_right_side,
    rational_cubic_interpolation)

from py_vollib_vectorized._model_calls import normalised_black_call  # noqa: E402,F401
from py_vollib_vectorized.util.greeks_helpers import _normalised_intrinsic  # noqa: E402,F401

implied_volatility_maximum_iterations = 2


@maybe_jit()
def _is_below_horizon(x):
    return fabs(x) < DENORMALIZATION_CUTOFF


@maybe_jit()
def _square(x):
    return x * x


@maybe_jit()
def _householder_factor(newton, halley, hh3):
    return (1 + 0.5 * halley * newton) / (1 + newton * (halley + hh3 * newton / 6))


@maybe_jit()
def normalised_vega(x, s):
    ax = fabs(x)
    if ax <= 0:
        return ONE_OVER_SQRT_TWO_PI * exp(-0.125 * s * s)
    if s <= 0 or s <= ax * SQRT_DBL_MIN:
        return 0.0
    return ONE_OVER_SQRT_TWO_PI * exp(-0.5 * (_square(x / s) + _square(0.5 * s)))


@maybe_jit()
def _compute_f_lower_map_and_first_two_derivatives(x, s):
    ax = fabs(x)
    z = SQRT_ONE_OVER_THREE * ax / s
    y = z * z
    s2 = s * s
    Phi = norm_cdf(-z)
    phi = norm_pdf(z)
    fpp = PI_OVER_SIX * y / (s2 * s) * Phi * (
            8 * SQRT_THREE * s * ax + (3 * s2 * (s2 - 8) - 8 * x * x) * Phi / phi) * exp(2 * y + 0.25 * s2)
    if _is_below_horizon(s):
        fp = 1.0
        f = 0.0
    else:
        Phi2 = Phi * Phi
        fp = TWO_PI * y * Phi2 * exp(y + 0.125 * s * s)
        if _is_below_horizon(x):
            f = 0.0
        else:
            f = TWO_PI_OVER_SQRT_TWENTY_SEVEN * ax * (Phi2 * Phi)
    return f, fp, fpp


@maybe_jit()
def _inverse_f_lower_map(x, f):
    if _is_below_horizon(f):
        return 0.0
    return fabs(x / (SQRT_THREE * inverse_norm_cdf(pow(f / (TWO_PI_OVER_SQRT_TWENTY_SEVEN * fabs(x)), 1. / 3.))))


@maybe_jit()
def _compute_f_upper_map_and_first_two_derivatives(x, s):
    f = norm_cdf(-0.5 * s)
    if _is_below_horizon(x):
        fp = -0.5
        fpp = 0.0
    else:
        w = _square(x / s)
        fp = -0.5 * exp(0.5 * w)
        fpp = SQRT_PI_OVER_TWO * exp(w + 0.125 * s * s) * w / s
    return f, fp, fpp


@maybe_jit()
def _inverse_f_upper_map(f):
    return -2. * inverse_norm_cdf(f)
