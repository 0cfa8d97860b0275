"""Delbourgo-Gregory rational cubic interpolation helpers used by "Let's be rational"."""
from math import fabs, sqrt

from py_lets_be_rational.constants import DBL_EPSILON, DBL_MAX, DBL_MIN
from py_lets_be_rational.numba_helper import maybe_jit

minimum_rational_cubic_control_parameter_value = -(1 - sqrt(DBL_EPSILON))
maximum_rational_cubic_control_parameter_value = 2 / (DBL_EPSILON * DBL_EPSILON)


@maybe_jit()
def _is_zero(x):
    return fabs(x) < DBL_MIN


@maybe_jit()
def rational_cubic_interpolation(x, x_l, x_r, y_l, y_r, d_l, d_r, r):
    h = (x_r - x_l)
    if fabs(h) <= 0:
        return 0.5 * (y_l + y_r)
    t = (x - x_l) / h
    if not (r >= maximum_rational_cubic_control_parameter_value):
        omt = 1 - t
        t2 = t * t
        omt2 = omt * omt
        return (y_r * t2 * t + (r * y_r - h * d_r) * t2 * omt + (r * y_l + h * d_l) * t * omt2 + y_l * omt2 * omt) / (
                1 + (r - 3) * t * omt)
    return y_r * t + y_l * (1 - t)


@maybe_jit()
def rational_cubic_control_parameter_to_fit_second_derivative_at_left_side(x_l, x_r, y_l, y_r, d_l, d_r,
                                                                           second_derivative_l):
    h = (x_r - x_l)
    numerator = 0.5 * h * second_derivative_l + (d_r - d_l)
    if _is_zero(numerator):
        return 0.0
    denominator = (y_r - y_l) / h - d_l
    if _is_zero(denominator):
        return maximum_rational_cubic_control_parameter_value if numerator > 0 else minimum_rational_cubic_control_parameter_value
    return numerator / denominator


@maybe_jit()
def rational_cubic_control_parameter_to_fit_second_derivative_at_right_side(x_l, x_r, y_l, y_r, d_l, d_r,
                                                                            second_derivative_r):
    h = (x_r - x_l)
    numerator = 0.5 * h * second_derivative_r + (d_r - d_l)
    if _is_zero(numerator):
        return 0.0
    denominator = d_r - (y_r - y_l) / h
    if _is_zero(denominator):
        return maximum_rational_cubic_control_parameter_value if numerator > 0 else minimum_rational_cubic_control_parameter_value
    return numerator / denominator


@maybe_jit()
def minimum_rational_cubic_control_parameter(d_l, d_r, s, preferShapePreservationOverSmoothness):
    monotonic = d_l * s >= 0 and d_r * s >= 0
    convex = d_l <= s and s <= d_r
    concave = d_l >= s and s >= d_r
    if not monotonic and not convex and not concave:
        return minimum_rational_cubic_control_parameter_value
    d_r_m_d_l = d_r - d_l
    d_r_m_s = d_r - s
    s_m_d_l = s - d_l
    r1 = -DBL_MAX
    r2 = r1
    if monotonic:
        if not _is_zero(s):
            r1 = (d_r + d_l) / s
        elif preferShapePreservationOverSmoothness:
            r1 = maximum_rational_cubic_control_parameter_value
    if convex or concave:
        if not (_is_zero(s_m_d_l) or _is_zero(d_r_m_s)):
            r2 = max(fabs(d_r_m_d_l / d_r_m_s), fabs(d_r_m_d_l / s_m_d_l))
        elif preferShapePreservationOverSmoothness:
            r2 = maximum_rational_cubic_control_parameter_value
    elif monotonic and preferShapePreservationOverSmoothness:
        r2 = maximum_rational_cubic_control_parameter_value
    return max(minimum_rational_cubic_control_parameter_value, max(r1, r2))


@maybe_jit()
def convex_rational_cubic_control_parameter_to_fit_second_derivative_at_left_side(x_l, x_r, y_l, y_r, d_l, d_r,
                                                                                  second_derivative_l,
                                                                                  preferShapePreservationOverSmoothness):
    r = rational_cubic_control_parameter_to_fit_second_derivative_at_left_side(x_l, x_r, y_l, y_r, d_l, d_r,
                                                                               second_derivative_l)
    r_min = minimum_rational_cubic_control_parameter(d_l, d_r, (y_r - y_l) / (x_r - x_l),
                                                     preferShapePreservationOverSmoothness)
    return max(r, r_min)


@maybe_jit()
def convex_rational_cubic_control_parameter_to_fit_second_derivative_at_right_side(x_l, x_r, y_l, y_r, d_l, d_r,
                                                                                   second_derivative_r,
                                                                                   preferShapePreservationOverSmoothness):
    r = rational_cubic_control_parameter_to_fit_second_derivative_at_right_side(x_l, x_r, y_l, y_r, d_l, d_r,
                                                                                second_derivative_r)
    r_min = minimum_rational_cubic_control_parameter(d_l, d_r, (y_r - y_l) / (x_r - x_l),
                                                     preferShapePreservationOverSmoothness)
    return max(r, r_min)
