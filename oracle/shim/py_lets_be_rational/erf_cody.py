"""W. J. Cody's rational Chebyshev erf / erfc / erfcx (Math. Comp. 1969, 631-638; CALERF)."""
from math import exp, fabs, floor

from py_lets_be_rational.numba_helper import maybe_jit


@maybe_jit()
def d_int(x):
    return floor(x) if x > 0 else -floor(-x)


A = (3.1611237438705656, 113.864154151050156, 377.485237685302021, 3209.37758913846947, .185777706184603153)
B = (23.6012909523441209, 244.024637934444173, 1282.61652607737228, 2844.23683343917062)
C = (.564188496988670089, 8.88314979438837594, 66.1191906371416295, 298.635138197400131, 881.95222124176909,
     1712.04761263407058, 2051.07837782607147, 1230.33935479799725, 2.15311535474403846e-8)
D = (15.7449261107098347, 117.693950891312499, 537.181101862009858, 1621.38957456669019, 3290.79923573345963,
     4362.61909014324716, 3439.36767414372164, 1230.33935480374942)
P = (.305326634961232344, .360344899949804439, .125781726111229246, .0160837851487422766, 6.58749161529837803e-4,
     .0163153871373020978)
Q = (2.56852019228982242, 1.87295284992346047, .527905102951428412, .0605183413124413191, .00233520497626869185)

SQRPI = 0.56418958354775628695
THRESH = .46875
XINF = 1.79e308
XNEG = -26.628
XSMALL = 1.11e-16
XBIG = 26.543
XHUGE = 6.71e7
XMAX = 2.53e307


@maybe_jit()
def fix_up_for_negative_argument_erf_etc(jint, result, x):
    if jint == 0:
        result = (0.5 - result) + 0.5
        if x < 0.0:
            result = -result
    elif jint == 1:
        if x < 0.0:
            result = 2.0 - result
    else:
        if x < 0.0:
            if x < XNEG:
                result = XINF
            else:
                ysq = d_int(x * 16.0) / 16.0
                _del = (x - ysq) * (x + ysq)
                y = exp(ysq * ysq) * exp(_del)
                result = y + y - result
    return result


@maybe_jit()
def calerf(x, jint):
    y = fabs(x)
    if y <= THRESH:
        ysq = 0.0
        if y > XSMALL:
            ysq = y * y
        xnum = A[4] * ysq
        xden = ysq
        for i in range(0, 3):
            xnum = (xnum + A[i]) * ysq
            xden = (xden + B[i]) * ysq
        result = x * (xnum + A[3]) / (xden + B[3])
        if jint != 0:
            result = 1.0 - result
        if jint == 2:
            result *= exp(ysq)
        return result
    elif y <= 4.0:
        xnum = C[8] * y
        xden = y
        for i in range(0, 7):
            xnum = (xnum + C[i]) * y
            xden = (xden + D[i]) * y
        result = (xnum + C[7]) / (xden + D[7])
        if jint != 2:
            ysq = d_int(y * 16.0) / 16.0
            _del = (y - ysq) * (y + ysq)
            result *= exp(-ysq * ysq) * exp(-_del)
    else:
        result = 0.0
        if y >= XBIG:
            if jint != 2 or y >= XMAX:
                return fix_up_for_negative_argument_erf_etc(jint, result, x)
            if y >= XHUGE:
                result = SQRPI / y
                return fix_up_for_negative_argument_erf_etc(jint, result, x)
        ysq = 1.0 / (y * y)
        xnum = P[5] * ysq
        xden = ysq
        for i in range(0, 4):
            xnum = (xnum + P[i]) * ysq
            xden = (xden + Q[i]) * ysq
        result = ysq * (xnum + P[4]) / (xden + Q[4])
        result = (SQRPI - result) / y
        if jint != 2:
            ysq = d_int(y * 16.0) / 16.0
            _del = (y - ysq) * (y + ysq)
            result *= exp(-ysq * ysq) * exp(-_del)
    return fix_up_for_negative_argument_erf_etc(jint, result, x)


@maybe_jit()
def erf_cody(x):
    return calerf(x, 0)


@maybe_jit()
def erfc_cody(x):
    return calerf(x, 1)


@maybe_jit()
def erfcx_cody(x):
    return calerf(x, 2)
