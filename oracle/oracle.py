"""ctypes front end of oracle/libpvv_oracle.so (plain C restatement of the reference's hot path).

TEST INFRASTRUCTURE ONLY: checker for tests/ and smoke(), timed CPU baseline for bench.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libpvv_oracle.so")

MODEL_BLACK, MODEL_BLACK_SCHOLES, MODEL_BLACK_SCHOLES_MERTON = 0, 1, 2
_MODELS = {"black": 0, "black_scholes": 1, "black_scholes_merton": 2, 0: 0, 1: 1, 2: 2}


def build_oracle(force=False):
    src = os.path.join(_HERE, "pvv_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "libpvv_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB


_c_dp = ctypes.POINTER(ctypes.c_double)


def _col(a, n):
    """float64 column + element stride (0 for a broadcast scalar), like the product C-ABI."""
    a = np.ascontiguousarray(np.ravel(np.asarray(a, dtype=np.float64)))
    if a.size == 1 and n != 1:
        return a, 0
    if a.size != n:
        raise ValueError(f"column of {a.size} elements in a call of {n} contracts")
    return a, 1


def _flag(f, n):
    f = np.ascontiguousarray(np.ravel(np.asarray(f)).astype(np.int8))
    if f.size == 1 and n != 1:
        return f, 0
    if f.size != n:
        raise ValueError("flag column length")
    return f, 1


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    def __init__(self, path=None):
        self.lib = ctypes.CDLL(path or build_oracle())
        L = self.lib
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        L.pvv_oracle_price.restype = i64
        L.pvv_oracle_price.argtypes = [i32, i64] + [vp] * 9 + [i32]
        L.pvv_oracle_greeks.restype = i64
        L.pvv_oracle_greeks.argtypes = [i32, i64] + [vp] * 14 + [i32]
        L.pvv_oracle_analytic_greeks.restype = None
        L.pvv_oracle_analytic_greeks.argtypes = [i32, i64] + [vp] * 14 + [i32]
        L.pvv_oracle_iv.restype = i64
        L.pvv_oracle_iv.argtypes = [i32, i64] + [vp] * 8 + [i32] + [vp] * 3 + [i32]
        L.pvv_oracle_iv_greeks.restype = i64
        L.pvv_oracle_iv_greeks.argtypes = [i32, i64] + [vp] * 8 + [i32] + [vp] * 7 + [i32]
        for name in ("erfcx", "erfc", "norm_cdf", "inverse_norm_cdf", "exp", "log"):
            fn = getattr(L, "pvv_oracle_" + name)
            fn.restype = ctypes.c_double
            fn.argtypes = [ctypes.c_double]
        L.pvv_oracle_normalised_black_call.restype = ctypes.c_double
        L.pvv_oracle_normalised_black_call.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.POINTER(i32)]
        L.pvv_oracle_normalised_iv.restype = ctypes.c_double
        L.pvv_oracle_normalised_iv.argtypes = [ctypes.c_double] * 3 + [i32, ctypes.POINTER(i32)]
        L.pvv_oracle_max_threads.restype = i32
        self.max_threads = L.pvv_oracle_max_threads()

    @staticmethod
    def _n(*cols):
        return max(np.size(c) for c in cols if c is not None)

    def price(self, model, flag, S, K, t, r, sigma, q=None, threads=1):
        """Returns (price, first_zero_division_index or -1).  model 'black': S is the forward, r ignored."""
        m = _MODELS[model]
        n = self._n(flag, S, K, t, r, sigma, q)
        f, sf = _flag(flag, n)
        cols = [_col(c if c is not None else 0.0, n) for c in (S, K, t, r, q, sigma)]
        stride = np.array([sf] + [c[1] for c in cols], dtype=np.int64)
        out = np.empty(n, dtype=np.float64)
        zde = self.lib.pvv_oracle_price(m, n, _p(f), *[_p(c[0]) for c in cols], _p(stride), _p(out), threads)
        return out, int(zde)

    def greeks(self, model, flag, S, K, t, r, sigma, q=None, threads=1, with_price=True):
        """Returns (dict(price, delta, gamma, theta, rho, vega), zde_index)."""
        m = _MODELS[model]
        n = self._n(flag, S, K, t, r, sigma, q)
        f, sf = _flag(flag, n)
        cols = [_col(c if c is not None else 0.0, n) for c in (S, K, t, r, q, sigma)]
        stride = np.array([sf] + [c[1] for c in cols], dtype=np.int64)
        names = ["price", "delta", "gamma", "theta", "rho", "vega"]
        outs = {k: np.empty(n, dtype=np.float64) for k in names}
        zde = self.lib.pvv_oracle_greeks(m, n, _p(f), *[_p(c[0]) for c in cols], _p(stride),
                                         *[_p(outs[k]) for k in names], threads)
        if not with_price:
            outs.pop("price")
        return outs, int(zde)

    def analytic_greeks(self, model, flag, S, K, t, r, sigma, q=None, threads=1):
        """Closed-form price and greeks (py_vollib greeks.analytical conventions)."""
        m = _MODELS[model]
        n = self._n(flag, S, K, t, r, sigma, q)
        f, sf = _flag(flag, n)
        cols = [_col(c if c is not None else 0.0, n) for c in (S, K, t, r, q, sigma)]
        stride = np.array([sf] + [c[1] for c in cols], dtype=np.int64)
        names = ["price", "delta", "gamma", "theta", "rho", "vega"]
        outs = {k: np.empty(n, dtype=np.float64) for k in names}
        self.lib.pvv_oracle_analytic_greeks(m, n, _p(f), *[_p(c[0]) for c in cols], _p(stride), *[_p(outs[k]) for k in names], threads)
        return outs

    def iv(self, model, price, S, K, t, r, flag, q=None, mask_errors=True, threads=1, want_branch=False):
        """Returns (iv, status, zde_index[, branch])."""
        m = _MODELS[model]
        n = self._n(flag, price, S, K, t, r, q)
        f, sf = _flag(flag, n)
        cols = [_col(c if c is not None else 0.0, n) for c in (price, S, K, t, r, q)]
        stride = np.array([sf] + [c[1] for c in cols], dtype=np.int64)
        iv = np.empty(n, dtype=np.float64)
        status = np.empty(n, dtype=np.uint8)
        branch = np.empty(n, dtype=np.int8) if want_branch else None
        zde = self.lib.pvv_oracle_iv(m, n, _p(f), *[_p(c[0]) for c in cols], _p(stride), int(bool(mask_errors)),
                                     _p(iv), _p(status), _p(branch), threads)
        if want_branch:
            return iv, status, int(zde), branch
        return iv, status, int(zde)

    def iv_greeks(self, model, price, S, K, t, r, flag, q=None, mask_errors=True, threads=1):
        """Returns (dict(iv, delta, gamma, theta, rho, vega), status, zde_index)."""
        m = _MODELS[model]
        n = self._n(flag, price, S, K, t, r, q)
        f, sf = _flag(flag, n)
        cols = [_col(c if c is not None else 0.0, n) for c in (price, S, K, t, r, q)]
        stride = np.array([sf] + [c[1] for c in cols], dtype=np.int64)
        names = ["iv", "delta", "gamma", "theta", "rho", "vega"]
        outs = {k: np.empty(n, dtype=np.float64) for k in names}
        status = np.empty(n, dtype=np.uint8)
        zde = self.lib.pvv_oracle_iv_greeks(m, n, _p(f), *[_p(c[0]) for c in cols], _p(stride), int(bool(mask_errors)),
                                            _p(outs["iv"]), _p(status), *[_p(outs[k]) for k in names[1:]], threads)
        return outs, status, int(zde)

    def normalised_black_call(self, x, s):
        reg = ctypes.c_int(-1)
        return self.lib.pvv_oracle_normalised_black_call(x, s, ctypes.byref(reg)), reg.value

    def normalised_iv(self, beta, x, q=1.0, N=2):
        br = ctypes.c_int(-1)
        return self.lib.pvv_oracle_normalised_iv(beta, x, q, N, ctypes.byref(br)), br.value


_ORACLE = None


def get_oracle():
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE
