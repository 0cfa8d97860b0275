"""Names the engine needs from `py_vollib` when that package is not installed.

The reference refuses to import without py_vollib (`__init__.py:50-53`).  This engine prefers the
real package (its exception classes are then the ones users already catch) and otherwise installs a
minimal namespace with the same module paths, so that `py_vollib.black_scholes.black_scholes(...)`
and friends resolve to the vectorised callables exactly as the reference's 21 monkey-patches do.
"""
import sys
import types

FLOAT_MAX = sys.float_info.max
MINUS_FLOAT_MAX = -FLOAT_MAX


class PriceIsAboveMaximum(Exception):
    def __init__(self):
        Exception.__init__(self, "The volatility is above the maximum price.")


class PriceIsBelowIntrinsic(Exception):
    def __init__(self):
        Exception.__init__(self, "The volatility is below the intrinsic value.")


def install():
    """Register stub `py_vollib.*` modules (patch targets only) in sys.modules."""
    def mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            m.__package__ = name
            m.__path__ = []
            sys.modules[name] = m
            parent, _, child = name.rpartition(".")
            if parent:
                setattr(mod(parent), child, m)
        return m

    root = mod("py_vollib")
    root.__pvv_stub__ = True
    h = mod("py_vollib.helpers")
    c = mod("py_vollib.helpers.constants")
    c.FLOAT_MAX, c.MINUS_FLOAT_MAX = FLOAT_MAX, MINUS_FLOAT_MAX
    e = mod("py_vollib.helpers.exceptions")
    e.PriceIsAboveMaximum, e.PriceIsBelowIntrinsic = PriceIsAboveMaximum, PriceIsBelowIntrinsic
    for m in ("black", "black_scholes", "black_scholes_merton"):
        mod(f"py_vollib.{m}")
        mod(f"py_vollib.{m}.implied_volatility")
        mod(f"py_vollib.{m}.greeks")
        mod(f"py_vollib.{m}.greeks.numerical")
    return root
