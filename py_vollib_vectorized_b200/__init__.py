"""py_vollib_vectorized_b200 — a B200-native engine behind the `py_vollib_vectorized` API.

Importing this package (like importing the reference, /root/reference/py_vollib_vectorized/
__init__.py:46-127) rebinds the 21 `py_vollib` entry points — `py_vollib.black.black`,
`py_vollib.black_scholes.implied_volatility.implied_volatility`,
`py_vollib.black_scholes_merton.greeks.numerical.delta`, ... — to vectorised callables.  Here those
callables run hand-written fp64 CUDA kernels for sm_100a through the C ABI in include/pvv.h.
There is no CPU fallback: without libpvv.so the import fails, without a GPU every call raises.
"""
import inspect
from functools import partial, update_wrapper

from . import _engine  # noqa: F401  (loads libpvv.so; fails loudly if it has not been built)
from .api import get_all_greeks, price_dataframe
from .greeks import delta as vectorized_delta, gamma as vectorized_gamma, rho as vectorized_rho, \
    theta as vectorized_theta, vega as vectorized_vega
from .implied_volatility import vectorized_implied_volatility, vectorized_implied_volatility_black
from .models import vectorized_black, vectorized_black_scholes, vectorized_black_scholes_merton

__version__ = '0.1'

__all__ = [
    'price_dataframe',
    'get_all_greeks',
    'vectorized_implied_volatility_black',
    'vectorized_implied_volatility',
    'vectorized_delta',
    'vectorized_gamma',
    'vectorized_rho',
    'vectorized_vega',
    'vectorized_theta',
    'vectorized_black',
    'vectorized_black_scholes',
    'vectorized_black_scholes_merton',
    '__version__'
]


class repr_partial(partial):
    """functools.partial whose repr names the vectorised function (reference __init__.py:34-43)."""

    def __repr__(self):
        sig = [p.name for p in inspect.signature(self.func).parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD]
        sig += [str(k) + "=" + str(v) for k, v in self.keywords.items()]
        return "Vectorized implementation of <{fn}({args}{kwargs})>".format(
            fn=self.func.__name__.replace("vectorized_", ""),
            args=", ".join(repr(a) for a in self.args),
            kwargs=", ".join(sig))


def _patch(module, attr, func, **keywords):
    module.__dict__[attr] = repr_partial(func, **keywords)
    update_wrapper(module.__dict__[attr], func)


def apply_monkeypatches():
    """The reference's 21 patches (__init__.py:55-127).  Uses the real py_vollib when it is
    installed, otherwise a stub namespace with the same module paths (_py_vollib_stub.py)."""
    import importlib
    try:
        import py_vollib  # noqa: F401
    except ImportError:
        from . import _py_vollib_stub
        _py_vollib_stub.install()

    def mod(name):
        return importlib.import_module(name)

    _patch(mod("py_vollib.black.implied_volatility"), "implied_volatility", vectorized_implied_volatility_black, model="black")
    _patch(mod("py_vollib.black_scholes.implied_volatility"), "implied_volatility", vectorized_implied_volatility, model="black_scholes")
    _patch(mod("py_vollib.black_scholes_merton.implied_volatility"), "implied_volatility", vectorized_implied_volatility,
           model="black_scholes_merton")
    for m in ("black", "black_scholes", "black_scholes_merton"):
        target = mod(f"py_vollib.{m}.greeks.numerical")
        for name, fn in (("delta", vectorized_delta), ("gamma", vectorized_gamma), ("rho", vectorized_rho),
                         ("theta", vectorized_theta), ("vega", vectorized_vega)):
            _patch(target, name, fn, model=m)
    _patch(mod("py_vollib.black"), "black", vectorized_black)
    _patch(mod("py_vollib.black_scholes"), "black_scholes", vectorized_black_scholes)
    _patch(mod("py_vollib.black_scholes_merton"), "black_scholes_merton", vectorized_black_scholes_merton)


apply_monkeypatches()
