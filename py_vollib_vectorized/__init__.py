"""Drop-in alias: `import py_vollib_vectorized` resolves to the B200 engine (py_vollib_vectorized_b200).

The reference's own monkey-patch test (tests/test_monkeypatch.py:21-42) checks that the patched
`py_vollib.*` callables report `__module__ == "py_vollib_vectorized.{implied_volatility,models,greeks}"`.
Importing the engine under the reference's name therefore (1) registers the engine's sub-modules under
the reference's module names, (2) relabels the public callables' `__module__`, and (3) re-applies the 21
patches so that the wrappers carry the relabelled names.  Nothing is computed here.
"""
import sys

import py_vollib_vectorized_b200 as _impl
from py_vollib_vectorized_b200 import api, greeks, implied_volatility, models, util  # noqa: F401
from py_vollib_vectorized_b200 import *  # noqa: F401,F403
from py_vollib_vectorized_b200 import __version__, apply_monkeypatches, repr_partial  # noqa: F401

for _name, _mod in (("api", api), ("greeks", greeks), ("implied_volatility", implied_volatility), ("models", models),
                    ("util", util), ("util.data_format", util.data_format)):
    sys.modules[__name__ + "." + _name] = _mod

for _mod, _names in ((models, ("vectorized_black", "vectorized_black_scholes", "vectorized_black_scholes_merton")),
                     (implied_volatility, ("vectorized_implied_volatility", "vectorized_implied_volatility_black")),
                     (greeks, ("delta", "gamma", "theta", "rho", "vega")),
                     (api, ("get_all_greeks", "price_dataframe"))):
    for _n in _names:
        getattr(_mod, _n).__module__ = __name__ + "." + _mod.__name__.rsplit(".", 1)[1]

apply_monkeypatches()
__all__ = _impl.__all__
