"""CPU tier: `import py_vollib_vectorized` (the reference's name) resolves to the engine, and the
reference's own monkey-patch test (tests/test_monkeypatch.py:5-42) holds verbatim."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r'''
import py_vollib_vectorized
import py_vollib.black.implied_volatility, py_vollib.black_scholes.implied_volatility, py_vollib.black_scholes_merton.implied_volatility
import py_vollib.black, py_vollib.black_scholes, py_vollib.black_scholes_merton
import py_vollib.black.greeks.numerical, py_vollib.black_scholes.greeks.numerical, py_vollib.black_scholes_merton.greeks.numerical
assert py_vollib.black.implied_volatility.implied_volatility.__module__ == "py_vollib_vectorized.implied_volatility"
assert py_vollib.black_scholes.implied_volatility.implied_volatility.__module__ == "py_vollib_vectorized.implied_volatility"
assert py_vollib.black_scholes_merton.implied_volatility.implied_volatility.__module__ == "py_vollib_vectorized.implied_volatility"
assert py_vollib.black.black.__module__ == "py_vollib_vectorized.models"
assert py_vollib.black_scholes.black_scholes.__module__ == "py_vollib_vectorized.models"
assert py_vollib.black_scholes_merton.black_scholes_merton.__module__ == "py_vollib_vectorized.models"
assert py_vollib.black.greeks.numerical.delta.__module__ == "py_vollib_vectorized.greeks"
assert py_vollib.black_scholes.greeks.numerical.delta.__module__ == "py_vollib_vectorized.greeks"
assert py_vollib.black_scholes_merton.greeks.numerical.delta.__module__ == "py_vollib_vectorized.greeks"
from py_vollib_vectorized.implied_volatility import vectorized_implied_volatility
from py_vollib_vectorized.models import vectorized_black_scholes, vectorized_black_scholes_merton
from py_vollib_vectorized.api import get_all_greeks, price_dataframe
from py_vollib_vectorized import vectorized_delta, vectorized_black, __version__
assert repr(py_vollib.black.black) == "Vectorized implementation of <black(flag, F, K, t, r, sigma)>"
print("alias ok")
'''


def test_reference_name_resolves_to_the_engine():
    # a fresh interpreter: the relabelling of __module__ is a property of the alias import
    out = subprocess.run([sys.executable, "-c", CODE], cwd=ROOT, capture_output=True, text=True)
    assert out.returncode == 0 and "alias ok" in out.stdout, out.stderr[-2000:]
