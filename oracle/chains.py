"""Re-export of the seeded workload generators (they live in the product package, numpy only)."""
import importlib.util
import os

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "py_vollib_vectorized_b200", "util", "chains.py")
_spec = importlib.util.spec_from_file_location("_pvv_chains", _p)
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
chain_cfg2, chain_wide, chain_edge, chain_cfg5 = _m.chain_cfg2, _m.chain_wide, _m.chain_edge, _m.chain_cfg5


def chain_cfg3(n, seed=1, oracle=None, pricer=None):
    """cfg 3 with the clean BSM prices supplied by the oracle (CPU-side scripts and tests)."""
    if pricer is None:
        def pricer(flag, S, K, t, r, sigma, q):
            return oracle.price("black_scholes_merton", flag, S, K, t, r, sigma, q, threads=oracle.max_threads)[0]
    return _m.chain_cfg3(n, seed, pricer=pricer)
