"""Re-export of the seeded workload generators (they live in the product package, numpy only)."""
import importlib.util
import os

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "py_vollib_vectorized_b200", "util", "chains.py")
_spec = importlib.util.spec_from_file_location("_pvv_chains", _p)
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
chain_cfg2, chain_cfg3, chain_wide, chain_edge, chain_cfg5 = _m.chain_cfg2, _m.chain_cfg3, _m.chain_wide, _m.chain_edge, _m.chain_cfg5
