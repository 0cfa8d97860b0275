"""CPU oracle of the hot path — TEST INFRASTRUCTURE ONLY (see oracle/pvv_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package never does.
"""
from .oracle import (Oracle, build_oracle, get_oracle, MODEL_BLACK, MODEL_BLACK_SCHOLES,  # noqa: F401
                     MODEL_BLACK_SCHOLES_MERTON)
