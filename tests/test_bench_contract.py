"""CPU tier: host-side logic of bench.py that the driver relies on — both arms print the same `config`, the strong-scaling
chain is independent of the rank count (contiguous shards of one seeded chain), and the reference arm's line has the
contract's keys (it runs the unmodified reference from baseline/_ref on a small sample when that is installed)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_config_is_the_same_object_in_both_arms():
    for wl in bench.WORKLOADS:
        for scaling, n in (("weak", bench.WORKLOADS[wl][2]), ("strong", bench.STRONG_TOTAL.get(wl, 1_000_000))):
            a = bench.config_of(wl, n, 8, scaling)
            b = bench.config_of(wl, n, 8, scaling)
            assert a == b and "workload" in a and "model" in a and not any(k in a for k in ("seq_len", "global_batch"))
            json.dumps(a)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_strong_scaling_shards_are_slices_of_one_chain(world, monkeypatch):
    monkeypatch.setattr(bench, "BLOCK", 1000)  # small blocks: the test covers shards that span block boundaries
    n = 4567
    whole = np.concatenate([bench.make_shard("iv_greeks", n, 1, 0, None)[k] for k in ("S",)])
    parts = [bench.make_shard("iv_greeks", n, world, r, None) for r in range(world)]
    for key in ("flag", "S", "K", "t", "r", "sigma", "q"):
        ref = bench.make_shard("iv_greeks", n, 1, 0, None)[key]
        got = np.concatenate([p[key] for p in parts])
        assert got.shape == (n,) and np.array_equal(ref, got), key
    sizes = [p["S"].size for p in parts]
    assert sum(sizes) == n == whole.size and max(sizes) - min(sizes) <= 1


def test_rotating_sets_exceed_the_l2():
    for wl, (_, _, n, bi, bo, _) in bench.WORKLOADS.items():
        assert bench.n_sets(wl, n) * n * (bi + bo) > 126e6, wl


def test_reference_arm_line_has_the_contract_keys():
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "py_vollib_vectorized")):
        pytest.skip("baseline/_ref is not installed")
    env = dict(os.environ, NUMBA_CACHE_DIR="/tmp/pvv_numba_cache")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_reference.py"), "--workload", "greeks", "--n", "2000", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd="/tmp")
    d = json.loads(res.stdout.strip().splitlines()[-1])
    assert d["cores"] == 1 and d["public"] > 0 and d["loop_only"] > 0 and d["n"] == 2000, d
