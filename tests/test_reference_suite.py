"""The reference's OWN test files (tests/test_entrypoints.py, tests/test_monkeypatch.py), unmodified, as installed
under the git-ignored baseline/_ref/ by oracle/install_reference.py (the drop-in proof).

  CPU tier  — against the reference itself on oracle/shim: pins the shim (the oracle's oracle) to the reference's suite;
  GPU tier  — against the drop-in alias package `py_vollib_vectorized` of this repo (the B200 engine under the
              reference's name): a user of the reference switches packages and their tests still pass.

Each run is a fresh interpreter with its own PYTHONPATH; --import-mode=importlib keeps pytest from putting
baseline/_ref (which holds the reference under the same package name) on sys.path."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
REF_TESTS = os.path.join(REF, "tests")
SHIM = os.path.join(ROOT, "oracle", "shim")


def _run(pythonpath):
    if not os.path.exists(os.path.join(REF_TESTS, "test_entrypoints.py")):
        pytest.skip("baseline/_ref is not installed (python oracle/install_reference.py in the build container)")
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    env.update(PYTHONPATH=os.pathsep.join(pythonpath), NUMBA_CACHE_DIR="/tmp/pvv_numba_cache", PYTHONDONTWRITEBYTECODE="1")
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--import-mode=importlib", "-W", "ignore",
                          "test_entrypoints.py", "test_monkeypatch.py"], cwd=REF_TESTS, env=env, capture_output=True, text=True, timeout=900)
    return res


def test_reference_suite_passes_on_the_reference_with_the_shim():
    res = _run([REF, SHIM])
    assert res.returncode == 0 and " passed" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]
    assert "9 passed" in res.stdout, res.stdout[-500:]


@pytest.mark.gpu
def test_reference_suite_passes_unmodified_on_the_drop_in_alias():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    # the alias package first; oracle/shim only supplies the absent third-party `py_vollib` the tests import
    res = _run([ROOT, SHIM])
    assert res.returncode == 0 and "9 passed" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]
    # and it really was the engine that ran
    probe = subprocess.run([sys.executable, "-c", "import py_vollib_vectorized as m, sys; print(m.__file__); print(m.vectorized_black_scholes.__module__)"],
                           cwd=REF_TESTS, env=dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, SHIM])), capture_output=True, text=True)
    assert os.path.join(ROOT, "py_vollib_vectorized", "__init__.py") in probe.stdout, probe.stdout + probe.stderr
