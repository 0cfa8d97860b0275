"""Installs the UNMODIFIED reference into the git-ignored baseline/_ref/ (TEST / BENCHMARK INFRASTRUCTURE).

    python oracle/install_reference.py            # needs /root/reference (the build container)

baseline/_ref/ is not tracked (the reference's sources never enter the history) but it is not gpurun-ignored
either, so it travels to the GPU box, where
  * bench.py --impl reference / cpu_baseline time the reference's own numba path (oracle/run_reference.py), and
  * tests/test_reference_suite.py runs the reference's own tests/ — unmodified — against the drop-in alias
    package `py_vollib_vectorized` (GPU tier) and against the reference itself on oracle/shim (CPU tier).
The install is the one the task statement prescribes: pip, offline, from a copy (the source tree is read-only
and setuptools writes an egg-info next to setup.py), without dependencies (py_lets_be_rational / py_vollib are
not in the wheelhouse: oracle/shim stands in for them).  The two data files the tests open are copied next to them.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def install(force=False):
    """Returns the install directory, or None when the reference tree is not available (GPU box)."""
    if not os.path.isdir(REF):
        return DST if os.path.isdir(os.path.join(DST, "py_vollib_vectorized")) else None
    marker = os.path.join(DST, ".installed_from")
    if not force and os.path.exists(marker) and os.path.isdir(os.path.join(DST, "py_vollib_vectorized")):
        return DST
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns(".git", "__pycache__"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", DST, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + res.stdout[-2000:] + res.stderr[-2000:])
    tests = os.path.join(DST, "tests")
    os.makedirs(tests, exist_ok=True)
    for name in os.listdir(os.path.join(REF, "tests")):  # pip installs tests/*.py only: the data files and anything missed
        s = os.path.join(REF, "tests", name)
        if os.path.isfile(s) and not os.path.exists(os.path.join(tests, name)):
            shutil.copy(s, os.path.join(tests, name))
    with open(marker, "w") as f:
        f.write(REF + "\n")
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
