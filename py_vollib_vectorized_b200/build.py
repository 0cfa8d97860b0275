"""Builds libpvv.so (the CUDA engine + C ABI) in-tree with nvcc for sm_100a.

    python py_vollib_vectorized_b200/build.py [--force] [--verbose]

-fmad=false is part of the numerical contract (see csrc/pvv_math.cuh): the reference's numba loops
never contract a*b+c, and bit-identical stencil prices are what makes the finite-difference greeks
match.  Explicit fma() calls (glibc's exp/log) are unaffected by the flag.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpvv.so")
# (source, extra flags): the parity kernels are compiled WITHOUT fmad, the opt-in analytical kernel with
UNITS = [(os.path.join(CSRC, "pvv_kernels.cu"), ["-fmad=false"]),
         (os.path.join(CSRC, "pvv_analytic.cu"), ["-fmad=true"])]
SOURCES = [u[0] for u in UNITS]
DEPS = SOURCES + [os.path.join(CSRC, "pvv_math.cuh"), os.path.join(CSRC, "pvv_tile.cuh"), os.path.join(CSRC, "glibc_tables.cuh"),
                  os.path.join(os.path.dirname(HERE), "include", "pvv.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA engine cannot be built (there is no CPU fallback)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_extension(force=False, verbose=False, variant=None, defines=(), extra_flags=()):
    """variant/defines/extra_flags: A/B builds for kernel experiments -> build/libpvv_<variant>.so (loaded with PVV_LIB=...)."""
    if variant is None and not force and not needs_build():
        return LIB
    # the image's default CC wrapper lacks some specs; the system g++ is the host compiler
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    objdir = os.path.join(HERE, "build") if variant is None else os.path.join(HERE, "build", variant)
    os.makedirs(objdir, exist_ok=True)
    objs = []
    lib = LIB if variant is None else os.path.join(HERE, "build", f"libpvv_{variant}.so")
    defs = [f"-D{d}" for d in defines] + list(extra_flags)

    def run(cmd):
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed building libpvv.so")

    for src, extra in UNITS:
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        run([_nvcc()] + ccbin + NVCC_FLAGS + extra + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src])
        objs.append(obj)
    run([_nvcc()] + ccbin + ["-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "shared", "-o", lib] + objs)
    return lib


if __name__ == "__main__":
    # python build.py [--force] [--verbose] [--variant NAME -DFOO -DBAR=1 ... --nvcc-flags "-Xptxas -dlcm=ca"]
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    xf = sys.argv[sys.argv.index("--nvcc-flags") + 1].split() if "--nvcc-flags" in sys.argv else []
    print(build_extension(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=var,
                          defines=[a[2:] for a in sys.argv if a.startswith("-D")], extra_flags=xf))
