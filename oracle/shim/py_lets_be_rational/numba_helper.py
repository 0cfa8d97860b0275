"""jit decorator of the stand-in package (upstream: nopython=True, cache=True, no fastmath)."""
from numba import jit


def maybe_jit(*jit_args, **jit_kwargs):
    jit_kwargs.update({"nopython": True, "cache": True})

    def wrapper(fun):
        return jit(*jit_args, **jit_kwargs)(fun)

    return wrapper
