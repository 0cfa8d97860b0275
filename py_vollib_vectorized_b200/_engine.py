"""ctypes binding of libpvv.so — the ONLY compute path of this package.

There is no CPU fallback: if the shared library is missing the import fails, and if no CUDA device
is present every call raises.  PyTorch is used for device memory, streams and pinned host memory
only (plumbing); the arithmetic lives in csrc/*.cu behind the C ABI declared in include/pvv.h.

Two families of calls:
  *  host path   (`price`, `greeks`, `iv`, `iv_greeks`): numpy arrays (pageable or pinned) in, numpy
     arrays out; the library runs its chunked H2D -> kernel -> D2H pipeline (pvv_*_host).
  *  device path (`price_dev`, ...): torch CUDA tensors in/out, asynchronous on the current torch
     stream (pvv_*_dev).
Columns are (array, stride) pairs: stride 0 marks a broadcast scalar that is never materialised.
"""
import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PVV_LIB: development hook — load an A/B build of the engine (build.py --variant) instead of the in-tree one
LIB_PATH = os.environ.get("PVV_LIB") or os.path.join(_HERE, "libpvv.so")

MODELS = {"black": 0, "black_scholes": 1, "black_scholes_merton": 2}

ST_BELOW_INTRINSIC, ST_ABOVE_MAXIMUM, ST_ZERO_DIVISION, ST_DEFLATER_ZERO = 1, 2, 4, 8

_INT64_MAX = np.iinfo(np.int64).max


class PvvError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA engine has not been built "
            "(run `python py_vollib_vectorized_b200/build.py`). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.pvv_version.restype = i32
    lib.pvv_error_string.restype = ctypes.c_char_p
    lib.pvv_error_string.argtypes = [i32]
    lib.pvv_launch_count.restype = i64
    lib.pvv_ctx_create.argtypes = [i32, i64, i32, ctypes.POINTER(vp)]
    lib.pvv_ctx_destroy.argtypes = [vp]
    lib.pvv_ctx_destroy.restype = None
    # (model, n, flag, 6 cols, stride, outputs..., first_zde[, stream])
    lib.pvv_price_dev.argtypes = [i32, i64] + [vp] * 7 + [vp, vp, vp, vp]
    lib.pvv_greeks_dev.argtypes = [i32, i64] + [vp] * 7 + [vp] + [vp] * 6 + [vp, vp]
    lib.pvv_iv_dev.argtypes = [i32, i64] + [vp] * 7 + [vp, i32, vp, vp, vp, vp]
    lib.pvv_iv_greeks_dev.argtypes = [i32, i64] + [vp] * 7 + [vp, i32, vp, vp] + [vp] * 5 + [vp, vp]
    lib.pvv_price_host.argtypes = [vp, i32, i64] + [vp] * 7 + [vp, vp, vp]
    lib.pvv_greeks_host.argtypes = [vp, i32, i64] + [vp] * 7 + [vp] + [vp] * 6 + [vp]
    lib.pvv_iv_host.argtypes = [vp, i32, i64] + [vp] * 7 + [vp, i32, vp, vp, vp]
    lib.pvv_iv_greeks_host.argtypes = [vp, i32, i64] + [vp] * 7 + [vp, i32, vp, vp] + [vp] * 5 + [vp]
    lib.pvv_analytic_greeks_dev.argtypes = [i32, i64] + [vp] * 7 + [vp] + [vp] * 6 + [vp]
    lib.pvv_analytic_greeks_host.argtypes = [vp, i32, i64] + [vp] * 7 + [vp] + [vp] * 6
    lib.pvv_lut_u8.restype = i64
    lib.pvv_lut_u8.argtypes = [i64, vp, vp, vp, i32]
    lib.pvv_match_u64.restype = i64
    lib.pvv_match_u64.argtypes = [i64, vp, i32, vp, vp, vp, i32]
    lib.pvv_unit_offsets.restype = i32
    lib.pvv_unit_offsets.argtypes = [i64, vp, i32, i32]
    lib.pvv_match_u32.restype = i64
    lib.pvv_match_u32.argtypes = [i64, vp, i32, vp, vp, vp, i32]
    lib.pvv_unary_dev.argtypes = [i32, i64, vp, vp, vp]
    lib.pvv_fp64_probe.argtypes = [i64, i32, i32, ctypes.POINTER(ctypes.c_float), vp]
    if hasattr(lib, "pvv_fp64_ilp_probe"):
        lib.pvv_fp64_ilp_probe.argtypes = [i32, i32, i32, ctypes.POINTER(ctypes.c_float), vp]
    return lib


_lib = _load()


def lib():
    return _lib


def launch_count():
    return int(_lib.pvv_launch_count())


def match_u64(addr, n, keys, vals, out):
    """out[i] = vals[j] where the uint64 at addr + 8 i equals keys[j]; returns the number of misses."""
    k = np.ascontiguousarray(keys, dtype=np.uint64)
    v = np.ascontiguousarray(vals, dtype=np.int8)
    return int(_lib.pvv_match_u64(n, ctypes.c_void_p(addr), k.size, _hp(k), _hp(v), _hp(out), 0))


def match_cells(arr, keys, vals, out):
    """out[i] = vals[j] where the 4- or 8-byte cell arr[i] has the bit pattern keys[j]; returns the number of misses."""
    if arr.dtype.itemsize == 8:
        k = np.ascontiguousarray(keys, dtype=np.uint64)
        fn = _lib.pvv_match_u64
    else:
        k = np.ascontiguousarray(keys, dtype=np.uint32)
        fn = _lib.pvv_match_u32
    v = np.ascontiguousarray(vals, dtype=np.int8)
    return int(fn(arr.size, ctypes.c_void_p(arr.ctypes.data), k.size, _hp(k), _hp(v), _hp(out), 0))


def unit_offsets(offsets):
    """True if every cell delimited by the (int32 / int64) Arrow offsets array is exactly one byte long."""
    return bool(_lib.pvv_unit_offsets(offsets.size - 1, ctypes.c_void_p(offsets.ctypes.data), offsets.dtype.itemsize, 0))


def lut_u8(data, lut, out):
    """out[i] = lut[data[i]] for a uint8 array; returns the number of zeros written."""
    return int(_lib.pvv_lut_u8(data.size, _hp(data), _hp(lut), _hp(out), 0))


def _check(code):
    if code != 0:
        raise PvvError(f"libpvv call failed ({code}): {_lib.pvv_error_string(code).decode()}")


# ------------------------------------------------------------------------------------------------
# host path
# ------------------------------------------------------------------------------------------------
class _Ctx:
    """One pvv_ctx (streams + device slots) per (device, chunk) — created on first use."""

    def __init__(self, device, chunk, nstreams):
        self.handle = ctypes.c_void_p()
        _check(_lib.pvv_ctx_create(device, chunk, nstreams, ctypes.byref(self.handle)))
        self.lock = threading.Lock()

    def close(self):
        if self.handle:
            _lib.pvv_ctx_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_ctxs = {}
_ctx_lock = threading.Lock()


def get_ctx(device=None, chunk=0, nstreams=0):
    if device is None:
        device = int(os.environ.get("PVV_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    key = (device, chunk, nstreams)
    with _ctx_lock:
        c = _ctxs.get(key)
        if c is None:
            c = _ctxs[key] = _Ctx(device, chunk, nstreams)
        return c


def _hp(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


# ---- pinned host memory (plumbing) ------------------------------------------------------------------
# A cudaMemcpyAsync from/to PAGEABLE memory is staged by the driver and blocks (measured on the B200
# box, 10 M contracts, K2: 127 ms pageable vs 12 ms pinned, tools/probe_ingest.py).  So:
#   * outputs are allocated pinned (torch's caching host allocator makes that free after first use) and
#     handed to the caller as ordinary numpy arrays — the DMA lands directly in the returned array;
#   * large pageable input columns are copied once into pinned staging arrays, one thread per column
#     (numpy's copy releases the GIL), before the pipeline starts.
_PIN_MIN_ELEMS = 1 << 15
_pool = None


def _pinned_empty(n, dtype):
    if n >= _PIN_MIN_ELEMS:
        try:
            import torch
            tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.uint8): torch.uint8, np.dtype(np.int8): torch.int8}[np.dtype(dtype)]
            return torch.empty(n, dtype=tdt, pin_memory=True).numpy()
        except Exception:
            pass
    return np.empty(n, dtype=dtype)


def _is_pinned(a):
    if not a.flags.writeable:  # torch.from_numpy refuses read-only views (pandas hands those out); stage them
        return False
    try:
        import torch
        return torch.from_numpy(a).is_pinned()
    except Exception:
        return False


def pin_columns(cols):
    """(array, stride) columns -> the same columns with every large pageable array replaced by a pinned
    copy, so that several engine calls on the same inputs (price_dataframe) stage them only once."""
    arrays = []
    for c in cols:
        if c is None:
            arrays.append(None)
        else:
            a = c[0]
            arrays.append(np.ascontiguousarray(a, dtype=np.int8 if a.dtype == np.int8 else np.float64))
    staged = _stage(arrays)
    return [None if c is None else (staged[i], c[1]) for i, c in enumerate(cols)]


def _stage(arrays):
    """arrays: list of contiguous numpy arrays or None.  Returns the list with every large pageable
    array replaced by a pinned copy."""
    global _pool
    todo = [i for i, a in enumerate(arrays) if a is not None and a.size >= _PIN_MIN_ELEMS and not _is_pinned(a)]
    if not todo:
        return arrays
    out = list(arrays)
    dst = {i: _pinned_empty(arrays[i].size, arrays[i].dtype) for i in todo}
    if len(todo) == 1 or max(arrays[i].nbytes for i in todo) < (4 << 20):  # the int8 flag column comes first: look at the largest
        for i in todo:
            np.copyto(dst[i], arrays[i])
    else:
        if _pool is None:
            from concurrent.futures import ThreadPoolExecutor
            _pool = ThreadPoolExecutor(max_workers=8)
        list(_pool.map(lambda i: np.copyto(dst[i], arrays[i]), todo))
    for i in todo:
        out[i] = dst[i]
    return out


def _hcol(col):
    """(array|None, stride) -> contiguous float64 host array (kept alive by the caller) and stride."""
    if col is None:
        return None, 0
    a, s = col
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, int(s)


def _hflag(col):
    a, s = col
    return np.ascontiguousarray(a, dtype=np.int8), int(s)


def _prepare(flag, cols):
    """Common front of the host entry points: contiguous typed columns, strides, pinned staging."""
    f, sf = _hflag(flag)
    cols = [_hcol(c) for c in cols]
    staged = _stage([f] + [c[0] for c in cols])
    stride = np.array([sf] + [c[1] for c in cols], dtype=np.int64)
    return staged[0], staged[1:], stride


def _raise_zde(first):
    if first >= 0:
        # numba's error model raises for the whole call
        raise ZeroDivisionError("division by zero")


GREEK_NAMES = ("delta", "gamma", "theta", "rho", "vega")

# ---- one process driving several GPUs (drop-in mode) -------------------------------------------------
# PVV_DEVICES="0,1,2,3" (or "all") shards every host-path call over those devices: contiguous ranges
# (keeps strike-sorted chains warp-coherent), one pvv_ctx and one Python thread per device (ctypes
# releases the GIL), each device's pipeline DMA-ing straight into its slice of the shared pinned
# output arrays.  No collective: contracts are independent.
_MULTI_MIN = 1 << 18
_dev_pool = None


def _devices():
    spec = os.environ.get("PVV_DEVICES", "").strip()
    if not spec:
        return None
    if spec == "all":
        import torch
        return list(range(torch.cuda.device_count()))
    return [int(x) for x in spec.split(",") if x.strip() != ""]


def _off(a, lo, stride=1):
    """Pointer to element `lo` of a host array (None stays NULL; broadcast scalars are not advanced)."""
    if a is None:
        return None
    return ctypes.c_void_p(a.ctypes.data + (lo * a.itemsize if stride else 0))


def _run_host(fn, model, n, f, cols, stride, mid, outs, ctx=None):
    """Call one pvv_*_host entry point, sharded over PVV_DEVICES when set.
    mid: extra scalar arguments between stride and the outputs; outs: list of output arrays (or None).
    Returns the first zero-division index (or -1)."""
    def call(c, lo, hi):
        first = ctypes.c_int64(-1)
        args = [c.handle, MODELS[model], hi - lo, _off(f, lo, stride[0])]
        args += [_off(a, lo, stride[1 + k]) for k, a in enumerate(cols)]
        args += [_hp(stride)] + list(mid) + [_off(o, lo) for o in outs]
        if fn is not _lib.pvv_analytic_greeks_host:
            args.append(ctypes.byref(first))
        with c.lock:
            _check(fn(*args))
        return -1 if first.value < 0 else first.value + lo

    devs = _devices() if ctx is None else None
    if not devs or len(devs) < 2 or n < _MULTI_MIN:
        return call(ctx or get_ctx(devs[0] if devs else None), 0, n)
    global _dev_pool
    if _dev_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _dev_pool = ThreadPoolExecutor(max_workers=16)
    g = len(devs)
    futs = [_dev_pool.submit(call, get_ctx(d), n * i // g, n * (i + 1) // g) for i, d in enumerate(devs)]
    firsts = [x for x in (fu.result() for fu in futs) if x >= 0]
    return min(firsts) if firsts else -1


def price(model, n, flag, S, K, t, r, q, sigma, *, ctx=None, out=None, raise_zde=True):
    """Host path of K1.  Columns are (array, stride) pairs; returns a float64 array of n prices."""
    f, cols, stride = _prepare(flag, (S, K, t, r, q, sigma))
    out = _pinned_empty(n, np.float64) if out is None else out
    first = _run_host(_lib.pvv_price_host, model, n, f, cols, stride, (), [out], ctx)
    if raise_zde:
        _raise_zde(first)
    return out


def greeks(model, n, flag, S, K, t, r, q, sigma, *, want=GREEK_NAMES, with_price=False, ctx=None, outs=None, raise_zde=True):
    """Host path of K2.  Returns dict name -> float64 array for the requested outputs."""
    f, cols, stride = _prepare(flag, (S, K, t, r, q, sigma))
    names = (("price",) if with_price else ()) + tuple(want)
    outs = outs or {k: _pinned_empty(n, np.float64) for k in names}
    first = _run_host(_lib.pvv_greeks_host, model, n, f, cols, stride, (), [outs.get(k) for k in ("price",) + GREEK_NAMES], ctx)
    if raise_zde:
        _raise_zde(first)
    return outs


def analytic_greeks(model, n, flag, S, K, t, r, q, sigma, *, want=GREEK_NAMES, with_price=False, ctx=None, outs=None):
    """Host path of the opt-in closed-form greeks kernel (pvv_analytic_greeks_host)."""
    f, cols, stride = _prepare(flag, (S, K, t, r, q, sigma))
    names = (("price",) if with_price else ()) + tuple(want)
    outs = outs or {k: _pinned_empty(n, np.float64) for k in names}
    _run_host(_lib.pvv_analytic_greeks_host, model, n, f, cols, stride, (), [outs.get(k) for k in ("price",) + GREEK_NAMES], ctx)
    return outs


def analytic_greeks_dev(model, n, flag, S, K, t, r, q, sigma, outs):
    torch = _torch()
    f, sf = _dcol(flag, n, torch.int8)
    cols = [_dcol(c, n, torch.float64) for c in (S, K, t, r, q, sigma)]
    stride = np.array([sf] + [c[1] for c in cols], dtype=np.int64)
    ptrs = [_dp(outs.get(k)) for k in ("price",) + GREEK_NAMES]
    _check(_lib.pvv_analytic_greeks_dev(MODELS[model], n, _dp(f), *[_dp(c[0]) for c in cols], _hp(stride), *ptrs, _stream_ptr()))
    return outs


def iv(model, n, flag, price_col, S, K, t, r, q, *, mask_errors=True, ctx=None, out=None, status=None):
    """Host path of K3.  Returns (iv, status, first_zero_division_index)."""
    f, cols, stride = _prepare(flag, (price_col, S, K, t, r, q))
    out = _pinned_empty(n, np.float64) if out is None else out
    status = _pinned_empty(n, np.uint8) if status is None else status
    first = _run_host(_lib.pvv_iv_host, model, n, f, cols, stride, (int(bool(mask_errors)),), [out, status], ctx)
    return out, status, first


def iv_greeks(model, n, flag, price_col, S, K, t, r, q, *, mask_errors=True, want=GREEK_NAMES, ctx=None, outs=None, status=None):
    """Host path of K4.  Returns (dict(iv, greeks...), status, first_zero_division_index)."""
    f, cols, stride = _prepare(flag, (price_col, S, K, t, r, q))
    outs = outs or {k: _pinned_empty(n, np.float64) for k in ("iv",) + tuple(want)}
    status = _pinned_empty(n, np.uint8) if status is None else status
    first = _run_host(_lib.pvv_iv_greeks_host, model, n, f, cols, stride, (int(bool(mask_errors)),),
                      [outs["iv"], status] + [outs.get(k) for k in GREEK_NAMES], ctx)
    return outs, status, first


# ------------------------------------------------------------------------------------------------
# device path (torch CUDA tensors; asynchronous on the current torch stream)
# ------------------------------------------------------------------------------------------------
def _torch():
    import torch
    return torch


def _dp(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _dcol(t, n, dtype):
    torch = _torch()
    if t is None:
        return None, 0
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise ValueError("device columns must be contiguous CUDA tensors of dtype %s" % dtype)
    if t.numel() == n:
        return t, 1
    if t.numel() == 1:
        return t, 0
    raise ValueError("device column has %d elements, expected %d or 1" % (t.numel(), n))


def _stream_ptr():
    torch = _torch()
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def new_zde_word(device):
    torch = _torch()
    return torch.full((1,), _INT64_MAX, dtype=torch.int64, device=device)


def price_dev(model, n, flag, S, K, t, r, q, sigma, out, first_zde=None):
    torch = _torch()
    f, sf = _dcol(flag, n, torch.int8)
    cols = [_dcol(c, n, torch.float64) for c in (S, K, t, r, q, sigma)]
    stride = np.array([sf] + [c[1] for c in cols], dtype=np.int64)
    _check(_lib.pvv_price_dev(MODELS[model], n, _dp(f), *[_dp(c[0]) for c in cols], _hp(stride), _dp(out), _dp(first_zde), _stream_ptr()))
    return out


def greeks_dev(model, n, flag, S, K, t, r, q, sigma, outs, first_zde=None):
    """outs: dict with any of price/delta/gamma/theta/rho/vega -> preallocated CUDA float64 tensors."""
    torch = _torch()
    f, sf = _dcol(flag, n, torch.int8)
    cols = [_dcol(c, n, torch.float64) for c in (S, K, t, r, q, sigma)]
    stride = np.array([sf] + [c[1] for c in cols], dtype=np.int64)
    ptrs = [_dp(outs.get(k)) for k in ("price",) + GREEK_NAMES]
    _check(_lib.pvv_greeks_dev(MODELS[model], n, _dp(f), *[_dp(c[0]) for c in cols], _hp(stride), *ptrs, _dp(first_zde), _stream_ptr()))
    return outs


def iv_dev(model, n, flag, price_col, S, K, t, r, q, out, status, mask_errors=True, first_zde=None):
    torch = _torch()
    f, sf = _dcol(flag, n, torch.int8)
    cols = [_dcol(c, n, torch.float64) for c in (price_col, S, K, t, r, q)]
    stride = np.array([sf] + [c[1] for c in cols], dtype=np.int64)
    _check(_lib.pvv_iv_dev(MODELS[model], n, _dp(f), *[_dp(c[0]) for c in cols], _hp(stride), int(bool(mask_errors)), _dp(out), _dp(status),
                           _dp(first_zde), _stream_ptr()))
    return out, status


def iv_greeks_dev(model, n, flag, price_col, S, K, t, r, q, iv_out, status, outs, mask_errors=True, first_zde=None):
    torch = _torch()
    f, sf = _dcol(flag, n, torch.int8)
    cols = [_dcol(c, n, torch.float64) for c in (price_col, S, K, t, r, q)]
    stride = np.array([sf] + [c[1] for c in cols], dtype=np.int64)
    ptrs = [_dp(outs.get(k)) for k in GREEK_NAMES]
    _check(_lib.pvv_iv_greeks_dev(MODELS[model], n, _dp(f), *[_dp(c[0]) for c in cols], _hp(stride), int(bool(mask_errors)), _dp(iv_out),
                                  _dp(status), *ptrs, _dp(first_zde), _stream_ptr()))
    return iv_out, status, outs


UNARY = {"exp": 0, "log": 1, "erfcx": 2, "erfc": 3, "norm_cdf": 4, "inverse_norm_cdf": 5, "cbrt_pow": 6, "sqrt_bounded": 7,
         "div_bounded_pairs": 8}


def unary_dev(which, x):
    """Diagnostic: evaluates one of the device scalar functions over a CUDA float64 tensor."""
    torch = _torch()
    y = torch.empty_like(x)
    _check(_lib.pvv_unary_dev(UNARY[which], x.numel(), _dp(x), _dp(y), _stream_ptr()))
    return y


def fp64_probe(threads=148 * 2048, iters=4096, use_fma=False):
    """Measured FP64 issue rate in DP instructions/s (dependent DMUL+DADD or DFMA chains)."""
    ms = ctypes.c_float(0)
    _check(_lib.pvv_fp64_probe(threads, iters, int(use_fma), ctypes.byref(ms), _stream_ptr()))
    instr = float(threads) * iters * 16
    return instr / (ms.value * 1e-3)


def fp64_ilp_probe(chains, warps_per_sm, iters=4096, sms=148):
    """DP instructions/s of `chains` dependent DMUL+DADD chains per thread at `warps_per_sm` resident warps."""
    ms = ctypes.c_float(0)
    _check(_lib.pvv_fp64_ilp_probe(int(chains), int(warps_per_sm), int(iters), ctypes.byref(ms), _stream_ptr()))
    return float(sms) * warps_per_sm * 32 * iters * 16 / (ms.value * 1e-3)
