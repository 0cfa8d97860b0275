"""Host-side data plumbing: the reference's `util/data_format.py` re-designed for one SoA upload.

Same accepted inputs, same exceptions and warning texts as the reference
(/root/reference/py_vollib_vectorized/util/data_format.py:8-100), but
  * flags become an int8 column with a vectorised encode instead of a Python loop over a dict
    (data_format.py:10-11 is >50 % of the reference's wall time at 1 M rows);
  * broadcasting never materialises a scalar: a column is an (array, stride) pair, stride 0 meaning
    "one value for every contract" (the reference copies every scalar out to full length,
    data_format.py:29-31);
  * the below-intrinsic / above-maximum masks come back from the IV kernel as a status byte per
    contract instead of ~10 numpy temporaries (data_format.py:42-72).
"""
import ctypes
import warnings

import numpy as np
import pandas as pd

try:  # the real py_vollib if the user has it, otherwise the names the reference imports from it
    from py_vollib.helpers.constants import FLOAT_MAX, MINUS_FLOAT_MAX
    from py_vollib.helpers.exceptions import PriceIsAboveMaximum, PriceIsBelowIntrinsic
except ImportError:  # pragma: no cover - depends on the environment
    from .._py_vollib_stub import FLOAT_MAX, MINUS_FLOAT_MAX, PriceIsAboveMaximum, PriceIsBelowIntrinsic

binary_flag = {'c': 1, 'p': -1, 1: 1, -1: -1, '1': 1, '-1': -1}


def _preprocess_flags(flags, dtype=np.float64):
    """Flags -> int8 array of +1/-1 (data_format.py:10-11 semantics: anything iterable whose items
    are keys of `binary_flag`; a bare int raises TypeError, an unknown item KeyError)."""
    if isinstance(flags, np.ndarray) and flags.dtype == np.int8 and flags.ndim == 1:
        bad = (flags != 1) & (flags != -1)
        if bad.any():
            raise KeyError(flags[bad][0].item())
        return flags
    if isinstance(flags, str):
        seq = list(flags)  # the reference iterates a bare string character by character
    elif isinstance(flags, (pd.Series, pd.Index)):
        if isinstance(flags.dtype, pd.StringDtype):
            return _encode_string_series(pd.Series(flags))
        seq = flags.to_numpy()
    elif isinstance(flags, pd.DataFrame):
        seq = list(flags)  # iterating a frame yields its column labels, as in the reference
    else:
        try:
            iter(flags)
        except TypeError:
            raise TypeError(f"'{type(flags).__name__}' object is not iterable")
        seq = flags
    arr = seq if isinstance(seq, np.ndarray) else np.asarray(list(seq) if not isinstance(seq, (list, tuple)) else seq, dtype=None)
    if arr.ndim != 1:
        raise TypeError(f"unhashable type: '{type(arr[0]).__name__}'")
    if arr.size == 0:
        return np.empty(0, dtype=np.int8)
    kind = arr.dtype.kind
    if arr.flags.c_contiguous and arr.dtype.itemsize in (4, 8) and arr.dtype.isnative and (kind in "iuf" or arr.dtype == np.dtype("<U1")):
        # one threaded C pass over the 4- or 8-byte cells (bit patterns of the accepted values) instead of
        # numpy's compare / mask-assign passes (1.0 ms of a 1.6 ms get_all_greeks call at 100 k rows)
        from .. import _engine
        bits = arr.dtype.itemsize * 8
        if kind == "U":
            keys, vals = [ord('c'), ord('1'), ord('p')], [1, 1, -1]
        elif kind == "f":
            one = np.array([1.0, -1.0], dtype=arr.dtype).view(f"u{arr.dtype.itemsize}")
            keys, vals = [int(one[0]), int(one[1])], [1, -1]
        else:
            keys, vals = [1, (1 << bits) - 1], [1, -1]
            if kind == "u":
                keys, vals = [1], [1]
        out = np.empty(arr.shape, dtype=np.int8)
        if _engine.match_cells(arr, keys, vals, out):
            raise KeyError(arr[out == 0][0].item())
        return out
    if kind in "iufb":
        out = np.zeros(arr.shape, dtype=np.int8)
        out[arr == 1] = 1
        out[arr == -1] = -1
        bad = out == 0
        if bad.any():
            raise KeyError(arr[bad][0].item())
        return out
    if kind == "U" and arr.dtype.itemsize == 4 and arr.flags.c_contiguous:  # '<U1' in a non-native layout: compare code points
        cp = arr.view(np.uint32)
        out = np.zeros(arr.shape, dtype=np.int8)
        out[(cp == ord('c')) | (cp == ord('1'))] = 1
        out[cp == ord('p')] = -1
        bad = out == 0
        if bad.any():
            raise KeyError(arr[bad][0].item())
        return out
    if kind in "US":
        a = arr.astype(str)
        out = np.zeros(a.shape, dtype=np.int8)
        out[(a == 'c') | (a == '1')] = 1
        out[(a == 'p') | (a == '-1')] = -1
        bad = out == 0
        if bad.any():
            raise KeyError(a[bad][0].item())
        return out
    # object column (the usual pandas string column, possibly mixed)
    return _encode_object_flags(arr)


def _singleton_ids():
    """ids of the objects a flag cell usually IS: CPython keeps one object per one-character Latin-1
    string (what numpy / pandas / csv readers produce), a second one for the interned literal, and
    one per small int."""
    ids = {}
    for key in ('c', 'p', '1'):
        ids[id(key)] = binary_flag[key]
        ids[id(chr(ord(key)))] = binary_flag[key]
    ids[id(1)] = 1
    ids[id(-1)] = -1
    return ids


def _encode_object_flags(arr):
    """Object array -> int8 flags: compare the PyObject* values vectorised (8 bytes per element) and
    send only the stragglers (e.g. '-1', float 1.0, freshly built strings) through the reference's dict."""
    from .. import _engine
    arr = np.ascontiguousarray(arr)
    n = arr.size
    ids = _singleton_ids()
    out = np.empty(n, dtype=np.int8)
    missing = _engine.match_u64(arr.ctypes.data, n, list(ids.keys()), list(ids.values()), out)
    if missing:
        rest = np.flatnonzero(out == 0)
        out[rest] = [binary_flag[f] for f in arr[rest]]
    return out


def _encode_string_series(s):
    """pandas string-dtype column (Arrow-backed in pandas 3).  Fast path: when every cell is one byte
    long the Arrow data buffer IS the flag column — compare bytes.  Otherwise vectorised equality."""
    n = len(s)
    try:
        import pyarrow as pa
        pa_arr = getattr(s.array, "_pa_array", None)
        if pa_arr is not None and n:
            if isinstance(pa_arr, pa.ChunkedArray):
                ca = pa_arr.chunk(0) if pa_arr.num_chunks == 1 else pa_arr.combine_chunks()
            else:
                ca = pa_arr
            if ca.null_count == 0 and (pa.types.is_large_string(ca.type) or pa.types.is_string(ca.type)):
                bufs = ca.buffers()
                off_dtype = np.int64 if pa.types.is_large_string(ca.type) else np.int32
                offsets = np.frombuffer(bufs[1], dtype=off_dtype)[ca.offset:ca.offset + n + 1]
                from .. import _engine
                if int(offsets[-1]) - int(offsets[0]) == n and _engine.unit_offsets(np.ascontiguousarray(offsets)):
                    data = np.frombuffer(bufs[2], dtype=np.uint8)[int(offsets[0]):int(offsets[-1])]
                    lut = np.zeros(256, dtype=np.int8)
                    lut[ord('c')] = lut[ord('1')] = 1
                    lut[ord('p')] = -1
                    out = np.empty(n, dtype=np.int8)
                    if _engine.lut_u8(np.ascontiguousarray(data), lut, out):
                        raise KeyError(s.iloc[int(np.flatnonzero(out == 0)[0])])
                    return out
    except (ImportError, AttributeError, TypeError):
        pass
    out = np.zeros(n, dtype=np.int8)
    for key in ('c', 'p', '1', '-1'):
        out[(s == key).to_numpy(dtype=bool, na_value=False)] = binary_flag[key]
    bad = out == 0
    if bad.any():
        raise KeyError(s.iloc[int(np.flatnonzero(bad)[0])])
    return out


def _maybe_format_data(data, dtype):
    """One input -> 1-D array of `dtype` (data_format.py:14-26: same accepted types, same errors)."""
    if isinstance(data, (int, float)):
        return np.array([data], dtype=dtype)
    elif isinstance(data, pd.Series):
        return np.ravel(data.values.astype(dtype, copy=False))
    elif isinstance(data, pd.DataFrame):
        assert data.shape[1] == 1, "You passed a `pandas.DataFrame` object that contains more than (1) column!"
        return np.ravel(data.values.astype(dtype, copy=False))
    elif isinstance(data, (list, tuple)):
        return np.ravel(np.array(data, dtype=dtype))
    elif isinstance(data, np.ndarray):
        return np.ravel(data.astype(dtype, copy=False))
    raise ValueError(f"Data type {type(data)} unsupported, must be in: list, tuple, np.array, pd.Series.")


class Columns:
    """The broadcast result: n contracts, each column an (array, stride) pair."""

    def __init__(self, n, cols):
        self.n = n
        self.cols = cols

    def __getitem__(self, i):
        return self.cols[i]

    def __iter__(self):
        return iter(self.cols)

    def dense(self, i):
        """Materialise one column to full length (only where a caller really needs an array)."""
        a, s = self.cols[i]
        return a if s == 1 or self.n == 1 else np.full(self.n, a[0], dtype=a.dtype)


def maybe_format_data_and_broadcast(*all_data, dtype=np.float64):
    """data_format.py:29-31 without the copies.  int8 flag arrays pass through untouched."""
    arrays = []
    for d in all_data:
        if isinstance(d, tuple) and len(d) == 2 and isinstance(d[0], np.ndarray) and isinstance(d[1], int):
            arrays.append(d[0] if d[1] == 1 or d[0].size != 1 else d[0][:1])  # already an (array, stride) column
        elif isinstance(d, np.ndarray) and d.dtype == np.int8:
            arrays.append(np.ravel(d))
        else:
            arrays.append(_maybe_format_data(d, dtype=dtype))
    # numpy's own broadcasting rule (and its ValueError on a shape mismatch)
    n = np.broadcast_shapes(*[a.shape for a in arrays])[0]
    cols = []
    for a in arrays:
        if a.size == n:
            cols.append((a, 1))
        else:  # size-1 column broadcast over n > 1 contracts
            cols.append((a, 0))
    return Columns(n, cols)


def _validate_data(*all_data):
    """Kept for API compatibility (data_format.py:34-39); broadcasting already guarantees it."""
    number_elems = [len(x) for x in all_data]
    if not all(x == number_elems[0] for x in number_elems):
        raise ValueError(
            f"All input values must contain the same number of elements. Found number of elements = {number_elems}")


def report_iv_status(status, on_error, deflater_checked=True):
    """Turn the per-contract status bytes of the IV kernel into the reference's warnings and
    exceptions, in the reference's order: deflater == 0 (implied_volatility.py:50-57), below
    intrinsic, above maximum (data_format.py:54-70), then numba's ZeroDivisionError."""
    if on_error != "ignore":
        if (status & 8).any():
            if on_error == "warn":
                warnings.warn(
                    "Unexpected value encountered in Interest Free Rate (r) or Annualized time to expiration (t). Are you sure the time to expiration is annualized?",
                    stacklevel=3)
            elif on_error == "raise":
                raise ValueError(
                    "Unexpected value encountered in Interest Free Rate (r) or Annualized time to expiration (t).")
        below = np.flatnonzero(status & 1)
        if below.size:
            if on_error == "warn":
                warnings.warn("Found Below Intrinsic contracts at index {}".format(below.tolist()), stacklevel=2)
            elif on_error == "raise":
                raise PriceIsBelowIntrinsic
        above = np.flatnonzero(status & 2)
        if above.size:
            if on_error == "warn":
                warnings.warn("Found Above Maximum Price contracts at index {}".format(above.tolist()), stacklevel=2)
            elif on_error == "raise":
                raise PriceIsAboveMaximum
    if (status & 4).any():
        raise ZeroDivisionError("division by zero")


def _validate_df_col(col, df):
    if not isinstance(col, str):
        raise ValueError(f"Column {col} must be a string!")
    elif col not in df:
        raise ValueError(f"Column {col} not found in dataframe!")
