import numpy as np


def bit_equal(a, b):
    """Element-wise float64 equality with NaNs matched by position (the reference's NaN payloads are
    not uniform — SURVEY.md Appendix B — so NaN bits are never compared)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    return (na == nb) & (na | (a == b))


def assert_bit_equal(a, b, what=""):
    ok = bit_equal(a, b)
    if not ok.all():
        i = np.flatnonzero(~ok)
        raise AssertionError(f"{what}: {i.size} of {ok.size} elements differ; first at {i[:5].tolist()}: "
                             f"{np.asarray(a)[i[:5]].tolist()} vs {np.asarray(b)[i[:5]].tolist()}")


def max_ulp_diff(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    m = np.isfinite(a) & np.isfinite(b)
    ia = a[m].view(np.int64)
    ib = b[m].view(np.int64)
    ia = np.where(ia < 0, np.int64(-2 ** 63) - ia, ia)
    ib = np.where(ib < 0, np.int64(-2 ** 63) - ib, ib)
    return int(np.abs(ia - ib).max()) if ia.size else 0
