"""Helpers shared by the parity tests."""
import numpy as np


def to_fe_params(P):
    """oracle Params -> product Params (same fields, separate ctypes classes)."""
    from feature_extraction_b200 import node_default
    q = node_default()
    for f, _ in q._fields_:
        setattr(q, f, getattr(P, f))
    return q


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def rel_err(a, b):
    """element-wise |a-b| / |b| with exact matches (incl. NaN==NaN, 0==0) counted as 0."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    e = np.where(same, 0.0, e)
    return np.where(np.isnan(e), np.inf, e)


# Descriptor tolerance of BASELINE.json:north_star: 1e-5 relative.  No whitelist: the device evaluates
# atan2f / acosf exactly like the host libm (csrc/glibc_f32.h), so a neighbour on a bin edge lands in the
# same bin on both sides and every row must meet the bar.
DESC_RTOL = 1e-5


def check_descriptors(d_gpu, d_ref):
    """-> (n_ok, n_bad): descriptor rows within / beyond 1e-5 relative of the oracle's."""
    assert d_gpu.shape == d_ref.shape
    if len(d_gpu) == 0:
        return 0, 0
    e = rel_err(d_gpu, d_ref).max(axis=1)
    ok = e <= DESC_RTOL
    return int(ok.sum()), int((~ok).sum())
