"""DeLong AUC variance -- host mirror of reference biscuit/delong.py (`delong_roc_variance` 96-107,
`compute_ground_truth_statistics` 89-93, `fastDeLong` 36-73).  The per-example placement values (the
midrank passes of fastDeLong) are computed on the GPU (`bq_delong_placements`, exact half-integer
arithmetic); the covariance is then taken with the same ``np.cov`` call on the same values in the same
order as the reference, so the result is bit-identical."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi


def compute_ground_truth_statistics(ground_truth):
    assert np.array_equal(np.unique(ground_truth), [0, 1])          # delong.py:90
    order = (-ground_truth).argsort()
    label_1_count = int(ground_truth.sum())
    return order, label_1_count


def delong_roc_variance(ground_truth, predictions, ctx=None):
    """(AUC, DeLong variance of the AUC) for one set of predictions; labels must be 0 / 1 with both present."""
    ground_truth = np.asarray(ground_truth)
    predictions = np.asarray(predictions)
    order, m = compute_ground_truth_statistics(ground_truth)
    n_all = int(ground_truth.shape[0])
    n = n_all - m
    if predictions.dtype not in (np.float32, np.float64):
        predictions = predictions.astype(np.float64)
    ctx = ctx or _ffi.default_context()
    scores = np.ascontiguousarray(predictions)
    labels = np.ascontiguousarray(ground_truth != 0, dtype=np.uint8)
    v = np.empty(n_all, np.float64)
    tz_sum = np.zeros(1, np.float64)
    n_pos = C.c_int64()
    code = _ffi.BQ_F32 if scores.dtype == np.float32 else _ffi.BQ_F64
    _ffi.check(ctx.handle, ctx.lib.bq_delong_placements(ctx.handle, _ffi.ptr(scores), C.c_int32(code), _ffi.ptr(labels),
                                                        n_all, _ffi.ptr(v), _ffi.ptr(tz_sum), C.byref(n_pos)),
               "bq_delong_placements")
    v_sorted = v[np.newaxis, order]                       # examples with label 1 first, in the reference's order
    v01, v10 = v_sorted[:, :m], v_sorted[:, m:]
    aucs = tz_sum / m / n - float(m + 1.0) / 2.0 / n       # delong.py:66
    sx = np.cov(v01)
    sy = np.cov(v10)
    delongcov = sx / m + sy / n
    return aucs[0], delongcov
