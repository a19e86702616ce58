"""DeLong AUC variance and the two-classifier DeLong test (reference biscuit/delong.py: `delong_roc_variance` 96-107,
`delong_roc_test` 110-123, `calc_pvalue` 76-86; adapted there from Netflix/vmaf, Sun & Xu 2014).

Sun & Xu's fast algorithm reduces DeLong's structural components to midranks: for classifier r and example i
    V10[r, i] = (rank_all(i) - rank_pos(i)) / n        for the m positives,
    V01[r, j] = 1 - (rank_all(j) - rank_neg(j)) / m    for the n negatives,
and AUC_r = sum_pos rank_all / (m n) - (m + 1) / (2 n).  The midrank passes (exact half-integer arithmetic) run on the
GPU per classifier (`bq_delong_placements`, csrc/metrics.cu); what is left is two `np.cov` calls on k x m and k x n
matrices, taken on the same values in the same order as the reference so the results are bit-identical."""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.stats

from . import _ffi


class _Placements:
    """Placement values of k classifiers on one labelled sample, positives first (the reference's example order)."""

    def __init__(self, ground_truth, predictions, ctx=None):
        labels = np.asarray(ground_truth)
        assert np.array_equal(np.unique(labels), [0, 1])               # delong.py:90
        scores = np.atleast_2d(np.asarray(predictions))
        if scores.dtype not in (np.float32, np.float64):
            scores = scores.astype(np.float64)
        self.m = int(labels.sum())
        self.n = int(labels.shape[0]) - self.m
        positives_first = np.argsort(-labels)                          # same permutation as delong.py:91
        ctx = ctx or _ffi.default_context()
        is_pos = np.ascontiguousarray(labels != 0, dtype=np.uint8)
        code = _ffi.BQ_F32 if scores.dtype == np.float32 else _ffi.BQ_F64
        k, total = scores.shape
        values = np.empty((k, total), np.float64)
        rank_sums = np.zeros(k, np.float64)
        for r in range(k):
            row = np.ascontiguousarray(scores[r])
            out = np.empty(total, np.float64)
            s = np.zeros(1, np.float64)
            n_pos = C.c_int64()
            _ffi.check(ctx.handle, ctx.lib.bq_delong_placements(ctx.handle, _ffi.ptr(row), C.c_int32(code), _ffi.ptr(is_pos),
                                                                total, _ffi.ptr(out), _ffi.ptr(s), C.byref(n_pos)),
                       "bq_delong_placements")
            values[r] = out[positives_first]
            rank_sums[r] = s[0]
        self.v_pos, self.v_neg = values[:, :self.m], values[:, self.m:]
        self.aucs = rank_sums / self.m / self.n - float(self.m + 1.0) / 2.0 / self.n        # delong.py:66

    def covariance(self):
        return np.cov(self.v_pos) / self.m + np.cov(self.v_neg) / self.n                    # delong.py:69-71


def delong_roc_variance(ground_truth, predictions, ctx=None):
    """(AUC, DeLong variance of the AUC) for one set of predictions; labels must be 0 / 1 with both present."""
    p = _Placements(ground_truth, np.asarray(predictions)[np.newaxis, :], ctx=ctx)
    return p.aucs[0], p.covariance()


def delong_roc_test(ground_truth, predictions_one, predictions_two, ctx=None):
    """log10(p-value) of the hypothesis that the two ROC AUCs differ (reference delong.py:76-123), returned as the 1 x 1
    array the reference returns.  Two-sided z test on the AUC difference with the DeLong covariance S of the pair:
    var = [1, -1] S [1, -1]^T, evaluated in the order of the reference's two dot products so the float64 result is the same."""
    placements = _Placements(ground_truth, np.vstack((predictions_one, predictions_two)), ctx=ctx)
    cov = np.asarray(placements.covariance(), dtype=np.float64)
    var_of_difference = (cov[0, 0] - cov[1, 0]) - (cov[0, 1] - cov[1, 1])
    z = np.abs(placements.aucs[1] - placements.aucs[0]) / np.sqrt(var_of_difference)
    log10_p = np.log10(2) + scipy.stats.norm.logsf(z, loc=0, scale=1) / np.log(10)
    return np.array([[log10_p]])
