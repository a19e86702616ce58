"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the evaluation metrics next to the hot path.

Follows /root/reference/biscuit/utils.py:400-464 (`prediction_metrics`) and /root/reference/biscuit/delong.py:5-107
(`compute_midrank`, `fastDeLong`, `delong_roc_variance`; adapted there from Netflix/vmaf, Sun & Xu 2014).  The reference file uses the
``np.float`` alias that NumPy removed in 1.24, so it does not run on this image's NumPy 2.3 as is; restoring the alias
(``np.float = float``, done for the duration of the call in :func:`reference_with_np_float`) runs it unmodified.

Parity status: PINNED -- tests/test_metrics_cpu.py compares this file with the unmodified reference (alias restored) when
/root/reference is present, and with tests/golden/metrics_golden.json (oracle/make_golden_metrics.py)."""
from __future__ import annotations

import contextlib
from statistics import mean, variance

import numpy as np
from scipy import stats


@contextlib.contextmanager
def reference_with_np_float():
    """Temporarily restores the removed ``np.float`` alias so the reference's delong.py runs unmodified."""
    had = hasattr(np, "float")
    if not had:
        np.float = float
    try:
        yield
    finally:
        if not had:
            del np.float


def midrank(x):
    """1-based midranks (delong.py:5-28) -- scipy's 'average' ranking is the same definition."""
    return stats.rankdata(x, method="average").astype(np.float64)


def delong_roc_variance(ground_truth, predictions):
    assert np.array_equal(np.unique(ground_truth), [0, 1])
    order = (-ground_truth).argsort()
    m = int(ground_truth.sum())
    p = predictions[np.newaxis, order]
    n = p.shape[1] - m
    tx, ty, tz = midrank(p[0, :m])[None], midrank(p[0, m:])[None], midrank(p[0])[None]
    aucs = tz[:, :m].sum(axis=1) / m / n - float(m + 1.0) / 2.0 / n
    v01 = (tz[:, :m] - tx[:, :]) / n
    v10 = 1.0 - (tz[:, m:] - ty[:, :]) / m
    return aucs[0], np.cov(v01) / m + np.cov(v10) / n


def delong_roc_test(ground_truth, predictions_one, predictions_two):
    """log10 p-value of the two-classifier DeLong test (delong.py:76-86, 110-123): 1 x 1 array"""
    assert np.array_equal(np.unique(ground_truth), [0, 1])
    order = (-ground_truth).argsort()
    m = int(ground_truth.sum())
    p = np.vstack((predictions_one, predictions_two))[:, order]
    n = p.shape[1] - m
    tx = np.stack([midrank(p[r, :m]) for r in range(2)])
    ty = np.stack([midrank(p[r, m:]) for r in range(2)])
    tz = np.stack([midrank(p[r]) for r in range(2)])
    aucs = tz[:, :m].sum(axis=1) / m / n - float(m + 1.0) / 2.0 / n
    cov = np.cov((tz[:, :m] - tx) / n) / m + np.cov(1.0 - (tz[:, m:] - ty) / m) / n
    lvec = np.array([[1, -1]])
    z = np.abs(np.diff(aucs)) / np.sqrt(np.dot(np.dot(lvec, cov), lvec.T))
    return np.log10(2) + stats.norm.logsf(z, loc=0, scale=1) / np.log(10)


# seeded inputs of the DeLong-test golden (tests/golden/delong_test_golden.json, oracle/make_golden_metrics.py)
DELONG_TEST_CASES = {"f32_400": dict(n=400, seed=11, dtype="float32"), "f64_900_ties": dict(n=900, seed=12, dtype="float64", ties=40),
                     "f32_3000": dict(n=3000, seed=13, dtype="float32")}


def make_delong_test_case(kw):
    rng = np.random.default_rng(kw["seed"])
    y = rng.integers(0, 2, kw["n"]).astype(np.int64)
    a = np.clip(0.5 + 0.15 * (2 * y - 1) + rng.normal(0, 0.3, kw["n"]), 0, 1)
    b = np.clip(0.5 + 0.08 * (2 * y - 1) + rng.normal(0, 0.3, kw["n"]), 0, 1)
    if kw.get("ties"):
        a, b = np.round(a * kw["ties"]) / kw["ties"], np.round(b * kw["ties"]) / kw["ties"]
    return y, a.astype(kw["dtype"]), b.astype(kw["dtype"])


def prediction_metrics(y_true, y_pred, threshold):
    yt = y_true.astype(bool)
    yp = y_pred > threshold
    alpha = 0.05
    z = stats.norm.ppf((1 - alpha / 2))
    tp = np.logical_and(yt, yp).sum(); fp = np.logical_and(~yt, yp).sum()
    tn = np.logical_and(~yt, ~yp).sum(); fn = np.logical_and(yt, ~yp).sum()
    all_jac = []
    for _ in range(500):
        i = np.random.choice(np.arange(yt.shape[0]), size=(150,))
        a, b = yt[i], yp[i]
        _tp = np.logical_and(a, b).sum(); _fp = np.logical_and(~a, b).sum()
        _tn = np.logical_and(~a, ~b).sum(); _fn = np.logical_and(a, ~b).sum()
        all_jac += [((_tn + 0.5 * z**2) / (_tn + _fp + z**2)) - ((_fn + 0.5 * z**2) / (_fn + _tp + z**2))]
    jac, jac_var = mean(all_jac), variance(all_jac)
    if not np.array_equal(np.unique(y_true), [0, 1]):
        ci = [None, None]
    else:
        auc, cov = delong_roc_variance(y_true, y_pred)
        ci = stats.norm.ppf(np.abs(np.array([0, 1]) - alpha / 2), loc=auc, scale=np.sqrt(cov))
        ci[ci > 1] = 1
    sens, spec = tp / (tp + fn), tn / (tn + fp)
    return {"auc_low": ci[0], "auc_high": ci[1], "acc": (tp + tn) / (tp + tn + fp + fn), "sens": sens, "spec": spec,
            "youden": sens + spec - 1, "youden_low": jac - z * np.sqrt(jac_var), "youden_high": jac + z * np.sqrt(jac_var)}


# seeded inputs shared by the golden generator and the tests
CASES = {
    "f64_120": dict(n=120, seed=31, dtype="float64", thr="py:0.5", ties=None),
    "f32_400_ties": dict(n=400, seed=32, dtype="float32", thr="np:0.4375", ties=40),
    "f32_37_pyfloat": dict(n=37, seed=33, dtype="float32", thr="py:0.55", ties=None),
    "f64_1000_single_class_ci": dict(n=1000, seed=34, dtype="float64", thr="py:0.5", ties=None, single=True),
}


def make_case(kw):
    rng = np.random.default_rng(kw["seed"])
    y = rng.integers(0, 2, kw["n"]).astype(np.int64)
    if kw.get("single"):
        y[:] = 1
    p = np.clip(0.5 + 0.18 * (2 * y - 1) + rng.normal(0, 0.25, kw["n"]), 0, 1)
    if kw["ties"]:
        p = np.round(p * kw["ties"]) / kw["ties"]
    p = p.astype(kw["dtype"])
    kind, val = kw["thr"].split(":")
    thr = float(val) if kind == "py" else np.float64(val)
    return y, p, thr
