"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's UQ-thresholding algorithm.

Follows /root/reference/biscuit/threshold.py (process_tile_predictions 125-177,
process_group_predictions 180-245, apply 248-361, detect 364-475, from_cv 478-557) and
biscuit/utils.py (auc 487-504).  Like the reference it delegates the two numerically
delicate primitives to the INSTALLED third-party libraries the reference itself calls:
``sklearn.metrics.roc_curve / auc`` (sklearn 1.9.0, metrics/_ranking.py) and
``DataFrame.groupby().mean()`` (pandas 3.0.2, row-order Kahan sum in the column dtype), so the
oracle's arithmetic is the reference's arithmetic.  Plotting is omitted (out of scope).

Parity status: PINNED.  `tests/test_oracle_pinning.py` compares every function here with the
unmodified reference executed through `oracle/ref_shim.py` (when /root/reference is present)
and with the committed outputs of that reference in `tests/golden/threshold_golden.json`.

A second, library-free tier (`roc_points`, `youden`, `kahan_group_mean`, `pairwise_sum`)
restates the sklearn / pandas / numpy algorithms themselves in plain numpy + python loops; it
documents exactly what the CUDA kernels implement and is pinned against the libraries in the
same test file.
"""
from __future__ import annotations

import warnings

import numpy as np
import pandas as pd
from sklearn import metrics
from sklearn.exceptions import UndefinedMetricWarning


class ThresholdError(Exception):      # reference biscuit/errors.py:17
    pass


class ROCFailedError(Exception):      # reference biscuit/errors.py:21
    pass


class PredsContainNaNError(Exception):  # reference biscuit/errors.py:25
    pass


# ----------------------------------------------------------------------------------------
# tier 1: restatement of biscuit/threshold.py on top of sklearn + pandas
# ----------------------------------------------------------------------------------------

def _roc(y_true, y_score):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=UndefinedMetricWarning)
        return metrics.roc_curve(y_true, y_score)


def _youden_pick(fpr, tpr, thr):
    """The reference's Youden idiom (threshold.py:151-152, 219-220, 423-424, 455-456):
    first maximum of tpr-fpr by python `max`, then `.index()` of that (tpr,fpr) pair.
    Raises ValueError when the pair holds NaN (single-class labels)."""
    pts = list(zip(tpr, fpr))
    best = max(pts, key=lambda q: q[0] - q[1])
    return thr[list(zip(tpr, fpr)).index(best)]


def auc(y_true, y_pred):
    """biscuit/utils.py:487-504 -- ROC AUC, NaN when sklearn raises ValueError."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=UndefinedMetricWarning)
        try:
            fpr, tpr, _ = metrics.roc_curve(y_true, y_pred)
            return metrics.auc(fpr, tpr)
        except ValueError:
            return np.nan


def process_tile_predictions(df, pred_thresh=0.5, patients=None):
    """threshold.py:125-177.  Mutates `df` (adds error/correct/incorrect/y_pred_bin[/patient])."""
    yp = df["y_pred"].to_numpy()
    if np.isnan(yp).sum():                                            # :141-142
        raise PredsContainNaNError
    fpr, tpr, thr = _roc(df["y_true"].to_numpy(), yp)                 # :145-148
    try:
        opt = _youden_pick(fpr, tpr, thr)                             # :151-152
    except ValueError:
        opt = 0.5                                                     # :153-155
    if isinstance(pred_thresh, str) and pred_thresh == "detect":      # :157-159
        pred_thresh = opt
    else:
        f"{pred_thresh:.4f}"                                          # :161 (TypeError on None)
    if patients is not None:                                          # :163-164
        df["patient"] = df["slide"].map(patients)
    df["error"] = abs(df["y_true"] - df["y_pred"])                    # :170
    lo = df["y_pred"] < pred_thresh
    hi = df["y_pred"] >= pred_thresh
    df["correct"] = (lo & (df["y_true"] == 0)) | (hi & (df["y_true"] == 1))  # :171-174
    df["incorrect"] = (~df["correct"]).astype(int)                    # :175
    df["y_pred_bin"] = hi.astype(int)                                 # :176
    return df, pred_thresh


def process_group_predictions(df, pred_thresh, level):
    """threshold.py:180-245.  One row per slide/patient in first-appearance order."""
    for c in ("y_true", "y_pred", "uncertainty"):                     # :184-186
        if c not in df.columns:
            raise UnboundLocalError("reference raises UnboundLocalError here (App. A.8)")
    levels = [v for v in pd.unique(df[level]) if v is not np.nan]     # :190
    means = df[[level, "y_pred", "y_true", "uncertainty"]].groupby(level, as_index=False).mean()
    means = means.set_index(level)                                    # :191-192 + lookups 193-204
    yp = np.array([means.at[v, "y_pred"] for v in levels])
    yt = np.array([means.at[v, "y_true"] for v in levels], dtype=np.uint8)  # :197-200 truncation
    u = np.array([means.at[v, "uncertainty"] for v in levels])
    if not len(yt):                                                   # :205-206
        raise ROCFailedError("Unable to generate ROC; preds are empty.")
    fpr, tpr, thr = _roc(yt, yp)                                      # :212
    metrics.auc(fpr, tpr)                                             # :214
    if isinstance(pred_thresh, str) and pred_thresh == "detect":      # :217-223
        try:
            pred_thresh = _youden_pick(fpr, tpr, thr)
        except ValueError:
            raise ROCFailedError(f"Unable to generate {level}-level ROC")
    else:
        f"{pred_thresh:.4f}"                                          # :225
    lo, hi = yp < pred_thresh, yp >= pred_thresh
    out = pd.DataFrame({                                              # :235-244
        level: pd.Series(levels),
        "error": pd.Series(abs(yt - yp)),
        "uncertainty": pd.Series(u),
        "correct": (lo & (yt == 0)) | (hi & (yt == 1)),
        "incorrect": pd.Series((lo & (yt == 1)) | (hi & (yt == 0))).astype(int),
        "y_true": pd.Series(yt),
        "y_pred": pd.Series(yp),
        "y_pred_bin": pd.Series(hi).astype(int),
    })
    return out, pred_thresh


_EMPTY_RESULTS = ("auc", "percent_incl", "acc", "sensitivity", "specificity")


def apply(df, tile_uq, slide_uq, tile_pred=0.5, slide_pred=0.5, plot=False,
          keep="high_confidence", title=None, patients=None, level="slide"):
    """threshold.py:248-361."""
    assert keep in ("high_confidence", "low_confidence")             # :281
    assert not (level == "patient" and patients is None)             # :282
    f"{tile_uq:.5f}"                                                  # :284 (TypeError on None)
    if patients:                                                      # :285-286
        df["patient"] = df["slide"].map(patients)
    pd.unique(df[level])                                              # :287 (KeyError if missing)
    df, _ = process_tile_predictions(df, pred_thresh=tile_pred, patients=patients)  # :290-294
    n_before = pd.unique(df[level]).shape[0]                          # :295
    if tile_uq:                                                       # :297-298
        df = df[df["uncertainty"] < tile_uq]
    try:
        s_df, _ = process_group_predictions(df, pred_thresh=slide_pred, level=level)  # :305-309
    except ROCFailedError:
        return {k: None for k in _EMPTY_RESULTS}, None                # :310-317
    if slide_uq:                                                      # :323-330
        f"{slide_uq:.5f}"
        if keep == "high_confidence":
            s_df = s_df.loc[s_df["uncertainty"] < slide_uq]
        else:
            s_df = s_df.loc[s_df["uncertainty"] >= slide_uq]
    area = auc(s_df["y_true"].to_numpy(), s_df["y_pred"].to_numpy())  # :333
    pct = len(s_df) / n_before                                        # :334-335
    t = s_df["y_true"].to_numpy().astype(bool)                        # :339
    p = s_df["y_pred"].to_numpy() > slide_pred                        # :340 (strict)
    tp = np.logical_and(t, p).sum()
    fp = np.logical_and(~t, p).sum()
    tn = np.logical_and(~t, ~p).sum()
    fn = np.logical_and(t, ~p).sum()
    with np.errstate(invalid="ignore", divide="ignore"):
        res = {"auc": area, "percent_incl": pct,
               "acc": (tp + tn) / (tp + tn + fp + fn),                # :346
               "sensitivity": tp / (tp + fn),                         # :347
               "specificity": tn / (tn + fp)}                         # :348
    return res, s_df


def detect(df, tile_uq="detect", slide_uq="detect", tile_pred="detect", slide_pred="detect",
           plot=False, patients=None):
    """threshold.py:364-475."""
    none4 = {k: None for k in ("tile_uq", "slide_uq", "tile_pred", "slide_pred")}
    try:
        df, found_tile_pred = process_tile_predictions(df, pred_thresh=tile_pred,
                                                       patients=patients)  # :398-402
    except PredsContainNaNError:
        return none4, None                                            # :403-405
    if isinstance(tile_pred, str) and tile_pred == "detect":          # :407-408
        tile_pred = found_tile_pred
    if isinstance(tile_uq, (float, np.float16, np.float32, np.float64)):  # :411-412
        df = df[df["uncertainty"] < tile_uq]
    elif not (isinstance(tile_uq, str) and tile_uq == "detect"):      # :413-415
        tile_uq = None
    else:                                                             # :416-426
        fpr, tpr, thr = _roc(df["incorrect"].to_numpy(), df["uncertainty"].to_numpy())
        tile_uq = _youden_pick(fpr, tpr, thr)       # ValueError propagates (App. A.1)
        df = df[df["uncertainty"] < tile_uq]
    try:
        s_df, slide_pred = process_group_predictions(df, pred_thresh=slide_pred,
                                                     level="slide")  # :433-438
    except ROCFailedError:
        return none4, None                                            # :439-441
    if isinstance(slide_uq, str) and slide_uq == "detect":            # :444-460
        if not s_df["incorrect"].to_numpy().sum():
            slide_uq = None
        else:
            fpr, tpr, thr = _roc(s_df["incorrect"], s_df["uncertainty"].to_numpy())
            slide_uq = _youden_pick(fpr, tpr, thr)
            s_df = s_df[s_df["uncertainty"] < slide_uq]
    else:
        slide_uq = 0.5                                                # :461-463
    area = auc(s_df["y_true"].to_numpy(), s_df["y_pred"].to_numpy())  # :468
    return {"tile_uq": tile_uq, "slide_uq": slide_uq,
            "tile_pred": tile_pred, "slide_pred": slide_pred}, area


def from_cv(dfs, **kwargs):
    """threshold.py:478-557: tile_uq=min, slide_uq=max, tile_pred/slide_pred=mean over folds."""
    need = ("y_true", "y_pred", "uncertainty", "slide", "patient")
    skip_tile = "tile_uq_thresh" in kwargs and kwargs["tile_uq_thresh"] is None     # :513-516
    skip_slide = "slide_uq_thresh" in kwargs and kwargs["slide_uq_thresh"] is None
    t_uq, s_uq, t_pred, s_pred = [], [], [], []
    for df in dfs:
        if not all(c in df.columns for c in need):                    # :520-524
            raise ValueError(f"DataFrame missing columns, expected {need}, got: "
                             f"{', '.join(df.columns.tolist())}")
        th, _ = detect(df, **kwargs)                                  # :525
        if th["tile_uq"] is None or th["slide_uq"] is None:           # :526-528
            continue
        t_pred.append(th["tile_pred"])
        s_pred.append(th["slide_pred"])
        if not skip_tile:
            t_uq.append(th["tile_uq"])
        if not skip_slide:
            s_uq.append(th["slide_uq"])
    if not skip_tile and not len(t_uq):                               # :539-542
        raise ThresholdError("Unable to detect tile UQ threshold.")
    if not skip_slide and not len(s_uq):
        raise ThresholdError("Unable to detect slide UQ threshold.")
    return {"tile_uq": np.min(t_uq) if not skip_tile else t_uq,       # :544-557
            "slide_uq": np.max(s_uq) if not skip_slide else s_uq,
            "tile_pred": np.mean(t_pred),
            "slide_pred": np.mean(s_pred)}


# ----------------------------------------------------------------------------------------
# tier 2: library-free restatement of the third-party primitives (what the kernels implement)
# ----------------------------------------------------------------------------------------

def roc_points(label, score, drop_intermediate=True):
    """sklearn 1.9.0 metrics/_ranking.py: _sort_inputs_and_compute_classification_thresholds
    (878-921), confusion_matrix_at_thresholds (1020-1043), roc_curve (1317-1372).

    Returns (fps, tps, thr) as float64 arrays INCLUDING the prepended (0, 0, inf) point;
    fpr = fps / fps[-1], tpr = tps / tps[-1] (NaN arrays when the divisor is 0)."""
    label = np.asarray(label).astype(np.int64)
    score = np.asarray(score)
    order = np.argsort(-score.astype(np.float64), kind="stable")      # stable descending (908)
    s, y = score[order], label[order]
    n = s.shape[0]
    if n == 0:
        raise ValueError("empty input")
    boundary = np.flatnonzero(s[1:] != s[:-1])                        # (917)
    idx = np.concatenate([boundary, [n - 1]])                         # (918-920)
    tps = np.cumsum(y)[idx].astype(np.float64)                        # (1034-1035)
    fps = 1.0 + idx.astype(np.float64) - tps                          # (1043)
    thr = s[idx].astype(np.float64)                                   # (1351) upcast
    if drop_intermediate and fps.shape[0] > 2:                        # (1331-1343)
        d2f = fps[2:] - 2 * fps[1:-1] + fps[:-2]
        d2t = tps[2:] - 2 * tps[1:-1] + tps[:-2]
        keep = np.concatenate([[True], (d2f != 0) | (d2t != 0), [True]])
        fps, tps, thr = fps[keep], tps[keep], thr[keep]
    return (np.concatenate([[0.0], fps]), np.concatenate([[0.0], tps]),
            np.concatenate([[np.inf], thr]))                          # (1347-1352)


def rates(fps, tps):
    with np.errstate(invalid="ignore", divide="ignore"):
        fpr = fps / fps[-1] if fps[-1] > 0 else np.full(fps.shape, np.nan)   # (1354-1361)
        tpr = tps / tps[-1] if tps[-1] > 0 else np.full(tps.shape, np.nan)   # (1363-1370)
    return fpr, tpr


def youden(fps, tps, thr):
    """First index maximising fp64 (tpr - fpr); raises ValueError for single-class labels
    exactly when the reference idiom does (the NaN pair is never `==` itself)."""
    fpr, tpr = rates(fps, tps)
    if np.isnan(fpr[0]) or np.isnan(tpr[0]):
        raise ValueError("(nan, nan) is not in list")
    j = tpr - fpr
    best, bi = j[0], 0
    for i in range(1, j.shape[0]):
        if j[i] > best:
            best, bi = j[i], i
    return thr[bi], bi


def pairwise_sum(a):
    """numpy's float pairwise summation (numpy/_core/src/umath/loops_utils.h.src,
    `@TYPE@_pairwise_sum`, as used by a contiguous 1-D `np.add.reduce`): n < 8 sequential from
    0; n <= 128: eight strided accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
    followed by the (n % 8) tail; otherwise split at n/2 rounded down to a multiple of 8.
    The order is machine independent; bit-equal to `np.sum` (tests/test_oracle_pinning.py)."""
    a = np.asarray(a)
    t = a.dtype.type

    def pw(lo, n):
        if n < 8:
            r = t(0.0)
            for i in range(n):
                r = t(r + a[lo + i])
            return r
        if n <= 128:
            r = [a[lo + k] for k in range(8)]
            i = 8
            while i < n - (n % 8):
                for k in range(8):
                    r[k] = t(r[k] + a[lo + i + k])
                i += 8
            res = t(t(t(r[0] + r[1]) + t(r[2] + r[3])) + t(t(r[4] + r[5]) + t(r[6] + r[7])))
            while i < n:
                res = t(res + a[lo + i])
                i += 1
            return res
        n2 = n // 2
        n2 -= n2 % 8
        return t(pw(lo, n2) + pw(lo + n2, n - n2))

    return t(t(0.0) + pw(0, a.shape[0]))


def trapezoid_auc(fps, tps):
    """sklearn.metrics.auc (_ranking.py:51-111) over the ROC points: np.trapezoid =
    sum(d * (y[1:] + y[:-1]) / 2.0) with numpy's pairwise sum; NaN for single-class."""
    fpr, tpr = rates(fps, tps)
    if fpr.shape[0] < 2:
        raise ValueError("At least 2 points are needed")
    d = np.diff(fpr)
    terms = d * (tpr[1:] + tpr[:-1]) / 2.0
    return float(pairwise_sum(terms))


def kahan_group_mean(values, codes, n_groups):
    """pandas 3.0.2 `_libs/groupby.pyx: group_mean`: per group, rows in table order,
    Kahan-compensated sum IN THE COLUMN DTYPE, then sum / count in that dtype.
    codes < 0 are skipped (NaN keys)."""
    values = np.asarray(values)
    t = values.dtype.type
    sumx = [t(0)] * n_groups
    comp = [t(0)] * n_groups
    cnt = [0] * n_groups
    with np.errstate(over="ignore", invalid="ignore"):
        for v, g in zip(values, codes):
            if g < 0:
                continue
            cnt[g] += 1
            y = t(v - comp[g])
            s = t(sumx[g] + y)
            c = t(t(s - sumx[g]) - y)
            comp[g] = t(0) if c != c else c
            sumx[g] = s
        out = np.array([t(sumx[g] / t(cnt[g])) if cnt[g] else t(np.nan) for g in range(n_groups)],
                       dtype=values.dtype)
    return out, np.array(cnt, dtype=np.int64)
