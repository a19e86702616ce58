"""UQ thresholding on the GPU -- drop-in for reference `biscuit.threshold`.

Same call surface, argument meaning, DataFrame mutation and error behaviour as reference
biscuit/threshold.py (`process_tile_predictions` 125-177, `process_group_predictions` 180-245,
`apply` 248-361, `detect` 364-475, `from_cv` 478-557); the arithmetic runs in
libbiscuit_b200.so (csrc/threshold.cu) through the C ABI of include/biscuit_b200.h:

    tile pass              -> bq_tile_process      (error / correct / incorrect / y_pred_bin)
    groupby(level).mean()  -> bq_group_reduce      (row-order Kahan sum in the column dtype)
    roc_curve + Youden + auc -> bq_tile_roc / bq_roc (radix sort + scan, fp64 J, first argmax)
    slide filter + confusion -> bq_group_apply

This module only factorises the string keys, resolves NumPy's scalar-promotion rule for each
comparison and rebuilds the DataFrames.  There is no CPU fallback: without the CUDA library and
a B200 every function raises `NativeLibraryError`.

Deliberate differences from the reference (all documented in DESIGN.md):
  * plotting (`plot=True`) is out of scope and ignored with a warning;
  * ROC curves whose result the reference discards (tile ROC when `tile_pred` is numeric, the
    group ROC when `slide_pred` is numeric) are not computed;
  * a missing y_true / y_pred / uncertainty column raises ValueError (the reference raises
    UnboundLocalError from a bug at threshold.py:184-186).
"""
from __future__ import annotations

import ctypes as C
import logging
import warnings

import numpy as np
import pandas as pd

from . import _ffi, errors

log = logging.getLogger("biscuit_b200")

_FLOAT_TYPES = (float, np.float16, np.float32, np.float64)   # threshold.py:411
_THRESH_KEYS = ("tile_uq", "slide_uq", "tile_pred", "slide_pred")
_RESULT_KEYS = ("auc", "percent_incl", "acc", "sensitivity", "specificity")


def _is_detect(x) -> bool:
    return isinstance(x, str) and x == "detect"


def _cmp_scalar(t, col_dtype) -> float:
    """The float64 value a column of `col_dtype` is effectively compared with under NumPy >= 2
    promotion (NEP 50), which pandas follows: python scalars are weak (rounded to the column
    dtype first), NumPy scalars are strong (float32 column vs np.float64 compares in float64)."""
    if isinstance(t, (np.floating, np.integer, np.bool_)) or (isinstance(t, np.ndarray) and t.ndim == 0):
        return float(t)
    if col_dtype == np.float32:
        with np.errstate(over="ignore"):
            return float(np.float32(t))
    return float(t)


def _float_col(a: np.ndarray) -> np.ndarray:
    if a.dtype == np.float32 or a.dtype == np.float64:
        return np.ascontiguousarray(a)
    return np.ascontiguousarray(a, dtype=np.float64)


def _label_col(a: np.ndarray, what="y_true") -> np.ndarray:
    """uint8 labels; anything sklearn's roc_curve would reject (non-binary, NaN) -> ValueError."""
    if a.dtype == np.bool_:
        return np.ascontiguousarray(a, dtype=np.uint8)
    with np.errstate(invalid="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b = a.astype(np.uint8)
    if not np.array_equal(b, a):
        raise ValueError(f"{what}: labels must be 0/1 (continuous / multiclass format is not supported)")
    return np.ascontiguousarray(b)


def _roc_struct():
    return _ffi.RocResult()


class _DeviceTable:
    """One tile table resident on the GPU (bq_table) for the duration of a call."""

    def __init__(self, y_pred, uncertainty, y_true, ctx=None):
        self.ctx = ctx or _ffi.default_context()
        self.lib = self.ctx.lib
        yp, un = _float_col(y_pred), _float_col(uncertainty)
        if yp.dtype != un.dtype:       # mixed float widths: promote both (documented deviation)
            yp, un = yp.astype(np.float64), un.astype(np.float64)
        self.dtype = yp.dtype
        self.code = _ffi.BQ_F32 if self.dtype == np.float32 else _ffi.BQ_F64
        self.n = int(yp.shape[0])
        yt = _label_col(y_true)
        if un.shape[0] != self.n or yt.shape[0] != self.n:
            raise ValueError("y_true, y_pred and uncertainty must have the same length")
        h = C.c_void_p()
        _ffi.check(self.ctx.handle,
                   self.lib.bq_table_create(self.ctx.handle, self.n, self.code, _ffi.ptr(yp),
                                            _ffi.ptr(un), _ffi.ptr(yt), C.byref(h)),
                   "bq_table_create")
        self.h = h
        self.n_groups = 0

    def close(self):
        if self.h:
            self.lib.bq_table_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- threshold.py:141 + sklearn input checks
    def validate(self):
        flags = (C.c_int64 * 4)()
        _ffi.check(self.ctx.handle, self.lib.bq_table_validate(self.h, flags), "bq_table_validate")
        return tuple(int(f) for f in flags)

    def tile_process(self, pred_thresh_eff: float, want_columns=True):
        n = self.n
        err = np.empty(n, np.float64) if want_columns else None
        cor = np.empty(n, np.uint8) if want_columns else None
        ybin = np.empty(n, np.uint8) if want_columns else None
        _ffi.check(self.ctx.handle,
                   self.lib.bq_tile_process(self.h, float(pred_thresh_eff), _ffi.ptr(err),
                                            _ffi.ptr(cor), _ffi.ptr(ybin)), "bq_tile_process")
        return err, cor, ybin

    def tile_roc(self, score_sel, label_sel):
        r = _roc_struct()
        _ffi.check(self.ctx.handle, self.lib.bq_tile_roc(self.h, score_sel, label_sel, C.byref(r)),
                   "bq_tile_roc")
        return r

    def set_groups(self, codes: np.ndarray, n_groups: int):
        codes = np.ascontiguousarray(codes, dtype=np.int32)
        _ffi.check(self.ctx.handle, self.lib.bq_table_set_groups(self.h, _ffi.ptr(codes), int(n_groups)),
                   "bq_table_set_groups")
        self.n_groups = int(n_groups)

    def set_filter(self, tile_uq_eff):
        if tile_uq_eff is None:
            rc = self.lib.bq_table_set_tile_filter(self.h, 0, 0.0)
        else:
            rc = self.lib.bq_table_set_tile_filter(self.h, 1, float(tile_uq_eff))
        _ffi.check(self.ctx.handle, rc, "bq_table_set_tile_filter")

    def group_reduce(self):
        L = self.n_groups
        gp, gu = np.empty(L, self.dtype), np.empty(L, self.dtype)
        gt = np.empty(L, np.float64)
        cnt, first = np.empty(L, np.int64), np.empty(L, np.int64)
        _ffi.check(self.ctx.handle,
                   self.lib.bq_group_reduce(self.h, _ffi.ptr(gp), _ffi.ptr(gu), _ffi.ptr(gt),
                                            _ffi.ptr(cnt), _ffi.ptr(first)), "bq_group_reduce")
        return gp, gu, gt, cnt, first


def _roc(ctx, score, label, include=None):
    """bq_roc on host arrays -> RocResult."""
    score = _float_col(np.asarray(score))
    label = np.ascontiguousarray(label, dtype=np.uint8)
    inc = None if include is None else np.ascontiguousarray(include, dtype=np.uint8)
    r = _roc_struct()
    code = _ffi.BQ_F32 if score.dtype == np.float32 else _ffi.BQ_F64
    _ffi.check(ctx.handle,
               ctx.lib.bq_roc(ctx.handle, _ffi.ptr(score), code, _ffi.ptr(label), _ffi.ptr(inc),
                              int(score.shape[0]), C.byref(r)), "bq_roc")
    return r


def _youden_or_raise(r):
    """The reference's `max(zip(tpr,fpr))` / `.index()` idiom raises ValueError for single-class
    labels; empty input makes sklearn raise ValueError too."""
    if r.status != 0:
        raise ValueError("(nan, nan) is not in list" if r.status == 1
                         else "Found array with 0 sample(s) while a minimum of 1 is required.")
    return np.float64(r.threshold)


def _check_columns(df):
    missing = [c for c in ("y_true", "y_pred", "uncertainty") if c not in df.columns]
    if missing:
        raise ValueError("Missing columns. Expected y_true, y_pred, uncertainty. "
                         f"Got: {', '.join(map(str, df.columns))}")


def _open_table(df, ctx=None) -> _DeviceTable:
    unc = df["uncertainty"].to_numpy() if "uncertainty" in df.columns else np.zeros(len(df), df["y_pred"].to_numpy().dtype)
    return _DeviceTable(df["y_pred"].to_numpy(), unc, df["y_true"].to_numpy(), ctx=ctx)


class ResidentTable:
    """A tile-prediction table kept RESIDENT on the GPU across calls.

    The reference re-derives everything from the DataFrame in every call: the nested-CV caller runs `detect` twice per
    inner fold and `apply` twice per outer fold on the same tables (experiment.py:966-1001).  Here the table is
    uploaded once (`bq_table`), and what does not depend on the thresholds being searched is computed once and kept:

      * the tile stage (validation, optional tile-ROC detection of `tile_pred`, the error / correct / incorrect /
        y_pred_bin columns added to the DataFrame, threshold.py:140-177) -- keyed by the `tile_pred` setting;
      * the first-appearance factorisation of the slide / patient keys (threshold.py:190).

    A later call with another tile-UQ threshold only re-runs the filter + reference-order slide reduction and the
    slide-level stage.  `apply_resident` / `detect_resident` are the entry points; `apply` / `detect` are the same code on
    a table that lives for one call.  The DataFrame is mutated exactly as the reference mutates it."""

    def __init__(self, df, ctx=None):
        _check_columns(df)
        self.df = df
        self.tab = _open_table(df, ctx=ctx)
        self.ctx = self.tab.ctx
        self._tile_done = {}
        self._keys = {}

    def close(self):
        if self.tab is not None:
            self.tab.close()
            self.tab = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def dtype(self):
        return self.tab.dtype

    def tile_stage(self, tile_pred, patients):
        """-> the tile prediction threshold in force (detected once per table when 'detect')."""
        key = "detect" if _is_detect(tile_pred) else (type(tile_pred).__name__, float(tile_pred))
        if key not in self._tile_done:
            # a different numeric threshold rewrites the derived columns; the incorrect flags on the device follow
            self._tile_done = {key: _tile_stage(self.tab, self.df, tile_pred, patients)}
        elif patients is not None and "patient" not in self.df.columns:
            self.df["patient"] = _map_slides(self.tab, self.df, patients)
        return self._tile_done[key]

    def keys(self, level):
        """(codes, uniques) of df[level] in first-appearance order, factorised once per level"""
        if level not in self._keys:
            fact = getattr(self.tab, "slide_factorized", None) if level == "slide" else None
            self._keys[level] = fact if fact is not None else _factorize(self.df[level])
        return self._keys[level]

    def groups(self, level, tile_uq_eff, pred_thresh):
        return _group_stage(self.tab, self.df, level, tile_uq_eff, pred_thresh, factorized=self.keys(level))


# ----------------------------------------------------------------------------------------
# tile level
# ----------------------------------------------------------------------------------------

def _tile_stage(tab: _DeviceTable, df, pred_thresh, patients):
    """threshold.py:140-177 on an open device table; mutates df; returns pred_thresh used."""
    n_nan, n_nonfinite, n_unc_nonfinite, n_badlabel = tab.validate()
    tab.n_unc_nonfinite = int(n_unc_nonfinite)      # sklearn's check fires only where the uncertainty feeds a ROC
    if n_nan:                                                         # :141-142
        raise errors.PredsContainNaNError
    if n_nonfinite:
        raise ValueError("Input contains infinity or a value too large for dtype.")
    if n_badlabel:
        raise ValueError("multiclass format is not supported")
    if _is_detect(pred_thresh):                                       # :145-159
        r = tab.tile_roc(_ffi.SCORE_Y_PRED, _ffi.LABEL_Y_TRUE)
        if r.status == 0:
            pred_thresh = np.float64(r.threshold)
        elif r.status == 1:
            log.debug("Unable to calculate tile prediction threshold; using 0.5")
            pred_thresh = 0.5                                         # :153-155
        else:
            raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required.")
        log.debug(f"Auto-detected tile prediction threshold: {pred_thresh:.4f}")
    else:
        if tab.n == 0:
            raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required.")
        log.debug(f"Using tile prediction threshold: {pred_thresh:.4f}")  # :161
    if patients is not None:                                          # :163-166
        df["patient"] = _map_slides(tab, df, patients)
    else:
        log.debug("Patients not provided; assuming 1:1 slide:patient mapping")
    err, cor, ybin = tab.tile_process(_cmp_scalar(pred_thresh, df["y_pred"].to_numpy().dtype))
    # dtypes as pandas gives them: |int64 - float| -> float64 (narrow ints keep the float width)
    err_dtype = np.result_type(df["y_true"].to_numpy().dtype, df["y_pred"].to_numpy().dtype)
    df["error"] = err if err_dtype == np.float64 else err.astype(err_dtype)   # :170
    df["correct"] = cor.view(np.bool_)                                # :171-174
    df["incorrect"] = (cor ^ 1).astype(int)                           # :175
    df["y_pred_bin"] = ybin.astype(int)                               # :176
    return pred_thresh


def _map_slides(tab, df, patients):
    """``df['slide'].map(patients)`` computed on the UNIQUE slide names and expanded through the factorisation
    codes (same `Series.map` on the same values, so the same dtype inference and NaN for unknown slides), instead
    of a hash lookup per tile row; the factorisation is kept on the table for the slide-level grouping."""
    codes, uniques = _factorize(df["slide"])
    tab.slide_factorized = (codes, uniques)
    if len(codes) == 0 or codes.min() < 0:                            # NaN slide names: leave it to pandas
        return df["slide"].map(patients)
    mapped = pd.Series(uniques, dtype=df["slide"].dtype).map(patients)
    return pd.Series(mapped.array.take(codes), index=df.index)


def process_tile_predictions(df, pred_thresh=0.5, patients=None):
    """Tile-level processing (reference threshold.py:125-177).

    Adds `error`, `correct`, `incorrect`, `y_pred_bin` (and `patient`) to `df` IN PLACE and returns
    ``(df, pred_thresh)``; ``pred_thresh='detect'`` picks Youden's J on the tile ROC."""
    with _open_table(df) as tab:
        pred_thresh = _tile_stage(tab, df, pred_thresh, patients)
    return df, pred_thresh


# ----------------------------------------------------------------------------------------
# group (slide / patient) level
# ----------------------------------------------------------------------------------------

def _factorize(keys: pd.Series):
    codes, uniques = pd.factorize(keys, use_na_sentinel=True)   # first-appearance order, NaN -> -1
    return codes.astype(np.int32, copy=False), uniques


def _group_stage(tab: _DeviceTable, df, level, tile_uq_eff, pred_thresh, factorized=None):
    """threshold.py:188-245 with the tile-UQ filter (threshold.py:298/412/426) applied on the
    device.  Returns (group DataFrame, pred_thresh used)."""
    if factorized is None and level == "slide":
        factorized = getattr(tab, "slide_factorized", None)          # already factorised for the patient mapping
    codes, uniques = factorized if factorized is not None else _factorize(df[level])
    tab.set_groups(codes, len(uniques))
    tab.set_filter(tile_uq_eff)
    gp, gu, gt, cnt, first = tab.group_reduce()
    alive = np.flatnonzero(cnt > 0)
    order = alive[np.argsort(first[alive], kind="stable")]           # first appearance after filter
    levels = [uniques[i] for i in order]                              # :190
    return _group_table(tab.ctx, level, levels, gp[order], gu[order], gt[order].astype(np.uint8), pred_thresh)


def _group_table(ctx, level, levels, yp, u, yt, pred_thresh):
    """threshold.py:197-245 on the per-group means (already in first-appearance order, `yt` truncated to uint8,
    :197-200): optional Youden detection on the group ROC, then the group DataFrame."""
    if not len(yt):                                                   # :205-206
        raise errors.ROCFailedError("Unable to generate ROC; preds are empty.")
    if _is_detect(pred_thresh):                                       # :217-223
        r = _roc(ctx, yp, yt)
        if r.status != 0:
            raise errors.ROCFailedError(f"Unable to generate {level}-level ROC")
        pred_thresh = np.float64(r.threshold)
        log.debug(f"Using detected prediction threshold: {pred_thresh:.4f}")
    else:
        log.debug(f"Using {level} prediction threshold: {pred_thresh:.4f}")   # :225
    cols = _group_columns(ctx, yp, u, yt, pred_thresh, pred_thresh, _ffi.KEEP_ALL, 0.0)
    l_df = pd.DataFrame({                                             # :235-244
        level: pd.Series(levels),
        "error": pd.Series(cols["error"]),
        "uncertainty": pd.Series(u),
        "correct": cols["correct"].view(np.bool_),
        "incorrect": pd.Series(cols["incorrect"]).astype(int),
        "y_true": pd.Series(yt),
        "y_pred": pd.Series(yp),
        "y_pred_bin": pd.Series(cols["y_pred_bin"].view(np.bool_)).astype(int),
    })
    return l_df, pred_thresh


def _group_columns(ctx, yp, u, yt, pred_thresh, strict_thresh, keep_mode, slide_uq_eff):
    L = int(yp.shape[0])
    dt = yp.dtype
    code = _ffi.BQ_F32 if dt == np.float32 else _ffi.BQ_F64
    out = {"error": np.empty(L, dt), "correct": np.empty(L, np.uint8), "incorrect": np.empty(L, np.uint8),
           "y_pred_bin": np.empty(L, np.uint8), "include": np.empty(L, np.uint8)}
    conf = (C.c_int64 * 4)()
    yp, u, yt = np.ascontiguousarray(yp), np.ascontiguousarray(u), np.ascontiguousarray(yt)
    _ffi.check(ctx.handle,
               ctx.lib.bq_group_apply(ctx.handle, L, code, _ffi.ptr(yp), _ffi.ptr(u), _ffi.ptr(yt),
                                      _cmp_scalar(pred_thresh, dt), _cmp_scalar(strict_thresh, dt),
                                      int(keep_mode), float(slide_uq_eff),
                                      _ffi.ptr(out["error"]), _ffi.ptr(out["correct"]),
                                      _ffi.ptr(out["incorrect"]), _ffi.ptr(out["y_pred_bin"]),
                                      _ffi.ptr(out["include"]), conf), "bq_group_apply")
    out["confusion"] = tuple(np.int64(c) for c in conf)
    return out


def process_group_predictions(df, pred_thresh, level):
    """Group-level (slide / patient) predictions and uncertainty from tile-level rows
    (reference threshold.py:180-245): per-group means in first-appearance order, `y_true`
    truncated to uint8, `correct / incorrect / y_pred_bin` with ``>= pred_thresh``.
    ``pred_thresh='detect'`` picks Youden's J on the group ROC (ROCFailedError if impossible)."""
    _check_columns(df)
    if len(df) == 0:
        df[level]                                                     # KeyError parity
        raise errors.ROCFailedError("Unable to generate ROC; preds are empty.")
    with _open_table(df) as tab:
        return _group_stage(tab, df, level, None, pred_thresh)


def _auc_included(ctx, s_yp, s_yt, include=None):
    """utils.py:487-504: ROC AUC of the surviving groups, NaN when undefined."""
    r = _roc(ctx, s_yp, s_yt, include)
    if r.status == 2:
        log.warning("Unable to calculate ROC")
        return np.nan
    return float(r.auc)


def apply(df, tile_uq, slide_uq, tile_pred=0.5, slide_pred=0.5, plot=False,
          keep="high_confidence", title=None, patients=None, level="slide"):
    """Apply pre-calculated tile- and group-level uncertainty thresholds
    (reference threshold.py:248-361).

    Args:
        df (pandas.DataFrame): columns 'y_true', 'y_pred', 'uncertainty', 'slide'. Mutated in place
            exactly like the reference (adds error / correct / incorrect / y_pred_bin / patient).
        tile_uq (float): tile-level uncertainty threshold (falsy: no tile filter).
        slide_uq (float): group-level uncertainty threshold (falsy: no group filter).
        tile_pred, slide_pred (float): prediction thresholds. Default 0.5.
        keep (str): 'high_confidence' (uncertainty < slide_uq) or 'low_confidence' (>=).
        patients (dict): slide -> patient; required for level='patient'.
        level (str): 'slide' or 'patient'.

    Returns:
        dict with auc, percent_incl, acc, sensitivity, specificity; DataFrame of the kept groups
        (original integer index labels).  ({...None}, None) if no group-level ROC is possible."""
    if plot:
        log.warning("plot=True ignored: plotting is outside the scope of biscuit_b200")
    assert keep in ("high_confidence", "low_confidence")             # :281
    assert not (level == "patient" and patients is None)             # :282
    log.debug(f"Applying tile UQ threshold of {tile_uq:.5f}")         # :284 (TypeError on None)
    if patients:                                                      # :285-286
        df["patient"] = df["slide"].map(patients)
    df[level]                                                         # :287 KeyError parity
    with ResidentTable(df) as table:
        return apply_resident(table, tile_uq, slide_uq, tile_pred=tile_pred, slide_pred=slide_pred, keep=keep,
                              patients=patients, level=level)


def apply_resident(table: ResidentTable, tile_uq, slide_uq, tile_pred=0.5, slide_pred=0.5, plot=False,
                   keep="high_confidence", title=None, patients=None, level="slide"):
    """`apply` on a table that is already resident on the GPU (see :class:`ResidentTable`): the tile stage and the key
    factorisation are reused from earlier calls, only the filter + slide reduction + slide-level stage run."""
    assert keep in ("high_confidence", "low_confidence")
    assert not (level == "patient" and patients is None)
    df = table.df
    if patients and "patient" not in df.columns:
        df["patient"] = df["slide"].map(patients)
    df[level]
    table.tile_stage(tile_pred, patients)                             # :290-294
    codes, uniques = table.keys(level)
    n_before = len(uniques) + int((codes < 0).any())                  # :295 (NaN counts as a key)
    tile_uq_eff = _cmp_scalar(tile_uq, table.dtype) if tile_uq else None   # :297-298
    try:
        s_df, _ = table.groups(level, tile_uq_eff, slide_pred)        # :305
    except errors.ROCFailedError:
        log.error("Unable to process slide predictions")
        return {k: None for k in _RESULT_KEYS}, None                  # :310-317
    return _apply_group_level(table.ctx, s_df, slide_uq, slide_pred, keep, level, n_before)


def _apply_group_level(ctx, s_df, slide_uq, slide_pred, keep, level, n_before):
    """threshold.py:323-361: group-level UQ filter, AUROC, percent included and the confusion counts."""
    yp, u, yt = s_df["y_pred"].to_numpy(), s_df["uncertainty"].to_numpy(), s_df["y_true"].to_numpy()
    if slide_uq:                                                      # :323-330
        log.debug(f"Using {level} uncertainty threshold of {slide_uq:.5f}")
        mode = _ffi.KEEP_HIGH if keep == "high_confidence" else _ffi.KEEP_LOW
        uq_eff = _cmp_scalar(slide_uq, u.dtype)
    else:
        mode, uq_eff = _ffi.KEEP_ALL, 0.0
    cols = _group_columns(ctx, yp, u, yt, slide_pred, slide_pred, mode, uq_eff)
    include = cols["include"].view(np.bool_)
    if slide_uq:
        s_df = s_df.loc[include]
    auc = _auc_included(ctx, yp, yt, cols["include"])                 # :333
    percent_incl = len(s_df) / n_before                               # :334-335
    tp, fp, tn, fn = cols["confusion"]                                # :339-345
    with np.errstate(invalid="ignore", divide="ignore"):
        results = {"auc": auc, "percent_incl": percent_incl,
                   "acc": (tp + tn) / (tp + tn + fp + fn),            # :346
                   "sensitivity": tp / (tp + fn),                     # :347
                   "specificity": tn / (tn + fp)}                     # :348
    return results, s_df


# ----------------------------------------------------------------------------------------
# cohort sharded over GPUs: one process per GPU, slide-aligned shards
# ----------------------------------------------------------------------------------------
_ERR_NONE, _ERR_NAN, _ERR_NONFINITE, _ERR_LABEL, _ERR_EMPTY = 0, 1, 2, 3, 4


def _raise_shard_error(code):
    """the same exception on every rank (a rank that raised alone would leave the others blocked in a collective)"""
    if code == _ERR_NAN:
        raise errors.PredsContainNaNError
    if code == _ERR_NONFINITE:
        raise ValueError("Input contains infinity or a value too large for dtype.")
    if code == _ERR_LABEL:
        raise ValueError("multiclass format is not supported")
    if code == _ERR_EMPTY:
        raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required.")


def _shard_local(df, level, patients, tile_pred, tile_uq_eff_fn):
    """Local half of a sharded call: validation, tile processing (mutates the shard like the reference mutates the
    table) and the reference-order per-slide reduction of THIS rank's slides.  Never raises: the error code travels
    with the size exchange so that every rank fails together.  An empty shard (more ranks than slides) contributes
    zero groups."""
    out = {"err": _ERR_NONE, "n_rows": len(df), "names": [], "n_keys": 0, "msg": np.zeros((0, 6), np.float64),
           "dtype": np.dtype(df["y_pred"].to_numpy().dtype) if "y_pred" in df.columns and len(df) else None,
           "ctx": _ffi.default_context()}
    if len(df) == 0:
        return out
    with _open_table(df) as tab:
        out["ctx"], out["dtype"] = tab.ctx, tab.dtype
        try:
            _tile_stage(tab, df, tile_pred, patients)
        except errors.PredsContainNaNError:
            out["err"] = _ERR_NAN
            return out
        except ValueError as e:
            out["err"] = _ERR_LABEL if "multiclass" in str(e) else _ERR_NONFINITE
            return out
        codes, uniques = _factorize(df[level])
        out["n_keys"] = len(uniques) + int((codes < 0).any())
        out["names"] = [str(u) for u in uniques]
        tab.set_groups(codes, len(uniques))
        tab.set_filter(tile_uq_eff_fn(tab.dtype))
        gp, gu, gt, cnt, first = tab.group_reduce()
    out["local"] = (gp, gu, gt, cnt, first)
    return out


def _exchange_groups(loc, group):
    """meta all-gather (sizes + error codes) and ONE payload all-gather (per-slide aggregates + slide names as utf-8
    tensors): -> (groups dict in global first-appearance order, all names, keys before the tile filter, dtype)."""
    from . import dist as bdist
    name_buf, name_len = bdist.pack_names(loc["names"])
    L = len(loc["names"])
    dcode = -1 if loc["dtype"] is None else (0 if loc["dtype"] == np.float32 else 1)
    device = f"cuda:{loc['ctx'].device}"
    meta = bdist.all_gather_meta([loc["n_rows"], L, loc["n_keys"], loc["err"], name_buf.shape[0], dcode],
                                 group=group, device=device)
    errs = meta[:, 3]
    if errs.any():
        _raise_shard_error(int(errs[errs != 0][0]))
    if int(meta[:, 0].sum()) == 0:
        _raise_shard_error(_ERR_EMPTY)
    _, _, rank = bdist._dist_state(group)
    code_off, row_off = int(meta[:rank, 1].sum()), int(meta[:rank, 0].sum())
    if L:
        gp, gu, gt, cnt, first = loc["local"]
        msg = bdist.pack_groups(code_off, cnt, first, row_off, gp, gu, gt)
    else:
        msg = np.zeros((0, 6), np.float64)
    payload = np.concatenate([msg.view(np.uint8).reshape(-1), name_len.view(np.uint8).reshape(-1), name_buf])
    sizes = [int(m[1]) * 52 + int(m[4]) for m in meta]              # 48 B of aggregates + 4 B name length per slide
    parts = bdist.all_gather_bytes(payload, sizes, group=group, device=device)
    msgs, names = [], []
    for m, part in zip(meta, parts):
        Lr, nb = int(m[1]), int(m[4])
        msgs.append(part[: Lr * 48].copy().view(np.float64).reshape(Lr, 6))
        lens = part[Lr * 48: Lr * 52].copy().view(np.int32)
        names += bdist.unpack_names(part[Lr * 52: Lr * 52 + nb], lens)
    dcodes = meta[:, 5][meta[:, 5] >= 0]
    dtype = np.dtype(np.float64 if (dcodes == 1).any() else np.float32)
    g = bdist.unpack_groups(np.concatenate(msgs, axis=0), dtype)
    return g, names, int(meta[:, 2].sum()), dtype


def apply_sharded(df, tile_uq, slide_uq, tile_pred=0.5, slide_pred=0.5, keep="high_confidence",
                  patients=None, level="slide", group=None):
    """`apply` for a cohort sharded over ranks (one process per GPU, torch.distributed initialised).

    Every rank passes ITS contiguous, slide-aligned shard of the tile table (see `dist.shard_bounds`;
    a slide / patient must not straddle ranks; a shard may be empty).  Tile processing and the
    reference-order slide reduction run locally on each GPU; the per-slide aggregates (48 B per slide)
    and the slide names (utf-8 tensors, no pickling) are all-gathered ONCE after a small size / error-code
    exchange, and the group-level thresholding runs replicated, so every rank returns the same
    (results, s_df) `apply` would return on the concatenated table -- and every rank raises the same
    exception when any shard fails validation.  Thresholds must be numeric: 'detect' is resolved on the
    whole cohort by `detect_sharded`."""
    assert keep in ("high_confidence", "low_confidence")
    assert not (level == "patient" and patients is None)
    if _is_detect(tile_pred) or _is_detect(slide_pred):
        raise ValueError("apply_sharded needs numeric prediction thresholds (a per-shard 'detect' would differ "
                         "between ranks): use detect_sharded / from_cv_sharded first")
    log.debug(f"Applying tile UQ threshold of {tile_uq:.5f}")
    if patients:
        df["patient"] = df["slide"].map(patients)
    df[level]
    _check_columns(df)
    loc = _shard_local(df, level, patients, tile_pred, lambda dt: _cmp_scalar(tile_uq, dt) if tile_uq else None)
    g, names, n_before, _ = _exchange_groups(loc, group)
    levels = [names[c] for c in g["code"]]
    try:
        s_df, _ = _group_table(loc["ctx"], level, levels, g["y_pred"], g["uncertainty"], g["y_true"], slide_pred)
    except errors.ROCFailedError:
        log.error("Unable to process slide predictions")
        return {k: None for k in _RESULT_KEYS}, None
    return _apply_group_level(loc["ctx"], s_df, slide_uq, slide_pred, keep, level, n_before)


def detect_sharded(df, tile_uq="detect", slide_uq="detect", tile_pred="detect", slide_pred="detect",
                   patients=None, group=None):
    """`detect` for a table sharded over ranks (reference threshold.py:364-475 on the concatenated table).

    The tile-level ROCs (`y_true ~ y_pred`, `incorrect ~ uncertainty`) run over ALL tiles of the cohort, so the
    per-tile `(pred, unc, label)` triples (9 B per tile for float32 tables) are all-gathered -- the one real
    exchange step of the path -- and the ROC kernels run on the gathered table on every GPU
    (deterministic integer/fp64 kernels: replicated, identical thresholds, no broadcast needed).  The slide
    reduction stays local and bit-exact (slide-aligned shards), its aggregates are all-gathered as in
    `apply_sharded`, and the slide-level detection runs replicated.  Returns what `detect` returns, on every rank."""
    from . import dist as bdist
    none4 = {k: None for k in _THRESH_KEYS}
    _check_columns(df)
    ctx = _ffi.default_context()
    device = f"cuda:{ctx.device}"
    n_local = len(df)
    yp_l = _float_col(df["y_pred"].to_numpy()) if n_local else np.zeros(0, np.float32)
    un_l = _float_col(df["uncertainty"].to_numpy()) if n_local else np.zeros(0, np.float32)
    if yp_l.dtype != un_l.dtype:
        yp_l, un_l = yp_l.astype(np.float64), un_l.astype(np.float64)
    need_tiles = _is_detect(tile_pred) or _is_detect(tile_uq)
    if need_tiles:
        # ---- exchange 1: sizes + column dtype, then the tile triples
        meta = bdist.all_gather_meta([n_local, -1 if not n_local else int(yp_l.dtype == np.float64)], group=group, device=device)
        wide = bool((meta[:, 1] == 1).any())
        dt = np.dtype(np.float64 if wide else np.float32)
        buf = bdist.pack_tiles(yp_l.astype(dt, copy=False), un_l.astype(dt, copy=False),
                               _label_col(df["y_true"].to_numpy()) if n_local else np.zeros(0, np.uint8))
        parts = bdist.all_gather_bytes(buf, [int(m[0]) * (2 * dt.itemsize + 1) for m in meta], group=group, device=device)
        cols = [bdist.unpack_tiles(part, int(m[0]), dt) for m, part in zip(meta, parts)]
        yp_g, un_g, yt_g = (np.concatenate([c[i] for c in cols]) for i in range(3))
        with _DeviceTable(yp_g, un_g, yt_g, ctx=ctx) as tab:
            n_nan, n_nonfinite, n_unc_nonfinite, n_badlabel = tab.validate()
            if n_nan:
                log.error("Tile-level predictions contain NaNs; unable to process.")
                return none4, None
            if n_nonfinite:
                raise ValueError("Input contains infinity or a value too large for dtype.")
            if n_badlabel:
                raise ValueError("multiclass format is not supported")
            if _is_detect(tile_pred):
                r = tab.tile_roc(_ffi.SCORE_Y_PRED, _ffi.LABEL_Y_TRUE)
                if r.status == 0:
                    tile_pred = np.float64(r.threshold)
                elif r.status == 1:
                    tile_pred = 0.5
                else:
                    raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required.")
            tab.tile_process(_cmp_scalar(tile_pred, tab.dtype))
            if _is_detect(tile_uq):
                if n_unc_nonfinite:
                    raise ValueError("Input contains NaN or infinity.")
                tile_uq = _youden_or_raise(tab.tile_roc(_ffi.SCORE_UNCERTAINTY, _ffi.LABEL_INCORRECT))
    if isinstance(tile_uq, _FLOAT_TYPES):
        def tile_uq_eff_fn(dtype, t=tile_uq):
            return _cmp_scalar(t, dtype) if not isinstance(t, np.float64) or True else float(t)
    else:
        if not _is_detect(tile_uq):
            log.debug("Not performing tile-level uncertainty thresholding.")
        tile_uq = None

        def tile_uq_eff_fn(dtype):
            return None
    # ---- local tile processing + slide reduction, exchange 2: per-slide aggregates
    loc = _shard_local(df, "slide", patients, tile_pred, tile_uq_eff_fn)
    try:
        g, names, _, _ = _exchange_groups(loc, group)
    except errors.PredsContainNaNError:
        log.error("Tile-level predictions contain NaNs; unable to process.")
        return none4, None
    levels = [names[c] for c in g["code"]]
    try:
        s_df, slide_pred = _group_table(loc["ctx"], "slide", levels, g["y_pred"], g["uncertainty"], g["y_true"], slide_pred)
    except errors.ROCFailedError:
        log.error("Unable to process slide predictions")
        return none4, None
    slide_uq, auc = _detect_group_level(loc["ctx"], s_df, slide_uq, slide_pred)
    return {"tile_uq": tile_uq, "slide_uq": slide_uq, "tile_pred": tile_pred, "slide_pred": slide_pred}, auc


def from_cv_sharded(dfs, group=None, **kwargs):
    """`from_cv` with every fold table sharded over the ranks: `detect_sharded` per fold, same fold-skipping and
    min / max / mean reduction over folds as the reference (threshold.py:478-557)."""
    return _from_cv(dfs, lambda df, **kw: detect_sharded(df, group=group, **kw), kwargs, require_patient=False)


def detect(df, tile_uq="detect", slide_uq="detect", tile_pred="detect", slide_pred="detect",
           plot=False, patients=None):
    """Detect optimal tile- and slide-level uncertainty thresholds (reference threshold.py:364-475).

    Each of tile_uq / slide_uq / tile_pred / slide_pred is 'detect' (Youden's J on the matching
    ROC) or a float to use as given.  Returns (dict of the four thresholds, slide-level AUROC), or
    (all-None dict, None) when predictions contain NaN or no slide-level ROC is possible."""
    if plot:
        log.warning("plot=True ignored: plotting is outside the scope of biscuit_b200")
    with ResidentTable(df) as table:
        return detect_resident(table, tile_uq=tile_uq, slide_uq=slide_uq, tile_pred=tile_pred, slide_pred=slide_pred,
                               patients=patients)


def detect_resident(table: ResidentTable, tile_uq="detect", slide_uq="detect", tile_pred="detect", slide_pred="detect",
                    plot=False, patients=None):
    """`detect` on a GPU-resident table (see :class:`ResidentTable`)."""
    none4 = {k: None for k in _THRESH_KEYS}
    tab = table.tab
    try:
        found_tile_pred = table.tile_stage(tile_pred, patients)       # :398-402
    except errors.PredsContainNaNError:
        log.error("Tile-level predictions contain NaNs; unable to process.")
        return none4, None                                            # :403-405
    if _is_detect(tile_pred):                                         # :407-408
        tile_pred = found_tile_pred
    if isinstance(tile_uq, _FLOAT_TYPES):                             # :411-412
        tile_uq_eff = _cmp_scalar(tile_uq, tab.dtype)
    elif not _is_detect(tile_uq):                                     # :413-415
        log.debug("Not performing tile-level uncertainty thresholding.")
        tile_uq, tile_uq_eff = None, None
    else:                                                             # :416-426
        if getattr(tab, "n_unc_nonfinite", 0):                         # roc_curve -> assert_all_finite (uncaught at :419)
            raise ValueError("Input contains NaN or infinity.")
        r = tab.tile_roc(_ffi.SCORE_UNCERTAINTY, _ffi.LABEL_INCORRECT)
        tile_uq = _youden_or_raise(r)
        log.debug(f"Tile-level optimal UQ threshold: {tile_uq:.4f}")
        tile_uq_eff = float(tile_uq)
    try:
        s_df, slide_pred = table.groups("slide", tile_uq_eff, slide_pred)   # :433-438
    except errors.ROCFailedError:
        log.error("Unable to process slide predictions")
        return none4, None                                            # :439-441
    slide_uq, auc = _detect_group_level(table.ctx, s_df, slide_uq, slide_pred)
    return {"tile_uq": tile_uq, "slide_uq": slide_uq,
            "tile_pred": tile_pred, "slide_pred": slide_pred}, auc


def _detect_group_level(ctx, s_df, slide_uq, slide_pred):
    """threshold.py:444-468: slide-level UQ threshold (Youden on incorrect ~ uncertainty) and the AUROC of the kept slides."""
    yp, u, yt = s_df["y_pred"].to_numpy(), s_df["uncertainty"].to_numpy(), s_df["y_true"].to_numpy()
    include = None
    if _is_detect(slide_uq):                                          # :444-460
        inc = s_df["incorrect"].to_numpy()
        if not inc.sum():
            log.debug("Unable to calculate slide UQ threshold; no incorrect predictions made")
            slide_uq = None
        else:
            slide_uq = _youden_or_raise(_roc(ctx, u, inc))
            log.debug(f"Slide-level optimal UQ threshold: {slide_uq:.4f}")
            include = _group_columns(ctx, yp, u, yt, slide_pred, slide_pred, _ffi.KEEP_HIGH,
                                     float(slide_uq))["include"]
    else:
        log.debug("Not performing slide-level uncertainty thresholding.")
        slide_uq = 0.5                                                # :461-463
    return slide_uq, _auc_included(ctx, yp, yt, include)              # :468


def from_cv(dfs, **kwargs):
    """Optimal tile- and slide-level thresholds from a set of (nested) cross-validation folds
    (reference threshold.py:478-557): `detect` per fold, folds without a tile_uq / slide_uq are
    skipped, then tile_uq = min, slide_uq = max, tile_pred / slide_pred = mean over folds.

    Args:
        dfs (list(DataFrame)): tile predictions with 'y_true', 'y_pred', 'uncertainty', 'slide',
            'patient'.
        **kwargs: forwarded to :func:`detect` (tile_uq, slide_uq, tile_pred, slide_pred, patients).

    Raises ValueError for missing columns and ThresholdError when no fold yields a threshold."""
    return _from_cv(dfs, detect, kwargs)


class FoldSet:
    """The tile tables of a set of cross-validation folds, uploaded once and kept resident on the GPU.

    `from_cv` may then be called repeatedly -- the nested-CV caller detects the tile-level UQ threshold first and the
    slide-level thresholds with that tile threshold fixed (reference experiment.py:966-977) -- and every call after the
    first reuses each fold's tile stage and key factorisation (:class:`ResidentTable`)."""

    def __init__(self, dfs, ctx=None):
        self.tables = []
        try:
            for df in dfs:
                missing = [c for c in _CV_COLUMNS if c not in df.columns]
                if missing:
                    raise ValueError(f"DataFrame missing columns, expected {_CV_COLUMNS}, got: "
                                     f"{', '.join(df.columns.tolist())}")
                self.tables.append(ResidentTable(df, ctx=ctx))
        except Exception:
            self.close()
            raise

    def close(self):
        for t in self.tables:
            t.close()
        self.tables = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __len__(self):
        return len(self.tables)

    def from_cv(self, **kwargs):
        """same result as ``threshold.from_cv([t.df for t in tables], **kwargs)`` (reference threshold.py:478-557)"""
        return _from_cv(self.tables, lambda t, **kw: detect_resident(t, **kw), kwargs, columns_of=lambda t: t.df.columns)


_CV_COLUMNS = ("y_true", "y_pred", "uncertainty", "slide", "patient")


def _from_cv(dfs, detect_fn, kwargs, require_patient=True, columns_of=lambda df: df.columns):
    """Fold pass + reduction of the reference's from_cv (threshold.py:478-557): every fold table is validated and run
    through `detect_fn`; folds without both UQ thresholds drop out (:526-528); the survivors reduce to
    min(tile_uq), max(slide_uq), mean(tile_pred), mean(slide_pred) (:544-557).  The two `*_uq_thresh=None` keywords
    switch a UQ reduction off and leave an empty list in its place, as the reference does (:513-516)."""
    required = ("y_true", "y_pred", "uncertainty", "slide") + (("patient",) if require_patient else ())
    kept = []
    for idx, df in enumerate(dfs):
        log.debug(f"Detecting thresholds from fold {idx}")
        present = list(columns_of(df))
        if any(col not in present for col in required):
            raise ValueError(f"DataFrame missing columns, expected {required}, got: {', '.join(present)}")
        fold, _ = detect_fn(df, **kwargs)
        if fold["tile_uq"] is None or fold["slide_uq"] is None:
            log.debug(f"Skipping CV #{idx}, unable to detect threshold")
            continue
        kept.append(fold)
    reducers = {"tile_uq": ("tile", np.min), "slide_uq": ("slide", np.max)}
    out = {}
    for key, (what, reduce_fn) in reducers.items():
        switched_off = kwargs.get(f"{key}_thresh", 0) is None and f"{key}_thresh" in kwargs
        if switched_off:
            out[key] = []
        elif not kept:
            raise errors.ThresholdError(f"Unable to detect {what} UQ threshold.")
        else:
            out[key] = reduce_fn([fold[key] for fold in kept])
    for key in ("tile_pred", "slide_pred"):
        out[key] = np.mean([fold[key] for fold in kept])
    return out
