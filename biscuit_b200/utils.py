"""Column contract and ROC helpers -- the parts of reference biscuit/utils.py that sit on the hot
path: `uncertainty_header / y_true_header / y_pred_header` (19-28), `rename_cols` (31-53),
`auc_and_threshold` (467-484) and `auc` (487-504).  The ROC arithmetic runs on the GPU (bq_roc)."""
from __future__ import annotations

import logging

import numpy as np

log = logging.getLogger("biscuit_b200")


def uncertainty_header(outcome, underscore=False):
    return f"{outcome}{'_' if underscore else '-'}uncertainty1"


def y_true_header(outcome, underscore=False):
    return f"{outcome}{'_' if underscore else '-'}y_true0"


def y_pred_header(outcome, underscore=False):
    return f"{outcome}{'_' if underscore else '-'}y_pred1"


def rename_cols(df, outcome, *, y_true=None, y_pred=None, uncertainty=None):
    """Renames Slideflow's per-outcome columns to y_true / y_pred / uncertainty, IN PLACE; dash and
    underscore spellings are both accepted, and '{outcome}-y_true' is the fallback label column
    (reference utils.py:31-53)."""
    def pick(header, given, fallback=None):
        if given is not None:
            return given
        name = header(outcome, underscore=(header(outcome, underscore=True) in df.columns))
        if fallback is not None and name not in df.columns:
            return fallback
        return name

    mapping = {
        pick(y_true_header, y_true, fallback=f"{outcome}-y_true"): "y_true",
        pick(y_pred_header, y_pred): "y_pred",
        pick(uncertainty_header, uncertainty): "uncertainty",
    }
    df.rename(columns=mapping, inplace=True)


def _roc(y_true, y_pred):
    from . import threshold
    from . import _ffi
    y_true = np.asarray(y_true)
    y_pred = np.asarray(y_pred)
    if y_true.shape[0] == 0:
        raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required.")
    if not np.isfinite(y_pred).all():
        raise ValueError("Input contains NaN or infinity.")
    return threshold._roc(_ffi.default_context(), y_pred, threshold._label_col(y_true))


def auc_and_threshold(y_true, y_pred):
    """AUROC and the Youden-J optimal threshold (reference utils.py:467-484)."""
    r = _roc(y_true, y_pred)
    if r.status != 0:
        raise ValueError("(nan, nan) is not in list")
    return float(r.auc), np.float64(r.threshold)


def auc(y_true, y_pred):
    """AUROC; NaN (with a warning) when it cannot be computed (reference utils.py:487-504)."""
    try:
        return float(_roc(y_true, y_pred).auc)
    except ValueError:
        log.warning("Unable to calculate ROC")
        return np.nan
