"""Column contract, prediction-table loaders and ROC helpers -- the parts of reference biscuit/utils.py
that sit on (or right next to) the hot path: `uncertainty_header / y_true_header / y_pred_header`
(19-28), `rename_cols` (31-53), `df_from_cv` (190-228), `find_model / model_exists / find_cv`
(233-311), `auc_and_threshold` (467-484) and `auc` (487-504).  The ROC arithmetic runs on the GPU
(bq_roc); the loaders are host-side pandas I/O exactly as in the reference, so a table read from CSV
is float64 and one read from parquet keeps its float32 columns -- the dtype the GPU path then
computes in."""
from __future__ import annotations

import csv
import logging
import os
from os.path import join

import numpy as np
import pandas as pd

from .errors import ModelNotFoundError, MultipleModelsFoundError

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


# --- prediction-table loaders and model lookup --------------------------------

def find_model(project, label, outcome, epoch=None, kfold=None):
    """Path of the trained model `{outcome}-{label}-HP0[-kfold{k}]` inside ``project.models_dir``
    (Slideflow prefixes every model folder with a 5-digit id and a dash, hence the ``[6:]``);
    with ``epoch`` the saved-model sub-folder, else the parent folder.  Raises
    ModelNotFoundError / MultipleModelsFoundError (reference utils.py:233-272)."""
    tail = "" if kfold is None else f"-kfold{kfold}"
    name = f"{outcome}-{label}-HP0{tail}"
    matching = [o for o in os.listdir(project.models_dir) if o[6:] == name]
    if len(matching) > 1:
        raise MultipleModelsFoundError(f"Multiple matching models found matching {name}")
    if not matching:
        raise ModelNotFoundError(f"No matching model found matching {name}.")
    if epoch is not None:
        return join(project.models_dir, matching[0], f"{name}_epoch{epoch}")
    return join(project.models_dir, matching[0])


def model_exists(project, label, outcome, epoch=None, kfold=None):
    """True if :func:`find_model` finds exactly one match (reference utils.py:275-292; more than
    one match still raises, as there)."""
    try:
        find_model(project, label, outcome, kfold=kfold, epoch=epoch)
        return True
    except ModelNotFoundError:
        return False


def find_cv(project, label, outcome, epoch=None, k=3):
    """Paths of the k cross-validation models of one experiment (reference utils.py:295-311)."""
    return [find_model(project, label, outcome, epoch=epoch, kfold=_k) for _k in range(1, k + 1)]


def read_tile_predictions(path):
    """One Slideflow tile-prediction table: ``.csv`` (slide column forced to str, reference
    experiment.py:980-981) or ``.parquet`` / ``.parquet.gzip`` (982-983); anything else is an
    OSError (984-985)."""
    ext = path.rsplit(".", 1)[-1].lower()
    if ext == "csv":
        return pd.read_csv(path, dtype={"slide": str})
    if ext in ("parquet", "gzip"):
        return pd.read_parquet(path)
    raise OSError(f"Unrecognized prediction filetype {path}")


def df_from_cv(project, label, outcome, epoch=None, k=3, y_true=None, y_pred=None, uncertainty=None):
    """Tile-prediction tables of the k cross-validation folds of `label`, columns renamed to
    y_true / y_pred / uncertainty and a ``patient`` column added from the project's slide ->
    patient map when the file has none (reference utils.py:190-228).  CSV wins over
    ``.parquet.gzip`` when both exist, as in the reference."""
    dfs = []
    folders = find_cv(project, label, epoch=epoch, k=k, outcome=outcome)
    patients = project.dataset().patients()
    e = "" if epoch is None else "../"
    for folder in folders:
        csv_path = join(folder, f"{e}tile_predictions_val_epoch1.csv")
        parquet_path = join(folder, f"{e}tile_predictions_val_epoch1.parquet.gzip")
        if os.path.exists(csv_path):
            df = pd.read_csv(csv_path)
        elif os.path.exists(parquet_path):
            df = pd.read_parquet(parquet_path)
        else:
            raise OSError(f"Could not find tile predictions file at {folder}")
        rename_cols(df, outcome, y_true=y_true, y_pred=y_pred, uncertainty=uncertainty)
        if "patient" not in df.columns:
            df["patient"] = df["slide"].map(patients)
        dfs.append(df)
    return dfs


def slides_from_model_manifest(model_path, dataset=None):
    """Slide names listed in a trained model's ``slide_manifest.csv`` (looked up in the model folder,
    then its parent), optionally restricted to one dataset ('training' / 'validation').  Stands in
    for ``sf.util.get_slides_from_model_manifest`` which reference experiment.py:1008 calls to
    report ``n_slides``; Slideflow is an external dependency, so this follows its documented file
    format (header row with 'slide' and 'dataset' columns)."""
    for folder in (model_path, os.path.dirname(os.path.normpath(model_path))):
        path = join(folder, "slide_manifest.csv")
        if os.path.exists(path):
            with open(path, newline="") as f:
                rows = list(csv.DictReader(f))
            return [r["slide"] for r in rows if dataset is None or r.get("dataset") == dataset]
    raise OSError(f"Could not find slide manifest for model {model_path}")


# --- evaluation metrics ------------------------------------------------------

def prediction_metrics(y_true, y_pred, threshold):
    """AUC confidence interval (DeLong), accuracy, sensitivity / specificity and Youden's J with its
    bootstrap confidence interval (reference utils.py:400-464).

    The 500 bootstrap samples of 150 rows are drawn with the same ``np.random.choice`` calls as the
    reference (so a seeded ``np.random`` gives the reference's numbers); their confusion matrices and the
    DeLong placement values are computed on the GPU, the order-sensitive finish (``statistics.mean`` /
    ``variance``, ``np.cov``, ``scipy.stats.norm``) by the same library calls as the reference."""
    import ctypes as C
    from statistics import mean, variance

    from scipy import stats

    from . import _ffi
    from .delong import delong_roc_variance

    yt = y_true.astype(bool)
    yp = y_pred > threshold
    alpha = 0.05
    z = stats.norm.ppf((1 - alpha / 2))
    tp = np.logical_and(yt, yp).sum()
    fp = np.logical_and(np.logical_not(yt), yp).sum()
    tn = np.logical_and(np.logical_not(yt), np.logical_not(yp)).sum()
    fn = np.logical_and(yt, np.logical_not(yp)).sum()
    acc = (tp + tn) / (tp + tn + fp + fn)
    sensitivity = tp / (tp + fn)
    specificity = tn / (tn + fp)

    n_boot, n_samp = 500, 150                                         # utils.py:430-431
    idx = np.empty((n_boot, n_samp), np.int64)
    population = np.arange(yt.shape[0])
    for b in range(n_boot):                                           # same RNG consumption as the reference
        idx[b] = np.random.choice(population, size=(n_samp,))
    ctx = _ffi.default_context()
    counts = np.empty((n_boot, 4), np.int64)
    yt8 = np.ascontiguousarray(yt, dtype=np.uint8)
    yp8 = np.ascontiguousarray(yp, dtype=np.uint8)
    _ffi.check(ctx.handle, ctx.lib.bq_bootstrap_confusion(ctx.handle, _ffi.ptr(yt8), _ffi.ptr(yp8), int(yt.shape[0]),
                                                          _ffi.ptr(idx), C.c_int32(n_boot), C.c_int32(n_samp),
                                                          _ffi.ptr(counts)), "bq_bootstrap_confusion")
    _tp, _fp, _tn, _fn = counts[:, 0], counts[:, 1], counts[:, 2], counts[:, 3]
    all_jac = (((_tn + 0.5 * z**2) / (_tn + _fp + z**2)) - ((_fn + 0.5 * z**2) / (_fn + _tp + z**2)))   # utils.py:438-439
    all_jac = list(all_jac)
    jac = mean(all_jac)
    jac_var = variance(all_jac)
    jac_low = jac - z * np.sqrt(jac_var)
    jac_high = jac + z * np.sqrt(jac_var)

    if not np.array_equal(np.unique(y_true), [0, 1]):                 # utils.py:448-450
        log.warning("Unable to calculate CI; NaNs exist")
        ci = [None, None]
    else:
        delong_auc, auc_cov = delong_roc_variance(y_true, y_pred)
        auc_std = np.sqrt(auc_cov)
        lower_upper_q = np.abs(np.array([0, 1]) - alpha / 2)
        ci = stats.norm.ppf(lower_upper_q, loc=delong_auc, scale=auc_std)
        ci[ci > 1] = 1
    return {"auc_low": ci[0], "auc_high": ci[1], "acc": acc, "sens": sensitivity, "spec": specificity,
            "youden": sensitivity + specificity - 1, "youden_low": jac_low, "youden_high": jac_high}
