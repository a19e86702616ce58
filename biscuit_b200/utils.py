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
#
# Slideflow lays a project out as  models_dir/<5-digit id>-<outcome>-<label>-HP0[-kfold<k>]/  with the saved epoch in a
# sub-folder of the same name + "_epoch<e>" and the tile predictions of the validation set next to it
# (`tile_predictions_val_epoch1.csv | .parquet.gzip`).  The reference lists the directory again for every lookup
# (utils.py:233-272 is called 2-3 times per fold); here the directory is indexed ONCE per project object and every lookup
# is a dictionary access.  Names, precedence (csv before parquet) and exceptions are the reference's.

_VAL_TABLE = "tile_predictions_val_epoch1"


class ProjectIndex:
    """name -> folder index of one Slideflow project's ``models_dir`` (built by a single directory listing)."""

    def __init__(self, project):
        self.root = project.models_dir
        self.by_name = {}
        for entry in os.listdir(self.root):
            self.by_name.setdefault(entry[6:], []).append(entry)       # strip the '00042-' run id
        self._patients = None
        self._project = project

    @staticmethod
    def model_name(label, outcome, kfold=None):
        return f"{outcome}-{label}-HP0" + ("" if kfold is None else f"-kfold{kfold}")

    def folder(self, label, outcome, kfold=None):
        name = self.model_name(label, outcome, kfold)
        hits = self.by_name.get(name, ())
        if len(hits) > 1:
            raise MultipleModelsFoundError(f"Multiple matching models found matching {name}")
        if not hits:
            raise ModelNotFoundError(f"No matching model found matching {name}.")
        return join(self.root, hits[0])

    def path(self, label, outcome, epoch=None, kfold=None):
        """the run folder, or the saved model of `epoch` inside it"""
        run = self.folder(label, outcome, kfold)
        return run if epoch is None else join(run, f"{self.model_name(label, outcome, kfold)}_epoch{epoch}")

    def patients(self):
        if self._patients is None:
            self._patients = self._project.dataset().patients()
        return self._patients

    def validation_table(self, label, outcome, kfold, epoch=None, headers=None):
        """tile predictions of one fold's validation set with the canonical column names and a `patient` column"""
        where = self.path(label, outcome, epoch=epoch, kfold=kfold)
        base = where if epoch is None else os.path.dirname(os.path.normpath(where))
        for suffix, reader in ((".csv", pd.read_csv), (".parquet.gzip", pd.read_parquet)):
            candidate = join(base, _VAL_TABLE + suffix)
            if os.path.exists(candidate):
                table = reader(candidate)
                break
        else:
            raise OSError(f"Could not find tile predictions file at {where}")
        rename_cols(table, outcome, **(headers or {}))
        if "patient" not in table.columns:
            table["patient"] = table["slide"].map(self.patients())
        return table


def _index(project) -> ProjectIndex:
    return ProjectIndex(project)


def find_model(project, label, outcome, epoch=None, kfold=None):
    """Path of the trained model `{outcome}-{label}-HP0[-kfold{k}]` inside ``project.models_dir``; with ``epoch`` the
    saved-model sub-folder, else the run folder.  ModelNotFoundError / MultipleModelsFoundError as the reference
    (utils.py:233-272)."""
    return _index(project).path(label, outcome, epoch=epoch, kfold=kfold)


def model_exists(project, label, outcome, epoch=None, kfold=None):
    """Whether exactly one such model exists; several matches still raise (reference utils.py:275-292)."""
    try:
        _index(project).folder(label, outcome, kfold)
    except ModelNotFoundError:
        return False
    return True


def find_cv(project, label, outcome, epoch=None, k=3):
    """Paths of the k cross-validation models of one experiment (reference utils.py:295-311)."""
    index = _index(project)
    return [index.path(label, outcome, epoch=epoch, kfold=fold) for fold in range(1, k + 1)]


def read_tile_predictions(path):
    """One Slideflow tile-prediction table: ``.csv`` (slide column forced to str, reference
    experiment.py:980-981) or ``.parquet`` / ``.parquet.gzip`` (982-983); anything else is an
    OSError (984-985)."""
    kind = path.rsplit(".", 1)[-1].lower()
    readers = {"csv": lambda f: pd.read_csv(f, dtype={"slide": str}), "parquet": pd.read_parquet, "gzip": pd.read_parquet}
    if kind not in readers:
        raise OSError(f"Unrecognized prediction filetype {path}")
    return readers[kind](path)


def df_from_cv(project, label, outcome, epoch=None, k=3, y_true=None, y_pred=None, uncertainty=None):
    """The k validation tables of cross-validated experiment `label`, ready for `threshold.from_cv`: columns renamed
    to y_true / y_pred / uncertainty, `patient` filled from the project's slide -> patient map when the file has none;
    a float64 table when it came from CSV, float32 from parquet (reference utils.py:190-228)."""
    index = _index(project)
    for fold in range(1, k + 1):
        index.folder(label, outcome, fold)                              # every fold is looked up before any file is read
    headers = dict(y_true=y_true, y_pred=y_pred, uncertainty=uncertainty)
    return [index.validation_table(label, outcome, fold, epoch=epoch, headers=headers) for fold in range(1, k + 1)]


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

_N_BOOTSTRAP, _BOOTSTRAP_ROWS, _ALPHA = 500, 150, 0.05            # reference utils.py:420,430-431


def _confusion(labels, calls):
    """(tp, fp, tn, fn) of boolean labels / calls as numpy integers"""
    cell = np.bincount(2 * labels.astype(np.int64) + calls.astype(np.int64), minlength=4)
    return cell[3], cell[1], cell[0], cell[2]


def _bootstrap_confusions(labels, calls, draws):
    """confusion counts [n_boot, 4] = (tp, fp, tn, fn) of every bootstrap resample, on the GPU (bq_bootstrap_confusion)"""
    import ctypes as C

    from . import _ffi
    ctx = _ffi.default_context()
    n_boot, n_rows = draws.shape
    counts = np.empty((n_boot, 4), np.int64)
    lab8, call8 = np.ascontiguousarray(labels, dtype=np.uint8), np.ascontiguousarray(calls, dtype=np.uint8)
    draws = np.ascontiguousarray(draws, dtype=np.int64)
    _ffi.check(ctx.handle, ctx.lib.bq_bootstrap_confusion(ctx.handle, _ffi.ptr(lab8), _ffi.ptr(call8), int(labels.shape[0]),
                                                          _ffi.ptr(draws), C.c_int32(n_boot), C.c_int32(n_rows),
                                                          _ffi.ptr(counts)), "bq_bootstrap_confusion")
    return counts


def prediction_metrics(y_true, y_pred, threshold):
    """AUC confidence interval (DeLong), accuracy, sensitivity / specificity and Youden's J with its bootstrap confidence
    interval (reference utils.py:400-464).

    The 500 x 150 bootstrap row draws come from ONE ``np.random.choice`` call -- the legacy generator fills it element by
    element, so a seeded ``np.random`` yields exactly the rows of the reference's 500 consecutive calls -- and all 500
    confusion matrices are counted by one kernel launch; the DeLong placement values are computed on the GPU as well.
    The order-sensitive floating-point finish (``statistics.mean`` / ``variance`` over the per-resample Wilson-adjusted J,
    ``np.cov``, ``scipy.stats.norm``) uses the library calls the reference uses, in its order: bit-identical results."""
    from statistics import mean, variance

    from scipy import stats

    from .delong import delong_roc_variance

    labels, calls = y_true.astype(bool), y_pred > threshold
    z = stats.norm.ppf(1 - _ALPHA / 2)
    tp, fp, tn, fn = _confusion(labels, calls)
    sensitivity, specificity = tp / (tp + fn), tn / (tn + fp)

    draws = np.random.choice(np.arange(labels.shape[0]), size=(_N_BOOTSTRAP, _BOOTSTRAP_ROWS))
    b = _bootstrap_confusions(labels, calls, draws)
    b_tp, b_fp, b_tn, b_fn = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    half = 0.5 * z**2                                                   # Wilson-type adjustment of both rates (:438-439)
    j_boot = list(((b_tn + half) / (b_tn + b_fp + z**2)) - ((b_fn + half) / (b_fn + b_tp + z**2)))
    j_mean, j_sd = mean(j_boot), np.sqrt(variance(j_boot))

    auc_low = auc_high = None
    if np.array_equal(np.unique(y_true), [0, 1]):
        auc, auc_var = delong_roc_variance(y_true, y_pred)
        interval = stats.norm.ppf(np.abs(np.array([0, 1]) - _ALPHA / 2), loc=auc, scale=np.sqrt(auc_var))
        interval[interval > 1] = 1
        auc_low, auc_high = interval
    else:                                                               # :448-450
        log.warning("Unable to calculate CI; NaNs exist")
    return {"auc_low": auc_low, "auc_high": auc_high, "acc": (tp + tn) / (tp + tn + fp + fn), "sens": sensitivity,
            "spec": specificity, "youden": sensitivity + specificity - 1,
            "youden_low": j_mean - z * j_sd, "youden_high": j_mean + z * j_sd}
