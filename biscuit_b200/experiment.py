"""The caller of the hot path: nested-cross-validation threshold detection
(reference biscuit/experiment.py:924-1026).

Only the part of the reference's ``Experiment`` that consumes tile-prediction tables is mirrored:
``thresholds_from_nested_cv`` -- per outer fold, two ``threshold.from_cv`` passes over the inner
folds (tile-level UQ threshold first, then the slide-level thresholds with that tile threshold
fixed) and two ``threshold.apply`` calls (patient and slide level) on the outer fold's validation
table.  Training, evaluation, plotting and the annotation-file book-keeping of the reference class
are Slideflow orchestration and stay out of scope.

`train_project` is any object with the two Slideflow ``Project`` members the path touches:
``models_dir`` (str) and ``dataset(verification=None).patients()`` (dict slide -> patient).  Every
threshold / ROC / slide reduction below runs in the CUDA library through :mod:`biscuit_b200.threshold`.
"""
from __future__ import annotations

import logging
from os.path import exists, join
from statistics import mean

import pandas as pd

from . import threshold, utils
from .errors import ModelNotFoundError

log = logging.getLogger("biscuit_b200")


class Experiment:
    """Holds the project and outcome names for threshold detection (reference experiment.py:49-85)."""

    def __init__(self, train_project, eval_projects=None, outcome="cohort", outcome1="LUAD", outcome2="LUSC",
                 outdir="results"):
        if isinstance(train_project, str) or not hasattr(train_project, "models_dir"):
            # the reference opens a path with sf.Project (experiment.py:63-64); Slideflow is not part of this
            # library, so a project object has to be supplied
            raise ValueError(f"Unrecognized value for train_project: {train_project}")
        self.train_project = train_project
        self.eval_projects = list(eval_projects or [])
        self.outcome = outcome
        self.outcome1 = outcome1
        self.outcome2 = outcome2
        self.outdir = outdir

    def thresholds_from_nested_cv(self, label, outer_k=3, inner_k=5, id=None, threshold_params=None, epoch=1,
                                  tile_filename="tile_predictions_val_epoch1.csv", y_true=None, y_pred=None,
                                  uncertainty=None):
        """Detects tile- and slide-level UQ thresholds and the slide-level prediction threshold from
        nested cross-validation (reference experiment.py:924-1026).

        Returns ``(df, thresholds)``: one row per usable outer fold (id, n_slides, fold, uq, patient_auc,
        patient_uq_perc, slide_auc, slide_uq_perc) and the mean over outer folds of tile_uq / slide_uq /
        slide_pred (None when no outer fold was usable).  Outer folds whose inner models or validation
        table are missing are skipped with a warning, as in the reference.

        Per outer fold the inner-fold tables are uploaded ONCE (`threshold.FoldSet`) and both detection passes run on the
        resident tables -- the second pass (tile threshold fixed, slide thresholds searched) reuses each fold's tile stage
        and slide factorisation and only re-runs the filter + slide reduction + slide ROCs; the outer fold's validation
        table is likewise resident for the patient- and the slide-level `apply`."""
        project = self.train_project
        index = utils.ProjectIndex(project)
        patients = project.dataset(verification=None).patients()
        params = dict(threshold_params) if threshold_params is not None else \
            {"tile_pred": "detect", "slide_pred": "detect", "plot": False, "patients": patients}
        headers = dict(y_true=y_true, y_pred=y_pred, uncertainty=uncertainty)
        found = {"tile_uq": [], "slide_uq": [], "slide_pred": []}
        report = []
        for fold in range(1, outer_k + 1):
            outer = self._outer_fold(index, label, fold, inner_k, tile_filename, headers)
            if outer is None:
                log.warning(f"Could not find {label} k-fold {fold}; skipping")       # :955-957, :963-965
                continue
            inner_tables, val_path = outer
            with threshold.FoldSet(inner_tables) as inner:
                tile_uq = inner.from_cv(tile_uq="detect", slide_uq=None, **params)["tile_uq"]      # :966-971
                cuts = inner.from_cv(tile_uq=tile_uq, slide_uq="detect", **params)                 # :972-977
            for key in found:
                found[key].append(tile_uq if key == "tile_uq" else cuts[key])
            validation = utils.read_tile_predictions(val_path)                        # :980-985
            utils.rename_cols(validation, self.outcome, **headers)
            with threshold.ResidentTable(_with_level_columns(validation, patients)) as table:
                per_level = {level: threshold.apply_resident(table, patients=patients, level=level, **cuts)[0]
                             for level in ("patient", "slide")}                       # :988-1001
            manifest = utils.slides_from_model_manifest(index.path(label, self.outcome, epoch=1, kfold=fold))   # :1003-1008
            report.append({"id": label if id is None else id, "n_slides": len(manifest), "fold": fold, "uq": "include",
                           "patient_auc": per_level["patient"]["auc"], "patient_uq_perc": per_level["patient"]["percent_incl"],
                           "slide_auc": per_level["slide"]["auc"], "slide_uq_perc": per_level["slide"]["percent_incl"]})
        df = pd.DataFrame()
        for row in report:                       # row-by-row outer concat: the dtypes pandas infers match the reference's (:1009-1018)
            df = pd.concat([df, pd.DataFrame([row])], axis=0, join="outer", ignore_index=True)
        return df, {key: (mean(vals) if vals else None) for key, vals in found.items()}          # :1021-1025

    def _outer_fold(self, index, label, fold, inner_k, tile_filename, headers):
        """-> (inner-fold tables, path of the outer fold's validation table), or None when a model / table is missing"""
        try:
            for k in range(1, inner_k + 1):
                index.folder(f"{label}-k{fold}", self.outcome, k)
            inner = [index.validation_table(f"{label}-k{fold}", self.outcome, k, headers=headers)
                     for k in range(1, inner_k + 1)]
            val_path = join(index.path(label, self.outcome, kfold=fold), tile_filename)
        except ModelNotFoundError:
            return None
        return (inner, val_path) if exists(val_path) else None


def _with_level_columns(table, patients):
    """the reference's `apply` adds the `patient` column itself (threshold.py:285-286); doing it before the table goes
    resident lets both levels share one upload"""
    if patients:
        table["patient"] = table["slide"].map(patients)
    return table
