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
        table are missing are skipped with a warning, as in the reference."""
        if id is None:
            id = label
        project = self.train_project
        patients = project.dataset(verification=None).patients()
        if threshold_params is None:
            threshold_params = {"tile_pred": "detect", "slide_pred": "detect", "plot": False, "patients": patients}
        all_tile_uq, all_slide_uq, all_slide_pred = [], [], []
        rows = []
        for k in range(1, outer_k + 1):
            try:
                dfs = utils.df_from_cv(project, f"{label}-k{k}", outcome=self.outcome, k=inner_k, y_true=y_true,
                                       y_pred=y_pred, uncertainty=uncertainty)        # :946-954
            except ModelNotFoundError:
                log.warning(f"Could not find {label} k-fold {k}; skipping")            # :955-957
                continue
            val_path = join(utils.find_model(project, f"{label}", kfold=k, outcome=self.outcome), tile_filename)
            if not exists(val_path):                                                   # :963-965
                log.warning(f"Could not find {label} k-fold {k}; skipping")
                continue
            tile_uq = threshold.from_cv(dfs, tile_uq="detect", slide_uq=None, **threshold_params)["tile_uq"]   # :966-971
            thresholds = threshold.from_cv(dfs, tile_uq=tile_uq, slide_uq="detect", **threshold_params)        # :972-977
            all_tile_uq.append(tile_uq)
            all_slide_uq.append(thresholds["slide_uq"])
            all_slide_pred.append(thresholds["slide_pred"])
            tile_pred_df = utils.read_tile_predictions(val_path)                       # :980-985
            utils.rename_cols(tile_pred_df, self.outcome, y_true=y_true, y_pred=y_pred, uncertainty=uncertainty)

            def uq_auc_by_level(level):                                                # :988-996
                results, _ = threshold.apply(tile_pred_df, plot=False, patients=patients, level=level, **thresholds)
                return results["auc"], results["percent_incl"]

            pt_auc, pt_perc = uq_auc_by_level("patient")
            slide_auc, slide_perc = uq_auc_by_level("slide")
            model = utils.find_model(project, f"{label}", kfold=k, epoch=1, outcome=self.outcome)
            m_slides = utils.slides_from_model_manifest(model, dataset=None)          # :1008
            rows.append({"id": id, "n_slides": len(m_slides), "fold": k, "uq": "include", "patient_auc": pt_auc,
                         "patient_uq_perc": pt_perc, "slide_auc": slide_auc, "slide_uq_perc": slide_perc})
        df = pd.DataFrame()
        for row in rows:                                                               # same concat as :1009-1018
            df = pd.concat([df, pd.DataFrame([row])], axis=0, join="outer", ignore_index=True)
        thresholds = {
            "tile_uq": None if not all_tile_uq else mean(all_tile_uq),
            "slide_uq": None if not all_slide_uq else mean(all_slide_uq),
            "slide_pred": None if not all_slide_pred else mean(all_slide_pred),
        }
        return df, thresholds
