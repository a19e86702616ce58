"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the nested-CV threshold caller.

Follows /root/reference/biscuit/experiment.py:924-1026 (`Experiment.thresholds_from_nested_cv`) and the
loaders it uses, /root/reference/biscuit/utils.py:31-53 (`rename_cols`), 190-228 (`df_from_cv`), 233-311
(`find_model`, `find_cv`), on top of the pinned thresholding oracle (oracle/threshold_oracle.py).
`sf.util.get_slides_from_model_manifest` is a Slideflow function (slideflow>=1.1, not vendored): restated
from its documented file format (slide_manifest.csv with 'slide','dataset' columns); only the COUNT of
slides reaches the output.

Parity status: PINNED -- tests/test_nested_cv_cpu.py compares it with the unmodified reference executed
through oracle/ref_shim.py (when /root/reference is present) and with tests/golden/nested_cv_golden.json.
"""
from __future__ import annotations

import csv
import os
from os.path import exists, join
from statistics import mean

import pandas as pd

from . import threshold_oracle as O


class ModelNotFoundError(Exception):       # reference biscuit/errors.py:5
    pass


class MultipleModelsFoundError(Exception):  # reference biscuit/errors.py:9
    pass


def rename_cols(df, outcome):               # utils.py:31-53 (default column names only)
    def hdr(kind, underscore):
        return f"{outcome}{'_' if underscore else '-'}{kind}"
    yt = hdr("y_true0", hdr("y_true0", True) in df.columns)
    if yt not in df.columns:
        yt = f"{outcome}-y_true"
    yp = hdr("y_pred1", hdr("y_pred1", True) in df.columns)
    un = hdr("uncertainty1", hdr("uncertainty1", True) in df.columns)
    df.rename(columns={yt: "y_true", yp: "y_pred", un: "uncertainty"}, inplace=True)


def find_model(project, label, outcome, epoch=None, kfold=None):   # utils.py:233-272
    tail = "" if kfold is None else f"-kfold{kfold}"
    name = f"{outcome}-{label}-HP0{tail}"
    hits = [o for o in os.listdir(project.models_dir) if o[6:] == name]
    if len(hits) > 1:
        raise MultipleModelsFoundError(name)
    if not hits:
        raise ModelNotFoundError(name)
    if epoch is not None:
        return join(project.models_dir, hits[0], f"{name}_epoch{epoch}")
    return join(project.models_dir, hits[0])


def df_from_cv(project, label, outcome, k):   # utils.py:190-228 with epoch=None
    out = []
    patients = project.dataset().patients()
    for j in range(1, k + 1):
        folder = find_model(project, label, outcome, kfold=j)
        c, p = join(folder, "tile_predictions_val_epoch1.csv"), join(folder, "tile_predictions_val_epoch1.parquet.gzip")
        if exists(c):
            df = pd.read_csv(c)
        elif exists(p):
            df = pd.read_parquet(p)
        else:
            raise OSError(folder)
        rename_cols(df, outcome)
        if "patient" not in df.columns:
            df["patient"] = df["slide"].map(patients)
        out.append(df)
    return out


def manifest_slides(model_path):
    for folder in (model_path, os.path.dirname(os.path.normpath(model_path))):
        path = join(folder, "slide_manifest.csv")
        if exists(path):
            with open(path, newline="") as f:
                return [r["slide"] for r in csv.DictReader(f)]
    raise OSError(model_path)


def thresholds_from_nested_cv(project, label, outcome="cohort", outer_k=3, inner_k=5,
                              tile_filename="tile_predictions_val_epoch1.csv"):   # experiment.py:924-1026
    patients = project.dataset(verification=None).patients()
    params = {"tile_pred": "detect", "slide_pred": "detect", "plot": False, "patients": patients}
    t_uq, s_uq, s_pred, rows = [], [], [], []
    for k in range(1, outer_k + 1):
        try:
            dfs = df_from_cv(project, f"{label}-k{k}", outcome, inner_k)
        except ModelNotFoundError:
            continue
        val_path = join(find_model(project, label, outcome, kfold=k), tile_filename)
        if not exists(val_path):
            continue
        tile_uq = O.from_cv(dfs, tile_uq="detect", slide_uq=None, **params)["tile_uq"]
        th = O.from_cv(dfs, tile_uq=tile_uq, slide_uq="detect", **params)
        t_uq.append(tile_uq); s_uq.append(th["slide_uq"]); s_pred.append(th["slide_pred"])
        ext = val_path.rsplit(".", 1)[-1].lower()
        if ext == "csv":
            val = pd.read_csv(val_path, dtype={"slide": str})
        elif ext in ("parquet", "gzip"):
            val = pd.read_parquet(val_path)
        else:
            raise OSError(val_path)
        rename_cols(val, outcome)
        res_p, _ = O.apply(val, plot=False, patients=patients, level="patient", **th)
        res_s, _ = O.apply(val, plot=False, patients=patients, level="slide", **th)
        n = len(manifest_slides(find_model(project, label, outcome, kfold=k, epoch=1)))
        rows.append({"id": label, "n_slides": n, "fold": k, "uq": "include", "patient_auc": res_p["auc"],
                     "patient_uq_perc": res_p["percent_incl"], "slide_auc": res_s["auc"],
                     "slide_uq_perc": res_s["percent_incl"]})
    df = pd.DataFrame()
    for r in rows:
        df = pd.concat([df, pd.DataFrame([r])], axis=0, join="outer", ignore_index=True)
    return df, {"tile_uq": mean(t_uq) if t_uq else None, "slide_uq": mean(s_uq) if s_uq else None,
                "slide_pred": mean(s_pred) if s_pred else None}
