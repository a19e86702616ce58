"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/nested_cv_golden.json.

Runs the UNMODIFIED reference `Experiment.thresholds_from_nested_cv` (/root/reference/biscuit/
experiment.py:924-1026, loaded through oracle/ref_shim.py) on seeded synthetic Slideflow-shaped project
trees (oracle/synth.py: nested_cv_project) and records its outputs bit-exactly (floats as hex).  Run in
the build container (the reference is not present on the GPU box):

    python -m oracle.make_golden_nested_cv
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import tempfile
import warnings

import numpy as np
import pandas as pd

from . import synth
from .make_golden import enc
from .ref_shim import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "nested_cv_golden.json")

CASES = {
    "csv_f64": dict(fmt="csv", dtype="float32", seed0=500),             # CSV round trip -> float64 columns
    "parquet_f32": dict(fmt="parquet", dtype="float32", seed0=700,     # parquet keeps float32
                        call=dict(tile_filename="tile_predictions_val_epoch1.parquet.gzip")),
    "parquet_default_filename": dict(fmt="parquet", dtype="float32", seed0=700),   # outer tables "missing" -> all folds skipped
    "csv_underscore_missing_fold": dict(fmt="csv", dtype="float64", seed0=900, underscore=True, missing_outer=[2]),
}


def build(root, kw):
    kw = dict(kw)
    kw.pop("call", None)
    kw["dtype"] = np.dtype(kw["dtype"]).type
    kw["missing_outer"] = tuple(kw.get("missing_outer", ()))
    return synth.nested_cv_project(root, **kw)


def run_reference(project_plain, call_kw):
    load_reference()
    import slideflow as sf
    ref_exp = importlib.import_module("biscuit.experiment")
    proj = type("P", (sf.Project, synth.FakeProject), {})(project_plain.models_dir, project_plain._patients)
    exp = ref_exp.Experiment(proj, outcome="cohort")
    return exp.thresholds_from_nested_cv("EXP_AA_UQ", outer_k=3, inner_k=5, **call_kw)


def main():
    warnings.simplefilter("ignore")
    import sklearn
    out = {"versions": {"numpy": np.__version__, "pandas": pd.__version__, "sklearn": sklearn.__version__,
                        "python": sys.version.split()[0]},
           "generator": "oracle/make_golden_nested_cv.py", "cases": {}}
    for name, kw in CASES.items():
        with tempfile.TemporaryDirectory() as root:
            df, th = run_reference(build(root, kw), kw.get("call", {}))
        rows = [{c: (enc(v) if not isinstance(v, str) else v) for c, v in r.items()} for r in df.to_dict("records")]
        out["cases"][name] = {"kwargs": kw, "thresholds": {k: enc(v) for k, v in th.items()}, "rows": rows,
                              "columns": list(df.columns), "dtypes": [str(t) for t in df.dtypes]}
        print(name, th, len(df))
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
