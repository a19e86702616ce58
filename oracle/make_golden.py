"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/threshold_golden.json.

Runs the UNMODIFIED reference (`/root/reference/biscuit/threshold.py`, loaded through
oracle/ref_shim.py) on seeded synthetic tile tables and records its outputs bit-exactly (floats
as hex).  Run in the build container (the reference is not present on the GPU box):

    python -m oracle.make_golden

Library versions are recorded because the arithmetic flows through sklearn / pandas / numpy.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import warnings

import numpy as np
import pandas as pd

from . import synth
from .ref_shim import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "threshold_golden.json")

# name -> (generator kwargs, extra)
CASES = {
    "f32_small": dict(n_slides=24, tiles_per_slide=60, seed=11, dtype="float32"),
    "f64_small": dict(n_slides=24, tiles_per_slide=60, seed=12, dtype="float64"),
    "f32_ties": dict(n_slides=30, tiles_per_slide=80, seed=13, dtype="float32", ties=50),
    "f64_ties": dict(n_slides=30, tiles_per_slide=80, seed=14, dtype="float64", ties=20),
    "f32_ragged_shuffled": dict(n_slides=20, tiles_per_slide=50, seed=15, dtype="float32", ragged=True,
                                shuffle=True),
    "f32_patients": dict(n_slides=36, tiles_per_slide=40, seed=16, dtype="float32", slides_per_patient=3),
    "f32_config5_fold": dict(n_slides=100, tiles_per_slide=2000, seed=0, dtype="float32"),
    "f64_config5_fold": dict(n_slides=100, tiles_per_slide=2000, seed=1, dtype="float64"),
}


def make_table(kw):
    kw = dict(kw)
    kw["dtype"] = np.dtype(kw["dtype"]).type
    return synth.tile_table(**kw)


def enc(v):
    """bit-exact JSON encoding of scalars"""
    if v is None:
        return None
    if isinstance(v, (float, np.floating)):
        return {"t": type(v).__name__, "hex": float(v).hex()}
    if isinstance(v, (int, np.integer)):
        return {"t": type(v).__name__, "int": int(v)}
    raise TypeError(type(v))


def enc_df(df, full):
    if df is None:
        return None
    d = {"columns": list(df.columns), "dtypes": [str(t) for t in df.dtypes], "n": len(df)}
    h = hashlib.sha256()
    h.update(np.asarray(df.index, dtype=np.int64).tobytes())
    cols = {}
    for c in df.columns:
        a = df[c].to_numpy()
        if a.dtype.kind in "OUT" or str(df[c].dtype) == "str":
            a = np.array([str(x) for x in a])
            h.update("|".join(a.tolist()).encode())
            if full:
                cols[c] = a.tolist()
        else:
            h.update(np.ascontiguousarray(a).tobytes())
            if full:
                cols[c] = [float(x).hex() for x in a] if a.dtype.kind == "f" else [int(x) for x in a]
    d["sha256"] = h.hexdigest()
    if full:
        d["index"] = [int(i) for i in df.index]
        d["values"] = cols
    return d


def main():
    warnings.simplefilter("ignore")
    R = load_reference().threshold
    import sklearn
    out = {"versions": {"numpy": np.__version__, "pandas": pd.__version__, "sklearn": sklearn.__version__,
                        "python": sys.version.split()[0]},
           "generator": "oracle/make_golden.py", "cases": {}}
    for name, kw in CASES.items():
        df = make_table(kw)
        full = len(df) <= 5000
        case = {"kwargs": kw, "n_rows": len(df),
                "input_sha256": hashlib.sha256(
                    df["y_pred"].to_numpy().tobytes() + df["uncertainty"].to_numpy().tobytes() +
                    df["y_true"].to_numpy().tobytes()).hexdigest()}
        th, auc = R.detect(df.copy())
        case["detect"] = {"thresholds": {k: enc(v) for k, v in th.items()}, "auc": enc(auc)}
        if th["tile_uq"] is None or th["slide_uq"] is None:
            th = {"tile_uq": 0.05, "slide_uq": 0.03, "tile_pred": 0.5, "slide_pred": 0.5}
        pats = synth.patients_map(df)
        case["apply"] = {}
        for level in ("slide", "patient"):
            for keep in ("high_confidence", "low_confidence"):
                d2 = df.copy()
                res, s_df = R.apply(d2, **th, keep=keep, patients=pats, level=level)
                case["apply"][f"{level}/{keep}"] = {
                    "results": {k: enc(v) for k, v in res.items()},
                    "s_df": enc_df(s_df, full),
                    "tile_df_sha256": enc_df(d2, False)["sha256"],
                }
        # python-float thresholds (weak scalars -> compared in the column dtype)
        d3 = df.copy()
        res, s_df = R.apply(d3, 0.045, 0.031, tile_pred=0.5, slide_pred=0.45)
        case["apply_pyfloat"] = {"results": {k: enc(v) for k, v in res.items()}, "s_df": enc_df(s_df, full)}
        out["cases"][name] = case
        print(name, len(df), case["detect"]["thresholds"]["tile_uq"])
    # from_cv over 4 small folds + the config-5 shape at reduced tile count
    dfs = synth.cv_tables(k=5, n_slides=40, tiles_per_slide=150, seed0=100)
    out["from_cv"] = {"kwargs": dict(k=5, n_slides=40, tiles_per_slide=150, seed0=100),
                      "all_detect": {k: enc(v) for k, v in R.from_cv([d.copy() for d in dfs]).items()}}
    r1 = R.from_cv([d.copy() for d in dfs], tile_uq="detect", slide_uq=None, tile_pred="detect",
                   slide_pred="detect")
    out["from_cv"]["tile_only"] = {k: enc(v) for k, v in r1.items()}
    r2 = R.from_cv([d.copy() for d in dfs], tile_uq=r1["tile_uq"], slide_uq="detect", tile_pred="detect",
                   slide_pred="detect")
    out["from_cv"]["nested_second"] = {k: enc(v) for k, v in r2.items()}
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
