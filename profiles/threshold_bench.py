#!/usr/bin/env python
"""Absolute time of the thresholding rows on the GPU next to the CPU restatement of the
reference (oracle/threshold_oracle.py = the reference's arithmetic on the installed sklearn / pandas, single-threaded
like the reference).  These kernels move 13-21 B per tile row and are launch-latency-bound by construction (8d), so
the honest figures are milliseconds and a speed-up, not a roofline fraction.  Run under gpurun:

    python profiles/threshold_bench.py > gpurun_out/threshold_bench.md
"""
import os
import sys
import tempfile
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from biscuit_b200 import threshold as T          # noqa: E402
from biscuit_b200.experiment import Experiment   # noqa: E402
from oracle import nested_cv_oracle as NO, synth, threshold_oracle as O   # noqa: E402


def best(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), out


def main():
    rows = []
    # config 5: 10 fold tables x 100 slides x 2000 tiles (2 M rows), from_cv
    dfs = synth.cv_tables(k=10, n_slides=100, tiles_per_slide=2000, seed0=0)
    pats = {}
    for d in dfs:
        pats.update(synth.patients_map(d))
    T.from_cv([d.copy() for d in dfs], patients=pats)       # warm-up (module load, CUB temp storage)
    g, tg = best(lambda: T.from_cv([d.copy() for d in dfs], patients=pats), 3)
    c, tc = best(lambda: O.from_cv([d.copy() for d in dfs], patients=pats), 1)
    assert all(tg[k] == tc[k] for k in tg), (tg, tc)
    rows.append(("config 5: from_cv, 10 folds x 200 k rows (f32)", g, c))
    # config 3 shape: apply on 1000 slides x 2000 tiles = 2 M rows
    big = synth.tile_table(n_slides=1000, tiles_per_slide=2000, seed=7)
    th = dict(tile_uq=0.05, slide_uq=0.03, tile_pred=0.5, slide_pred=0.5)
    T.apply(big.copy(), **th)
    g, rg = best(lambda: T.apply(big.copy(), **th), 3)
    c, rc = best(lambda: O.apply(big.copy(), **th), 1)
    assert all(rg[0][k] == rc[0][k] for k in rg[0]), (rg[0], rc[0])
    rows.append(("config 3: apply, 1000 slides x 2000 tiles (2 M rows, f32)", g, c))
    # config 2 shape: apply on one 10 k-tile slide (what bench.py runs per step) -- single class, so detect a 4-slide table instead
    small = synth.tile_table(n_slides=4, tiles_per_slide=128, seed=3)
    T.detect(small.copy())
    g, _ = best(lambda: T.detect(small.copy()), 5)
    c, _ = best(lambda: O.detect(small.copy()), 3)
    rows.append(("config 1: detect, 4 slides x 128 tiles (512 rows)", g, c))
    # the nested-CV caller on a synthetic project tree (3 outer x 5 inner folds x 100 slides x 500 tiles, parquet)
    with tempfile.TemporaryDirectory() as root:
        project = synth.nested_cv_project(root, n_slides=100, tiles_per_slide=500, seed0=4000, fmt="parquet")
        fname = "tile_predictions_val_epoch1.parquet.gzip"
        exp = Experiment(project, outcome="cohort")
        exp.thresholds_from_nested_cv("EXP_AA_UQ", tile_filename=fname)
        g, og = best(lambda: exp.thresholds_from_nested_cv("EXP_AA_UQ", tile_filename=fname), 2)
        c, oc = best(lambda: NO.thresholds_from_nested_cv(project, "EXP_AA_UQ", tile_filename=fname), 1)
        assert all(og[1][k] == oc[1][k] for k in og[1])
    rows.append(("f1: thresholds_from_nested_cv, 3 x (5 + 1) tables x 50 k rows (incl. parquet I/O)", g, c))
    print("# thresholding rows: GPU (C ABI, host DataFrames in, DataFrames out) vs CPU restatement of the reference\n")
    print(f"host: {os.cpu_count()} cores; the reference's thresholding code is single-threaded Python + sklearn + pandas\n")
    print("| case | GPU path ms | reference arithmetic on CPU ms | ratio | results |")
    print("|---|---|---|---|---|")
    for name, g, c in rows:
        print(f"| {name} | {g * 1e3:.1f} | {c * 1e3:.1f} | {c / g:.1f}x | identical |")


if __name__ == "__main__":
    main()
