"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic inputs for the hot path.

Tile-prediction tables shaped like the ones the reference reads back from Slideflow
(`tile_predictions_val_epoch1.csv|.parquet.gzip`, reference biscuit/utils.py:216-223, columns
renamed by biscuit/utils.py:31-53 to ``y_true, y_pred, uncertainty`` next to ``slide`` /
``patient``), synthetic uint8 WSI tiles and random-init Xception-UQ weights.

All generators are pure functions of their seed (numpy PCG64) so tests, the golden-vector
script and bench.py regenerate identical inputs on any box with this image.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

# ----------------------------------------------------------------------------------------
# tile tables
# ----------------------------------------------------------------------------------------


def tile_table(n_slides=100, tiles_per_slide=2000, seed=0, dtype=np.float32, ties=None,
               slides_per_patient=1, shuffle=False, ragged=False, prefix="s"):
    """One fold of tile-level predictions (BASELINE config 5 generator).

    y_true ~ Bernoulli(.5) per slide; slide mean mu = .5 +- .12 + N(0,.15);
    y_pred = clip(mu + N(0,.2), 0, 1); unc = .08 exp(-((p-.5)/.18)^2) + |N(0,.01)|.
    The noise level guarantees some misclassified slides so `from_cv` does not skip the fold.

    ties: if an int q, scores are rounded to multiples of 1/q (heavy-ties variant).
    ragged: tile counts vary per slide (1 .. 2*tiles_per_slide).
    shuffle: rows are permuted so slides are NOT contiguous (exercises the sorted reduce path).
    """
    rng = np.random.default_rng(seed)
    if ragged:
        counts = rng.integers(1, 2 * tiles_per_slide + 1, n_slides)
    else:
        counts = np.full(n_slides, tiles_per_slide)
    y_slide = rng.integers(0, 2, n_slides)
    if n_slides >= 2 and y_slide.min() == y_slide.max():
        y_slide[0] = 1 - y_slide[0]
    mu = 0.5 + 0.12 * (2 * y_slide - 1) + rng.normal(0, 0.15, n_slides)
    sid = np.repeat(np.arange(n_slides), counts)
    n = sid.shape[0]
    p = np.clip(mu[sid] + rng.normal(0, 0.2, n), 0, 1)
    u = 0.08 * np.exp(-((p - 0.5) / 0.18) ** 2) + np.abs(rng.normal(0, 0.01, n))
    if ties:
        p = np.round(p * ties) / ties
        u = np.round(u * ties * 10) / (ties * 10)
    names = np.array([f"{prefix}{i:05d}" for i in range(n_slides)], dtype=object)
    pat = np.array([f"p{i // slides_per_patient:05d}" for i in range(n_slides)], dtype=object)
    df = pd.DataFrame({
        "slide": names[sid],
        "y_true": y_slide[sid].astype(np.int64),
        "y_pred": p.astype(dtype),
        "uncertainty": u.astype(dtype),
        "patient": pat[sid],
    })
    if shuffle:
        df = df.iloc[rng.permutation(n)].reset_index(drop=True)
    return df


def cv_tables(k=10, n_slides=100, tiles_per_slide=2000, seed0=0, **kw):
    """k fold tables with seeds seed0..seed0+k-1 (config 5: 10 x 100 x 2000 = 2 M rows)."""
    return [tile_table(n_slides, tiles_per_slide, seed=seed0 + i, prefix=f"f{i}s", **kw)
            for i in range(k)]


def patients_map(df):
    """slide -> patient dict as `project.dataset().patients()` would give (utils.py:213)."""
    sub = df.drop_duplicates("slide")
    return dict(zip(sub["slide"], sub["patient"]))


# ----------------------------------------------------------------------------------------
# a Slideflow-shaped project tree for the nested-CV caller (reference experiment.py:924-1026)
# ----------------------------------------------------------------------------------------


class FakeDataset:
    def __init__(self, patients):
        self._patients = patients

    def patients(self):
        return dict(self._patients)


class FakeProject:
    """The two members of ``sf.Project`` the path touches: ``models_dir`` and ``dataset().patients()``."""

    def __init__(self, models_dir, patients):
        self.models_dir = models_dir
        self._patients = patients

    def dataset(self, verification=None):
        return FakeDataset(self._patients)


def _write_model(models_dir, idx, name, table, outcome, fmt, underscore):
    """One trained-model folder as Slideflow lays it out: `<5-digit id>-<name>/` holding the validation
    tile predictions (per-outcome column names, utils.py:19-28) and `slide_manifest.csv`."""
    import os
    folder = os.path.join(models_dir, f"{idx:05d}-{name}")
    os.makedirs(folder, exist_ok=True)
    sep = "_" if underscore else "-"
    out = pd.DataFrame({
        "slide": table["slide"],
        f"{outcome}{sep}y_true0": table["y_true"],
        f"{outcome}{sep}y_pred0": (1 - table["y_pred"]).astype(table["y_pred"].dtype),
        f"{outcome}{sep}y_pred1": table["y_pred"],
        f"{outcome}{sep}uncertainty0": table["uncertainty"],
        f"{outcome}{sep}uncertainty1": table["uncertainty"],
    })
    if fmt == "csv":
        out.to_csv(os.path.join(folder, "tile_predictions_val_epoch1.csv"), index=False)
    else:
        out.to_parquet(os.path.join(folder, "tile_predictions_val_epoch1.parquet.gzip"), compression="gzip")
    slides = table["slide"].drop_duplicates().tolist()
    with open(os.path.join(folder, "slide_manifest.csv"), "w") as f:
        f.write("slide,dataset\n")
        for i in range(3 * len(slides)):            # a model trains on more slides than it validates on
            f.write(f"train{i:05d},training\n")
        for sname in slides:
            f.write(f"{sname},validation\n")
    return folder


def nested_cv_project(root, label="EXP_AA_UQ", outcome="cohort", outer_k=3, inner_k=5, n_slides=40,
                      tiles_per_slide=50, seed0=500, fmt="csv", dtype=np.float32, underscore=False,
                      slides_per_patient=2, missing_outer=()):
    """Writes `outer_k` x `inner_k` inner-fold models (`{label}-k{k}`, kfold 1..inner_k) and `outer_k` outer
    models (`{label}`, kfold k) under `root`/models and returns a FakeProject.  Tables come from
    :func:`tile_table` with seeds seed0 + 100*k + j (j = 0 for the outer validation table).
    Outer folds listed in `missing_outer` get no inner models (the reference skips them)."""
    import os
    models_dir = os.path.join(root, "models")
    os.makedirs(models_dir, exist_ok=True)
    patients = {}
    idx = 1
    for k in range(1, outer_k + 1):
        outer = tile_table(n_slides, tiles_per_slide, seed=seed0 + 100 * k, dtype=dtype, prefix=f"o{k}s",
                           slides_per_patient=slides_per_patient)
        patients.update(patients_map(outer))
        _write_model(models_dir, idx, f"{outcome}-{label}-HP0-kfold{k}", outer, outcome, fmt, underscore)
        idx += 1
        if k in missing_outer:
            continue
        for j in range(1, inner_k + 1):
            inner = tile_table(n_slides, tiles_per_slide, seed=seed0 + 100 * k + j, dtype=dtype, prefix=f"i{k}{j}s",
                               slides_per_patient=slides_per_patient)
            patients.update(patients_map(inner))
            _write_model(models_dir, idx, f"{outcome}-{label}-k{k}-HP0-kfold{j}", inner, outcome, fmt, underscore)
            idx += 1
    return FakeProject(models_dir, patients)


# ----------------------------------------------------------------------------------------
# image tiles
# ----------------------------------------------------------------------------------------

TILE_PX = 299  # reference biscuit/hp.py:5


def tiles_u8(n_tiles, seed=0, n_slides=1, px=TILE_PX):
    """uint8 NHWC tiles [n,px,px,3]; per-slide colour bias + smooth structure + noise so that
    tiles (and slides) differ in mean/contrast and per-image standardisation matters."""
    rng = np.random.default_rng(seed)
    out = np.empty((n_tiles, px, px, 3), dtype=np.uint8)
    per = max(1, (n_tiles + n_slides - 1) // n_slides)
    bias = rng.uniform(90, 170, (n_slides, 3))
    yy, xx = np.meshgrid(np.arange(px), np.arange(px), indexing="ij")
    for i in range(n_tiles):
        s = min(i // per, n_slides - 1)
        fx, fy = rng.uniform(0.01, 0.08, 2)
        ph = rng.uniform(0, 6.28, 3)
        amp = rng.uniform(10, 60)
        base = np.stack([np.sin(fx * xx + fy * yy + ph[c]) for c in range(3)], -1) * amp
        img = bias[s] + base + rng.normal(0, 12, (px, px, 3))
        out[i] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out
