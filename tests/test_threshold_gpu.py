"""GPU tier: biscuit_b200.threshold (CUDA through the C ABI) must be BIT-EXACT against the oracle
(oracle/threshold_oracle.py, itself pinned to the unmodified reference) and against the committed
outputs of the reference in tests/golden/threshold_golden.json."""
import warnings

import numpy as np
import pandas as pd
import pytest

from oracle import synth, threshold_oracle as O
from oracle.make_golden import CASES, make_table

from helpers import assert_same_df, assert_same_results, dec, df_sha, load_golden, same_scalar

pytestmark = pytest.mark.gpu
warnings.simplefilter("ignore")


@pytest.fixture(scope="module")
def T():
    from biscuit_b200 import threshold
    return threshold


@pytest.fixture(scope="module")
def GOLD():
    return load_golden()


def _cases():
    out = []
    for seed in range(24):
        out.append(dict(n_slides=6 + 3 * seed, tiles_per_slide=20 + 7 * seed, seed=1000 + seed,
                        dtype=[np.float32, np.float64][seed % 2], ties=[None, 40, 7][seed % 3],
                        shuffle=seed % 4 == 1, ragged=seed % 5 == 2, slides_per_patient=1 + seed % 3))
    return out


@pytest.mark.parametrize("kw", _cases(), ids=lambda k: f"seed{k['seed']}")
def test_detect_and_apply_bit_exact_vs_oracle(T, kw):
    df = synth.tile_table(**kw)
    a, b = df.copy(), df.copy()
    ra, rb = O.detect(a), T.detect(b)
    assert_same_results(ra[0], rb[0], "detect thresholds")
    assert same_scalar(ra[1], rb[1]), ("detect auc", ra[1], rb[1])
    assert_same_df(a, b, "tile frame mutated by detect")
    th = ra[0]
    if th["tile_uq"] is None or th["slide_uq"] is None:
        th = dict(tile_uq=0.05, slide_uq=0.03, tile_pred=0.5, slide_pred=0.5)
    pats = synth.patients_map(df)
    for level in ("slide", "patient"):
        for keep in ("high_confidence", "low_confidence"):
            a, b = df.copy(), df.copy()
            x = O.apply(a, **th, keep=keep, patients=pats, level=level)
            y = T.apply(b, **th, keep=keep, patients=pats, level=level)
            assert_same_results(x[0], y[0], f"apply {level}/{keep}")
            assert_same_df(x[1], y[1], f"apply s_df {level}/{keep}")
            assert_same_df(a, b, "tile frame mutated by apply")


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_reference_outputs(T, GOLD, name):
    case = GOLD["cases"][name]
    df = make_table(CASES[name])
    th, auc = T.detect(df.copy())
    for k, v in case["detect"]["thresholds"].items():
        assert same_scalar(th[k], dec(v)), (name, k, th[k], dec(v))
    assert same_scalar(auc, dec(case["detect"]["auc"]))
    if th["tile_uq"] is None or th["slide_uq"] is None:
        th = {"tile_uq": 0.05, "slide_uq": 0.03, "tile_pred": 0.5, "slide_pred": 0.5}
    pats = synth.patients_map(df)
    for key, g in case["apply"].items():
        level, keep = key.split("/")
        d2 = df.copy()
        res, s_df = T.apply(d2, **th, keep=keep, patients=pats, level=level)
        assert_same_results(res, {k: dec(v) for k, v in g["results"].items()}, f"{name}/{key}")
        assert df_sha(s_df) == g["s_df"]["sha256"], f"{name}/{key}: group frame differs from the reference"
        assert df_sha(d2) == g["tile_df_sha256"], f"{name}/{key}: mutated tile frame differs"
        assert [str(t) for t in s_df.dtypes] == g["s_df"]["dtypes"]
    res, s_df = T.apply(df.copy(), 0.045, 0.031, tile_pred=0.5, slide_pred=0.45)
    assert_same_results(res, {k: dec(v) for k, v in case["apply_pyfloat"]["results"].items()}, name)
    assert df_sha(s_df) == case["apply_pyfloat"]["s_df"]["sha256"]


def test_from_cv_golden(T, GOLD):
    g = GOLD["from_cv"]
    dfs = synth.cv_tables(**g["kwargs"])
    assert_same_results(T.from_cv([d.copy() for d in dfs]), {k: dec(v) for k, v in g["all_detect"].items()})
    r1 = T.from_cv([d.copy() for d in dfs], tile_uq="detect", slide_uq=None, tile_pred="detect",
                   slide_pred="detect")
    assert_same_results(r1, {k: dec(v) for k, v in g["tile_only"].items()})
    r2 = T.from_cv([d.copy() for d in dfs], tile_uq=r1["tile_uq"], slide_uq="detect", tile_pred="detect",
                   slide_pred="detect")
    assert_same_results(r2, {k: dec(v) for k, v in g["nested_second"].items()})


def test_scalar_promotion_boundary(T):
    """float32 column vs python float (weak: compared in float32) and np.float64 (compared in float64)"""
    df = synth.tile_table(12, 40, seed=77)
    u = df["uncertainty"].to_numpy()
    pivot = float(np.sort(u)[len(u) // 2])                  # exactly a float32 value
    just_above = np.nextafter(np.float64(pivot), 1.0)        # rounds back to pivot in float32
    for t in (pivot, float(just_above), np.float64(just_above), np.float32(pivot)):
        a, b = df.copy(), df.copy()
        x, y = O.apply(a, t, 0.03), T.apply(b, t, 0.03)
        assert_same_results(x[0], y[0], f"tile_uq={t!r}")
        assert_same_df(x[1], y[1], f"tile_uq={t!r}")
    s = O.apply(df.copy(), 0.05, 0.03)[1]["uncertainty"].to_numpy()
    sp = float(np.sort(s)[len(s) // 2])
    for t in (sp, float(np.nextafter(np.float64(sp), 1.0)), np.float64(np.nextafter(np.float64(sp), 1.0))):
        x, y = O.apply(df.copy(), 0.05, t), T.apply(df.copy(), 0.05, t)
        assert_same_results(x[0], y[0], f"slide_uq={t!r}")
        assert_same_df(x[1], y[1], f"slide_uq={t!r}")


def test_error_behaviour(T):
    from biscuit_b200 import errors
    df = synth.tile_table(8, 20, seed=3)
    bad = df.copy()
    bad.loc[3, "y_pred"] = np.nan
    with pytest.raises(errors.PredsContainNaNError):
        T.apply(bad.copy(), 0.05, 0.03)
    assert T.detect(bad.copy()) == ({k: None for k in ("tile_uq", "slide_uq", "tile_pred", "slide_pred")}, None)
    with pytest.raises(TypeError):
        T.apply(df.copy(), None, 0.03)
    with pytest.raises(AssertionError):
        T.apply(df.copy(), 0.05, 0.03, keep="medium")
    with pytest.raises(AssertionError):
        T.apply(df.copy(), 0.05, 0.03, level="patient")
    res, s = T.apply(df.copy(), 1e-9, 0.03)                  # every tile filtered -> no ROC
    assert s is None and all(v is None for v in res.values())
    with pytest.raises(ValueError):
        T.from_cv([df.drop(columns=["patient"])])
    sep = df.copy()
    sep["y_pred"] = sep["y_true"].astype(np.float32) * 0.8 + 0.1
    with pytest.raises(ValueError):                          # single-class tile-UQ ROC (App. A.1)
        T.from_cv([sep.copy()])
    easy = df.copy()
    rng = np.random.default_rng(5)
    easy["y_pred"] = (0.5 + 0.1 * (2 * easy["y_true"] - 1) + rng.normal(0, 0.12, len(easy))).astype(np.float32)
    with pytest.raises(errors.ThresholdError):
        T.from_cv([easy.copy()])
    # falsy thresholds disable the filters (threshold.py:297,323)
    x, y = O.apply(df.copy(), 0.0, 0), T.apply(df.copy(), 0.0, 0)
    assert_same_results(x[0], y[0])
    assert_same_df(x[1], y[1])
    # single-class cohort: AUC is NaN, sens or spec is 0/0
    one = df[df["y_true"] == df["y_true"].iloc[0]].copy()
    x, y = O.apply(one.copy(), 0.05, 0.03), T.apply(one.copy(), 0.05, 0.03)
    assert_same_results(x[0], y[0])
    assert_same_df(x[1], y[1])


def test_process_helpers(T):
    df = synth.tile_table(15, 33, seed=9, shuffle=True)
    a, b = df.copy(), df.copy()
    (_, ta), (_, tb) = O.process_tile_predictions(a, "detect"), T.process_tile_predictions(b, "detect")
    assert same_scalar(ta, tb)
    assert_same_df(a, b)
    pats = synth.patients_map(df)
    a, b = df.copy(), df.copy()
    O.process_tile_predictions(a, 0.4, pats), T.process_tile_predictions(b, 0.4, pats)
    assert_same_df(a, b)
    for thr in (0.5, "detect", np.float64(0.43)):
        (ga, pa), (gb, pb) = O.process_group_predictions(a, thr, "slide"), T.process_group_predictions(b, thr, "slide")
        assert same_scalar(pa, pb)
        assert_same_df(ga, gb)


def test_config5_full_size_properties(T):
    """2 M rows (10 folds x 100 slides x 2000 tiles): every fold against the oracle (the reference's
    arithmetic needs ~0.1 s per 200 k-row fold), then size-independent properties: from_cv == (min,
    max, mean, mean) of per-fold detect; apply idempotent; percent_incl consistent with the
    kept-slide frame; row-permutation invariance of detected thresholds."""
    dfs = synth.cv_tables(k=10, n_slides=100, tiles_per_slide=2000)
    per = [T.detect(d.copy())[0] for d in dfs]
    for k, d in enumerate(dfs):
        assert_same_results(O.detect(d.copy())[0], per[k], f"fold {k} vs oracle")
    assert_same_results(O.from_cv([d.copy() for d in dfs]), T.from_cv([d.copy() for d in dfs]), "from_cv vs oracle")
    cv = T.from_cv([d.copy() for d in dfs])
    ok = [p for p in per if p["tile_uq"] is not None and p["slide_uq"] is not None]
    assert cv["tile_uq"] == min(p["tile_uq"] for p in ok)
    assert cv["slide_uq"] == max(p["slide_uq"] for p in ok)
    assert cv["tile_pred"] == np.mean([p["tile_pred"] for p in ok])
    assert cv["slide_pred"] == np.mean([p["slide_pred"] for p in ok])
    big = pd.concat(dfs, ignore_index=True)
    r1, s1 = T.apply(big.copy(), **cv)
    r2, s2 = T.apply(big.copy(), **cv)
    assert_same_results(r1, r2)
    assert_same_df(s1, s2)
    assert r1["percent_incl"] == len(s1) / 1000
    assert (s1["uncertainty"] < cv["slide_uq"]).all()
    # ROC thresholds do not depend on row order
    perm = dfs[1].sample(frac=1.0, random_state=0).reset_index(drop=True)
    pa = T.detect(perm)[0]
    assert pa["tile_uq"] == per[1]["tile_uq"] and pa["tile_pred"] == per[1]["tile_pred"]


@pytest.mark.parametrize("dtype,ties", [(np.float32, 50), (np.float64, 50), (np.float64, None)],
                         ids=["f32_ties", "f64_ties", "f64"])
def test_config5_full_size_variants_vs_oracle(T, dtype, ties):
    """SURVEY 8d's other config-5 tables at FULL fold size (200 k rows): the float64 (CSV-loaded) variant
    and the heavy-ties variant (scores rounded to 1/50: thousands of equal scores per ROC boundary, equal
    slide means), detect + apply on the detected thresholds, bit-for-bit against the oracle."""
    dfs = synth.cv_tables(k=3, n_slides=100, tiles_per_slide=2000, seed0=40, dtype=dtype, ties=ties)
    for k, d in enumerate(dfs):
        want, got = O.detect(d.copy()), T.detect(d.copy())
        assert_same_results(want[0], got[0], f"fold {k} thresholds")
        assert same_scalar(want[1], got[1])
    try:
        cv_o = O.from_cv([d.copy() for d in dfs])
    except Exception as e:  # every fold skipped: the product must raise the same error type
        with pytest.raises(type(e)):
            T.from_cv([d.copy() for d in dfs])
        return
    cv = T.from_cv([d.copy() for d in dfs])
    assert_same_results(cv_o, cv, "from_cv")
    big = pd.concat(dfs, ignore_index=True)
    (ro, so), (rg, sg) = O.apply(big.copy(), **cv_o), T.apply(big.copy(), **cv)
    assert_same_results(ro, rg, "apply")
    assert_same_df(so, sg)


def test_apply_sharded_single_rank_equals_apply(T):
    for kw in (dict(n_slides=20, tiles_per_slide=60, seed=31), dict(n_slides=14, tiles_per_slide=45, seed=32, ragged=True)):
        df = synth.tile_table(**kw)
        th = dict(tile_uq=0.05, slide_uq=np.float64(0.031), tile_pred=0.5, slide_pred=0.47)
        a, b = df.copy(), df.copy()
        x, y = O.apply(a, **th), T.apply_sharded(b, **th)
        assert_same_results(x[0], y[0])
        assert_same_df(x[1], y[1])
        assert_same_df(a, b)


def test_detect_sharded_single_rank_equals_detect(T):
    """world == 1: the sharded entry points degenerate to the single-process functions (collectives pass through), bit-exact
    vs the oracle incl. the mutation of the table; heavy-ties and f64 tables; from_cv_sharded over folds."""
    for kw in (dict(n_slides=30, tiles_per_slide=80, seed=51), dict(n_slides=18, tiles_per_slide=50, seed=52, ties=50),
               dict(n_slides=22, tiles_per_slide=40, seed=53, dtype=np.float64, ragged=True)):
        df = synth.tile_table(**kw)
        for dkw in ({}, dict(tile_uq=0.05), dict(tile_uq=None, slide_uq=None), dict(tile_pred=0.5, slide_pred=np.float64(0.45))):
            a, b = df.copy(), df.copy()
            x, y = O.detect(a, **dkw), T.detect_sharded(b, **dkw)
            assert_same_results(x[0], y[0], f"{kw} {dkw}")
            assert same_scalar(x[1], y[1]), (x[1], y[1])
            assert_same_df(a, b)
    folds = synth.cv_tables(k=4, n_slides=25, tiles_per_slide=60, seed0=70)
    assert_same_results(O.from_cv([f.copy() for f in folds]), T.from_cv_sharded([f.copy() for f in folds]))
    # NaN predictions: the reference logs and returns all-None (threshold.py:403-405)
    bad = synth.tile_table(n_slides=6, tiles_per_slide=20, seed=54)
    bad.loc[5, "y_pred"] = np.nan
    th, auc = T.detect_sharded(bad)
    assert all(v is None for v in th.values()) and auc is None
    with pytest.raises(ValueError):
        T.apply_sharded(synth.tile_table(n_slides=4, tiles_per_slide=10, seed=1), 0.05, 0.03, tile_pred="detect")


def test_native_comm_world1_passthrough(T):
    """bq_comm_* / bq_allgather_bytes without a communicator behave as a world of one (the 2-rank NCCL run lives in
    tests/multi_gpu_worker.py)."""
    import ctypes as C
    from biscuit_b200 import _ffi
    ctx = _ffi.default_context()
    r, w = C.c_int32(-1), C.c_int32(-1)
    assert ctx.lib.bq_comm_size(ctx.handle, C.byref(r), C.byref(w)) == 0 and (r.value, w.value) == (0, 1)
    src = np.arange(37, dtype=np.uint8)
    dst = np.zeros(37, np.uint8)
    assert ctx.lib.bq_allgather_bytes(ctx.handle, _ffi.ptr(src), 37, _ffi.ptr(dst)) == 0
    assert np.array_equal(src, dst)
    uid = _ffi.load_library().bq_comm_unique_id
    buf = np.zeros(128, np.uint8)
    assert uid(_ffi.ptr(buf)) == 0 and buf.any()
