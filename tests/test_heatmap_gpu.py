"""GPU tier: slide grid heat map + uncertainty masking (reference results.py:216-227, 257-264)."""
import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def iface():
    from biscuit_b200 import weights
    from biscuit_b200.uq import UncertaintyInterface
    it = UncertaintyInterface(weights.random_init(seed=1), max_batch=16)
    yield it
    it.close()


def test_grid_scatter_and_mask(iface):
    from biscuit_b200.heatmap import EMPTY, UQHeatmap
    n = 11
    tiles = synth.tiles_u8(n, seed=5)
    rng = np.random.default_rng(0)
    cells = rng.permutation(5 * 4)[:n]
    grid = np.stack([cells % 5, cells // 5], axis=1)            # (x, y) on a 5 x 4 grid
    hm = UQHeatmap(iface, tiles, grid, grid_shape=(5, 4), T=12, seed=3)
    mean, std = iface.predict(tiles, T=12, seed=3)
    assert hm.logits.shape == (4, 5, 2) and hm.uncertainty.shape == (4, 5, 2)
    filled = np.zeros((4, 5), bool)
    for i, (x, y) in enumerate(grid):
        assert hm.logits[y, x].tobytes() == mean[i].tobytes()
        assert hm.uncertainty[y, x].tobytes() == std[i].tobytes()
        filled[y, x] = True
    assert (hm.logits[~filled] == EMPTY).all() and (hm.uncertainty[~filled] == EMPTY).all()
    thr = np.float64(np.median(std[:, 0]))
    before = hm.logits.copy()
    mask = hm.mask_uncertain(thr)
    want = np.zeros((4, 5), bool)
    for i, (x, y) in enumerate(grid):
        want[y, x] = np.float64(std[i, 0]) > thr                 # strict >, float64 compare for an np.float64 threshold
    assert (mask == want).all() and mask.sum() == (std[:, 0].astype(np.float64) > thr).sum()
    assert (hm.logits[mask] == EMPTY).all() and (hm.logits[~mask] == before[~mask]).all()
    excl, incl = hm.split_tiles(float(thr))
    assert sorted(np.concatenate([excl, incl]).tolist()) == list(range(n))
    assert (std[excl, 0] > np.float32(thr)).all() and not (std[incl, 0] > np.float32(thr)).any()
    assert hm.tile_names()[0] == f"{std[0, 0]:.4f}-{grid[0, 0]}-{grid[0, 1]}.png"


def test_grid_errors(iface):
    from biscuit_b200.heatmap import UQHeatmap
    tiles = np.zeros((2, 299, 299, 3), np.uint8)
    with pytest.raises(ValueError):
        UQHeatmap(iface, tiles, [[0, 0]])
    with pytest.raises(ValueError):
        UQHeatmap(iface, tiles, [[0, 0], [3, 1]], grid_shape=(2, 2))
    with pytest.raises(ValueError):
        UQHeatmap(iface, tiles, [[0, 0], [-1, 1]])


def test_tile_uq_threshold_from_nested_cv(tmp_path):
    from statistics import mean
    from biscuit_b200.heatmap import tile_uq_threshold_from_nested_cv
    from oracle import nested_cv_oracle as NO, threshold_oracle as O
    project = synth.nested_cv_project(str(tmp_path), seed0=3100, n_slides=30, tiles_per_slide=60, fmt="parquet")
    got = tile_uq_threshold_from_nested_cv(project, "cohort")
    want = mean(O.from_cv(NO.df_from_cv(project, f"EXP_AA_UQ-k{k}", "cohort", 5), tile_uq="detect", slide_uq=None,
                          patients=project.dataset().patients())["tile_uq"] for k in (1, 2, 3))
    assert type(got) is type(want) and got == want


def test_heatmap_vs_oracle_and_generator_feed(iface):
    """The grid built from the CPU oracle's predictions (numpy scatter, `uncertainty[:, :, 0] > thresh`, logits := -1:
    results.py:216-227) against the GPU heat map: values within the model tolerance, mask equal wherever the oracle's
    uncertainty is not within tolerance of the threshold.  The streaming generator feed (records as Slideflow's
    `wsi.build_generator()` yields them, micro-batches of 4) must equal the array constructor bit for bit."""
    from biscuit_b200 import weights
    from biscuit_b200.heatmap import EMPTY, UQHeatmap
    from oracle import xception_uq as X
    n = 7
    tiles = synth.tiles_u8(n, seed=8, n_slides=2)
    cells = np.random.default_rng(1).permutation(12)[:n]
    grid = np.stack([cells % 4, cells // 4], axis=1)
    T, seed = 20, 5
    hm = UQHeatmap(iface, tiles, grid, grid_shape=(4, 3), T=T, seed=seed)
    m_ref, s_ref = X.XceptionUQOracle(weights.random_init(seed=1), emulate_bf16=True).predict_uq(tiles, T=T, seed=seed)
    logits_ref = np.full((3, 4, 2), EMPTY, np.float32)
    unc_ref = np.full((3, 4, 2), EMPTY, np.float32)
    logits_ref[grid[:, 1], grid[:, 0]] = m_ref
    unc_ref[grid[:, 1], grid[:, 0]] = s_ref
    assert np.abs(hm.logits - logits_ref).max() <= 4e-3 and np.abs(hm.uncertainty - unc_ref).max() <= 4e-3
    thr = float(np.median(s_ref[:, 0]))
    mask_ref = unc_ref[:, :, 0] > thr
    mask = hm.mask_uncertain(thr)
    decided = np.abs(unc_ref[:, :, 0] - thr) > 4e-3
    assert (mask[decided] == mask_ref[decided]).all() and not mask[unc_ref[:, :, 0] == EMPTY].any()
    assert (hm.logits[mask] == EMPTY).all()
    records = ({"image": tiles[i], "grid": tuple(int(v) for v in grid[i])} for i in range(n))
    hg = UQHeatmap.from_generator(iface, records, grid_shape=(4, 3), batch=4, T=T, seed=seed)
    hm2 = UQHeatmap(iface, tiles, grid, grid_shape=(4, 3), T=T, seed=seed)
    assert hg.logits.tobytes() == hm2.logits.tobytes() and hg.uncertainty.tobytes() == hm2.uncertainty.tobytes()
    assert hg.tile_names() == hm2.tile_names()
