"""GPU tier: Reinhard-fast stain normalisation (csrc/stain_sm100.cuh) against the numpy restatement of Slideflow's
published algorithm (oracle/reinhard.py -- parity with TensorFlow unpinned, see its header).

Tolerance, stated: the uint8 outputs must be equal except for differences of exactly 1 LSB on at most 0.1 % of the
values (float32 powf / cbrtf rounding differs between numpy and CUDA by an ulp, which flips the int32 truncation for
values that land within ~1e-4 of an integer); per-tile LAB statistics within 2e-4 absolute."""
import numpy as np
import pytest

from oracle import reinhard as R, synth

pytestmark = pytest.mark.gpu


def _check_u8(got, want, what):
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert d.max() <= 1, (what, int(d.max()))
    frac = float((d != 0).mean())
    assert frac <= 1e-3, (what, frac)


def test_rgb_to_rgb_matches_oracle():
    from biscuit_b200.norm import ReinhardFastNormalizer
    tiles = synth.tiles_u8(6, seed=21, n_slides=3)
    nz = ReinhardFastNormalizer()
    out = nz.rgb_to_rgb(tiles)
    want = R.reinhard_fast(tiles, R.SLIDEFLOW_V1_FIT["target_means"], R.SLIDEFLOW_V1_FIT["target_stds"])
    assert out.shape == tiles.shape and out.dtype == np.uint8
    _check_u8(out, want, "v1 fit")
    assert (out != tiles).mean() > 0.5                      # it actually changes the image
    single = nz.rgb_to_rgb(tiles[2])
    assert single.tobytes() == out[2].tobytes()              # per-tile statistics: batching does not matter
    stats = nz.lab_stats(tiles)
    want_stats = np.stack([R.lab_stats(t) for t in tiles])
    assert np.abs(stats - want_stats).max() <= 2e-4, np.abs(stats - want_stats).max()


def test_extreme_tiles_and_custom_fit():
    from biscuit_b200.norm import ReinhardFastNormalizer
    rng = np.random.default_rng(3)
    tiles = np.stack([
        rng.integers(0, 256, (64, 64, 3), dtype=np.uint8),               # white noise: full gamut, clipping both ends
        np.clip(rng.normal(235, 8, (64, 64, 3)), 0, 255).astype(np.uint8),   # near-white background tile
        np.clip(rng.normal(12, 6, (64, 64, 3)), 0, 255).astype(np.uint8),    # near-black (linear segment of the gamma)
    ])
    fit = dict(target_means=(60.0, 12.5, -9.0), target_stds=(22.0, 9.0, 7.5))
    out = ReinhardFastNormalizer(**fit).rgb_to_rgb(tiles)
    want = R.reinhard_fast(tiles, fit["target_means"], fit["target_stds"])
    _check_u8(out, want, "custom fit")


def test_fit_roundtrip_and_errors():
    from biscuit_b200.norm import ReinhardFastNormalizer, autoselect
    tiles = synth.tiles_u8(2, seed=4)
    nz = ReinhardFastNormalizer().fit(tiles[0])
    want = R.lab_stats(tiles[0])
    assert np.abs(np.concatenate([nz.target_means, nz.target_stds]) - want).max() <= 2e-4
    # normalising the fit image onto its own statistics is (almost) the identity: u8 -> LAB -> u8 round trip
    back = nz.rgb_to_rgb(tiles[0])
    assert np.abs(back.astype(np.int16) - tiles[0].astype(np.int16)).max() <= 1
    with pytest.raises(ValueError):
        ReinhardFastNormalizer(target_stds=(1.0, 0.0, 1.0))
    with pytest.raises(ValueError):
        autoselect("macenko")
    with pytest.raises(TypeError):
        nz.rgb_to_rgb(tiles[0].astype(np.float32))


def test_interface_normalizer_equals_explicit_prepass():
    """predict(normalizer=...) on raw tiles == predict() on tiles normalised through rgb_to_rgb (same kernels, so
    bit-identical), and differs from the un-normalised prediction."""
    from biscuit_b200 import weights
    from biscuit_b200.norm import ReinhardFastNormalizer
    from biscuit_b200.uq import UncertaintyInterface
    tiles = synth.tiles_u8(5, seed=8)
    w = weights.random_init(seed=1)
    plain = UncertaintyInterface(w, max_batch=4)
    fused = UncertaintyInterface(w, max_batch=4, normalizer="reinhard_fast")
    try:
        pre = ReinhardFastNormalizer().rgb_to_rgb(tiles)
        m0, s0 = plain.predict(pre, T=8, seed=5)
        m1, s1 = fused.predict(tiles, T=8, seed=5)
        assert m0.tobytes() == m1.tobytes() and s0.tobytes() == s1.tobytes()
        m2, _ = plain.predict(tiles, T=8, seed=5)
        assert np.abs(m2 - m1).max() > 0
        fused.set_normalizer(None)
        m3, _ = fused.predict(tiles, T=8, seed=5)
        assert m3.tobytes() == m2.tobytes()
    finally:
        plain.close()
        fused.close()
