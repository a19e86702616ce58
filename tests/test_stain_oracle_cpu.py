"""CPU tier: known-answer checks for the restated sRGB <-> CIE-LAB conversion behind the Reinhard-fast oracle
(oracle/reinhard.py).  Slideflow / TensorFlow are not installable here (parity unpinned), so the restatement is anchored on
published CIE values instead: D65 white, the sRGB primaries, mid grey -- and on its own round-trip / identity properties."""
import numpy as np

from oracle import reinhard as R, synth


def lab_of(rgb):
    return R.rgb_to_lab((np.array(rgb, np.float32) / np.float32(255.0)).reshape(1, 1, 3)).reshape(3)


def test_known_lab_values():
    # CIE L*a*b* (D65, 2 degree) of sRGB colours, as tabulated by any colour-science reference
    known = {
        (255, 255, 255): (100.0, 0.0, 0.0),
        (0, 0, 0): (0.0, 0.0, 0.0),
        (255, 0, 0): (53.24, 80.09, 67.20),
        (0, 255, 0): (87.73, -86.18, 83.18),
        (0, 0, 255): (32.30, 79.19, -107.86),
        (119, 119, 119): (50.03, 0.0, 0.0),
    }
    for rgb, lab in known.items():
        got = lab_of(rgb)
        assert np.abs(got - np.array(lab)).max() < 0.06, (rgb, got, lab)


def test_round_trip_and_identity_fit():
    tiles = synth.tiles_u8(2, seed=3)
    lab = R.rgb_to_lab(tiles[0].astype(np.float32) / np.float32(255.0))
    back = R.lab_to_rgb(lab) * np.float32(255.0)
    assert np.abs(back - tiles[0]).max() < 0.02                      # u8 -> LAB -> RGB is the identity to 1e-4 relative
    st = R.lab_stats(tiles[0])
    same = R.reinhard_fast(tiles[:1], st[:3], st[3:])               # normalising onto its own statistics
    assert np.abs(same[0].astype(np.int16) - tiles[0].astype(np.int16)).max() <= 1


def test_transfer_moves_statistics_onto_target():
    tiles = synth.tiles_u8(3, seed=4, n_slides=3)
    tm, ts = R.SLIDEFLOW_V1_FIT["target_means"], R.SLIDEFLOW_V1_FIT["target_stds"]
    out = R.reinhard_fast(tiles, tm, ts)
    assert out.dtype == np.uint8 and out.shape == tiles.shape
    for o in out:
        st = R.lab_stats(o)
        # clipping to the sRGB gamut and the uint8 truncation keep it from being exact
        assert np.abs(st[:3] - tm).max() < 6.0, st
