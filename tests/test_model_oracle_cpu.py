"""CPU tier: known-answer anchors for the pieces of the restated model oracle (oracle/xception_uq.py) that have a
published definition independent of TensorFlow: the Philox4x32-10 generator behind the dropout masks (Random123's
known-answer vectors, Salmon et al. SC'11), the dropout keep-rate and mask layout, and the architecture's arithmetic size
(Keras Xception, 299 x 299: 8,355.4 M multiply-accumulates)."""
import numpy as np

from oracle import xception_uq as X


def philox(c, k):
    out = X.philox4x32_10(*[np.array([v], np.uint32) for v in c], *k)
    return tuple(int(o[0]) for o in out)


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    assert philox((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert philox((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_keep_mask_layout_and_rate():
    m = X.keep_masks(n_tiles=3, T=5, width=1024, rate=0.1, seed=77)
    assert m.shape == (3, 5, 2, 1024) and m.dtype == np.uint8 and set(np.unique(m)) <= {0, 1}
    assert abs(m.mean() - 0.9) < 0.01                                   # keep probability 1 - rate
    # counter-based: a tile's mask depends on its GLOBAL index only, so shards reproduce the single-process stream
    shard = X.keep_masks(n_tiles=1, T=5, width=1024, rate=0.1, seed=77, tile_index_base=2)
    assert np.array_equal(shard[0], m[2])
    assert not np.array_equal(m[0], m[1]) and not np.array_equal(m[0, 0, 0], m[0, 0, 1])
    assert not np.array_equal(m, X.keep_masks(n_tiles=3, T=5, width=1024, rate=0.1, seed=78))


def test_architecture_size_matches_keras_xception():
    from biscuit_b200 import weights
    assert abs(weights.backbone_macs_per_tile() / 1e6 - 8355.4) < 0.1
    table = weights.layer_table()
    assert sum(1 for kind, *_ in table if kind == "sep") == 34          # 34 SeparableConv2D layers in Keras Xception
    assert sum(1 for kind, *_ in table if kind == "res") == 4           # 4 strided 1x1 residual convolutions
