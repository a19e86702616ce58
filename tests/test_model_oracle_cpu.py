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


# ----------------------------------------------------------------------------------------------------------------
# The oracle owns its layer list (oracle/xception_arch.py); anchors that do not come from this repository
# ----------------------------------------------------------------------------------------------------------------
def test_oracle_layer_list_matches_keras_published_parameter_count():
    """Keras publishes Xception(include_top=False): 20,861,480 parameters, 20,806,952 trainable, 54,528 non-trainable
    (the BatchNorm moving statistics).  Counted from the oracle's own weight dict."""
    from oracle import xception_arch as A
    w = A.make_weights(seed=3)
    total, trainable = A.backbone_param_counts(w)
    assert total == A.KERAS_XCEPTION_NOTOP_PARAMS == 20_861_480
    assert trainable == A.KERAS_XCEPTION_NOTOP_TRAINABLE == 20_806_952
    assert total - trainable == 54_528
    assert abs(A.backbone_macs_per_tile() / 1e6 - 8355.4) < 0.1
    # head of reference biscuit/hp.py:13,21: 2 x Dense(1024) + Dense(2)
    assert w["hidden_0/kernel"].shape == (2048, 1024) and w["hidden_1/kernel"].shape == (1024, 1024)
    assert w["prelogits/kernel"].shape == (1024, 2)


def test_product_layer_table_and_random_init_agree_with_the_oracles_own():
    """two independently written statements (biscuit_b200/weights.py and oracle/xception_arch.py) of the layer list
    and of the random-init generator: same table, bit-identical weights for the same seed"""
    from biscuit_b200 import weights as P
    from oracle import xception_arch as A
    assert [tuple(r) for r in P.layer_table()] == [tuple(r) for r in A.layer_table()]
    assert P.backbone_macs_per_tile() == A.backbone_macs_per_tile()
    a, b = P.random_init(seed=5), A.make_weights(seed=5)
    assert list(a) == list(b)
    for k in a:
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and a[k].tobytes() == b[k].tobytes(), k


def test_stage_shapes_at_299():
    """147 / 74 / 37 / 19 / 10 stage geometry of Keras Xception at 299 x 299, checked on the
    oracle's forward pass with a tiny-channel stand-in is not possible (channels are fixed), so one real tile"""
    import torch
    from oracle import synth, xception_arch as A
    o = X.XceptionUQOracle(A.make_weights(seed=1))
    stages = {}
    with torch.no_grad():
        f = o.backbone(synth.tiles_u8(1, seed=0), stages=stages)
    assert tuple(f.shape) == (1, 2048)
    for name, (hw, c) in A.STAGE_SHAPES.items():
        assert stages[name].shape == (1, hw, hw, c), (name, stages[name].shape)


def test_tf_same_padding_rules_hand_computed():
    """TF 'SAME' with k=3, s=2: 147->74 pads (1,1); 74->37 pads (0,1) -- ASYMMETRIC; 37->19 (1,1); 19->10 (1,1).
    Hand-computed 1-channel example for the asymmetric case: a 4x4 map pads (0,1): windows start at rows/cols 0, 2."""
    import torch
    assert X._same_pad_s2(147) == (1, 1) and X._same_pad_s2(74) == (0, 1)
    assert X._same_pad_s2(37) == (1, 1) and X._same_pad_s2(19) == (1, 1)
    assert X._same_pad_s2(4) == (0, 1)
    o = X.XceptionUQOracle.__new__(X.XceptionUQOracle)
    x = torch.arange(16, dtype=torch.float32).reshape(1, 1, 4, 4)
    # rows 0-2 x cols 0-2 -> 10; rows 0-2 x cols 2-4(pad) -> 11; rows 2-4 x cols 0-2 -> 14; rows 2-4 x cols 2-4 -> 15
    assert o._pool(x).reshape(-1).tolist() == [10.0, 11.0, 14.0, 15.0]
    # a symmetric (PyTorch-style) pad of 1 would instead give windows centred on rows/cols 0 and 2: [5, 7, 13, 15]
    sym = torch.nn.functional.max_pool2d(x, 3, 2, padding=1).reshape(-1).tolist()
    assert sym == [5.0, 7.0, 13.0, 15.0] and sym != [10.0, 11.0, 14.0, 15.0]
    # odd size: 5x5 pads (1,1): windows centred on 0, 2, 4
    y = torch.arange(25, dtype=torch.float32).reshape(1, 1, 5, 5)
    assert o._pool(y).reshape(-1).tolist() == [6.0, 8.0, 9.0, 16.0, 18.0, 19.0, 21.0, 23.0, 24.0]
    # strided 1x1 residual convolutions sample pixels 0, 2, 4, ... without padding
    assert y[:, :, ::2, ::2].reshape(-1).tolist() == [0.0, 2.0, 4.0, 10.0, 12.0, 14.0, 20.0, 22.0, 24.0]
