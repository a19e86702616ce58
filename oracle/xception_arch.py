"""TEST INFRASTRUCTURE ONLY -- the oracle's OWN statement of the Keras Xception(include_top=False) layer list
and of the random-init weight generator.

Written out from SURVEY.md Appendix B (Keras `applications/xception.py`, recalled; frozen here) and kept
independent of `biscuit_b200.weights` on purpose: an architecture error in the product's table would otherwise be
shared by the checker.  `tests/test_model_oracle_cpu.py` ties the two statements together (same table, same weights
for the same seed) and anchors this one on what Keras publishes for `Xception(include_top=False)`:
20,861,480 parameters, 20,806,952 trainable, 54,528 non-trainable (BatchNorm moving statistics).

Architecture contract: reference biscuit/hp.py:3-24 (`model='xception'`, `pooling='avg'`, `include_top=False`,
`hidden_layers=2`, `hidden_layer_width=1024`).
"""
from __future__ import annotations

import numpy as np

FEATURES = 2048
KERAS_XCEPTION_NOTOP_PARAMS = 20_861_480
KERAS_XCEPTION_NOTOP_TRAINABLE = 20_806_952

# (block, cin, cout): residual 1x1 s2 conv cin->cout, sepconv1 cin->cout, sepconv2 cout->cout, max-pool 3x3 s2
ENTRY_BLOCKS = [(2, 64, 128), (3, 128, 256), (4, 256, 728)]
MIDDLE_BLOCKS = [5, 6, 7, 8, 9, 10, 11, 12]

# every convolution of the backbone in execution order: (kind, name, cin, cout)
#   conv3x3s2: 3x3 stride 2 'valid';  conv3x3: 3x3 'valid';  res: 1x1 stride 2 (Keras auto-names these conv2d_N);
#   sep: SeparableConv2D = depthwise 3x3 'same' directly followed by pointwise 1x1.  All bias-free, all followed by BN.
LAYERS = [
    ("conv3x3s2", "block1_conv1", 3, 32),
    ("conv3x3", "block1_conv2", 32, 64),
    ("res", "block2_res", 64, 128),
    ("sep", "block2_sepconv1", 64, 128),
    ("sep", "block2_sepconv2", 128, 128),
    ("res", "block3_res", 128, 256),
    ("sep", "block3_sepconv1", 128, 256),
    ("sep", "block3_sepconv2", 256, 256),
    ("res", "block4_res", 256, 728),
    ("sep", "block4_sepconv1", 256, 728),
    ("sep", "block4_sepconv2", 728, 728),
    ("sep", "block5_sepconv1", 728, 728), ("sep", "block5_sepconv2", 728, 728), ("sep", "block5_sepconv3", 728, 728),
    ("sep", "block6_sepconv1", 728, 728), ("sep", "block6_sepconv2", 728, 728), ("sep", "block6_sepconv3", 728, 728),
    ("sep", "block7_sepconv1", 728, 728), ("sep", "block7_sepconv2", 728, 728), ("sep", "block7_sepconv3", 728, 728),
    ("sep", "block8_sepconv1", 728, 728), ("sep", "block8_sepconv2", 728, 728), ("sep", "block8_sepconv3", 728, 728),
    ("sep", "block9_sepconv1", 728, 728), ("sep", "block9_sepconv2", 728, 728), ("sep", "block9_sepconv3", 728, 728),
    ("sep", "block10_sepconv1", 728, 728), ("sep", "block10_sepconv2", 728, 728), ("sep", "block10_sepconv3", 728, 728),
    ("sep", "block11_sepconv1", 728, 728), ("sep", "block11_sepconv2", 728, 728), ("sep", "block11_sepconv3", 728, 728),
    ("sep", "block12_sepconv1", 728, 728), ("sep", "block12_sepconv2", 728, 728), ("sep", "block12_sepconv3", 728, 728),
    ("res", "block13_res", 728, 1024),
    ("sep", "block13_sepconv1", 728, 728),
    ("sep", "block13_sepconv2", 728, 1024),
    ("sep", "block14_sepconv1", 1024, 1536),
    ("sep", "block14_sepconv2", 1536, 2048),
]

# spatial size of every named stage output at a 299 x 299 input
STAGE_SHAPES = {
    "block1_conv1": (149, 32), "block1_conv2": (147, 64), "block2": (74, 128), "block3": (37, 256),
    "block4": (19, 728), "block5": (19, 728), "block12": (19, 728), "block13": (10, 1024), "block14": (10, 2048),
}

# sepconvs whose BatchNorm feeds a residual sum (gain kept < 1 in the random init)
_SKIP_FEEDERS = {"block2_sepconv2", "block3_sepconv2", "block4_sepconv2", "block13_sepconv2"} | \
                {f"block{b}_sepconv3" for b in MIDDLE_BLOCKS}


def layer_table():
    return list(LAYERS)


def backbone_param_counts(weights):
    """(total, trainable) parameter counts of the backbone part of a weight dict"""
    total = trainable = 0
    for name, v in weights.items():
        if name.startswith(("hidden_", "prelogits")):
            continue
        total += int(np.asarray(v).size)
        if "/moving_" not in name:
            trainable += int(np.asarray(v).size)
    return total, trainable


def backbone_macs_per_tile(px=299):
    """multiply-accumulates of one backbone pass, from the layer list and the stage geometry"""
    s1 = (px - 3) // 2 + 1
    s2 = s1 - 2
    h = s2
    macs = 0
    for kind, name, cin, cout in LAYERS:
        if kind == "conv3x3s2":
            macs += s1 * s1 * 9 * cin * cout
        elif kind == "conv3x3":
            macs += s2 * s2 * 9 * cin * cout
        elif kind == "res":
            ho = -(-h // 2)
            macs += ho * ho * cin * cout
        else:
            macs += h * h * cin * 9 + h * h * cin * cout
            if name in ("block2_sepconv2", "block3_sepconv2", "block4_sepconv2", "block13_sepconv2"):
                h = -(-h // 2)          # the max-pool that follows
    return macs


def make_weights(seed=1, hidden_width=1024, hidden_layers=2, n_classes=2):
    """Random-init Xception-UQ weights under Keras variable names (HWIO kernels, [in, out] dense): He-scaled
    convolutions, RANDOMISED BatchNorm statistics (a random 36-layer net with identity BN collapses every
    prediction to 0.5 and parity would be vacuous), head gains chosen so the softmax is neither
    saturated nor constant.  Draw order: per layer kernel(s) then gamma, beta, moving_mean, moving_variance."""
    rng = np.random.default_rng(seed)
    w = {}
    f32 = np.float32

    def batch_norm(prefix, c, lo, hi):
        w[prefix + "/gamma"] = rng.uniform(lo, hi, c).astype(f32)
        w[prefix + "/beta"] = rng.normal(0, 0.1, c).astype(f32)
        w[prefix + "/moving_mean"] = rng.normal(0, 0.1, c).astype(f32)
        w[prefix + "/moving_variance"] = rng.uniform(0.5, 1.5, c).astype(f32)

    for kind, name, cin, cout in LAYERS:
        if kind in ("conv3x3s2", "conv3x3"):
            w[name + "/kernel"] = rng.normal(0, np.sqrt(2.0 / (9 * cin)), (3, 3, cin, cout)).astype(f32)
            batch_norm(name + "_bn", cout, 0.5, 1.5)
        elif kind == "res":
            w[name + "/kernel"] = rng.normal(0, np.sqrt(1.0 / cin), (1, 1, cin, cout)).astype(f32)
            batch_norm(name + "_bn", cout, 0.4, 0.8)
        else:
            w[name + "/depthwise_kernel"] = rng.normal(0, np.sqrt(2.0 / 9), (3, 3, cin, 1)).astype(f32)
            w[name + "/pointwise_kernel"] = rng.normal(0, np.sqrt(1.0 / cin), (1, 1, cin, cout)).astype(f32)
            if name in _SKIP_FEEDERS:
                batch_norm(name + "_bn", cout, 0.3, 0.7)
            else:
                batch_norm(name + "_bn", cout, 0.5, 1.5)
    fan_in = FEATURES
    for i in range(hidden_layers):
        gain = 0.27 if i == 0 else np.sqrt(2.0)
        w[f"hidden_{i}/kernel"] = rng.normal(0, gain / np.sqrt(fan_in), (fan_in, hidden_width)).astype(f32)
        w[f"hidden_{i}/bias"] = rng.normal(0, 0.05, hidden_width).astype(f32)
        fan_in = hidden_width
    w["prelogits/kernel"] = rng.normal(0, 1.0 / np.sqrt(fan_in), (fan_in, n_classes)).astype(f32)
    w["prelogits/bias"] = rng.normal(0, 0.05, n_classes).astype(f32)
    return w
