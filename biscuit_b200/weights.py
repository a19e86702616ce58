"""Xception-UQ weight naming and random initialisation.

Weights are exchanged as a dict of float32 arrays under Keras variable names (HWIO convolution kernels,
[in, out] dense kernels) -- what `model.get_weights()` of the Slideflow/Keras model holds; the residual
1x1 stride-2 convolutions, which Keras auto-names `conv2d_N`, are called `block{2,3,4,13}_res` here.

`random_init` gives the random-init network BASELINE.json's configs ask for ("Xception-UQ (random init,
dropout head)"): He-scaled convolutions with RANDOMISED BatchNorm statistics, gains chosen so a 36-layer
random net neither explodes nor collapses every prediction to 0.5.
"""
from __future__ import annotations

import numpy as np

# (name, cin, cout) of every SeparableConv2D, in execution order
ENTRY_BLOCKS = [(2, 64, 128), (3, 128, 256), (4, 256, 728)]
MIDDLE_BLOCKS = list(range(5, 13))
FEATURES = 2048


def layer_table():
    """[(kind, name, cin, cout)] -- kind in conv3x3s2/conv3x3/sep/res."""
    t = [("conv3x3s2", "block1_conv1", 3, 32), ("conv3x3", "block1_conv2", 32, 64)]
    for b, cin, cout in ENTRY_BLOCKS:
        t += [("res", f"block{b}_res", cin, cout), ("sep", f"block{b}_sepconv1", cin, cout),
              ("sep", f"block{b}_sepconv2", cout, cout)]
    for b in MIDDLE_BLOCKS:
        t += [("sep", f"block{b}_sepconv{i}", 728, 728) for i in (1, 2, 3)]
    t += [("res", "block13_res", 728, 1024), ("sep", "block13_sepconv1", 728, 728),
          ("sep", "block13_sepconv2", 728, 1024), ("sep", "block14_sepconv1", 1024, 1536),
          ("sep", "block14_sepconv2", 1536, 2048)]
    return t


# ----------------------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------------------

def random_init(seed=1, hidden_width=1024, hidden_layers=2, n_classes=2):
    """Random-init weights under Keras variable names (HWIO conv kernels, [in,out] dense).
    He-scaled convolutions and RANDOMISED BatchNorm statistics (otherwise a random 36-layer net
    collapses every prediction to 0.5 and parity would be vacuous)."""
    rng = np.random.default_rng(seed)
    w = {}

    def bn(name, c, gamma_lo=0.5, gamma_hi=1.5):
        w[f"{name}/gamma"] = rng.uniform(gamma_lo, gamma_hi, c).astype(np.float32)
        w[f"{name}/beta"] = rng.normal(0, 0.1, c).astype(np.float32)
        w[f"{name}/moving_mean"] = rng.normal(0, 0.1, c).astype(np.float32)
        w[f"{name}/moving_variance"] = rng.uniform(0.5, 1.5, c).astype(np.float32)

    for kind, name, cin, cout in layer_table():
        if kind.startswith("conv3x3"):
            w[f"{name}/kernel"] = rng.normal(0, np.sqrt(2.0 / (9 * cin)), (3, 3, cin, cout)).astype(np.float32)
            bn(f"{name}_bn", cout)
        elif kind == "res":
            w[f"{name}/kernel"] = rng.normal(0, np.sqrt(1.0 / cin), (1, 1, cin, cout)).astype(np.float32)
            bn(f"{name}_bn", cout, 0.4, 0.8)
        else:
            w[f"{name}/depthwise_kernel"] = rng.normal(0, np.sqrt(2.0 / 9), (3, 3, cin, 1)).astype(np.float32)
            w[f"{name}/pointwise_kernel"] = rng.normal(0, np.sqrt(1.0 / cin), (1, 1, cin, cout)).astype(np.float32)
            # the last BN of a residual block feeds the skip sum: keep its gain < 1 so the stream
            # neither explodes nor vanishes over 12 blocks
            last = name.endswith("sepconv3") or name in ("block2_sepconv2", "block3_sepconv2",
                                                         "block4_sepconv2", "block13_sepconv2")
            bn(f"{name}_bn", cout, 0.3 if last else 0.5, 0.7 if last else 1.5)
    # head: gains chosen so that the softmax is neither saturated nor constant (class-1 mean ~0.6-0.75,
    # dropout std ~0.05-0.1 on synthetic tiles) -- a degenerate head would make parity vacuous
    cin = FEATURES
    for i in range(hidden_layers):
        gain = 0.27 if i == 0 else np.sqrt(2.0)
        w[f"hidden_{i}/kernel"] = rng.normal(0, gain / np.sqrt(cin), (cin, hidden_width)).astype(np.float32)
        w[f"hidden_{i}/bias"] = rng.normal(0, 0.05, hidden_width).astype(np.float32)
        cin = hidden_width
    w["prelogits/kernel"] = rng.normal(0, 1.0 / np.sqrt(cin), (cin, n_classes)).astype(np.float32)
    w["prelogits/bias"] = rng.normal(0, 0.05, n_classes).astype(np.float32)
    return w


def backbone_macs_per_tile(px=299):
    """Algorithmic multiply-accumulates of one backbone pass (8,355.4 MMAC)."""
    s1 = (px - 3) // 2 + 1
    s2 = s1 - 2
    macs = s1 * s1 * 27 * 32 + s2 * s2 * 288 * 64
    h = s2
    for kind, name, cin, cout in layer_table()[2:]:
        if kind == "res":
            ho = (h + 1) // 2
            macs += ho * ho * cin * cout
            res_h = ho
        else:
            macs += h * h * cin * (9 + cout)
            if name.endswith("sepconv2") and name.split("_")[0] in ("block2", "block3", "block4", "block13"):
                h = (h + 1) // 2
    return macs
