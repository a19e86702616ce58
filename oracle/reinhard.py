"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float32) of Slideflow's 'reinhard_fast' stain normaliser.

The reference selects it with ``normalizer='reinhard_fast'`` (biscuit/hp.py:19) and applies it per tile with
``interface.wsi_normalizer.rgb_to_rgb(image)`` in front of ``tf.image.per_image_standardization``
(/root/reference/results.py:251-255).  The arithmetic is in a third-party dependency that is NOT under
/root/reference and is not installable here: ``slideflow>=1.1.0rc1`` (requirements.txt:1),
``slideflow/norm/tensorflow/reinhard.py`` (lab_split / get_mean_std / transform / merge_back) and
``slideflow/norm/tensorflow/color.py`` (rgb_to_lab / lab_to_rgb, the pix2pix-tensorflow sRGB <-> CIE-LAB code, D65).
This file restates that published algorithm:

    I1, I2, I3 = unstack(rgb_to_lab(float32(I) / 255))
    mean_c, std_c = reduce_mean / reduce_std (population) of each LAB channel over the tile
    norm_c = (I_c - mean_c) * (target_std_c / std_c) + target_mean_c
    out = uint8(clip(int32(lab_to_rgb(stack(norm)) * 255), 0, 255))        # the int32 cast truncates toward zero

"fast" = the brightness standardisation (90th-percentile rescale) of the plain Reinhard normaliser is skipped.
The fit (target_means / target_stds) is data: the statistics of whatever reference image the model was trained with;
`SLIDEFLOW_V1_FIT` are the values of Slideflow's built-in 'v1' preset as recalled -- treat as an example, not a constant
of the algorithm.

Parity status: PARITY UNPINNED -- no golden vectors for this step exist in the reference and TensorFlow / Slideflow
cannot be installed in this image, so the restatement is the definition the CUDA kernels are tested against
(tests/test_stain_gpu.py: uint8 output equal except for <= 1 LSB on <= 0.1 % of the values, which is the float32
transcendental-rounding slack between numpy's and CUDA's powf / cbrtf at the truncation boundaries).
"""
from __future__ import annotations

import numpy as np

SLIDEFLOW_V1_FIT = {
    "target_means": np.array([72.909996, 20.8268, -4.9465137], np.float32),
    "target_stds": np.array([18.560713, 14.889295, 5.6756697], np.float32),
}

_F = np.float32
_RGB2XYZ = np.array([[0.412453, 0.212671, 0.019334],
                     [0.357580, 0.715160, 0.119193],
                     [0.180423, 0.072169, 0.950227]], _F)
_XYZ2RGB = np.array([[3.2404542, -0.9692660, 0.0556434],
                     [-1.5371385, 1.8760108, -0.2040259],
                     [-0.4985314, 0.0415560, 1.0572252]], _F)
_FXFYFZ2LAB = np.array([[0.0, 500.0, 0.0],
                        [116.0, -500.0, 200.0],
                        [0.0, 0.0, -200.0]], _F)
_LAB2FXFYFZ = np.array([[1 / 116.0, 1 / 116.0, 1 / 116.0],
                        [1 / 500.0, 0.0, 0.0],
                        [0.0, 0.0, -1 / 200.0]], _F)
_EPS = _F(6.0 / 29.0)


def _matmul3(p, m):
    """[n, 3] @ [3, 3] as three fp32 multiply-adds in row order (what a 3-wide fp32 matmul does)."""
    return (p[:, 0:1] * m[0] + p[:, 1:2] * m[1]) + p[:, 2:3] * m[2]


def rgb_to_lab(srgb):
    """srgb float32 [..., 3] in 0..1 -> LAB (color.py: rgb_to_lab)."""
    shape = srgb.shape
    p = srgb.reshape(-1, 3).astype(_F)
    lin = np.where(p <= _F(0.04045), p / _F(12.92), ((p + _F(0.055)) / _F(1.055)) ** _F(2.4)).astype(_F)
    xyz = _matmul3(lin, _RGB2XYZ)
    xyz = xyz * np.array([1 / 0.950456, 1.0, 1 / 1.088754], _F)
    f = np.where(xyz <= _EPS ** 3, xyz / (_F(3.0) * _EPS ** 2) + _F(4.0 / 29.0), np.cbrt(xyz)).astype(_F)
    lab = _matmul3(f, _FXFYFZ2LAB) + np.array([-16.0, 0.0, 0.0], _F)
    return lab.reshape(shape).astype(_F)


def lab_to_rgb(lab):
    """LAB float32 [..., 3] -> srgb in 0..1 (color.py: lab_to_rgb)."""
    shape = lab.shape
    p = lab.reshape(-1, 3).astype(_F)
    f = _matmul3(p + np.array([16.0, 0.0, 0.0], _F), _LAB2FXFYFZ)
    xyz = np.where(f <= _EPS, _F(3.0) * _EPS ** 2 * (f - _F(4.0 / 29.0)), f * f * f).astype(_F)
    xyz = xyz * np.array([0.950456, 1.0, 1.088754], _F)
    rgb = np.clip(_matmul3(xyz, _XYZ2RGB), _F(0.0), _F(1.0))
    srgb = np.where(rgb <= _F(0.0031308), rgb * _F(12.92), (rgb ** _F(1 / 2.4)) * _F(1.055) - _F(0.055))
    return srgb.reshape(shape).astype(_F)


def lab_stats(tile_u8):
    """{mean L, a, b, std L, a, b} (population std) of one uint8 RGB tile (reinhard.py: lab_split + get_mean_std)."""
    lab = rgb_to_lab(tile_u8.astype(_F) / _F(255.0))
    flat = lab.reshape(-1, 3)
    return np.concatenate([flat.mean(axis=0, dtype=np.float64), flat.std(axis=0, dtype=np.float64)]).astype(_F)


def reinhard_fast(tiles_u8, target_means, target_stds):
    """uint8 [n, H, W, 3] -> uint8 [n, H, W, 3] (reinhard.py: transform, without standardize_brightness)."""
    tm, ts = np.asarray(target_means, _F), np.asarray(target_stds, _F)
    out = np.empty_like(tiles_u8)
    for i, tile in enumerate(tiles_u8):
        lab = rgb_to_lab(tile.astype(_F) / _F(255.0))
        st = lab_stats(tile)
        norm = (lab - st[:3]) * (ts / st[3:]) + tm
        merged = (lab_to_rgb(norm.astype(_F)) * _F(255.0)).astype(np.int32)     # truncation toward zero
        out[i] = np.clip(merged, 0, 255).astype(np.uint8)
    return out
