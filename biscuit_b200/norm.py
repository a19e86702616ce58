"""Stain normalisation in front of per-image standardisation -- host mirror of the Slideflow normaliser
object the reference uses (`normalizer='reinhard_fast'`, biscuit/hp.py:19; applied per tile with
``interface.wsi_normalizer.rgb_to_rgb(image)``, results.py:251-254).  The pixels are transformed by the
CUDA library (csrc/stain_sm100.cuh); nothing here computes on the CPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi

NORM_NONE, NORM_REINHARD_FAST = 0, 1

# statistics of Slideflow's built-in 'v1' reference image as recalled; a model must be run with the fit it was
# trained with, so pass your own `target_means` / `target_stds` when they differ
SLIDEFLOW_V1_FIT = {"target_means": (72.909996, 20.8268, -4.9465137), "target_stds": (18.560713, 14.889295, 5.6756697)}


class ReinhardFastNormalizer:
    """``rgb_to_rgb(image)`` like Slideflow's normaliser; also accepted by
    ``UncertaintyInterface(..., normalizer=...)`` which then normalises every tile on the GPU inside
    ``predict`` (no extra host round trip)."""

    kind = NORM_REINHARD_FAST

    def __init__(self, target_means=None, target_stds=None, ctx=None):
        fit = SLIDEFLOW_V1_FIT
        self.target_means = np.asarray(fit["target_means"] if target_means is None else target_means, np.float32)
        self.target_stds = np.asarray(fit["target_stds"] if target_stds is None else target_stds, np.float32)
        if self.target_means.shape != (3,) or self.target_stds.shape != (3,):
            raise ValueError("target_means / target_stds must have 3 entries (L, a, b)")
        if not (self.target_stds > 0).all():
            raise ValueError("target_stds must be positive")
        self._ctx = ctx

    def fit(self, image):
        """Takes the target statistics from a reference RGB image (uint8 [H, W, 3]); computed on the GPU."""
        _, stats = self._run(np.ascontiguousarray(image[None], dtype=np.uint8), want_stats=True, identity=True)
        self.target_means, self.target_stds = stats[0, :3].copy(), stats[0, 3:].copy()
        return self

    def _run(self, tiles, want_stats=False, identity=False):
        ctx = self._ctx or _ffi.default_context()
        n, h, w, c = tiles.shape
        if c != 3 or h != w:
            raise ValueError("tiles must be uint8 [n, px, px, 3]")
        out = np.empty_like(tiles)
        stats = np.empty((n, 6), np.float32) if want_stats else None
        tm, ts = self.target_means, self.target_stds
        if identity:
            tm, ts = np.zeros(3, np.float32), np.ones(3, np.float32)
        _ffi.check(ctx.handle,
                   ctx.lib.bq_stain_normalize(ctx.handle, C.c_int32(self.kind), _ffi.ptr(tiles), n, C.c_int32(h),
                                              _ffi.ptr(np.ascontiguousarray(tm)), _ffi.ptr(np.ascontiguousarray(ts)),
                                              _ffi.ptr(out), _ffi.ptr(stats)), "bq_stain_normalize")
        return out, stats

    def rgb_to_rgb(self, image):
        """uint8 [H, W, 3] or [n, H, W, 3] -> normalised uint8 of the same shape."""
        image = np.asarray(image)
        if image.dtype != np.uint8:
            raise TypeError("image must be uint8 RGB")
        single = image.ndim == 3
        tiles = np.ascontiguousarray(image[None] if single else image)
        out, _ = self._run(tiles)
        return out[0] if single else out

    def lab_stats(self, tiles):
        """Per-tile {mean L, a, b, std L, a, b} of uint8 [n, px, px, 3] tiles."""
        return self._run(np.ascontiguousarray(tiles, dtype=np.uint8), want_stats=True)[1]


def autoselect(name, **kw):
    """`sf.norm.autoselect`-style factory: 'reinhard_fast' is the one the reference uses (hp.py:19)."""
    if name in (None, "none"):
        return None
    if name == "reinhard_fast":
        return ReinhardFastNormalizer(**kw)
    raise ValueError(f"Unknown / unsupported normalizer {name!r} (only 'reinhard_fast' is on the reference's path)")
