"""Whole-slide tile-grid heat map with uncertainty masking -- the step right after the hot path for
single-slide use (reference results.py:179-188, 216-227, 234-265).

The reference builds ``sf.Heatmap(slide, model)`` (Slideflow: one MC-dropout prediction per grid
location, arrays ``logits [gy, gx, C]`` and ``uncertainty [gy, gx, C]``, -1 where no tile was
extracted), masks locations whose uncertainty exceeds the nested-CV tile threshold with
``hm.logits[uq_mask, :] = [-1, -1]`` (results.py:222-223) and sorts the tiles into uq_incl / uq_excl
with the SAME strict ``>`` (results.py:261).  Tile extraction and rendering are Slideflow / matplotlib
and out of scope; this module takes the extracted uint8 tiles with their grid coordinates.
All predictions come from :class:`biscuit_b200.uq.UncertaintyInterface` (CUDA library)."""
from __future__ import annotations

from statistics import mean

import numpy as np

from . import _ffi, threshold, utils

EMPTY = -1.0     # Slideflow's value for grid cells without a prediction, and the reference's mask value


def tile_uq_threshold_from_nested_cv(project, outcome, label="EXP_AA_UQ", outer_k=3, inner_k=5):
    """Mean over the outer folds of the tile-level UQ threshold detected on each fold's inner CV tables
    (reference results.py:179-188)."""
    found = []
    for k in range(1, outer_k + 1):
        dfs = utils.df_from_cv(project, f"{label}-k{k}", outcome=outcome, k=inner_k)
        found.append(threshold.from_cv(dfs, tile_uq="detect", slide_uq=None,
                                       patients=project.dataset().patients())["tile_uq"])
    return mean(found)


class UQHeatmap:
    """logits / uncertainty grids of one slide.

    Args:
        interface: :class:`biscuit_b200.uq.UncertaintyInterface`.
        tiles: uint8 [n, 299, 299, 3] tiles of the slide (raw RGB), in generator order.
        grid: int [n, 2] grid coordinates (x, y) of each tile (Slideflow's ``tile['grid']`` / ``'loc'``).
        grid_shape: (gx, gy) size of the slide's tile grid; default: tight bounding box of `grid`.

    The grids are built on the GPU (`bq_heatmap_build`: fill with -1, scatter) and masked there (`bq_heatmap_mask`).
    `UQHeatmap.from_generator` streams a Slideflow-style tile generator instead of a pre-assembled array."""

    def __init__(self, interface, tiles, grid, grid_shape=None, T=None, seed=0, tile_index_base=0):
        grid = self._check_grid(grid, tiles.shape[0])
        tile_pred, tile_std = interface.predict(tiles, T=T, seed=seed, tile_index_base=tile_index_base)
        self._assemble(interface, grid, grid_shape, tile_pred, tile_std)

    @classmethod
    def from_generator(cls, interface, generator, grid_shape=None, batch=None, T=None, seed=0, tile_index_base=0):
        """Heat map from an iterable of tile records as ``wsi.build_generator(shuffle=False)()`` yields them
        (reference results.py:249): dicts with ``'image'`` (uint8 [299, 299, 3]) and ``'grid'`` (x, y).  Tiles are packed
        into pinned micro-batches of `batch` tiles (default: the interface's micro-batch) and predicted as they arrive --
        the reference pushes one tile per call through the model (results.py:255-257).  Philox counters carry the running
        tile index, so the result equals one `predict` over all tiles."""
        self = cls.__new__(cls)
        batch = int(batch or getattr(interface, "max_batch", 64))
        px = interface.config.tile_px
        try:
            import torch
            hold = torch.empty((batch, px, px, 3), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
            buf = hold.numpy()
        except Exception:
            buf = np.empty((batch, px, px, 3), np.uint8)
        grids, preds, stds, fill, done = [], [], [], 0, 0

        def flush():
            nonlocal fill, done
            if fill:
                m, s = interface.predict(buf[:fill], T=T, seed=seed, tile_index_base=tile_index_base + done)
                preds.append(m)
                stds.append(s)
                done += fill
                fill = 0

        for rec in generator:
            img = np.asarray(rec["image"])
            if img.shape != (px, px, 3) or img.dtype != np.uint8:
                raise ValueError(f"tile images must be uint8 [{px}, {px}, 3]")
            buf[fill] = img
            grids.append(rec["grid"] if "grid" in rec else rec["loc"])
            fill += 1
            if fill == batch:
                flush()
        flush()
        nc = interface.config.n_classes
        grid = self._check_grid(np.asarray(grids, dtype=np.int64).reshape(-1, 2), done)
        self._assemble(interface, grid, grid_shape, np.concatenate(preds) if preds else np.empty((0, nc), np.float32),
                       np.concatenate(stds) if stds else np.empty((0, nc), np.float32))
        return self

    @staticmethod
    def _check_grid(grid, n):
        grid = np.asarray(grid, dtype=np.int64)
        if grid.ndim != 2 or grid.shape[1] != 2 or grid.shape[0] != n:
            raise ValueError("grid must be [n, 2] (x, y), one row per tile")
        if (grid < 0).any():
            raise ValueError("grid coordinates must be non-negative")
        return grid

    def _assemble(self, interface, grid, grid_shape, tile_pred, tile_std):
        if grid_shape is None:
            grid_shape = (int(grid[:, 0].max()) + 1, int(grid[:, 1].max()) + 1) if len(grid) else (0, 0)
        gx, gy = int(grid_shape[0]), int(grid_shape[1])
        if len(grid) and (grid[:, 0].max() >= gx or grid[:, 1].max() >= gy):
            raise ValueError("grid coordinates outside grid_shape")
        self.grid, self.tile_pred, self.tile_std = grid, tile_pred, tile_std
        self.ctx = interface.ctx
        nc = tile_pred.shape[1]
        self.logits = np.empty((gy, gx, nc), np.float32)
        self.uncertainty = np.empty((gy, gx, nc), np.float32)
        g32 = np.ascontiguousarray(grid, dtype=np.int32)
        _ffi.check(self.ctx.handle,
                   self.ctx.lib.bq_heatmap_build(self.ctx.handle, len(grid), nc, _ffi.ptr(np.ascontiguousarray(tile_pred)),
                                                 _ffi.ptr(np.ascontiguousarray(tile_std)), _ffi.ptr(g32), gx, gy,
                                                 _ffi.ptr(self.logits), _ffi.ptr(self.uncertainty)), "bq_heatmap_build")

    def uq_mask(self, tile_uq_thresh):
        """``hm.uncertainty[:, :, 0] > thresh`` (results.py:222): strict, class-0 std, compared after the
        NumPy promotion of the threshold (python float -> float32 compare, np.float64 -> float64)."""
        return self._mask(tile_uq_thresh, self.logits.copy())

    def mask_uncertain(self, tile_uq_thresh):
        """In place: logits of low-confidence locations := -1 (results.py:222-223). Returns the mask."""
        return self._mask(tile_uq_thresh, self.logits)

    def _mask(self, tile_uq_thresh, logits):
        gy, gx, nc = self.uncertainty.shape
        mask = np.zeros((gy, gx), np.uint8)
        eff = threshold._cmp_scalar(tile_uq_thresh, np.float32)
        _ffi.check(self.ctx.handle,
                   self.ctx.lib.bq_heatmap_mask(self.ctx.handle, gy * gx, nc, _ffi.ptr(self.uncertainty), float(eff),
                                                _ffi.ptr(logits), _ffi.ptr(mask)), "bq_heatmap_mask")
        return mask.view(np.bool_)

    def split_tiles(self, tile_uq_thresh):
        """Indices of (excluded, included) tiles by ``uncertainty[0][0] > thresh`` where `uncertainty` is what
        ``interface(batch)`` returns (results.py:257-264)."""
        u = self.tile_std[:, 0]
        excl = u > tile_uq_thresh
        return np.nonzero(excl)[0], np.nonzero(~excl)[0]

    def tile_names(self):
        """'{u:.4f}-{gx}-{gy}.png' for every tile (results.py:259)."""
        return [f"{u:.4f}-{x}-{y}.png" for u, (x, y) in zip(self.tile_std[:, 0], self.grid)]
