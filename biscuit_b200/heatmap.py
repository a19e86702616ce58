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

from . import threshold, utils

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
    """

    def __init__(self, interface, tiles, grid, grid_shape=None, T=None, seed=0, tile_index_base=0):
        grid = np.asarray(grid, dtype=np.int64)
        if grid.ndim != 2 or grid.shape[1] != 2 or grid.shape[0] != tiles.shape[0]:
            raise ValueError("grid must be [n, 2] (x, y), one row per tile")
        if (grid < 0).any():
            raise ValueError("grid coordinates must be non-negative")
        if grid_shape is None:
            grid_shape = (int(grid[:, 0].max()) + 1, int(grid[:, 1].max()) + 1) if len(grid) else (0, 0)
        gx, gy = int(grid_shape[0]), int(grid_shape[1])
        if len(grid) and (grid[:, 0].max() >= gx or grid[:, 1].max() >= gy):
            raise ValueError("grid coordinates outside grid_shape")
        self.grid = grid
        self.tile_pred, self.tile_std = interface.predict(tiles, T=T, seed=seed, tile_index_base=tile_index_base)
        nc = self.tile_pred.shape[1]
        self.logits = np.full((gy, gx, nc), EMPTY, dtype=np.float32)
        self.uncertainty = np.full((gy, gx, nc), EMPTY, dtype=np.float32)
        self.logits[grid[:, 1], grid[:, 0]] = self.tile_pred
        self.uncertainty[grid[:, 1], grid[:, 0]] = self.tile_std

    def uq_mask(self, tile_uq_thresh):
        """``hm.uncertainty[:, :, 0] > thresh`` (results.py:222): strict, class-0 std, compared after the
        NumPy promotion of the threshold (python float -> float32 compare, np.float64 -> float64)."""
        return self.uncertainty[:, :, 0] > tile_uq_thresh

    def mask_uncertain(self, tile_uq_thresh):
        """In place: logits of low-confidence locations := -1 (results.py:222-223). Returns the mask."""
        m = self.uq_mask(tile_uq_thresh)
        self.logits[m, :] = EMPTY
        return m

    def split_tiles(self, tile_uq_thresh):
        """Indices of (excluded, included) tiles by ``uncertainty[0][0] > thresh`` where `uncertainty` is what
        ``interface(batch)`` returns, i.e. the class-1 std (results.py:257-264)."""
        u = self.tile_std[:, 1]
        excl = u > tile_uq_thresh
        return np.nonzero(excl)[0], np.nonzero(~excl)[0]

    def tile_names(self):
        """'{u:.4f}-{gx}-{gy}.png' for every tile (results.py:259)."""
        return [f"{u:.4f}-{x}-{y}.png" for u, (x, y) in zip(self.tile_std[:, 1], self.grid)]
