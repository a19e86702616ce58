"""Developer helper (run under gpurun): where do two runs of the same backbone stage differ? (tile, rows, columns, channels)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biscuit_b200 import weights
from biscuit_b200.uq import UncertaintyInterface
from oracle import synth
B = 256
it = UncertaintyInterface(weights.random_init(seed=1), max_batch=B)
tiles = np.ascontiguousarray(np.tile(synth.tiles_u8(64, seed=1), (4, 1, 1, 1)))
for stage in sys.argv[1:] or ["block4", "block5", "block13", "block14"]:
    ref = it.debug_stage(tiles, stage)
    for rep in range(6):
        cur = it.debug_stage(tiles, stage)
        d = np.abs(cur - ref)
        bad_tiles = np.nonzero(d.reshape(B, -1).max(1) > 0)[0]
        if len(bad_tiles):
            t = bad_tiles[0]
            dt = d[t]
            ys = np.nonzero(dt.max(axis=(1, 2)) > 0)[0]; xs = np.nonzero(dt.max(axis=(0, 2)) > 0)[0]; cs = np.nonzero(dt.max(axis=(0, 1)) > 0)[0]
            print(stage, "rep", rep, "bad tiles", bad_tiles[:10], "tile", t, "shape", dt.shape, "rows", ys[:6], "..", ys[-2:], "cols", xs[:6], "..", xs[-2:],
                  "chan", cs[:6], "..", cs[-3:], "n_chan", len(cs), "max", float(dt.max()))
            break
    else:
        print(stage, "identical over 6 repeats")
