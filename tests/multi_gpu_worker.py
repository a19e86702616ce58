"""torchrun worker for tests/test_multi_gpu.py: every rank runs MC-dropout inference on its slide-aligned shard
and `threshold.apply_sharded`; rank 0 compares with the single-process oracle on the concatenated table."""
import os
import sys
import warnings

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    from biscuit_b200 import dist as bdist, threshold
    from biscuit_b200.uq import UncertaintyInterface
    from biscuit_b200.weights import random_init
    from oracle import synth, threshold_oracle as O
    from helpers import assert_same_df, assert_same_results

    # (1) thresholding on a sharded synthetic cohort, bit-exact vs the oracle on the whole table
    df = synth.tile_table(n_slides=37, tiles_per_slide=90, seed=41, ragged=True)
    counts = df.groupby("slide", sort=False).size().to_numpy()
    lo, hi = bdist.shard_bounds(counts, world)[rank]
    r0, r1 = int(counts[:lo].sum()), int(counts[:hi].sum())
    local_df = df.iloc[r0:r1].reset_index(drop=True)
    th = dict(tile_uq=0.05, slide_uq=np.float64(0.03), tile_pred=0.5, slide_pred=0.48)
    res, s_df = threshold.apply_sharded(local_df, **th)
    ref_res, ref_s = O.apply(df.copy(), **th)
    assert_same_results(ref_res, res, f"rank {rank}")
    assert_same_df(ref_s, s_df, f"rank {rank}")

    # (2) inference: shards draw disjoint Philox streams (tile_index_base) == one process over all tiles
    tiles = synth.tiles_u8(8, seed=3, n_slides=4)
    iface = UncertaintyInterface(random_init(seed=1), max_batch=4, device=local)
    per = 8 // world
    mean, std = iface.predict(tiles[rank * per:(rank + 1) * per], T=30, seed=9, tile_index_base=rank * per)
    full_mean, full_std = iface.predict(tiles, T=30, seed=9)
    assert np.array_equal(mean, full_mean[rank * per:(rank + 1) * per])
    assert np.array_equal(std, full_std[rank * per:(rank + 1) * per])
    gathered = [torch.zeros(per, 2, device="cuda") for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mean).cuda())
    assert np.array_equal(torch.cat(gathered).cpu().numpy(), full_mean)
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
