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

    # (1b) cohort-wide detection on the sharded table: tile ROCs over ALL tiles (triples all-gathered), slide level replicated
    ref_th, ref_auc = O.detect(df.copy())
    th2, auc2 = threshold.detect_sharded(df.iloc[r0:r1].reset_index(drop=True))
    assert_same_results(ref_th, th2, f"detect rank {rank}")
    assert (ref_auc == auc2) or (ref_auc != ref_auc and auc2 != auc2), (ref_auc, auc2)
    # ... and with numeric tile thresholds / no tile-UQ filter (no triple exchange needed)
    for kw in (dict(tile_uq=0.05, tile_pred=0.5), dict(tile_uq=None, slide_uq=None, tile_pred=np.float64(0.4), slide_pred=0.5)):
        a, _ = O.detect(df.copy(), **kw)
        b, _ = threshold.detect_sharded(df.iloc[r0:r1].reset_index(drop=True), **kw)
        assert_same_results(a, b, f"detect {kw} rank {rank}")
    # from_cv over sharded folds
    folds = synth.cv_tables(k=3, n_slides=24, tiles_per_slide=70, seed0=300)
    shards = []
    for f in folds:
        c = f.groupby("slide", sort=False).size().to_numpy()
        a, b = bdist.shard_bounds(c, world)[rank]
        shards.append(f.iloc[int(c[:a].sum()):int(c[:b].sum())].reset_index(drop=True))
    assert_same_results(O.from_cv([f.copy() for f in folds]), threshold.from_cv_sharded(shards), f"from_cv rank {rank}")

    # (1c) the same exchange through the C ABI communicator (bq_comm_init / bq_allgather_bytes) instead of torch.distributed
    from biscuit_b200 import _ffi
    ctx = _ffi.default_context(local)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(bdist.NativeComm.unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    comm = bdist.NativeComm(ctx, rank, world, bytes(idt.cpu().numpy().tobytes()))
    res3, s3 = threshold.apply_sharded(df.iloc[r0:r1].reset_index(drop=True), group=comm, **th)
    assert_same_results(ref_res, res3, f"native comm rank {rank}")
    assert_same_df(ref_s, s3, f"native comm rank {rank}")
    th3, _ = threshold.detect_sharded(df.iloc[r0:r1].reset_index(drop=True), group=comm)
    assert_same_results(ref_th, th3, f"native comm detect rank {rank}")
    comm.close()

    # (1d) an EMPTY shard (more ranks than slides) and a validation error on ONE rank: nobody blocks, everyone agrees
    one = synth.tile_table(n_slides=1, tiles_per_slide=50, seed=7)
    mine = one if rank == 0 else one.iloc[0:0]
    r_one, s_one = threshold.apply_sharded(mine.reset_index(drop=True), tile_uq=0.05, slide_uq=0.03)
    a_one, b_one = O.apply(one.copy(), tile_uq=0.05, slide_uq=0.03)
    assert_same_results(a_one, r_one, f"empty shard rank {rank}")
    bad = local_df.copy()
    if rank == world - 1:
        bad.loc[bad.index[3], "y_pred"] = np.nan
    from biscuit_b200.errors import PredsContainNaNError
    try:
        threshold.apply_sharded(bad, **th)
        raise AssertionError("expected PredsContainNaNError on every rank")
    except PredsContainNaNError:
        pass
    try:
        threshold.apply_sharded(local_df.copy(), tile_uq=0.05, slide_uq=0.03, tile_pred="detect")
        raise AssertionError("expected ValueError for a per-shard 'detect'")
    except ValueError:
        pass

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
