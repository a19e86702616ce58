"""Developer helper (run under gpurun): are repeated predictions of the same tiles bit-identical?  Prints the number of
tiles whose backbone features differ between runs.  Env switches (BQ_DW, BQ_SEP2D, ...) select kernel generations."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biscuit_b200 import weights
from biscuit_b200.uq import UncertaintyInterface
from oracle import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
it = UncertaintyInterface(weights.random_init(seed=1), max_batch=B)
base = torch.from_numpy(synth.tiles_u8(64, seed=1)).cuda()
t = base.repeat((n + 63) // 64, 1, 1, 1)[:n].contiguous()
outs = [it.predict(t, T=4, seed=1, return_features=True)[2] for _ in range(4)]
bad = set()
for o in outs[1:]:
    bad |= set(np.nonzero(np.abs(outs[0] - o).max(1) > 0)[0].tolist())
print({k: os.environ[k] for k in os.environ if k.startswith("BQ_")}, f"B={B} n={n}: {len(bad)} tiles differ", sorted(bad)[:12])
