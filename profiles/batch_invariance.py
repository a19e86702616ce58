"""Developer helper (run under gpurun): a tile's features / mean / std must not depend on the micro-batch size, on its
position inside the micro-batch or on how many times the call is repeated.  Prints the number of tiles that differ."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biscuit_b200 import weights
from biscuit_b200.uq import UncertaintyInterface
from oracle import synth
w = weights.random_init(seed=1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2100
base = torch.from_numpy(synth.tiles_u8(100, seed=9)).cuda()
t = base.repeat((n + 99) // 100, 1, 1, 1)[:n].contiguous()
ref = None
for B in (512, 128, 300, 512):
    it = UncertaintyInterface(w, max_batch=B)
    m, s, f = it.predict(t, T=8, seed=4, return_features=True)
    it.close()
    if ref is None:
        ref = (m, s, f)
        copies = f[: (n // 100) * 100].reshape(n // 100, 100, -1)
        print(f"B={B}: copies of the same tile identical: {all(copies[0].tobytes() == c.tobytes() for c in copies[1:])}")
        continue
    bad = {name: int((np.abs(a - b).reshape(n, -1).max(1) > 0).sum()) for name, a, b in zip(("mean", "std", "feat"), ref, (m, s, f))}
    print(f"B={B} vs B=512: tiles that differ {bad}")
