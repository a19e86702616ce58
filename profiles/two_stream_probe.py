"""Developer probe (run under gpurun): do kernels of two independent inference pipelines (two contexts = two streams,
two host threads) overlap on the GPU?  Compares 1 pipeline x N tiles with 2 pipelines x N/2 tiles."""
import sys, threading, time
import numpy as np, torch
sys.path.insert(0, ".")
from biscuit_b200 import _ffi
from biscuit_b200.uq import UncertaintyInterface
from biscuit_b200.weights import random_init

N, B = 4096, 256
w = random_init(seed=1)
tiles = (torch.rand((N, 299, 299, 3), device="cuda") * 255).to(torch.uint8)
ifaces = [UncertaintyInterface(w, max_batch=B, ctx=_ffi.Context(0)) for _ in range(2)]
torch.cuda.synchronize()

def run(i, lo, hi):
    ifaces[i].predict(tiles[lo:hi], T=30, seed=1)

for _ in range(2):
    run(0, 0, N); run(1, 0, 512)
t0 = time.perf_counter(); run(0, 0, N); t1 = time.perf_counter()
print("1 pipeline : %.1f ms  %.0f tiles/s" % ((t1 - t0) * 1e3, N / (t1 - t0)))
for rep in range(2):
    th = [threading.Thread(target=run, args=(i, i * N // 2, (i + 1) * N // 2)) for i in range(2)]
    t0 = time.perf_counter()
    [x.start() for x in th]; [x.join() for x in th]
    t1 = time.perf_counter()
    print("2 pipelines: %.1f ms  %.0f tiles/s" % ((t1 - t0) * 1e3, N / (t1 - t0)))
