#!/usr/bin/env python
"""Profiling driver: one warm-up + one profiled `predict` of N synthetic tiles at micro-batch B (run under ncu).

    python profiles/run_predict.py [N=512] [B=512]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from biscuit_b200.uq import UncertaintyInterface  # noqa: E402
from biscuit_b200.weights import random_init  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
it = UncertaintyInterface(random_init(seed=1), max_batch=B)
g = torch.Generator(device="cuda").manual_seed(0)
t = (torch.rand((n, 299, 299, 3), generator=g, device="cuda") * 255).to(torch.uint8)
it.set_profiling(1)          # plain stream launches (no graph replay) so ncu sees every kernel
it.predict(t, T=30, seed=1)
torch.cuda.synchronize()
