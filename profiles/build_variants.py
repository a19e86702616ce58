#!/usr/bin/env python
"""Developer tool: A/B copies of libbiscuit_b200.so with -D overrides of the fused middle-flow kernel's tunables.

    python profiles/build_variants.py name1:-DBQ_SM_PW=8,-DBQ_SM_SWPIPE=0 name2:...

Each variant recompiles csrc/model.cu only and links it with the objects of the regular build into
profiles/variants/lib_<name>.so (git-ignored, travels with the gpurun snapshot).  A profiling job copies a variant over
biscuit_b200/libbiscuit_b200.so on the (scratch) GPU box before a bench run -- the product never selects a library at run
time."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from biscuit_b200 import build as B  # noqa: E402

out_dir = os.path.join(ROOT, "profiles", "variants")
os.makedirs(out_dir, exist_ok=True)
B.build()
objs = [os.path.join(B.OBJ, f[:-3] + ".o") for f in B._sources() if f != "model.cu"]
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    flags = [f for f in flags.split(",") if f]
    obj = os.path.join(out_dir, f"model_{name}.o")
    lib = os.path.join(out_dir, f"lib_{name}.so")
    cmd = f"{B._nvcc()} {' '.join(B.NVCC_FLAGS)} {' '.join(flags)} -c {os.path.join(B.CSRC, 'model.cu')} -o {obj} && " \
          f"{B._nvcc()} -shared -o {lib} {obj} {' '.join(objs)} -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC"
    procs.append((name, subprocess.Popen(cmd, shell=True)))
for name, p in procs:
    print(name, "ok" if p.wait() == 0 else "FAILED")
