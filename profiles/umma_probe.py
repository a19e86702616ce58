"""Developer probe (run under gpurun): does tcgen05.mma read row-shifted windows of a 128B-swizzled tile correctly?"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from biscuit_b200 import _ffi

ctx = _ffi.default_context(0)
rows = 384
x = (np.arange(rows)[:, None] * 64 + np.arange(64)[None, :]).astype(np.float32) % 251 - 125     # exact in bf16
xb = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16).copy()
out = np.zeros((128, 16), np.float32)
for mode in (0,):
    for shift in (0, 1, 2, 3, 7, 8, 9, 20, 21, 22, 23, 43, 64, 100):
        for cg in (0, 1, 3):
            rc = ctx.lib.bq_debug_umma_probe(ctx.handle, rows, shift, cg, mode, _ffi.ptr(xb), _ffi.ptr(out))
            if rc:
                print("mode", mode, "shift", shift, "cg", cg, "rc", rc, ctx.lib.bq_last_error(ctx.handle))
                continue
            ref = x[shift:shift + 128, cg * 16:cg * 16 + 16]
            ok = np.array_equal(out, ref)
            bad_rows = int((out != ref).any(axis=1).sum())
            print(f"base_offset_mode={mode} shift={shift:3d} cg={cg} -> {'OK' if ok else 'MISMATCH'} ({bad_rows} bad rows)")

# ---- 64-byte swizzle (rows of 32 bf16: block1_conv2's input pixels)
x64 = (np.arange(rows)[:, None] * 32 + np.arange(32)[None, :]).astype(np.float32) % 251 - 125
xb64 = torch.from_numpy(x64).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16).copy()
for shift in (0, 1, 2, 3, 4, 7, 8, 9, 148, 149, 150, 151, 200):
    for cg in (0, 1):
        rc = ctx.lib.bq_debug_umma_probe(ctx.handle, rows, shift, cg, 2, _ffi.ptr(xb64), _ffi.ptr(out))
        ref = x64[shift:shift + 128, cg * 16:cg * 16 + 16]
        ok = rc == 0 and np.array_equal(out, ref)
        print(f"SW64 shift={shift:3d} cg={cg} -> {'OK' if ok else 'MISMATCH rc=%d' % rc} ({int((out != ref).any(axis=1).sum())} bad rows)")
