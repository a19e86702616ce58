"""CPU tier: the C-ABI library loads here (no GPU) and exports every symbol include/biscuit_b200.h
declares; compute entry points fail loudly without a device instead of falling back."""
import os
import re

import pytest

from biscuit_b200 import _ffi
from biscuit_b200.errors import NativeLibraryError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "biscuit_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bq_[a-z0-9_]+)\s*\(", src)))


def test_library_built_and_loads():
    assert os.path.exists(_ffi.LIB_PATH), "run `python -m biscuit_b200.build` (or __graft_entry__.build())"
    lib = _ffi.load_library()
    assert lib.bq_abi_version() == 1


def test_every_declared_symbol_is_exported_and_bound():
    lib = _ffi.load_library()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/biscuit_b200.h but not exported"
        assert s in _ffi.SIGNATURES, f"{s} has no ctypes signature in biscuit_b200/_ffi.py"
    for s in _ffi.SIGNATURES:
        assert s in syms, f"{s} bound in _ffi.py but not declared in the header"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(NativeLibraryError):
        _ffi.Context(0)
    import numpy as np
    import pandas as pd
    from biscuit_b200 import threshold
    df = pd.DataFrame({"slide": ["a", "b"], "y_true": [0, 1], "y_pred": np.float32([0.2, 0.7]),
                       "uncertainty": np.float32([0.01, 0.02])})
    with pytest.raises(NativeLibraryError):
        threshold.apply(df, 0.05, 0.05)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "biscuit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "/root/reference" not in text, f


def test_bench_micro_batch_fills_the_middle_flow_rounds():
    """uq.BENCH_MICRO_BATCH is chosen so that the fused middle-flow kernel's work items (160 rows of the 400-row padded
    image layout) divide evenly over the 74 CTA pairs of a 148-SM B200 (sepmid_sm100.cuh); it must also respect the
    library's micro-batch cap."""
    from biscuit_b200.uq import BENCH_MICRO_BATCH as B
    src = open(os.path.join(ROOT, "biscuit_b200", "csrc", "sepmid_sm100.cuh")).read()
    rows_per_image = int(re.search(r"kPitch = (\d+)", src).group(1)) ** 2
    item = int(re.search(r"kItemPx = (\d+)", src).group(1))
    assert (rows_per_image, item) == (400, 160)
    items = -(-rows_per_image * B // item)
    assert items % 74 == 0 and B <= 512
