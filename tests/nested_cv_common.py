"""Shared by the CPU and GPU nested-CV tests: rebuild a golden case's project tree and compare outputs bit for bit."""
import numpy as np

from oracle import synth
from helpers import dec, same_scalar


def build_case(root, kw):
    kw = dict(kw)
    call = kw.pop("call", {})
    kw["dtype"] = np.dtype(kw["dtype"]).type
    kw["missing_outer"] = tuple(kw.get("missing_outer", ()))
    return synth.nested_cv_project(str(root), **kw), call


def assert_matches_golden(df, th, case, what):
    for k, v in case["thresholds"].items():
        assert same_scalar(th[k], dec(v)), (what, k, th[k], dec(v))
    assert len(df) == len(case["rows"]), (what, len(df), len(case["rows"]))
    assert list(df.columns) == case["columns"], (what, list(df.columns))
    assert [str(t) for t in df.dtypes] == case["dtypes"], (what, [str(t) for t in df.dtypes])
    for got, exp in zip(df.to_dict("records"), case["rows"]):
        for c, v in exp.items():
            if isinstance(v, str):
                assert got[c] == v, (what, c, got[c], v)
            else:
                assert same_scalar(got[c], dec(v), check_type=False), (what, c, got[c], dec(v))


def assert_same_outputs(a, b, what):
    (dfa, tha), (dfb, thb) = a, b
    assert set(tha) == set(thb)
    for k in tha:
        assert same_scalar(tha[k], thb[k]), (what, k, tha[k], thb[k])
    assert list(dfa.columns) == list(dfb.columns) and len(dfa) == len(dfb), what
    for c in dfa.columns:
        x, y = dfa[c].to_numpy(), dfb[c].to_numpy()
        assert dfa[c].dtype == dfb[c].dtype, (what, c)
        if x.dtype.kind == "f":
            assert x.tobytes() == y.tobytes(), (what, c, x, y)
        else:
            assert (x == y).all(), (what, c)
