"""GPU tier for the evaluation metrics: biscuit_b200.utils.prediction_metrics / biscuit_b200.delong.delong_roc_variance (bootstrap
confusion matrices and DeLong placement values on the GPU through the C ABI) against the committed outputs of the unmodified
reference and against the pinned CPU oracle.  Bit-exact: the kernels do integer / half-integer arithmetic only and the
floating-point finish uses the reference's own library calls in the reference's order."""
import warnings

import numpy as np
import pytest

from oracle import metrics_oracle as MO

from helpers import load_golden
from test_metrics_cpu import check

pytestmark = pytest.mark.gpu
warnings.simplefilter("ignore")
GOLD = load_golden("metrics_golden.json")


@pytest.mark.parametrize("name", sorted(MO.CASES))
def test_matches_reference_golden(name):
    from biscuit_b200 import utils
    from biscuit_b200.delong import delong_roc_variance
    kw = MO.CASES[name]
    y, p, thr = MO.make_case(kw)
    np.random.seed(kw["seed"])
    res = utils.prediction_metrics(y, p, thr)
    dl = None if kw.get("single") else delong_roc_variance(y, p)
    check(res, dl, GOLD["cases"][name], name)


@pytest.mark.parametrize("n,dtype,ties", [(5000, np.float32, 200), (2047, np.float64, None), (64, np.float32, 4)])
def test_delong_matches_oracle(n, dtype, ties):
    from biscuit_b200.delong import delong_roc_variance
    rng = np.random.default_rng(n)
    y = rng.integers(0, 2, n).astype(np.int64)
    p = np.clip(0.5 + 0.1 * (2 * y - 1) + rng.normal(0, 0.3, n), 0, 1)
    if ties:
        p = np.round(p * ties) / ties
    p = p.astype(dtype)
    a, v = delong_roc_variance(y, p)
    ra, rv = MO.delong_roc_variance(y, p)
    assert float(a) == float(ra) and float(v) == float(rv), (a, ra, v, rv)
    assert np.ndim(v) == 0 and type(v) is type(rv)              # same scalar kind as the reference arithmetic gives


def test_errors():
    from biscuit_b200.delong import delong_roc_variance
    with pytest.raises(AssertionError):
        delong_roc_variance(np.ones(10, np.int64), np.linspace(0, 1, 10))


@pytest.mark.parametrize("name", sorted(MO.DELONG_TEST_CASES))
def test_delong_roc_test_matches_reference_golden(name):
    """`delong_roc_test` (reference delong.py:110-123): placement values of both classifiers on the GPU, bit-exact log10 p"""
    from biscuit_b200.delong import delong_roc_test
    from helpers import dec, same_scalar
    g = load_golden("delong_test_golden.json")["cases"][name]
    y, a, b = MO.make_delong_test_case(MO.DELONG_TEST_CASES[name])
    lp = delong_roc_test(y, a, b)
    assert list(lp.shape) == g["shape"] and same_scalar(np.float64(lp[0, 0]), dec(g["log10_p"])), (lp, dec(g["log10_p"]))
