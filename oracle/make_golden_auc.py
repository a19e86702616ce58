"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/auc_golden.json by running the UNMODIFIED reference
`biscuit.utils.auc` / `biscuit.utils.auc_and_threshold` (reference utils.py:467-504) through oracle/ref_shim.py on
seeded inputs (installed sklearn / numpy recorded in the file).

    python -m oracle.make_golden_auc
"""
import json
import os
import sys
import warnings

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "auc_golden.json")

# name -> generator kwargs.  `big` exceeds the 131,072 ROC points up to which the CUDA path sums the trapezoid terms in
# numpy's pairwise order (bq_roc_result.auc_exact == 0 beyond it).
AUC_CASES = {
    "f32_500": dict(n=500, seed=1, dtype="float32"),
    "f64_500": dict(n=500, seed=2, dtype="float64"),
    "f32_ties_2000": dict(n=2000, seed=3, dtype="float32", ties=50),
    "f64_ties_2000": dict(n=2000, seed=4, dtype="float64", ties=20),
    "f32_separable_64": dict(n=64, seed=5, dtype="float32", separable=True),
    "f32_inverted_300": dict(n=300, seed=6, dtype="float32", inverted=True),
    "f32_single_class_100": dict(n=100, seed=7, dtype="float32", single=True),
    "f32_two_rows": dict(n=2, seed=8, dtype="float32"),
    "f32_big_200k": dict(n=200_000, seed=9, dtype="float32", big=True),
    "f64_big_150k": dict(n=150_000, seed=10, dtype="float64", big=True),
}


def make_auc_case(kw):
    rng = np.random.default_rng(kw["seed"])
    n = kw["n"]
    y = rng.integers(0, 2, n).astype(np.int64)
    if n >= 2 and y.min() == y.max():
        y[0] = 1 - y[0]
    if kw.get("single"):
        y[:] = 1
    p = np.clip(0.5 + 0.15 * (2 * y - 1) + rng.normal(0, 0.25, n), 0, 1)
    if kw.get("separable"):
        p = 0.25 + 0.5 * y + rng.uniform(-0.2, 0.2, n)
    if kw.get("inverted"):
        p = 1.0 - p
    if kw.get("ties"):
        p = np.round(p * kw["ties"]) / kw["ties"]
    return y, p.astype(kw["dtype"])


def main():
    from .make_golden import enc
    from .ref_shim import load_reference
    warnings.simplefilter("ignore")
    R = load_reference()
    import sklearn
    out = {"versions": {"numpy": np.__version__, "sklearn": sklearn.__version__, "python": sys.version.split()[0]},
           "generator": "oracle/make_golden_auc.py", "cases": {}}
    for name, kw in AUC_CASES.items():
        y, p = make_auc_case(kw)
        a = R.utils.auc(y, p)
        try:
            a2, thr = R.utils.auc_and_threshold(y, p)
            at = {"auc": enc(np.float64(a2)), "threshold": enc(np.float64(thr)), "threshold_type": type(thr).__name__}
        except ValueError as e:
            at = {"raises": "ValueError", "message": str(e)}
        out["cases"][name] = {"kwargs": kw, "auc": enc(np.float64(a)), "auc_type": type(a).__name__,
                              "auc_and_threshold": at}
        print(name, a, at)
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
