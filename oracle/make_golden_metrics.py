"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/metrics_golden.json by running the UNMODIFIED reference
`utils.prediction_metrics` / `delong.delong_roc_variance` (np.float alias restored for the call) on seeded inputs.

    python -m oracle.make_golden_metrics
"""
import json
import os
import sys
import warnings

import numpy as np

from .make_golden import enc
from .metrics_oracle import CASES, make_case, reference_with_np_float
from .ref_shim import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "metrics_golden.json")


def main():
    warnings.simplefilter("ignore")
    R = load_reference()
    import scipy
    out = {"versions": {"numpy": np.__version__, "scipy": scipy.__version__, "python": sys.version.split()[0]},
           "generator": "oracle/make_golden_metrics.py", "cases": {}}
    for name, kw in CASES.items():
        y, p, thr = make_case(kw)
        np.random.seed(kw["seed"])
        with reference_with_np_float():
            res = R.utils.prediction_metrics(y, p, thr)
            dl = None if kw.get("single") else R.delong.delong_roc_variance(y, p)
        out["cases"][name] = {"kwargs": kw, "metrics": {k: enc(None if v is None else np.float64(v)) for k, v in res.items()},
                              "delong": None if dl is None else [enc(np.float64(dl[0])), enc(np.float64(dl[1]))]}
        print(name, {k: (None if v is None else float(v)) for k, v in res.items()})
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT)
    # two-classifier DeLong test (reference delong.py:110-123)
    from .metrics_oracle import DELONG_TEST_CASES, make_delong_test_case
    dt = {"versions": out["versions"], "generator": "oracle/make_golden_metrics.py (reference delong.delong_roc_test, np.float restored)",
          "cases": {}}
    for name, kw in DELONG_TEST_CASES.items():
        y, a, b = make_delong_test_case(kw)
        with reference_with_np_float():
            lp = R.delong.delong_roc_test(y, a, b)
        dt["cases"][name] = {"kwargs": kw, "log10_p": enc(np.float64(lp[0, 0])), "shape": list(lp.shape)}
        print(name, lp)
    with open(OUT.replace("metrics_golden", "delong_test_golden"), "w") as f:
        json.dump(dt, f, indent=1)


if __name__ == "__main__":
    main()
