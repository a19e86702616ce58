"""GPU tier for the nested-CV caller: biscuit_b200.experiment.Experiment.thresholds_from_nested_cv (every ROC / Youden /
slide reduction on the GPU through the C ABI) against the committed outputs of the unmodified reference and against the
pinned CPU oracle on further seeds.  Bit-exact."""
import warnings

import numpy as np
import pytest

from oracle import nested_cv_oracle as NO, synth

from helpers import load_golden
from nested_cv_common import assert_matches_golden, assert_same_outputs, build_case

pytestmark = pytest.mark.gpu
warnings.simplefilter("ignore")
GOLD = load_golden("nested_cv_golden.json")


@pytest.mark.parametrize("name", sorted(GOLD["cases"]))
def test_matches_reference_golden(name, tmp_path):
    from biscuit_b200.experiment import Experiment
    case = GOLD["cases"][name]
    project, call = build_case(tmp_path, case["kwargs"])
    df, th = Experiment(project, outcome="cohort").thresholds_from_nested_cv("EXP_AA_UQ", **call)
    assert_matches_golden(df, th, case, name)


@pytest.mark.parametrize("seed0,fmt,dtype", [(2100, "csv", np.float32), (2300, "parquet", np.float32),
                                             (2500, "parquet", np.float64)])
def test_matches_oracle(seed0, fmt, dtype, tmp_path):
    from biscuit_b200.experiment import Experiment
    project = synth.nested_cv_project(str(tmp_path), seed0=seed0, fmt=fmt, dtype=dtype, n_slides=60, tiles_per_slide=120)
    fname = "tile_predictions_val_epoch1." + ("csv" if fmt == "csv" else "parquet.gzip")
    mine = Experiment(project, outcome="cohort").thresholds_from_nested_cv("EXP_AA_UQ", tile_filename=fname)
    ref = NO.thresholds_from_nested_cv(project, "EXP_AA_UQ", tile_filename=fname)
    assert len(mine[0]) == 3
    assert_same_outputs(mine, ref, f"seed {seed0} {fmt}")
