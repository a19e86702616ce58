"""biscuit_b200 -- B200-native implementation of BISCUIT's data-parallel hot path:
MC-dropout Xception-UQ inference (`uq.UncertaintyInterface`) -> per-tile mean/std -> per-slide
aggregation and uncertainty thresholding (`threshold.apply / detect / from_cv`), plus the caller of
that path, `experiment.Experiment.thresholds_from_nested_cv`, with its prediction-table loaders.

Python here is only the host mirror of the reference's call surface; every number is computed by
hand-written sm_100a kernels in libbiscuit_b200.so through the C ABI in include/biscuit_b200.h."""
from . import errors, experiment, hp, threshold, utils  # noqa: F401
from .experiment import Experiment  # noqa: F401
from .hp import ModelConfig, nature2022  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name == "uq":
        import importlib
        return importlib.import_module(".uq", __name__)
    raise AttributeError(name)
