"""Exception types of the thresholding API -- same names as reference biscuit/errors.py:17-26."""


class MatchError(Exception):
    """Base of the model-lookup errors (reference errors.py:1-2)."""


class ModelNotFoundError(MatchError):
    """No trained model folder matches the experiment label (reference errors.py:5-6, utils.py:261-263)."""


class MultipleModelsFoundError(MatchError):
    """More than one model folder matches the experiment label (reference errors.py:9-10, utils.py:258-260)."""


class ThresholdError(Exception):
    """No UQ threshold could be detected in any cross-validation fold (threshold.py:539-542)."""


class ROCFailedError(Exception):
    """A group-level ROC could not be generated (threshold.py:205-206, 221-222)."""


class PredsContainNaNError(Exception):
    """Tile-level predictions contain NaN (threshold.py:141-142)."""


class NativeLibraryError(RuntimeError):
    """libbiscuit_b200.so is missing, failed to load, or a call into it failed.

    There is no CPU fallback: the product path is the CUDA library or nothing."""
