"""multipoint_b200: B200-native keypoint extract-and-match hot path of ethz-asl/multipoint.

Host-side mirror of the reference interface (``models.MultiPoint``, the ``utils`` helpers) over
hand-written sm_100a CUDA kernels behind a C ABI (include/multipoint_b200.h).  See DESIGN.md.
"""
from . import _lib, synthetic  # noqa: F401

__all__ = ["models", "utils", "ops", "pipeline", "parallel", "synthetic"]


def __getattr__(name):
    if name in ("models", "utils", "ops", "pipeline", "parallel"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
