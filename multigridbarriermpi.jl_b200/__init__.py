"""mgb_b200: B200-native Newton-step assembly behind the MultiGridBarrierMPI.jl operator API."""
from . import geometry  # noqa: F401
from .geometry import Geometry, fem1d, fem2d, fem3d  # noqa: F401
from . import amg  # noqa: F401


def __getattr__(name):
    # api / solver import torch: load them lazily so the symbolic (CPU) pieces stay light
    if name in ("api", "solver", "hpc", "capi", "dist"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
