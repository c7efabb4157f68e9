"""mgb_b200: B200-native Newton-step assembly behind the MultiGridBarrierMPI.jl operator API."""
from . import geometry  # noqa: F401
from .geometry import Geometry, fem1d, fem2d, fem3d  # noqa: F401
from . import amg  # noqa: F401
