"""fruid_b200 -- B200-native stable-fluids time step behind the fruid crate API.

Importing this package loads fruid_b200/lib/libfruid_b200.so (hand-written sm_100a
CUDA behind the C ABI of include/fruid_b200.h) and raises if it is missing: there is
no CPU fallback.
"""
from . import _lib, ops, render  # noqa: F401
from ._lib import FruidError, launch_count, LIB_PATH  # noqa: F401
from .api import Array2D, DensitySim, FluidSim  # noqa: F401

__all__ = ["Array2D", "DensitySim", "FluidSim", "FruidError", "ops", "launch_count", "LIB_PATH"]
