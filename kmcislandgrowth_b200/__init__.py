"""kmcislandgrowth_b200 -- B200 (sm_100a) implementation of the data-parallel hot path of
YouCantCodeWithUs/KMCIslandGrowth behind the reference's own module-level API.

    from kmcislandgrowth_b200 import Periodic, Bins, Energy, Growth, Constants

Importing the package touches neither CUDA nor the shared library; the first call does.
"""
from . import Constants, Periodic, Bins, Energy, Growth, runtime  # noqa: F401

__all__ = ["Constants", "Periodic", "Bins", "Energy", "Growth", "runtime"]
