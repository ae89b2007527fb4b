"""cooler_b200 -- B200-native (sm_100a) ICE balancing and coarsening for
cooler-format Hi-C matrices, behind the reference's Python API.

    import cooler_b200
    bias, stats = cooler_b200.balance_cooler(clr)         # == cooler.balance_cooler

All compute runs in hand-written CUDA kernels (cooler_b200/csrc, C ABI in
include/cooler_b200.h).  There is no CPU fallback: importing the compute
entry points without the built library or without a CUDA device raises.
"""
from .balance import (ConvergenceWarning, balance_arrays, balance_cooler, balance_table,  # noqa: F401
                      iterative_correction)
from .api import matrix  # noqa: F401
from .store import H5Cooler, MemCooler, open_cooler  # noqa: F401
from .table import PixelTable  # noqa: F401

__version__ = "0.1.0"
