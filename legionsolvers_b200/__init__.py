"""legionsolvers_b200 -- B200-native Krylov inner loop behind the LegionSolvers task API.

The product is liblsk.so (hand-written sm_100a CUDA kernels reached through the C ABI in
include/lsk.h, plus the C++ host layer that mirrors PartitionedVector / CSRMatrix / COOMatrix /
SquarePlanner / CGSolver / BiCGStabSolver / GMRESSolver).  This Python package is a thin ctypes
driver used by the tests and the benchmark; PyTorch only supplies device memory, streams and
torch.distributed rendezvous.
"""
from .build import LIB_PATH, build_library  # noqa: F401

__all__ = ["LIB_PATH", "build_library"]
__version__ = "0.1.0"
