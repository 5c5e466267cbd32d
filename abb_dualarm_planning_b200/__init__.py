"""abb_dualarm_planning_b200 -- B200-native batched state validity for the closed-chain dual ABB IRB120 + rod system.

The product is the C-ABI shared library `csrc/libckc.so` (include/ckc.h) with hand-written sm_100a CUDA kernels, and
the C++ drop-in `StateValidityChecker` classes in `host/`.  This Python package is only a thin ctypes binding used by
the tests and bench.py; it never computes anything itself and fails loudly when the library is missing.
"""
from .binding import Checker, CkcError, lib_path, load_library, MESH_PACK, Stats  # noqa: F401
