"""dint_b200 — B200-native decoder for DINT-compressed posting lists (decode path of jermp/dint).

The product is the C-ABI CUDA library `libdint_b200.so` (include/dint_b200.h); this package
holds its sources (csrc/), the in-tree build, the ctypes stub (capi) and the host-side
mirror of the reference's decode drivers (driver).
"""
from . import capi  # noqa: F401
from ._build import build  # noqa: F401

__all__ = ["capi", "build"]
