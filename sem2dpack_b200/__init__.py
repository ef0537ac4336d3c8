"""sem2dpack_b200 -- B200-native engine for SEM2DPACK's explicit time-stepping path.

The package is a thin host layer over the CUDA library `lib/libsem2d_b200.so` (C-ABI in
include/sem2d_b200.h).  There is no CPU implementation here: importing works without a GPU (so the
C-ABI can be inspected), but creating an engine without a CUDA device raises S2DError(S2D_ENODEV).
"""
from .capi import (LEAPFROG, NEWMARK, S2D_ASM_ATOMIC, S2D_ASM_COLOR, S2D_ASM_PATCH, S2DError, lib)
from .engine import CartEngine, Engine

__all__ = ["Engine", "CartEngine", "S2DError", "lib", "LEAPFROG", "NEWMARK", "S2D_ASM_PATCH", "S2D_ASM_COLOR", "S2D_ASM_ATOMIC"]
