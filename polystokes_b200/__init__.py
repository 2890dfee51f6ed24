"""polystokes_b200 -- B200-native per-step viscous Stokes solve behind the HDK_PolyStokesSolver entry point.

Only the hot path of panuelosj/polystokes lives here: classify -> assemble -> PCG,
as hand-written sm_100a CUDA kernels behind the C ABI in ``include/polystokes_b200.h``.
"""
from .solver import PolyStokesSolver, PolyStokesError, SLOTS, slot_shape  # noqa: F401
from ._capi import (PS_SUCCESS, PS_NOCONVERGE, PS_FAILED, PS_INVALID, PS_UNSUPPORTED_SOLVER, PS_INCOMPLETE,  # noqa: F401
                    PS_MEM_HOST, PS_MEM_DEVICE, PRODUCT_LIB)
from . import scenes  # noqa: F401
