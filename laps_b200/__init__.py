"""laps_b200 — B200-native implementation of the LAPS pseudo-spectral RHS + RK3 hot path.

Only what the path needs lives here: ``csrc/`` (hand-written CUDA for sm_100a + the C ABI
implementation), ``capi.py`` (ctypes binding of ``include/laps_b200.h``) and ``solver.py`` (the
host-side mirror of the reference's module interface).  There is no CPU implementation.
"""
from .capi import LapsError, make_params  # noqa: F401
from .solver import Solver  # noqa: F401
