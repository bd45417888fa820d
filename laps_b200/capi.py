"""ctypes binding of the C ABI in ``include/laps_b200.h``.

This is what a driver binds instead of the reference's FFTW/MPI calls.  There is exactly one
implementation behind it: the nvcc-built CUDA library ``laps_b200/_lib/liblaps_b200.so``.  If that
library is missing or no CUDA device is present, loading / ``laps_create`` fails loudly — there is
no CPU path.  (``load(path=...)`` exists so that the unit tests can point the same binding at the
test-only kernel emulator built under ``tests/_build``; the package never does that itself.)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

ABI_VERSION = 6
MAX_RANKS = 8
PEER_BLOB_BYTES = 256

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "_lib", "liblaps_b200.so")


class LapsParams(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("Lx", C.c_double), ("Ly", C.c_double), ("Lz", C.c_double),
        ("adiabatic_index", C.c_double),
        ("if_resis", C.c_int32), ("if_resis_exp", C.c_int32),
        ("resistivity", C.c_double),
        ("if_visc", C.c_int32), ("if_visc_exp", C.c_int32),
        ("viscosity", C.c_double),
        ("if_conserve_background", C.c_int32),
        ("cfl", C.c_double),
        ("dealias_option", C.c_int32),
        ("afx", C.c_double), ("afy", C.c_double), ("afz", C.c_double),
        ("if_AEB", C.c_int32), ("if_corotating", C.c_int32),
        ("radius0", C.c_double), ("Ur0", C.c_double), ("corotating_angle", C.c_double),
        ("if_hall", C.c_int32),
        ("ion_inertial_length", C.c_double),
        ("rank", C.c_int32), ("nranks", C.c_int32),
        ("device", C.c_int32),
        ("ndim", C.c_int32), ("if_z_radial", C.c_int32), ("if_limit_dt_increase", C.c_int32),
        ("incompressible", C.c_int32), ("rho0", C.c_double),
        ("if_external_force", C.c_int32),
    ]


class LapsExtents(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("nxh", C.c_int32),
                ("z_offset", C.c_int32), ("z_size", C.c_int32),
                ("y_offset", C.c_int32), ("y_size", C.c_int32), ("y_stride", C.c_int32)]


# every symbol include/laps_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "laps_create", "laps_destroy", "laps_last_error", "laps_get_extents",
    "laps_export_peer_blob", "laps_import_peer_blobs", "laps_connect_local",
    "laps_set_primitive", "laps_set_time", "laps_vardt", "laps_rkt_init", "laps_evolve", "laps_step",
    "laps_sync", "laps_get_stream", "laps_max_divb", "laps_rms", "laps_invariants", "laps_get_state", "laps_get_spectral",
    "laps_fft_forward", "laps_fft_inverse", "laps_transpose_yz_indexmap", "laps_transpose_zy_indexmap",
    "laps_last_step_ms", "laps_set_profiling", "laps_get_profile", "laps_get_pruning",
    "laps_max_divv", "laps_max_div_real", "laps_get_rho0", "laps_get_field_counts", "laps_get_output", "laps_get_pruning_counts", "laps_set_primitive_modes",
    "laps_check_nan", "laps_set_external_force", "laps_get_profile_bytes", "laps_get_footprint", "laps_set_tune",
    "laps_get_output_async", "laps_output_wait",
]


class LapsError(RuntimeError):
    pass


_libs = {}


def load(path: Optional[str] = None) -> C.CDLL:
    """Load the CUDA library (or, for tests only, an explicit ``path``)."""
    path = os.path.abspath(path or DEFAULT_LIB)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise LapsError(
            f"{path} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
            "laps_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    H = C.c_void_p
    lib.laps_create.argtypes = [C.POINTER(LapsParams), C.POINTER(H)]
    lib.laps_destroy.argtypes = [H]
    lib.laps_last_error.argtypes = [H]
    lib.laps_last_error.restype = C.c_char_p
    lib.laps_get_extents.argtypes = [H, C.POINTER(LapsExtents)]
    lib.laps_export_peer_blob.argtypes = [H, C.c_void_p]
    lib.laps_import_peer_blobs.argtypes = [H, C.c_void_p]
    lib.laps_connect_local.argtypes = [C.POINTER(H), C.c_int32]
    lib.laps_set_primitive.argtypes = [H, dp]
    lib.laps_set_primitive_modes.argtypes = [H, C.c_int32, C.POINTER(C.c_int32), dp, dp]
    lib.laps_set_time.argtypes = [H, C.c_double]
    lib.laps_vardt.argtypes = [H, dp]
    lib.laps_rkt_init.argtypes = [H, C.c_double]
    lib.laps_evolve.argtypes = [H]
    lib.laps_step.argtypes = [H, dp, dp]
    lib.laps_sync.argtypes = [H]
    lib.laps_get_stream.argtypes = [H, C.POINTER(C.c_void_p)]
    lib.laps_max_divb.argtypes = [H, dp]
    lib.laps_max_divv.argtypes = [H, dp]
    lib.laps_max_div_real.argtypes = [H, dp]
    lib.laps_get_rho0.argtypes = [H, dp]
    lib.laps_check_nan.argtypes = [H, C.POINTER(C.c_int32)]
    lib.laps_set_external_force.argtypes = [H, dp]
    lib.laps_rms.argtypes = [H, dp]
    lib.laps_invariants.argtypes = [H, dp]
    lib.laps_get_state.argtypes = [H, dp, dp]
    lib.laps_get_spectral.argtypes = [H, dp]
    lib.laps_get_output.argtypes = [H, dp, C.c_int32]
    lib.laps_fft_forward.argtypes = [H, dp, C.c_int32, dp]
    lib.laps_fft_inverse.argtypes = [H, dp, C.c_int32, dp]
    lib.laps_transpose_yz_indexmap.argtypes = [H, C.POINTER(C.c_int64)]
    lib.laps_transpose_zy_indexmap.argtypes = [H, C.POINTER(C.c_int64)]
    lib.laps_last_step_ms.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    lib.laps_get_pruning.argtypes = [H, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.laps_get_field_counts.argtypes = [H, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.laps_get_pruning_counts.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.laps_set_profiling.argtypes = [H, C.c_int32]
    lib.laps_get_profile.argtypes = [H, C.c_char_p, C.POINTER(C.c_float), C.c_int32, C.POINTER(C.c_int32)]
    lib.laps_get_profile_bytes.argtypes = [H, dp, C.c_int32, C.POINTER(C.c_int32)]
    lib.laps_get_footprint.argtypes = [H, C.POINTER(C.c_int64)]
    lib.laps_set_tune.argtypes = [H, C.c_char_p, C.c_int32]
    lib.laps_get_output_async.argtypes = [H, dp, C.c_int32]
    lib.laps_output_wait.argtypes = [H]
    for name in SYMBOLS:
        if name != "laps_last_error":
            getattr(lib, name).restype = C.c_int
    _libs[path] = lib
    return lib


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_params(**kw) -> LapsParams:
    """Defaults are the module-variable initialisers of the reference
    (mhdinit.f90:5-54, dealiasing.f90:9-10, AEBmod.f90:10-12, mhd.f90:21)."""
    p = LapsParams()
    p.abi_version = ABI_VERSION
    p.nx, p.ny, p.nz = 128, 128, 64
    p.Lx = p.Ly = p.Lz = 1.0
    p.adiabatic_index = 5.0 / 3.0
    p.cfl = 0.5
    p.dealias_option = 2
    p.afx = p.afy = p.afz = 0.495
    p.radius0 = 30.0
    p.rank, p.nranks, p.device = 0, 1, 0
    p.ndim = 3
    p.incompressible, p.rho0 = 0, 1.0
    names = {f[0] for f in LapsParams._fields_}
    for k, v in kw.items():
        if k not in names:
            raise KeyError(f"unknown laps_params field {k!r}")
        setattr(p, k, int(v) if isinstance(v, (bool, np.bool_)) else v)
    return p
