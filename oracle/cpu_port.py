"""ctypes wrapper of oracle/laps_cpu.c, the C + OpenMP restatement of the reference's RK step that bench.py
times as the CPU baseline.  TEST / BASELINE INFRASTRUCTURE ONLY (same rule as laps_oracle.py: only tests/,
__graft_entry__ and bench.py's CPU legs may import it)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "laps_cpu.c")
LIB = os.path.join(HERE, "_build", "liblaps_cpu.so")


class CpuParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("Lx", C.c_double), ("Ly", C.c_double), ("Lz", C.c_double), ("gamma", C.c_double),
                ("if_resis", C.c_int), ("if_resis_exp", C.c_int), ("eta", C.c_double),
                ("if_visc", C.c_int), ("if_visc_exp", C.c_int), ("nu", C.c_double),
                ("if_conserve_background", C.c_int), ("cfl", C.c_double),
                ("dealias_option", C.c_int), ("afx", C.c_double), ("afy", C.c_double), ("afz", C.c_double),
                ("if_AEB", C.c_int), ("if_corotating", C.c_int), ("radius0", C.c_double), ("Ur0", C.c_double),
                ("corotating_angle", C.c_double), ("if_hall", C.c_int), ("di", C.c_double)]


def build(force: bool = False) -> str:
    """gcc -O3 -fopenmp -fcx-limited-range -shared -> oracle/_build/liblaps_cpu.so"""
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        # generic x86-64 code: the library is built in one container and may run on another host
        # -fcx-limited-range: plain (ac - bd, ad + bc) complex products, without the NaN-recovery call of C99 Annex G
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-fcx-limited-range", "-std=c11", "-fPIC", "-shared", SRC, "-o", LIB, "-lm"])
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        lib.cpu_create.argtypes = [C.POINTER(CpuParams)]
        lib.cpu_create.restype = C.c_void_p
        lib.cpu_destroy.argtypes = [C.c_void_p]
        lib.cpu_set_primitive.argtypes = [C.c_void_p, dp]
        lib.cpu_vardt.argtypes = [C.c_void_p]
        lib.cpu_vardt.restype = C.c_double
        lib.cpu_evolve.argtypes = [C.c_void_p]
        lib.cpu_step.argtypes = [C.c_void_p]
        lib.cpu_step.restype = C.c_double
        lib.cpu_get_state.argtypes = [C.c_void_p, dp, dp, dp, dp]
        lib.cpu_threads.restype = C.c_int
        lib.cpu_set_threads.argtypes = [C.c_int]
        lib.cpu_set_fft.argtypes = [C.c_int]
        _lib = lib
    return _lib


class CpuPort:
    """Driven like the oracle's State: set_primitive; vardt; step ..."""

    def __init__(self, p):
        """``p``: an oracle ``Params`` (3D compressible tree) or any object with the same attributes."""
        self._lib = _load()
        cp = CpuParams(p.nx, p.ny, p.nz, p.Lx, p.Ly, p.Lz, p.adiabatic_index, int(p.if_resis), int(p.if_resis_exp), p.resistivity,
                       int(p.if_visc), int(p.if_visc_exp), p.viscosity, int(p.if_conserve_background), p.cfl,
                       p.dealias_option, p.afx, p.afy, p.afz, int(p.if_AEB), int(p.if_corotating), p.radius0, p.Ur0,
                       p.corotating_angle, int(p.if_hall), p.ion_inertial_length)
        self.shape = (p.nz, p.ny, p.nx)
        self._h = self._lib.cpu_create(C.byref(cp))
        if not self._h:
            raise MemoryError("cpu_create failed")
        self.dt = 0.0
        self.time = 0.0

    @property
    def threads(self) -> int:
        return int(self._lib.cpu_threads())

    def set_threads(self, n: int):
        """OpenMP threads for the calls that follow (overrides the launcher's OMP_NUM_THREADS)."""
        self._lib.cpu_set_threads(int(n))

    def set_fft(self, batched: bool):
        """True (default): batched SIMD transforms; False: the per-line transforms (for comparison)."""
        self._lib.cpu_set_fft(1 if batched else 0)

    def set_primitive(self, prim):
        a = np.ascontiguousarray(prim, dtype=np.float64)
        assert a.shape == (8,) + self.shape
        self._lib.cpu_set_primitive(self._h, a.ctypes.data_as(C.POINTER(C.c_double)))

    def vardt(self) -> float:
        self.dt = float(self._lib.cpu_vardt(self._h))
        return self.dt

    def step(self) -> float:
        self.dt = float(self._lib.cpu_step(self._h))
        return self.dt

    def get_state(self):
        uu = np.empty((8,) + self.shape)
        prim = np.empty((4,) + self.shape)
        t, dt = C.c_double(), C.c_double()
        dp = C.POINTER(C.c_double)
        self._lib.cpu_get_state(self._h, uu.ctypes.data_as(dp), prim.ctypes.data_as(dp), C.byref(t), C.byref(dt))
        self.time, self.dt = t.value, dt.value
        return uu, prim

    def close(self):
        if self._h:
            self._lib.cpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
