"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the reference's own C reader of its output files,
/root/reference/data_process/3D_C/iofunctions.c, compiled where it lies by oracle/Makefile into
oracle/_ref/libref_io.so and called through ctypes.  It pins the file formats laps_b200.lapsio writes
(grid.dat, parallel_info.dat, EBM_info.dat, outNNN.dat: mhdoutput.f90:51-131, AEBmod.f90:75-85) against
reference code executed here — the only part of the reference this image can build (no Fortran, MPI, FFTW)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libref_io.so")
REF_SRC = "/root/reference/data_process/3D_C/iofunctions.c"


def build(force: bool = False) -> str | None:
    """make -C oracle (needs /root/reference); returns the library path, or None when neither the reference
    sources nor a previously built library exist (the GPU box only ever uses the prebuilt file)."""
    if os.path.exists(REF_SRC) and (force or not os.path.exists(LIB) or os.path.getmtime(REF_SRC) > os.path.getmtime(LIB)):
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, stdout=subprocess.DEVNULL)
    return LIB if os.path.exists(LIB) else None


class ReferenceReader:
    """iofunctions.h: read_grid, read_parallel_info, read_EBM, read_output.  The C functions malloc their results;
    they are copied into NumPy arrays and freed."""

    def __init__(self, path: str | None = None):
        path = path or build()
        if path is None:
            raise FileNotFoundError("oracle/_ref/libref_io.so is not built and /root/reference is absent")
        self.lib = C.CDLL(path)
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]
        fpp, dpp, ip = C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_int)
        self.lib.read_grid.argtypes = [C.c_char_p, ip, ip, ip, fpp, fpp, fpp]
        self.lib.read_parallel_info.argtypes = [C.c_char_p, ip, ip, ip, ip]
        self.lib.read_EBM.argtypes = [C.c_char_p, dpp, dpp, dpp, ip]
        self.lib.read_output.argtypes = [C.c_char_p, dpp, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int]
        for f in (self.lib.read_grid, self.lib.read_parallel_info, self.lib.read_EBM, self.lib.read_output):
            f.restype = None

    def _take(self, ptr, n, dtype):
        a = np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)
        self.libc.free(C.cast(ptr, C.c_void_p))
        return a

    def read_grid(self, filename):
        nx, ny, nz = C.c_int(), C.c_int(), C.c_int()
        x, y, z = (C.POINTER(C.c_float)() for _ in range(3))
        self.lib.read_grid(filename.encode(), C.byref(nx), C.byref(ny), C.byref(nz), C.byref(x), C.byref(y), C.byref(z))
        return (nx.value, ny.value, nz.value, self._take(x, nx.value, np.float32), self._take(y, ny.value, np.float32),
                self._take(z, nz.value, np.float32))

    def read_parallel_info(self, filename):
        v = [C.c_int() for _ in range(4)]
        self.lib.read_parallel_info(filename.encode(), *[C.byref(a) for a in v])
        return tuple(a.value for a in v)          # npe, iproc, jproc, nvar

    def read_EBM(self, filename):
        t, r, u = (C.POINTER(C.c_double)() for _ in range(3))
        n = C.c_int()
        self.lib.read_EBM(filename.encode(), C.byref(t), C.byref(r), C.byref(u), C.byref(n))
        return tuple(self._take(p, n.value, np.float64) for p in (t, r, u))

    def read_output(self, filename, nx, ny, nz, nvar=8):
        """-> (t, uu[ivar, ix, iy, iz]): the reader's own index order (macros.h IDXIJ)."""
        uu = C.POINTER(C.c_double)()
        t = C.c_float()
        self.lib.read_output(filename.encode(), C.byref(uu), C.byref(t), nx, ny, nz, nvar)
        return t.value, self._take(uu, nx * ny * nz * nvar, np.float64).reshape(nvar, nx, ny, nz)
