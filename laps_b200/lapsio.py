"""Readers and writers of the reference's file formats — the data either side of the hot path
(SURVEY.md 8(f) rank 1).  Citations are relative to ``/root/reference/src_compressible/``.

* ``mhd.input``       Fortran namelists read by mhd.f90:30-53 (&genr, &numerical, &prl, &grid, &field,
                      &pert, &phys, &AEB, &Hall)
* ``outNNN.dat``      mhdoutput.f90:72-131 / restart.f90:17-63: a 12-byte sequential record holding
                      ``real(time,4)`` followed, at byte 12, by the GLOBAL array ``uu(nx,ny,nz,nvar)`` in
                      Fortran order (x fastest) as float64; every rank writes its block through an MPI
                      subarray view — here every rank writes its z-slab of each variable at its offset
* ``grid.dat``        mhdoutput.f90:51-60: two sequential records of float32
* ``parallel_info.dat``  mhdoutput.f90:62-69
* ``rms.dat``         mhdrms.f90:25,39-51: one line ``(f12.6,2x,19(1pe16.8))`` per call
* ``EBM_info.dat``    AEBmod.f90:75-85: ``(3(1pe16.8))``
* ``log``             mhd.f90:431-457

The reference's own post-processing reader (data_process/3D_Python/read_output.py) reads these files
back in ``tests/golden/make_io_fixtures.py``; the fixtures under ``tests/golden/`` pin the formats.
"""
from __future__ import annotations

import os
import re
import struct
from typing import Dict

import numpy as np

# ----------------------------------------------------------------------------------------------
# Fortran namelists (mhd.input)
# ----------------------------------------------------------------------------------------------
_BOOL = {"t": True, ".true.": True, "true": True, ".t.": True, "f": False, ".false.": False, "false": False, ".f.": False}


def _value(tok: str):
    t = tok.strip().rstrip(",")
    low = t.lower()
    if low in _BOOL:
        return _BOOL[low]
    if (t.startswith("'") and t.endswith("'")) or (t.startswith('"') and t.endswith('"')):
        return t[1:-1]
    try:
        return int(t)
    except ValueError:
        pass
    return float(re.sub(r"[dD]", "e", t))   # 1d-4 -> 1e-4; 1e-4 and 0. parse as they are


def read_namelists(path: str) -> Dict[str, Dict[str, object]]:
    """All ``&group ... /`` blocks of a namelist file; group and variable names are lower-cased
    (Fortran namelist input is case-insensitive); ``!`` starts a comment."""
    groups: Dict[str, Dict[str, object]] = {}
    cur = None
    with open(path) as f:
        for raw in f:
            line = raw.split("!", 1)[0].strip()
            if not line:
                continue
            if line.startswith("&"):
                cur = line[1:].split()[0].lower()
                groups[cur] = {}
                line = line[1 + len(cur):].strip()
                if not line:
                    continue
            if line in ("/", "&end", "$end"):
                cur = None
                continue
            if cur is None:
                continue
            if line.endswith("/"):
                line, end = line[:-1], True
            else:
                end = False
            for m in re.finditer(r"(\w+)\s*=\s*([^=]+?)(?=(?:,?\s*\w+\s*=)|$)", line):
                vals = [v for v in re.split(r"[,\s]+", m.group(2).strip()) if v]
                groups[cur][m.group(1).lower()] = _value(vals[0]) if len(vals) == 1 else [_value(v) for v in vals]
            if end:
                cur = None
    return groups


# ----------------------------------------------------------------------------------------------
# Fortran formatted output helpers
# ----------------------------------------------------------------------------------------------
def fmt_1pe16_8(x: float) -> str:
    """``1pe16.8``: one digit before the point, 8 after, exponent ``E+dd`` (``+ddd`` without the E for
    three-digit exponents, as Fortran prints them)."""
    s = "%.8E" % x
    mant, exp = s.split("E")
    e = int(exp)
    if abs(e) >= 100:
        s = "%s%+04d" % (mant, e)
    return s.rjust(16)


def rms_line(time: float, uu_ave, uu_rms, rho_u2) -> str:
    """mhdrms.f90:25,48 — format ``(f12.6,2x,19(1pe16.8))``."""
    vals = list(uu_ave) + list(uu_rms) + list(rho_u2)
    return "%12.6f  " % time + "".join(fmt_1pe16_8(v) for v in vals)


def ebm_line(time: float, radius: float, ur: float) -> str:
    """AEBmod.f90:81 — format ``(3(1pe16.8))``."""
    return "".join(fmt_1pe16_8(v) for v in (time, radius, ur))


def out_name(iout: int) -> str:
    """mhdoutput.f90:84-87: out000.dat ... out999.dat."""
    if not 0 <= iout < 1000:
        raise ValueError("iout must be in 0..999")
    return "out%03d.dat" % iout


def _record(payload: bytes) -> bytes:
    """One Fortran sequential unformatted record: 4-byte length, payload, 4-byte length."""
    n = struct.pack("<i", len(payload))
    return n + payload + n


# ----------------------------------------------------------------------------------------------
# writers
# ----------------------------------------------------------------------------------------------
def write_grid(path: str, xgrid, ygrid, zgrid=None):
    """mhdoutput.f90:51-60; with ``zgrid=None`` the file of the 2D trees (2D/mhdoutput.f90:45-54): nx, ny and the
    two grids only."""
    grids = [g for g in (xgrid, ygrid, zgrid) if g is not None]
    with open(path, "wb") as f:
        f.write(_record(np.array([len(g) for g in grids], dtype="<f4").tobytes()))
        f.write(_record(np.concatenate([np.asarray(g, dtype="<f4") for g in grids]).tobytes()))


def write_parallel_info(path: str, npe: int, iproc: int | None = None, jproc: int | None = None, nvar: int = 8):
    """mhdoutput.f90:62-69; with ``iproc=None`` the file of the 2D trees (2D/mhdoutput.f90:56-63): npe, nvar."""
    vals = [npe, nvar] if iproc is None else [npe, iproc, jproc, nvar]
    with open(path, "wb") as f:
        f.write(_record(np.array(vals, dtype="<f4").tobytes()))


OUT_DISPLACEMENT = 12   # mhdoutput.f90:10


def write_out_header(path: str, time: float):
    """mhdoutput.f90:88-93 (rank 0): ``status='replace'`` + one record with real(time,4)."""
    with open(path, "wb") as f:
        f.write(_record(struct.pack("<f", time)))


def write_out_slab(path: str, fields: np.ndarray, nz: int, z_offset: int):
    """mhdoutput.f90:107-118: this rank's block of the global ``uu(nx,ny,nz,nvar)`` (Fortran order).
    ``fields`` = [nvar, z_size, ny, nx] (the layout of ``Solver.get_state``); slab decomposition, so
    each variable's block is one contiguous run of z_size*ny*nx float64 at
    12 + 8*((v*nz + z_offset)*ny*nx)."""
    a = np.ascontiguousarray(fields, dtype="<f8")
    nvar, zs, ny, nx = a.shape
    with open(path, "r+b") as f:
        for v in range(nvar):
            f.seek(OUT_DISPLACEMENT + 8 * ((v * nz + z_offset) * ny * nx))
            f.write(a[v].tobytes())


def primitive_for_output(uu: np.ndarray, uu_prim: np.ndarray) -> np.ndarray:
    """mhdoutput.f90:95-103 (output_primitive = .true.): rho u -> u, e -> p."""
    out = np.array(uu, dtype=np.float64, copy=True)
    out[1:4] = uu_prim[0:3]
    out[7] = uu_prim[3]
    return out


# ----------------------------------------------------------------------------------------------
# readers
# ----------------------------------------------------------------------------------------------
def read_out_header(path: str) -> float:
    """restart.f90:31-36: the time record."""
    with open(path, "rb") as f:
        n0, t, n1 = struct.unpack("<ifi", f.read(12))
    if n0 != 4 or n1 != 4:
        raise ValueError(f"{path}: not a LAPS output file (record markers {n0}, {n1})")
    return float(t)


def read_out_slab(path: str, nx: int, ny: int, nz: int, z_offset: int = 0, z_size: int | None = None, nvar: int = 8):
    """restart.f90:50-61: this rank's block [nvar, z_size, ny, nx] of the global array."""
    z_size = nz - z_offset if z_size is None else z_size
    out = np.empty((nvar, z_size, ny, nx), dtype=np.float64)
    need = OUT_DISPLACEMENT + 8 * nvar * nz * ny * nx
    if os.path.getsize(path) < need:
        raise ValueError(f"{path}: {os.path.getsize(path)} bytes, expected at least {need}")
    with open(path, "rb") as f:
        for v in range(nvar):
            f.seek(OUT_DISPLACEMENT + 8 * ((v * nz + z_offset) * ny * nx))
            out[v] = np.fromfile(f, dtype="<f8", count=z_size * ny * nx).reshape(z_size, ny, nx)
    return out


def read_grid(path: str):
    """-> (xgrid, ygrid, zgrid), or (xgrid, ygrid) for a file of the 2D trees (the first record says which)."""
    with open(path, "rb") as f:
        raw = f.read()
    n0 = struct.unpack_from("<i", raw, 0)[0]
    dims = [int(v) for v in np.frombuffer(raw, dtype="<f4", count=n0 // 4, offset=4)]
    off = 4 + n0 + 4 + 4
    g = np.frombuffer(raw, dtype="<f4", count=sum(dims), offset=off).astype(np.float64)
    edges = np.cumsum([0] + dims)
    return tuple(g[a:b] for a, b in zip(edges[:-1], edges[1:]))


def read_parallel_info(path: str):
    """-> (npe, iproc, jproc, nvar), or (npe, nvar) for a file of the 2D trees."""
    with open(path, "rb") as f:
        raw = f.read()
    n0 = struct.unpack_from("<i", raw, 0)[0]
    return tuple(int(v) for v in np.frombuffer(raw, dtype="<f4", count=n0 // 4, offset=4))
