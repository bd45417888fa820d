"""Host-side mirror of the reference's module interface for the hot path.

Method names follow the Fortran subroutines they stand in for (``evolve``, ``vardt``,
``evolve_radius``, ``calc_rms``, ``calc_max_divB`` ...; file:line in each docstring, relative to
``src_compressible/``) so that parity tests read like a LAPS driver.  All work is done by the CUDA
library through the C ABI (``capi.py``); arrays here are only the host copies a driver owns.

Array convention (same memory layout as the Fortran arrays): real fields ``a[v, iz_local, iy, ix]``
(= ``uu(ix,iy,iz,v)``), spectral fields ``a[v, kz, ky_local, kx]`` (= ``uu_fourier(ix,iy,iz,v)``).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import capi


class Solver:
    def __init__(self, lib_path: Optional[str] = None, **params):
        self._lib = capi.load(lib_path)
        self.params = capi.make_params(**params)
        self._h = C.c_void_p()
        rc = self._lib.laps_create(C.byref(self.params), C.byref(self._h))
        if rc:
            raise capi.LapsError("laps_create: " + self._lib.laps_last_error(None).decode())
        ext = capi.LapsExtents()
        self._ck(self._lib.laps_get_extents(self._h, C.byref(ext)))
        self.ext = ext
        self.nx, self.ny, self.nz, self.nxh = ext.nx, ext.ny, ext.nz, ext.nxh
        self.nzl, self.nyl = ext.z_size, ext.y_size
        # global ky of this rank's Fourier rows (contiguous slab by default, see laps_extents.y_stride)
        self.ky_rows = ext.y_offset + np.arange(ext.y_size) * max(1, ext.y_stride)
        self.time = 0.0
        self.dt = 0.0

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc:
            raise capi.LapsError(self._lib.laps_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._lib.laps_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def real_shape(self):
        return (self.nzl, self.ny, self.nx)

    # ------------------------------------------------------------------ multi-rank wiring
    def export_peer_blob(self) -> bytes:
        """This rank's exchange-buffer descriptor (CUDA IPC handles); all-gather and import."""
        buf = C.create_string_buffer(capi.PEER_BLOB_BYTES)
        self._ck(self._lib.laps_export_peer_blob(self._h, buf))
        return buf.raw

    def import_peer_blobs(self, blobs: bytes):
        assert len(blobs) == capi.PEER_BLOB_BYTES * self.params.nranks
        self._ck(self._lib.laps_import_peer_blobs(self._h, C.create_string_buffer(blobs, len(blobs))))

    @staticmethod
    def connect_local(solvers):
        """laps_connect_local: wire the handles of all ranks created in THIS process (ordered by rank) to each other.
        Every handle must afterwards be driven by its own host thread (the collectives wait for each other)."""
        arr = (C.c_void_p * len(solvers))(*[g._h for g in solvers])
        rc = solvers[0]._lib.laps_connect_local(arr, len(solvers))
        if rc:
            msgs = [g._lib.laps_last_error(g._h).decode() for g in solvers]
            raise capi.LapsError("laps_connect_local: " + "; ".join(m for m in msgs if m))

    # ------------------------------------------------------------------ driver-facing calls
    def set_primitive(self, prim: np.ndarray):
        """initial_calc_conserve_variable + transform_uu_real_to_fourier (mhd.f90:121-122)."""
        a = np.ascontiguousarray(prim, dtype=np.float64)
        assert a.shape == (8,) + self.real_shape, (a.shape, self.real_shape)
        self._ck(self._lib.laps_set_primitive(self._h, capi._dptr(a)))

    def set_primitive_modes(self, ks, coefs, background):
        """laps_set_primitive_modes: ``ks`` int [nmodes, 3] (kx >= 0), ``coefs`` complex [7, nmodes]
        (``synthetic.mode_table``), ``background`` = 8 uniform values (rho, u, B, p)."""
        k = np.ascontiguousarray(ks, dtype=np.int32).reshape(-1, 3)
        c = np.ascontiguousarray(coefs, dtype=np.complex128)
        assert c.shape == (7, k.shape[0])
        b = np.ascontiguousarray(background, dtype=np.float64)
        assert b.shape == (8,)
        self._ck(self._lib.laps_set_primitive_modes(self._h, k.shape[0], k.ctypes.data_as(C.POINTER(C.c_int32)),
                                                    c.ctypes.data_as(C.POINTER(C.c_double)), capi._dptr(b)))

    def evolve_radius(self, time: float):
        """AEBmod.f90:56-73."""
        self._ck(self._lib.laps_set_time(self._h, float(time)))

    def vardt(self) -> float:
        """mhd.f90:328-429."""
        dt = C.c_double(self.dt)
        self._ck(self._lib.laps_vardt(self._h, C.byref(dt)))
        self.dt = dt.value
        return self.dt

    def rkt_init(self, dt: float):
        """rktmod.f90:15-32."""
        self._ck(self._lib.laps_rkt_init(self._h, float(dt)))
        self.dt = float(dt)

    def evolve(self):
        """mhd.f90:298-326 (asynchronous)."""
        self._ck(self._lib.laps_evolve(self._h))

    def step(self, calc_dt: bool = True) -> float:
        """One pass of the Principal loop body (mhd.f90:245-248,285).  ``calc_dt=False`` leaves out
        vardt, as the 2D tree's driver does on 19 steps out of 20 (2D/mhd.f90:237-240)."""
        if not calc_dt:
            self.evolve()
            self.time = self.time + self.dt
            self.evolve_radius(self.time)
            return self.dt
        t = C.c_double(self.time)
        dt = C.c_double(self.dt)
        self._ck(self._lib.laps_step(self._h, C.byref(t), C.byref(dt)))
        self.time, self.dt = t.value, dt.value
        return self.dt

    def sync(self):
        self._ck(self._lib.laps_sync(self._h))

    def cuda_stream(self) -> int:
        """cudaStream_t of this handle as an integer (for torch.cuda.ExternalStream)."""
        out = C.c_void_p()
        self._ck(self._lib.laps_get_stream(self._h, C.byref(out)))
        return out.value or 0

    def calc_max_divB(self) -> float:
        """mhd.f90:522-570."""
        out = C.c_double()
        self._ck(self._lib.laps_max_divb(self._h, C.byref(out)))
        return out.value

    def calc_max_divV(self) -> float:
        """src_incompressible/mhd.f90:620-668."""
        out = C.c_double()
        self._ck(self._lib.laps_max_divv(self._h, C.byref(out)))
        return out.value

    def calc_max_div_real(self):
        """calc_divB_real/calc_divV_real + calc_max_div*_real (src_incompressible/mhdrhs.f90:532-648,
        mhd.f90:672-732) -> (max |div B|, max |div u|) in real space."""
        out = np.zeros(2)
        self._ck(self._lib.laps_max_div_real(self._h, capi._dptr(out)))
        return float(out[0]), float(out[1])

    def checkNan(self) -> bool:
        """2D/mhd.f90:563-591 (checkNan): is any uu value a NaN on any rank."""
        out = C.c_int32()
        self._ck(self._lib.laps_check_nan(self._h, C.byref(out)))
        return bool(out.value)

    def set_external_force(self, force: np.ndarray):
        """The field of calc_external_force_real (2D/mhdrhs.f90:480-531), shape (1, ny, nx) like one variable of uu;
        it is transformed and added to fnl(7) in every stage until it is replaced."""
        f = np.ascontiguousarray(force, dtype=np.float64)
        assert f.size == int(np.prod(self.real_shape)), (f.shape, self.real_shape)
        self._ck(self._lib.laps_set_external_force(self._h, capi._dptr(f)))

    @property
    def rho0(self) -> float:
        """Background density of the incompressible tree after update_rho_p (AEBmod.f90:123-134)."""
        out = C.c_double()
        self._ck(self._lib.laps_get_rho0(self._h, C.byref(out)))
        return out.value

    def calc_rms(self):
        """mhdrms.f90:53-126 -> (uu_ave[8], uu_rms[8], rho_u2[3])."""
        out = np.zeros(19)
        self._ck(self._lib.laps_rms(self._h, capi._dptr(out)))
        return out[:8].copy(), out[8:16].copy(), out[16:].copy()

    def invariants(self) -> np.ndarray:
        out = np.zeros(3)
        self._ck(self._lib.laps_invariants(self._h, capi._dptr(out)))
        return out

    def get_state(self, want_prim=True, out_uu=None, out_prim=None):
        """Host copies of uu and uu_prim for output_uu / restart (mhdoutput.f90:95-123)."""
        uu = np.empty((8,) + self.real_shape) if out_uu is None else out_uu
        prim = (np.empty((4,) + self.real_shape) if out_prim is None else out_prim) if want_prim else None
        assert uu.shape == (8,) + self.real_shape and uu.dtype == np.float64 and uu.flags.c_contiguous
        self._ck(self._lib.laps_get_state(self._h, capi._dptr(uu), capi._dptr(prim) if want_prim else None))
        return uu, prim

    def get_output(self, primitive=True, out=None):
        """The array output_uu writes (mhdoutput.f90:95-123): [rho, u, B, p] or the conserved uu."""
        a = np.empty((8,) + self.real_shape) if out is None else out
        assert a.shape == (8,) + self.real_shape and a.dtype == np.float64 and a.flags.c_contiguous
        self._ck(self._lib.laps_get_output(self._h, capi._dptr(a), 1 if primitive else 0))
        return a

    def get_output_async(self, out: np.ndarray, primitive=True):
        """laps_get_output_async: the dump leaves on a copy stream while the following steps run; ``out`` (pinned, to
        overlap) must stay alive until ``output_wait`` returns."""
        assert out.shape == (8,) + self.real_shape and out.dtype == np.float64 and out.flags.c_contiguous
        self._ck(self._lib.laps_get_output_async(self._h, capi._dptr(out), 1 if primitive else 0))
        return out

    def output_wait(self):
        self._ck(self._lib.laps_output_wait(self._h))

    def uu_fourier(self) -> np.ndarray:
        """Spectral state in the reference index order [v, kz, ky_local, kx]."""
        raw = np.empty((8, self.nxh, self.nyl, self.nz), dtype=np.complex128)
        self._ck(self._lib.laps_get_spectral(self._h, raw.ctypes.data_as(C.POINTER(C.c_double))))
        return np.ascontiguousarray(raw.transpose(0, 3, 2, 1))

    def fft_forward(self, fields: np.ndarray) -> np.ndarray:
        """fftw.f90:42-71 + 136-180 for up to 8 fields; result [f, kz, ky_local, kx]."""
        a = np.ascontiguousarray(fields, dtype=np.float64)
        nf = a.shape[0]
        assert a.shape == (nf,) + self.real_shape
        raw = np.empty((nf, self.nxh, self.nyl, self.nz), dtype=np.complex128)
        self._ck(self._lib.laps_fft_forward(self._h, capi._dptr(a), nf, raw.ctypes.data_as(C.POINTER(C.c_double))))
        return np.ascontiguousarray(raw.transpose(0, 3, 2, 1))

    def fft_inverse(self, spec: np.ndarray) -> np.ndarray:
        """fftw.f90:73-103 + 182-222 for up to 8 fields given as [f, kz, ky_local, kx]."""
        nf = spec.shape[0]
        assert spec.shape == (nf, self.nz, self.nyl, self.nxh)
        raw = np.ascontiguousarray(np.asarray(spec, dtype=np.complex128).transpose(0, 3, 2, 1))
        out = np.empty((nf,) + self.real_shape)
        self._ck(self._lib.laps_fft_inverse(self._h, raw.ctypes.data_as(C.POINTER(C.c_double)), nf, capi._dptr(out)))
        return out

    def transpose_yz_indexmap(self) -> np.ndarray:
        out = np.empty((self.nxh * self.ny * self.nzl, 2), dtype=np.int64)
        self._ck(self._lib.laps_transpose_yz_indexmap(self._h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    def transpose_zy_indexmap(self) -> np.ndarray:
        out = np.empty((self.nxh * self.nyl * self.nz, 2), dtype=np.int64)
        self._ck(self._lib.laps_transpose_zy_indexmap(self._h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    # ------------------------------------------------------------------ measurement helpers
    def last_step_ms(self):
        ms = C.c_float()
        n = C.c_int32()
        self._ck(self._lib.laps_last_step_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def pruning(self):
        """(nkx, kymax, nky_local): the columns that survive the dealiasing mask (laps_get_pruning)."""
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self._lib.laps_get_pruning(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def pruning_counts(self):
        """(live_columns, live_modes) of this rank (laps_get_pruning_counts)."""
        a, b = C.c_int64(), C.c_int64()
        self._ck(self._lib.laps_get_pruning_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def field_counts(self):
        """(nf, ni, spec_rows): fields transformed forward / inverse per RK stage and the state rows of the
        main z-pass launch (laps_get_field_counts)."""
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self._lib.laps_get_field_counts(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def set_profiling(self, on: bool):
        self._ck(self._lib.laps_set_profiling(self._h, 1 if on else 0))

    def get_profile(self, cap=512, with_bytes=False):
        """(name, ms) of every launch of the last instrumented evolve/step; with_bytes: (name, ms, algorithmic bytes)."""
        names = C.create_string_buffer(cap * 32)
        ms = (C.c_float * cap)()
        cnt = C.c_int32()
        self._ck(self._lib.laps_get_profile(self._h, names, ms, cap, C.byref(cnt)))
        by = (C.c_double * cap)()
        if with_bytes:
            c2 = C.c_int32()
            self._ck(self._lib.laps_get_profile_bytes(self._h, by, cap, C.byref(c2)))
        out = []
        for i in range(cnt.value):
            nm = names.raw[i * 32:(i + 1) * 32].split(b"\0", 1)[0].decode()
            out.append((nm, ms[i], by[i]) if with_bytes else (nm, ms[i]))
        return out

    def footprint(self) -> int:
        """Device bytes this handle allocated (laps_get_footprint)."""
        out = C.c_int64()
        self._ck(self._lib.laps_get_footprint(self._h, C.byref(out)))
        return out.value

    def set_tune(self, name: str, value: int):
        """laps_set_tune: switch between equivalent kernels / launch shapes on a live handle (measurement helper)."""
        self._ck(self._lib.laps_set_tune(self._h, name.encode(), int(value)))
