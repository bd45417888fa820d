"""CPU oracle for the LAPS 3D compressible Hall-MHD + expanding-box hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``laps_b200/``) may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do.

This is a NumPy/SciPy FP64 restatement of the reference algorithm, function by function.
All citations are ``file:line`` relative to ``/root/reference/src_compressible/``.

PARITY: the reference ships no tests, golden vectors or fixtures, and it cannot be compiled in this image (no
Fortran compiler, MPI or FFTW).  What pins this restatement is the reference's own source EXECUTED here:
oracle/fortran_exec.py translates the hot-path subroutines of all four source trees statement by statement
(src_*/{mhdinit,dealiasing,AEBmod,rktmod,mhdrhs,fftw,parallel,mhd,mhdrms}.f90, read where they lie) and
tests/golden/make_ref_exec_fixtures.py runs program mhd's sequence on one rank — grid/dealias/AEB initialisation,
vardt, evolve with every loop around the FFTW calls and the single-rank branches of the transposes, diagnostics —
storing golden vectors under tests/golden/ref_exec/ (tests/test_reference_source_pins.py: this oracle agrees to
1e-12..1e-14).  Substituted, and therefore NOT pinned that way: FFTW's 1-D executions (third-party, un-vendored,
version unpinned: makefile:7,13 mention 3.3.4/3.3.8; numpy.fft computes the same DFT definition, here restated with
``scipy.fft``), MPI (one rank: the sendrecv loops have zero trips), and compiler-specific evaluation order.  In
addition analytic known answers (tests/test_oracle_analytic.py): Alfven-wave translation with third-order
convergence, the exact RK3-polynomial decay of the k=0 mode in the expanding box, bit-exact conservation of the k=0
mode, div B at round-off, Hall-MHD dispersion relation.

Array conventions: every array is NumPy C-order with the *last* axis = x, i.e. ``a[v, iz, iy, ix]``
has exactly the memory layout of the Fortran ``a(ix, iy, iz, v)`` (x fastest).  Spectral arrays
are ``[v, kz, ky, kx]`` with ``kx`` in ``0..nx/2``.  Indices are 0-based here.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import Optional

import numpy as np
import scipy.fft as sfft

PI = 3.141592653589793  # mhdinit.f90:7 (literal; FP64 under -r8)

_WORKERS = int(os.environ.get("LAPS_ORACLE_WORKERS", "0")) or (os.cpu_count() or 1)


# --------------------------------------------------------------------------------------
# parameters (the namelists of mhd.f90:30-53 that matter to the hot path)
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class Params:
    nx: int = 64
    ny: int = 64
    nz: int = 64
    Lx: float = 24.0
    Ly: float = 24.0
    Lz: float = 24.0
    adiabatic_index: float = 5.0 / 3.0          # mhdinit.f90:21
    if_resis: bool = False
    if_resis_exp: bool = False
    resistivity: float = 0.0
    if_visc: bool = False
    if_visc_exp: bool = False
    viscosity: float = 0.0
    if_conserve_background: bool = False
    if_AEB: bool = False
    if_corotating: bool = False
    corotating_angle: float = 0.0
    radius0: float = 30.0                        # AEBmod.f90:12
    Ur0: float = 0.0
    if_hall: bool = False
    ion_inertial_length: float = 0.0
    cfl: float = 0.5                             # mhd.f90:21
    dealias_option: int = 2                      # dealiasing.f90:9 (module default)
    afx: float = 0.495
    afy: float = 0.495
    afz: float = 0.495
    # 2D tree only (src_compressible/2D/mhd.f90:23,35,44; 2D/mhdinit.f90:23)
    if_z_radial: bool = False
    if_limit_dt_increase: bool = False
    if_external_force: bool = False              # 2D/mhdinit.f90:24 (&pert, 2D/mhd.f90:43)
    # incompressible tree only (src_incompressible/mhdinit.f90:15): which State class applies, and rho0
    incompressible: bool = False
    rho0: float = 1.0


# --------------------------------------------------------------------------------------
# decomposition and index maps (parallel.f90)
# --------------------------------------------------------------------------------------
def decompose_1d(ngrid: int, nproc: int):
    """parallel.f90:326-349 — every rank gets ngrid//nproc, the last rank the remainder."""
    normal = ngrid // nproc
    offset = np.zeros(nproc, dtype=np.int64)
    size = np.zeros(nproc, dtype=np.int64)
    offset[0] = 0
    size[0] = normal
    for ip in range(1, nproc):
        offset[ip] = offset[ip - 1] + size[ip - 1]
        size[ip] = normal if ip < nproc - 1 else ngrid - offset[ip]
    return offset, size


class Decomp:
    """Index tables of parallel_start (parallel.f90:100-103) for an iproc x jproc grid.

    Slab mode (ndim_parallel=1, parallel.f90:56-58) is iproc=1, jproc=npe.
    """

    def __init__(self, nx, ny, nz, iproc, jproc):
        self.nx, self.ny, self.nz = nx, ny, nz
        self.iproc, self.jproc = iproc, jproc
        self.nxh = nx // 2 + 1
        self.yi_offset, self.yi_size = decompose_1d(ny, iproc)
        self.yj_offset, self.yj_size = decompose_1d(ny, jproc)
        self.zj_offset, self.zj_size = decompose_1d(nz, jproc)
        self.xi_offset, self.xi_size = decompose_1d(self.nxh, iproc)

    # local linear offsets (units of one element) — SURVEY 9.9 / parallel.f90:105-143
    def off_w_xyz(self, mi, mj, gx, gy, gz):
        return gx + self.nxh * ((gy - self.yi_offset[mi]) + self.yi_size[mi] * (gz - self.zj_offset[mj]))

    def off_w_yxz(self, mi, mj, gx, gy, gz):
        return (gx - self.xi_offset[mi]) + self.xi_size[mi] * (gy + self.ny * (gz - self.zj_offset[mj]))

    def off_w_zxy(self, mi, mj, gx, gy, gz):
        return (gx - self.xi_offset[mi]) + self.xi_size[mi] * ((gy - self.yj_offset[mj]) + self.yj_size[mj] * gz)

    def transpose_yz_sendmap(self, mi, mj, q):
        """Global (gx,gy,gz) triples that rank (mi,mj) sends to peer q in comm1d_j by
        transpose_yz (parallel.f90:185-196,273-297): [all local gx] x [gy in Yj(q)] x [gz in Zj(me)],
        in the element order of the MPI subarray datatype (x fastest)."""
        gx = np.arange(self.xi_offset[mi], self.xi_offset[mi] + self.xi_size[mi])
        gy = np.arange(self.yj_offset[q], self.yj_offset[q] + self.yj_size[q])
        gz = np.arange(self.zj_offset[mj], self.zj_offset[mj] + self.zj_size[mj])
        Z, Y, X = np.meshgrid(gz, gy, gx, indexing="ij")
        return X.ravel(), Y.ravel(), Z.ravel()


def transpose_yz_distributed(dec: Decomp, w_yxz_blocks):
    """Emulate transpose_yz (parallel.f90:273-297) for slab mode over all ranks at once.

    ``w_yxz_blocks[mj]`` is rank mj's local w_yxz as a flat array (layout off_w_yxz).  Returns the
    list of flat local w_zxy arrays (layout off_w_zxy).  Pure index shuffling: bit-exact.
    """
    assert dec.iproc == 1
    P = dec.jproc
    out = [np.zeros(dec.xi_size[0] * dec.yj_size[mj] * dec.nz, dtype=w_yxz_blocks[0].dtype) for mj in range(P)]
    for src in range(P):
        for dst in range(P):
            X, Y, Z = dec.transpose_yz_sendmap(0, src, dst)
            out[dst][dec.off_w_zxy(0, dst, X, Y, Z)] = w_yxz_blocks[src][dec.off_w_yxz(0, src, X, Y, Z)]
    return out


# --------------------------------------------------------------------------------------
# grid / wave numbers (mhdinit.f90:58-124)
# --------------------------------------------------------------------------------------
def wave_numbers(n: int, L: float) -> np.ndarray:
    """mhdinit.f90:79-110 — 2*pi*(i-1)/L for i<=n/2+1 (Nyquist kept POSITIVE), else 2*pi*(i-1-n)/L."""
    k = np.empty(n, dtype=np.float64)
    for i in range(1, n + 1):
        if i <= n // 2 + 1:
            k[i - 1] = 2 * PI * (i - 1) / L
        else:
            k[i - 1] = 2 * PI * (i - 1 - n) / L
    return k


class Grid:
    def __init__(self, p: Params):
        self.p = p
        self.nxh = p.nx // 2 + 1
        self.dx, self.dy, self.dz = p.Lx / p.nx, p.Ly / p.ny, p.Lz / p.nz  # mhdinit.f90:75-77
        self.xgrid = np.arange(p.nx) * self.dx
        self.ygrid = np.arange(p.ny) * self.dy
        self.zgrid = np.arange(p.nz) * self.dz
        self.wnx = wave_numbers(p.nx, p.Lx)
        self.wny = wave_numbers(p.ny, p.Ly)
        self.wnz = wave_numbers(p.nz, p.Lz)
        # broadcast views over the half spectrum [kz, ky, kx]
        self.KX = self.wnx[: self.nxh][None, None, :]
        self.KY = self.wny[None, :, None]
        self.KZ = self.wnz[:, None, None]


# --------------------------------------------------------------------------------------
# FFTs (fftw.f90)
# --------------------------------------------------------------------------------------
def fft_forward(a: np.ndarray) -> np.ndarray:
    """fftw.f90:42-71 + 136-180: r2c along x (/nx), c2c along y (/ny), c2c along z (/nz).
    ``a`` is [..., nz, ny, nx] real; result [..., nz, ny, nx/2+1] complex."""
    nz, ny, nx = a.shape[-3:]
    w = sfft.rfft(a, axis=-1, workers=_WORKERS) / nx
    w = sfft.fft(w, axis=-2, workers=_WORKERS) / ny
    w = sfft.fft(w, axis=-3, workers=_WORKERS) / nz
    return w


def fft_inverse(w: np.ndarray, nx: int) -> np.ndarray:
    """fftw.f90:73-103 + 182-222: unnormalised backward c2c along z, then y, then c2r along x.
    FFTW's c2r ignores the imaginary parts of the DC and Nyquist bins; pocketfft's c2r does the
    same (verified in tests/test_oracle_analytic.py)."""
    nz, ny = w.shape[-3], w.shape[-2]
    a = sfft.ifft(w, axis=-3, workers=_WORKERS) * nz
    a = sfft.ifft(a, axis=-2, workers=_WORKERS) * ny
    return sfft.irfft(a, n=nx, axis=-1, workers=_WORKERS) * nx


# --------------------------------------------------------------------------------------
# state
# --------------------------------------------------------------------------------------
class State:
    """Module-level arrays of mhdinit.f90:34-43 for one (undistributed) domain."""

    def __init__(self, p: Params):
        self.p = p
        self.g = Grid(p)
        shp = (p.nz, p.ny, p.nx)
        self.uu = np.zeros((8,) + shp)
        self.uu_prim = np.zeros((4,) + shp)
        sshp = (p.nz, p.ny, self.g.nxh)
        self.uu_fourier = np.zeros((8,) + sshp, dtype=np.complex128)
        self.fnl = np.zeros((8,) + sshp, dtype=np.complex128)
        self.fnl_rk = np.zeros((8,) + sshp, dtype=np.complex128)
        self.current_density = np.zeros((3,) + shp) if p.if_hall else None
        # AEB_initialize (AEBmod.f90:16-44); mhd.f90:88-90 forces Ur0=0 when AEB is off
        self.Ur0 = p.Ur0 if p.if_AEB else 0.0
        self.radius = p.radius0
        self._aeb_calc()
        ang = p.corotating_angle if p.if_corotating else 0.0
        self.cos_cor_ang = math.cos(ang)
        self.sin_cor_ang = math.sin(ang)
        self.k_square = self.g.KX ** 2 + self.g.KY ** 2 + self.g.KZ ** 2  # mhdinit.f90:114-122
        self.time = 0.0
        self.dt = 0.0
        self.cc1 = np.zeros(3)
        self.dd1 = np.zeros(3)
        self.time_step = np.zeros(3)
        self._filters = None
        if p.dealias_option == 2:
            self._filters = dealias_filters(p, self.g)

    # AEBmod.f90:46-54
    def _aeb_calc(self):
        self.Ur = self.Ur0
        with np.errstate(divide="ignore"):
            self.tau_exp = np.float64(self.radius) / np.float64(self.Ur)

    # AEBmod.f90:56-73
    def evolve_radius(self, t: float):
        self.radius = self.p.radius0 + self.Ur * t
        self._aeb_calc()
        self.update_ksquare()

    # AEBmod.f90:87-124
    def update_ksquare(self):
        p, g = self.p, self.g
        kx, ky, kz = g.KX, g.KY, g.KZ
        r0, r = p.radius0, self.radius
        if p.if_corotating:
            c, s = self.cos_cor_ang, self.sin_cor_ang
            self.k_square = (kx ** 2 * (c ** 2 + (s * r0 / r) ** 2)
                             + ky ** 2 * (s ** 2 + (c * r0 / r) ** 2)
                             + kx * ky * 2 * c * s * (1 - (r0 / r) ** 2)
                             + (kz * r0 / r) ** 2)
        else:
            self.k_square = kx ** 2 + (ky * r0 / r) ** 2 + (kz * r0 / r) ** 2

    # derivative vectors, mhdrhs.f90:191-204 (also :313-326, mhd.f90:542-555)
    def kvec(self):
        """Returns (kx, ky, kz) as *imaginary parts* (the Fortran values are cmplx(0, .))."""
        p, g = self.p, self.g
        r0, r = p.radius0, self.radius
        kz = g.KZ * r0 / r
        ky = g.KY * r0 / r
        kx = g.KX + 0.0 * g.KY
        if p.if_AEB and p.if_corotating:
            c, s = self.cos_cor_ang, self.sin_cor_ang
            kx = g.KX * c + g.KY * s
            ky = (-g.KX * s + g.KY * c) * r0 / r
        return kx, ky, kz

    # ---------------------------------------------------------------- initial data
    def set_primitive(self, prim: np.ndarray):
        """initial_calc_conserve_variable (mhdinit.f90:1038-1056) + transform_uu_real_to_fourier
        (fftw.f90:42-71).  ``prim`` = [rho, ux, uy, uz, bx, by, bz, p]."""
        gam = self.p.adiabatic_index
        uu = np.array(prim, dtype=np.float64, copy=True)
        self.uu_prim[0:3] = uu[1:4]
        self.uu_prim[3] = uu[7]
        uu[1] = uu[0] * self.uu_prim[0]
        uu[2] = uu[0] * self.uu_prim[1]
        uu[3] = uu[0] * self.uu_prim[2]
        uu[7] = self.uu_prim[3] / (gam - 1) + 0.5 * (
            uu[0] * (self.uu_prim[0] ** 2 + self.uu_prim[1] ** 2 + self.uu_prim[2] ** 2)
            + uu[4] ** 2 + uu[5] ** 2 + uu[6] ** 2)
        self.uu = uu
        self.uu_fourier = fft_forward(self.uu)

    # ---------------------------------------------------------------- hot path pieces
    def calc_current_density_real(self):
        """mhdrhs.f90:296-362 — J = i k x B (stretched/rotated k), three inverse 3D FFTs."""
        kx, ky, kz = self.kvec()
        uf = self.uu_fourier
        jf = np.empty((3,) + uf.shape[1:], dtype=np.complex128)
        jf[0] = 1j * ky * uf[6] - 1j * kz * uf[5]
        jf[1] = 1j * kz * uf[4] - 1j * kx * uf[6]
        jf[2] = 1j * kx * uf[5] - 1j * ky * uf[4]
        self.current_density = fft_inverse(jf, self.p.nx)
        return jf

    def calc_flux(self):
        """mhdrhs.f90:21-124 — the 18 real-space fluxes (+ EBM energy source)."""
        p = self.p
        if p.if_hall:
            self.calc_current_density_real()
        uu, pr = self.uu, self.uu_prim
        P = pr[3]
        Bx, By, Bz = uu[4], uu[5], uu[6]
        ux, uy, uz = pr[0], pr[1], pr[2]
        rho = uu[0]
        ptot = P + 0.5 * (Bx ** 2 + By ** 2 + Bz ** 2)
        udotb = ux * Bx + uy * By + uz * Bz
        flux = np.empty((18,) + uu.shape[1:])
        flux[0:3] = uu[1:4]
        flux[3] = uu[1] * ux - Bx * Bx + ptot
        flux[4] = uu[2] * ux - By * Bx
        flux[5] = uu[3] * ux - Bz * Bx
        flux[6] = uu[1] * uy - Bx * By
        flux[7] = uu[2] * uy - By * By + ptot
        flux[8] = uu[3] * uy - Bz * By
        flux[9] = uu[1] * uz - Bx * Bz
        flux[10] = uu[2] * uz - By * Bz
        flux[11] = uu[3] * uz - Bz * Bz + ptot
        flux[12] = uz * By - uy * Bz
        flux[13] = ux * Bz - uz * Bx
        flux[14] = uy * Bx - ux * By
        flux[15] = (uu[7] + ptot) * ux - udotb * Bx
        flux[16] = (uu[7] + ptot) * uy - udotb * By
        flux[17] = (uu[7] + ptot) * uz - udotb * Bz
        expand = None
        if p.if_AEB:
            gam, tau = p.adiabatic_index, self.tau_exp
            expand = (-2 * gam / (gam - 1) * P / tau
                      - (2.0 * Bx ** 2 + By ** 2 + Bz ** 2) / tau
                      - (uu[1] * ux + 2 * uu[2] * uy + 2 * uu[3] * uz) / tau)
        if p.if_hall:
            J = self.current_density
            di = p.ion_inertial_length
            flux[12] = flux[12] + di / rho * (J[1] * uu[6] - J[2] * uu[5])
            flux[13] = flux[13] + di / rho * (J[2] * uu[4] - J[0] * uu[6])
            flux[14] = flux[14] + di / rho * (J[0] * uu[5] - J[1] * uu[4])
        return flux, expand

    def calc_rhs(self, flux_fourier, expand_fourier):
        """mhdrhs.f90:174-279."""
        p = self.p
        kxi, kyi, kzi = self.kvec()
        kx, ky, kz = 1j * kxi, 1j * kyi, 1j * kzi
        ff, uf = flux_fourier, self.uu_fourier
        fnl = np.empty_like(uf)
        fnl[0] = -(kx * ff[0] + ky * ff[1] + kz * ff[2])
        fnl[1] = -(kx * ff[3] + ky * ff[4] + kz * ff[5])
        fnl[2] = -(kx * ff[6] + ky * ff[7] + kz * ff[8])
        fnl[3] = -(kx * ff[9] + ky * ff[10] + kz * ff[11])
        fnl[4] = kz * ff[13] - ky * ff[14]
        fnl[5] = kx * ff[14] - kz * ff[12]
        fnl[6] = ky * ff[12] - kx * ff[13]
        fnl[7] = -(kx * ff[15] + ky * ff[16] + kz * ff[17])
        if p.if_AEB:
            tau = self.tau_exp
            for v, c in enumerate((2.0, 2.0, 3.0, 3.0, 2.0, 1.0, 1.0)):
                fnl[v] = fnl[v] - c * uf[v] / tau
            fnl[7] = fnl[7] + expand_fourier
        if p.if_visc and p.if_visc_exp:
            for v in (1, 2, 3):
                fnl[v] = fnl[v] - p.viscosity * uf[v] * self.k_square
        if p.if_resis and p.if_resis_exp:
            ksq = np.broadcast_to(self.k_square, uf[0].shape).copy()
            if p.if_conserve_background:
                ksq[0, :, 0] = 0.0  # `cycle` where ix==1 .and. iz==1 (mhdrhs.f90:265-267)
            for v in (4, 5, 6):
                fnl[v] = fnl[v] - p.resistivity * uf[v] * ksq
        self.fnl = fnl
        return fnl

    def rkt_init(self, dt):
        """rktmod.f90:15-32."""
        self.fnl_rk[...] = 0.0
        cc10, cc20, cc30 = 8.0 / 15.0, 5.0 / 12.0, 0.75
        dd20, dd30 = -17.0 / 60.0, -5.0 / 12.0
        self.cc1[:] = (cc10 * dt, cc20 * dt, cc30 * dt)
        self.dd1[:] = (0.0, dd20 * dt, dd30 * dt)
        self.time_step[:] = ((8.0 / 15.0) * dt, (2.0 / 15.0) * dt, (1.0 / 3.0) * dt)

    def rkt(self, irk):
        """rktmod.f90:34-62 (irk is 0-based here)."""
        p = self.p
        self.uu_fourier = self.cc1[irk] * self.fnl + self.dd1[irk] * self.fnl_rk + self.uu_fourier
        self.fnl_rk = self.fnl.copy()
        if p.if_visc and not p.if_visc_exp:
            for v in (1, 2, 3):
                self.uu_fourier[v] = self.uu_fourier[v] / (self.time_step[irk] * self.k_square * p.viscosity + 1.0)
        if p.if_resis and not p.if_resis_exp:
            for v in (4, 5, 6):
                self.uu_fourier[v] = self.uu_fourier[v] / (self.time_step[irk] * self.k_square * p.resistivity + 1.0)

    def dealias(self):
        """dealiasing.f90:70-112."""
        p = self.p
        if p.dealias_option == 1:
            self.uu_fourier[:, dealias_mask(p, self.g)] = 0.0
        elif p.dealias_option == 2:
            fx, fy, fz = self._filters
            self.uu_fourier = self.uu_fourier * fx[None, None, None, :] * fy[None, None, :, None] * fz[None, :, None, None]

    def update_uu_prim_from_uu(self):
        """mhdrhs.f90:282-294."""
        uu, pr = self.uu, self.uu_prim
        pr[0] = uu[1] / uu[0]
        pr[1] = uu[2] / uu[0]
        pr[2] = uu[3] / uu[0]
        pr[3] = (uu[7] - 0.5 * (uu[1] * pr[0] + uu[2] * pr[1] + uu[3] * pr[2]
                                + uu[4] ** 2 + uu[5] ** 2 + uu[6] ** 2)) * (self.p.adiabatic_index - 1)

    def stage(self, irk):
        """One pass of the loop body of evolve (mhd.f90:303-325)."""
        flux, expand = self.calc_flux()
        ff = fft_forward(flux)                               # mhdrhs.f90:128-172
        ef = fft_forward(expand) if expand is not None else None
        self.calc_rhs(ff, ef)
        self.rkt(irk)
        self.dealias()
        self.uu = fft_inverse(self.uu_fourier, self.p.nx)    # fftw.f90:73-103
        self.update_uu_prim_from_uu()

    def evolve(self):
        for irk in range(3):
            self.stage(irk)

    def vardt(self):
        """mhd.f90:328-429 — CFL time step with 2 % hysteresis, then rkt_init."""
        dtmin = cfl_dtmin(self.p, self.g, self.uu, self.uu_prim, self.radius)
        dtmin = dtmin * self.p.cfl
        if self.dt < 0.98 * dtmin or self.dt > 1.02 * dtmin:
            self.dt = dtmin
        self.rkt_init(self.dt)
        return self.dt

    def step(self):
        """mhd.f90:244-248,285: evolve; time+=dt; evolve_radius(time); vardt."""
        self.evolve()
        self.time = self.time + self.dt
        self.evolve_radius(self.time)
        self.vardt()

    # ---------------------------------------------------------------- diagnostics
    def calc_max_divB(self):
        """mhd.f90:522-570."""
        kx, ky, kz = self.kvec()
        uf = self.uu_fourier
        return float(np.max(np.abs(1j * kx * uf[4] + 1j * ky * uf[5] + 1j * kz * uf[6])))

    def calc_rms(self):
        """mhdrms.f90:53-126 — returns (uu_ave[8], uu_rms[8], rho_u2[3])."""
        uu, pr = self.uu, self.uu_prim
        fields = [uu[0], pr[0], pr[1], pr[2], uu[4], uu[5], uu[6], pr[3]]
        n = float(self.p.nx * self.p.ny * self.p.nz)
        ave = np.array([f.sum() for f in fields]) / n
        sq = np.array([(f ** 2).sum() for f in fields]) / n
        rms = sq - ave ** 2
        rho_u2 = np.array([(uu[0] * (pr[i] - ave[1 + i]) ** 2).sum() for i in range(3)]) / n
        return ave, rms, rho_u2

    def invariants(self):
        """Not in the reference: mean total energy density uu(8), mean
        cross helicity u.B, and max |k.B^| — defined here and computed identically on the GPU."""
        uu, pr = self.uu, self.uu_prim
        n = float(self.p.nx * self.p.ny * self.p.nz)
        return np.array([uu[7].sum() / n,
                         (pr[0] * uu[4] + pr[1] * uu[5] + pr[2] * uu[6]).sum() / n,
                         self.calc_max_divB()])


# --------------------------------------------------------------------------------------
# 2D tree (src_compressible/2D/): the same algorithm on an (nx, ny, 1) grid with kz = 0
# --------------------------------------------------------------------------------------
class State2D(State):
    """Restatement of the 2D compressible tree; citations are relative to src_compressible/2D/.
    Arrays keep the 3D shapes with nz = 1 ([v, 0, iy, ix]), exactly like the Fortran arrays
    uu(ix,iy,1,v); the z transform of length 1 is the identity (2D/fftw.f90 has only x and y passes).
    The user routine calc_external_force_real is restated as shipped (the moving Gaussian forcing of B_z)."""

    def __init__(self, p: Params):
        assert p.nz == 1, "the 2D tree has nz = 1"
        assert not (p.if_z_radial and p.if_corotating)          # 2D/mhd.f90:62-67
        super().__init__(p)
        self.k_square = self.g.KX ** 2 + self.g.KY ** 2 + 0.0 * self.g.KZ   # 2D/mhdinit.f90 grid_initialize

    # 2D/AEBmod.f90:66-118
    def update_ksquare(self):
        p, g = self.p, self.g
        kx, ky = g.KX, g.KY
        r0, r = p.radius0, self.radius
        if p.if_corotating:
            c, s = self.cos_cor_ang, self.sin_cor_ang
            self.k_square = (kx ** 2 * (c ** 2 + (s * r0 / r) ** 2)
                             + ky ** 2 * (s ** 2 + (c * r0 / r) ** 2)
                             + kx * ky * 2 * c * s * (1 - (r0 / r) ** 2)) + 0.0 * g.KZ
        elif p.if_z_radial:
            self.k_square = (kx * r0 / r) ** 2 + (ky * r0 / r) ** 2 + 0.0 * g.KZ
        else:
            self.k_square = kx ** 2 + (ky * r0 / r) ** 2 + 0.0 * g.KZ

    # 2D/mhdrhs.f90:272-288 (also :412-428, 2D/mhd.f90:527-543)
    def kvec(self):
        p, g = self.p, self.g
        r0, r = p.radius0, self.radius
        kz = 0.0 * g.KZ
        ky = g.KY * r0 / r
        kx = g.KX + 0.0 * g.KY
        if p.if_AEB and p.if_z_radial:
            kx = kx * r0 / r
        if p.if_AEB and p.if_corotating:
            c, s = self.cos_cor_ang, self.sin_cor_ang
            kx = g.KX * c + g.KY * s
            ky = (-g.KX * s + g.KY * c) * r0 / r
        return kx, ky, kz

    def calc_flux(self):
        """2D/mhdrhs.f90:23-128: as the 3D fluxes; the EBM source differs when the radial direction is z."""
        flux, expand = super().calc_flux()
        p = self.p
        if p.if_AEB and p.if_z_radial:                           # 2D/mhdrhs.f90:96-100
            uu, pr = self.uu, self.uu_prim
            gam, tau = p.adiabatic_index, self.tau_exp
            Bx, By, Bz = uu[4], uu[5], uu[6]
            expand = (-2 * gam / (gam - 1) * pr[3] / tau
                      - (Bx ** 2 + By ** 2 + 2.0 * Bz ** 2) / tau
                      - (2 * uu[1] * pr[0] + 2 * uu[2] * pr[1] + uu[3] * pr[2]) / tau)
        return flux, expand

    def calc_external_force_real(self):
        """2D/mhdrhs.f90:480-531 as shipped: dBz/dt forcing, Gaussian in x around Lx/2, Gaussian in y around a
        centre that moves at 0.3 along y (with its two periodic images); `time` is the module variable, i.e. the
        time at the start of the step for all three stages (it advances after evolve, 2D/mhd.f90:232)."""
        p, g = self.p, self.g
        dBdt = 0.2
        xc = 0.5 * p.Lx
        x_width = 0.05 * p.Ly
        yc = math.fmod(0.2 * p.Ly + 0.3 * self.time, p.Ly)
        y_width = 0.05 * p.Ly
        func_x = np.exp(-((g.xgrid - xc) / x_width) ** 2)[None, None, :]
        y = g.ygrid[None, :, None]
        f = dBdt * func_x * np.exp(-((y - yc) / y_width) ** 2)
        f = f + dBdt * func_x * np.exp(-((y - (yc + p.Ly)) / y_width) ** 2)
        f = f + dBdt * func_x * np.exp(-((y - (yc - p.Ly)) / y_width) ** 2)
        return f

    def stage(self, irk):
        """2D/mhd.f90 evolve: as the 3D loop body; calc_flux ends with calc_external_force_real (2D/mhdrhs.f90:129-131),
        transform_flux_real_to_fourier transforms it (:216-251) and calc_rhs adds it to fnl(7) (:370-372)."""
        if not self.p.if_external_force:
            return super().stage(irk)
        flux, expand = self.calc_flux()
        self.external_force = self.calc_external_force_real()
        ff = fft_forward(flux)
        ef = fft_forward(expand) if expand is not None else None
        xf = fft_forward(self.external_force)
        self.calc_rhs(ff, ef)
        self.fnl[6] = self.fnl[6] + xf
        self.rkt(irk)
        self.dealias()
        self.uu = fft_inverse(self.uu_fourier, self.p.nx)
        self.update_uu_prim_from_uu()

    def calc_rhs(self, flux_fourier, expand_fourier):
        """2D/mhdrhs.f90:255-392."""
        p = self.p
        kxi, kyi, kzi = self.kvec()
        kx, ky, kz = 1j * kxi, 1j * kyi, 1j * kzi
        ff, uf = flux_fourier, self.uu_fourier
        fnl = np.empty_like(uf)
        fnl[0] = -(kx * ff[0] + ky * ff[1] + kz * ff[2])
        fnl[1] = -(kx * ff[3] + ky * ff[4] + kz * ff[5])
        fnl[2] = -(kx * ff[6] + ky * ff[7] + kz * ff[8])
        fnl[3] = -(kx * ff[9] + ky * ff[10] + kz * ff[11])
        fnl[4] = kz * ff[13] - ky * ff[14]
        fnl[5] = kx * ff[14] - kz * ff[12]
        fnl[6] = ky * ff[12] - kx * ff[13]
        fnl[7] = -(kx * ff[15] + ky * ff[16] + kz * ff[17])
        if p.if_AEB:
            tau = self.tau_exp
            coef = (2.0, 3.0, 3.0, 2.0, 1.0, 1.0, 2.0) if p.if_z_radial else (2.0, 2.0, 3.0, 3.0, 2.0, 1.0, 1.0)
            for v, c in enumerate(coef):
                fnl[v] = fnl[v] - c * uf[v] / tau
            fnl[7] = fnl[7] + expand_fourier
        if p.if_visc and p.if_visc_exp:
            for v in (1, 2, 3):
                fnl[v] = fnl[v] - p.viscosity * uf[v] * self.k_square
        if p.if_resis and p.if_resis_exp:
            ksq = np.broadcast_to(self.k_square, uf[0].shape).copy()
            if p.if_conserve_background:
                ksq[:, :, 0] = 0.0  # `cycle` where ix==1 (2D/mhdrhs.f90:372-374)
            for v in (4, 5, 6):
                fnl[v] = fnl[v] - p.resistivity * uf[v] * ksq
        self.fnl = fnl
        return fnl

    def dealias(self):
        """2D/dealiasing.f90:62-119 (option 3 = square truncation)."""
        p, g = self.p, self.g
        if p.dealias_option == 1:
            tx = (g.wnx[: g.nxh] * p.Lx / (2 * PI * p.nx)) ** 2
            ty = (g.wny * p.Ly / (2 * PI * p.ny)) ** 2
            mask = np.sqrt(tx[None, None, :] + ty[None, :, None]) > (1.0 / 3.0)
            self.uu_fourier[:, mask] = 0.0
        elif p.dealias_option == 2:
            fx, fy, _ = self._filters
            self.uu_fourier = self.uu_fourier * fx[None, None, None, :] * fy[None, None, :, None]
        elif p.dealias_option == 3:
            rx = np.abs(g.wnx[: g.nxh] * p.Lx / (2 * PI * p.nx))
            ry = np.abs(g.wny * p.Ly / (2 * PI * p.ny))
            mask = (rx[None, None, :] > (1.0 / 3.0)) | (ry[None, :, None] > (1.0 / 3.0))
            self.uu_fourier[:, mask] = 0.0

    def vardt(self):
        """2D/mhd.f90:296-406."""
        p, g = self.p, self.g
        uu, pr = self.uu, self.uu_prim
        rho = uu[0]
        csound2 = p.adiabatic_index * pr[3] / rho
        sq = np.sqrt(rho)
        ca = [uu[4] / sq, uu[5] / sq, uu[6] / sq]
        cms2 = csound2 + (ca[0] ** 2 + ca[1] ** 2 + ca[2] ** 2)
        s2 = math.sqrt(2.0)
        cmax = []
        for d in (0, 1):
            cns2 = np.sqrt(np.maximum(cms2 ** 2 - 4 * csound2 * ca[d] ** 2, 0.0))
            cfast = np.sqrt(cms2 + cns2) / s2
            cslow = np.sqrt(np.maximum(cms2 - cns2, 0.0)) / s2
            u = pr[d]
            c = np.abs(u + cfast)
            for t in (np.abs(u + cslow), np.abs(u + ca[d]), np.abs(u - cfast), np.abs(u - cslow),
                      np.abs(u - ca[d]), np.abs(u)):
                c = np.maximum(c, t)
            cmax.append(c)
        if p.if_resis and p.if_resis_exp:                        # 2D/mhd.f90:361-364
            cmax[0] = np.maximum(cmax[0], p.resistivity / g.dx)
            cmax[1] = np.maximum(cmax[1], p.resistivity / g.dy)
        if p.if_hall:                                            # 2D/mhd.f90:366-374
            ch = p.ion_inertial_length / rho * np.maximum(np.maximum(uu[4], uu[5]), uu[6]) / min(g.dx, g.dy)
            cmax[0] = np.maximum(cmax[0], ch)
            cmax[1] = np.maximum(cmax[1], ch)
        dtx = g.dx / cmax[0]
        if p.if_AEB and p.if_z_radial:
            dtx = dtx * (self.radius / p.radius0)
        dty = g.dy / cmax[1] * (self.radius / p.radius0)
        dtmin = float(np.minimum(dtx, dty).min()) * p.cfl
        if p.if_limit_dt_increase:                               # 2D/mhd.f90:396-404
            if self.dt == 0.0 or self.dt > 1.02 * dtmin:
                self.dt = dtmin
        elif self.dt < 0.98 * dtmin or self.dt > 1.02 * dtmin:
            self.dt = dtmin
        self.rkt_init(self.dt)
        return self.dt

    def step(self, calc_dt: bool = True):
        """2D/mhd.f90:209-240: evolve; time+=dt; evolve_radius; vardt (the driver calls vardt only every
        dstep_calcdt = 20 steps: pass calc_dt=False for the steps in between)."""
        self.evolve()
        self.time = self.time + self.dt
        self.evolve_radius(self.time)
        if calc_dt:
            self.vardt()


# --------------------------------------------------------------------------------------
# incompressible tree (src_incompressible/): pressure projection instead of an energy equation
# --------------------------------------------------------------------------------------
class StateIncompressible(State):
    """Restatement of the 3D incompressible tree; citations are relative to src_incompressible/.
    uu = [rho, rho*u (3), B (3), p] (uu(8) is the PRESSURE here, mhdinit.f90:210), uu_prim = u (3
    components, mhdinit.f90:5,142; the 4th row of the base-class array is unused).  ``rho0`` is the
    namelist background density (mhdinit.f90:15) that calc_gradient_velocity_real divides by
    (mhdrhs.f90:369); update_rho_p compounds it every step in the expanding box (AEBmod.f90:123-134).
    Every stage re-derives the spectrum from the real fields (mhd.f90:325)."""

    def __init__(self, p: Params, p0: float = 1.0):
        super().__init__(p)
        self.rho0 = float(p.rho0)
        self.p0 = float(p0)
        self.current_density = np.zeros((3, p.nz, p.ny, p.nx))
        self.grad_velocity = np.zeros((9, p.nz, p.ny, p.nx))

    def set_primitive(self, prim: np.ndarray):
        """initial_calc_conserve_variable (mhdinit.f90:1031-1042) + transform_uu_real_to_fourier."""
        uu = np.array(prim, dtype=np.float64, copy=True)
        self.uu_prim[0:3] = uu[1:4]
        uu[1] = uu[0] * self.uu_prim[0]
        uu[2] = uu[0] * self.uu_prim[1]
        uu[3] = uu[0] * self.uu_prim[2]
        self.uu = uu
        self.uu_fourier = fft_forward(self.uu)

    def calc_gradient_velocity_real(self):
        """mhdrhs.f90:312-391 — d u_b / d x_a = IFFT( k_a (rho u_b)^ / rho0 ), slot 3*b + a."""
        kx, ky, kz = self.kvec()
        uf = self.uu_fourier
        gf = np.empty((9,) + uf.shape[1:], dtype=np.complex128)
        for b in range(3):
            gf[3 * b + 0] = 1j * kx * uf[1 + b]
            gf[3 * b + 1] = 1j * ky * uf[1 + b]
            gf[3 * b + 2] = 1j * kz * uf[1 + b]
        gf = gf / self.rho0                                     # mhdrhs.f90:369
        self.grad_velocity = fft_inverse(gf, self.p.nx)

    def calc_flux_for_pressure(self):
        """mhdrhs.f90:393-437 — -(rho u . grad) u + J x B."""
        uu, G, J = self.uu, self.grad_velocity, self.current_density
        fp = np.empty((3,) + uu.shape[1:])
        fp[0] = (-uu[1] * G[0] - uu[2] * G[1] - uu[3] * G[2] + J[1] * uu[6] - J[2] * uu[5])
        fp[1] = (-uu[1] * G[3] - uu[2] * G[4] - uu[3] * G[5] + J[2] * uu[4] - J[0] * uu[6])
        fp[2] = (-uu[1] * G[6] - uu[2] * G[7] - uu[3] * G[8] + J[0] * uu[5] - J[1] * uu[4])
        return fp

    def calc_pressure_fourier(self, fpf):
        """mhdrhs.f90:468-518 — p^ = -(k . Fp^)/k^2, zero where k^2 < 1e-10."""
        kxi, kyi, kzi = self.kvec()
        kx, ky, kz = 1j * kxi, 1j * kyi, 1j * kzi
        k2 = np.broadcast_to(self.k_square, fpf[0].shape)
        with np.errstate(divide="ignore", invalid="ignore"):
            ph = -(kx * fpf[0] + ky * fpf[1] + kz * fpf[2]) / k2
        self.uu_fourier[7] = np.where(k2 < 1e-10, 0.0, ph)

    def calc_flux(self):
        """mhdrhs.f90:25-84 — the electric field -u x B (+ Hall term)."""
        p = self.p
        uu, pr = self.uu, self.uu_prim
        ux, uy, uz = pr[0], pr[1], pr[2]
        Bx, By, Bz = uu[4], uu[5], uu[6]
        flux = np.empty((3,) + uu.shape[1:])
        flux[0] = uz * By - uy * Bz
        flux[1] = ux * Bz - uz * Bx
        flux[2] = uy * Bx - ux * By
        if p.if_hall:
            J, di, rho = self.current_density, p.ion_inertial_length, uu[0]
            flux[0] = flux[0] + di / rho * (J[1] * uu[6] - J[2] * uu[5])
            flux[1] = flux[1] + di / rho * (J[2] * uu[4] - J[0] * uu[6])
            flux[2] = flux[2] + di / rho * (J[0] * uu[5] - J[1] * uu[4])
        return flux

    def calc_rhs(self, flux_fourier, fpf):
        """mhdrhs.f90:117-232."""
        p = self.p
        kxi, kyi, kzi = self.kvec()
        kx, ky, kz = 1j * kxi, 1j * kyi, 1j * kzi
        ff, uf = flux_fourier, self.uu_fourier
        k2 = np.broadcast_to(self.k_square, uf[0].shape)
        bg = k2 < 1e-10
        fnl = np.zeros_like(uf)
        with np.errstate(divide="ignore", invalid="ignore"):
            kdotfp = (kx * fpf[0] + ky * fpf[1] + kz * fpf[2]) / k2
        kdotfp = np.where(bg, 0.0, kdotfp)
        fnl[1] = np.where(bg, 0.0, fpf[0] + kdotfp * kx)
        fnl[2] = np.where(bg, 0.0, fpf[1] + kdotfp * ky)
        fnl[3] = np.where(bg, 0.0, fpf[2] + kdotfp * kz)
        fnl[4] = kz * ff[1] - ky * ff[2]
        fnl[5] = kx * ff[2] - kz * ff[0]
        fnl[6] = ky * ff[0] - kx * ff[1]
        if p.if_AEB:
            tau = self.tau_exp
            for v, c in enumerate((2.0, 2.0, 3.0, 3.0, 2.0, 1.0, 1.0)):
                fnl[v] = fnl[v] - c * uf[v] / tau
            fnl[7] = fnl[7] - 2.0 * p.adiabatic_index * uf[7] / tau
        if p.if_visc and p.if_visc_exp:
            for v in (1, 2, 3):
                fnl[v] = fnl[v] - p.viscosity * uf[v] * self.k_square
        if p.if_resis and p.if_resis_exp:
            ksq = k2.copy()
            if p.if_conserve_background:
                ksq[0, :, 0] = 0.0                              # `cycle` where ix==1 .and. iz==1 (mhdrhs.f90:219-221)
            for v in (4, 5, 6):
                fnl[v] = fnl[v] - p.resistivity * uf[v] * ksq
        self.fnl = fnl
        return fnl

    def update_uu_prim_from_uu(self):
        """mhdrhs.f90:235-242."""
        uu, pr = self.uu, self.uu_prim
        pr[0] = uu[1] / uu[0]
        pr[1] = uu[2] / uu[0]
        pr[2] = uu[3] / uu[0]

    def stage(self, irk, retransform: bool = True):
        """One pass of the loop body of evolve (mhd.f90:323-364)."""
        if retransform:
            self.uu_fourier = fft_forward(self.uu)               # mhd.f90:325
        self.calc_current_density_real()                         # :328 (always, J x B needs it)
        self.calc_gradient_velocity_real()                       # :329
        fpf = fft_forward(self.calc_flux_for_pressure())         # :332-335
        self.calc_pressure_fourier(fpf)                          # :338
        ff = fft_forward(self.calc_flux())                       # :341-344
        self.calc_rhs(ff, fpf)
        self.rkt(irk)
        self.dealias()
        self.uu = fft_inverse(self.uu_fourier, self.p.nx)
        self.update_uu_prim_from_uu()

    def update_rho_p(self):
        """AEBmod.f90:123-134 — compounds rho0 and p0 with the CURRENT radius on every call."""
        q = self.p.radius0 / self.radius
        self.rho0 = self.rho0 * q ** 2
        self.p0 = self.p0 * q ** (2 * self.p.adiabatic_index)

    def evolve(self, retransform: bool = True):
        for irk in range(3):
            self.stage(irk, retransform)
        self.update_rho_p()                                      # mhd.f90:366

    def vardt(self):
        """mhd.f90:369-476 — Alfven and flow speeds only."""
        p, g = self.p, self.g
        uu, pr = self.uu, self.uu_prim
        sq = np.sqrt(uu[0])
        dmin = min(g.dx, g.dy, g.dz)
        cm = []
        for d in range(3):
            ca = uu[4 + d] / sq
            u = pr[d]
            cm.append(np.maximum(np.maximum(np.abs(u + ca), np.abs(u - ca)), np.abs(u)))
        if p.if_hall:
            ch = p.ion_inertial_length / uu[0] * np.maximum(np.maximum(uu[4], uu[5]), uu[6]) / dmin
            cm = [np.maximum(c, ch) for c in cm]
        rr = self.radius / p.radius0
        with np.errstate(divide="ignore"):
            dtx = g.dx / cm[0]
            dty = g.dy / cm[1] * rr
            dtz = g.dz / cm[2] * rr
        dtmin = float(np.minimum(np.minimum(dtx, dty), dtz).min()) * p.cfl
        if self.dt < 0.98 * dtmin or self.dt > 1.02 * dtmin:
            self.dt = dtmin
        self.rkt_init(self.dt)
        return self.dt

    # ---------------------------------------------------------------- diagnostics
    def calc_max_divV(self):
        """mhd.f90:620-668 (Fourier-space maximum of |k . (rho u)^| / rho0)."""
        kx, ky, kz = self.kvec()
        uf = self.uu_fourier
        return float(np.max(np.abs(1j * kx * uf[1] + 1j * ky * uf[2] + 1j * kz * uf[3]))) / self.rho0

    def calc_max_div_real(self):
        """calc_divB_real/calc_divV_real + calc_max_div*_real (mhdrhs.f90:532-648, mhd.f90:672-732):
        maxima of |div B| and |div u| in real space, as the driver prints them."""
        kx, ky, kz = self.kvec()
        uf = self.uu_fourier
        db = fft_inverse(1j * kx * uf[4] + 1j * ky * uf[5] + 1j * kz * uf[6], self.p.nx)
        dv = fft_inverse((1j * kx * uf[1] + 1j * ky * uf[2] + 1j * kz * uf[3]) / self.rho0, self.p.nx)
        return float(np.abs(db).max()), float(np.abs(dv).max())

    def calc_rms(self):
        """mhdrms.f90:47-120.  The reference reads uu_prim(:,:,:,4), which does not exist in this tree
        (uu_prim has 3 components): entry 8 is taken from the pressure uu(8) here."""
        uu, pr = self.uu, self.uu_prim
        fields = [uu[0], pr[0], pr[1], pr[2], uu[4], uu[5], uu[6], uu[7]]
        n = float(self.p.nx * self.p.ny * self.p.nz)
        ave = np.array([f.sum() for f in fields]) / n
        sq = np.array([(f ** 2).sum() for f in fields]) / n
        rho_u2 = np.array([(uu[0] * (pr[i] - ave[1 + i]) ** 2).sum() for i in range(3)]) / n
        return ave, sq - ave ** 2, rho_u2

    def invariants(self):
        uu, pr = self.uu, self.uu_prim
        n = float(self.p.nx * self.p.ny * self.p.nz)
        e = 0.5 * (uu[1] * pr[0] + uu[2] * pr[1] + uu[3] * pr[2] + uu[4] ** 2 + uu[5] ** 2 + uu[6] ** 2)
        return np.array([e.sum() / n, (pr[0] * uu[4] + pr[1] * uu[5] + pr[2] * uu[6]).sum() / n, self.calc_max_divB()])


class StateIncompressible2D(StateIncompressible):
    """src_incompressible/2D/: the incompressible tree on an (nx, ny, 1) grid with kz = 0 (2D/mhdrhs.f90:189,
    312,396,583).  All nine velocity gradients and three currents are still formed (the z derivatives are
    transforms of zeros).  Wave vectors, k_square and dealiasing (incl. option 3) as in the compressible 2D
    tree — the files differ only in names; vardt has its own two-direction form (2D/mhd.f90:380-478)."""

    def __init__(self, p: Params, p0: float = 1.0):
        assert p.nz == 1 and not p.if_z_radial
        super().__init__(p, p0)
        self.k_square = self.g.KX ** 2 + self.g.KY ** 2 + 0.0 * self.g.KZ

    kvec = State2D.kvec
    update_ksquare = State2D.update_ksquare
    dealias = State2D.dealias

    def calc_rhs(self, flux_fourier, fpf):
        """2D/mhdrhs.f90:170-290: as the 3D routine; if_conserve_background skips every mode with ix == 1 (:268)."""
        p = self.p
        if not (p.if_resis and p.if_resis_exp and p.if_conserve_background):
            return super().calc_rhs(flux_fourier, fpf)
        q = dataclasses.replace(p, if_conserve_background=False)
        self.p = q
        try:
            fnl = super().calc_rhs(flux_fourier, fpf)
        finally:
            self.p = p
        ksq = np.broadcast_to(self.k_square, fnl[0].shape)
        for v in (4, 5, 6):   # undo the resistive term where ix == 1
            fnl[v][:, :, 0] = fnl[v][:, :, 0] + p.resistivity * self.uu_fourier[v][:, :, 0] * ksq[:, :, 0]
        self.fnl = fnl
        return fnl

    def vardt(self):
        """2D/mhd.f90:380-478."""
        p, g = self.p, self.g
        uu, pr = self.uu, self.uu_prim
        sq = np.sqrt(uu[0])
        cm = []
        for d in range(2):
            ca = uu[4 + d] / sq
            u = pr[d]
            cm.append(np.maximum(np.maximum(np.abs(u + ca), np.abs(u - ca)), np.abs(u)))
        if p.if_resis and p.if_resis_exp:
            cm[0] = np.maximum(cm[0], p.resistivity / g.dx)
            cm[1] = np.maximum(cm[1], p.resistivity / g.dy)
        if p.if_hall:
            ch = p.ion_inertial_length / uu[0] * np.maximum(np.maximum(uu[4], uu[5]), uu[6]) / min(g.dx, g.dy)
            cm = [np.maximum(c, ch) for c in cm]
        with np.errstate(divide="ignore"):
            dtx = g.dx / cm[0]
            dty = g.dy / cm[1] * (self.radius / p.radius0)
        dtmin = float(np.minimum(dtx, dty).min()) * p.cfl
        if p.if_limit_dt_increase:
            if self.dt == 0.0 or self.dt > 1.02 * dtmin:
                self.dt = dtmin
        elif self.dt < 0.98 * dtmin or self.dt > 1.02 * dtmin:
            self.dt = dtmin
        self.rkt_init(self.dt)
        return self.dt

    def step(self, calc_dt: bool = True):
        """2D/mhd.f90:259-290: evolve; time+=dt; evolve_radius; vardt only every dstep_calcdt = 20 steps."""
        self.evolve()
        self.time = self.time + self.dt
        self.evolve_radius(self.time)
        if calc_dt:
            self.vardt()


# --------------------------------------------------------------------------------------
# dealiasing tables (dealiasing.f90)
# --------------------------------------------------------------------------------------
def dealias_axis_terms(p: Params, g: Grid):
    """The three squared 1-D terms of the radius test, dealiasing.f90:91-93, evaluated in the
    reference's FP64 operation order: (k*L/(2*pi*n))**2."""
    tx = (g.wnx[: g.nxh] * p.Lx / (2 * PI * p.nx)) ** 2
    ty = (g.wny * p.Ly / (2 * PI * p.ny)) ** 2
    tz = (g.wnz * p.Lz / (2 * PI * p.nz)) ** 2
    return tx, ty, tz


def dealias_mask(p: Params, g: Grid) -> np.ndarray:
    """True where option 1 zeroes the mode: sqrt(tx+ty+tz) > 1./3. (dealiasing.f90:87-99)."""
    tx, ty, tz = dealias_axis_terms(p, g)
    radius = np.sqrt((tx[None, None, :] + ty[None, :, None]) + tz[:, None, None])
    return radius > (1.0 / 3.0)


def _filter_1d(k, L, n, af):
    aj = (5.0 + 6.0 * af) / 8.0
    bj = (1.0 + 2.0 * af) / 2.0
    cj = -(1.0 - 2 * af) / 8.0
    w = k * L / n
    return (aj + bj * np.cos(w) + cj * np.cos(2 * w)) / (1 + 2 * af * np.cos(w))


def dealias_filters(p: Params, g: Grid):
    """dealiasing.f90:31-64 — separable compact-filter transfer functions."""
    return (_filter_1d(g.wnx[: g.nxh], p.Lx, p.nx, p.afx),
            _filter_1d(g.wny, p.Ly, p.ny, p.afy),
            _filter_1d(g.wnz, p.Lz, p.nz, p.afz))


# --------------------------------------------------------------------------------------
# CFL (mhd.f90:328-429)
# --------------------------------------------------------------------------------------
def cfl_dt_pointwise(p: Params, g: Grid, uu, uu_prim, radius):
    rho = uu[0]
    csound2 = p.adiabatic_index * uu_prim[3] / rho
    sq = np.sqrt(rho)
    cax, cay, caz = uu[4] / sq, uu[5] / sq, uu[6] / sq
    calfven2 = cax ** 2 + cay ** 2 + caz ** 2
    cms2 = csound2 + calfven2
    out = []
    s2 = math.sqrt(2.0)
    cmaxhall = None
    if p.if_hall:
        cmaxhall = (p.ion_inertial_length / rho * np.maximum(np.maximum(uu[4], uu[5]), uu[6])
                    / min(g.dx, g.dy, g.dz))
    for ca, u in ((cax, uu_prim[0]), (cay, uu_prim[1]), (caz, uu_prim[2])):
        cns2 = np.sqrt(np.maximum(cms2 ** 2 - 4 * csound2 * ca ** 2, 0.0))
        cfast = np.sqrt(cms2 + cns2) / s2
        cslow = np.sqrt(np.maximum(cms2 - cns2, 0.0)) / s2
        cmax = np.abs(u + cfast)
        for c in (np.abs(u + cslow), np.abs(u + ca), np.abs(u - cfast),
                  np.abs(u - cslow), np.abs(u - ca), np.abs(u)):
            cmax = np.maximum(cmax, c)
        if cmaxhall is not None:
            cmax = np.maximum(cmax, cmaxhall)
        out.append(cmax)
    dtx = g.dx / out[0]
    dty = g.dy / out[1] * (radius / p.radius0)
    dtz = g.dz / out[2] * (radius / p.radius0)
    return np.minimum(np.minimum(dtx, dty), dtz)


def cfl_dtmin(p, g, uu, uu_prim, radius) -> float:
    return float(cfl_dt_pointwise(p, g, uu, uu_prim, radius).min())


# --------------------------------------------------------------------------------------
# initial conditions (the subset used by BASELINE configs; mhdinit.f90)
# --------------------------------------------------------------------------------------
def ic_uniform_background(p: Params, bx0=0.0, by0=0.0, bz0=0.0, press0=1.0, rho0=1.0):
    """ifield=3 (mhdinit.f90:193-194,251-256): returns primitive [rho,u,B,p]."""
    prim = np.zeros((8, p.nz, p.ny, p.nx))
    prim[0] = rho0
    prim[4], prim[5], prim[6] = bx0, by0, bz0
    prim[7] = press0
    return prim


def ic_alfven_wave(p: Params, prim, db0=0.1, wave_number_jet=1, cor_angle=None):
    """ipert=1 (mhdinit.f90:328-342): circularly polarised Alfven wave along x."""
    g = Grid(p)
    ang = (p.corotating_angle if p.if_corotating else 0.0) if cor_angle is None else cor_angle
    ca, sa = math.cos(ang), math.sin(ang)
    kx = 2 * PI / p.Lx * wave_number_jet
    s = np.sin(kx * g.xgrid)[None, None, :]
    c = np.cos(kx * g.xgrid)[None, None, :]
    rs = np.sqrt(prim[0])
    prim[6] = prim[6] - db0 * s
    prim[3] = prim[3] + db0 / rs * s
    prim[1] = prim[1] + db0 / rs * c * sa
    prim[4] = prim[4] - db0 * c * sa
    prim[2] = prim[2] + db0 / rs * c * ca
    prim[5] = prim[5] - db0 * c * ca
    return prim


def ic_turbulence(p: Params, prim, bx0, by0, bz0, db0=0.1, dv0=0.1, drho0=0.01,
                  nmodex=8, nmodey=8, nmodez=8, seeds=(101, 116, 132)):
    """ipert=7 (mhdinit.f90:695-823): isotropic random-phase modes, amplitudes ~ k^-3/2,
    polarised along k x B0.  The reference's phases come from a compiler-specific generator
    (random_seed(PUT=ir+100), :705-727); here three seeded NumPy generators stand in
    (SURVEY 8(c)).  Everything else follows the Fortran, mode table included."""
    g = Grid(p)
    correlation_vb = -0.05
    nmode = (2 * nmodex + 1) * (2 * nmodey + 1) * (2 * nmodez + 1)
    phs = np.random.default_rng(seeds[0]).random(nmode) * 2 * PI
    phs1 = np.random.default_rng(seeds[1]).random(nmode) * 2 * PI
    phs2 = np.random.default_rng(seeds[2]).random(nmode) * 2 * PI
    X = g.xgrid[None, None, :]
    Y = g.ygrid[None, :, None]
    Z = g.zgrid[:, None, None]
    B0mod = math.sqrt(bx0 ** 2 + by0 ** 2 + bz0 ** 2)
    cu = math.sqrt(1 - correlation_vb ** 2)
    for ikx in range(0, nmodex + 1):
        kx = ikx * 2 * PI / p.Lx
        for iky in range(-nmodey, nmodey + 1):
            ky = iky * 2 * PI / p.Ly
            for ikz in range(-nmodez, nmodez + 1):
                kz = ikz * 2 * PI / p.Lz
                if (ikx == 0 and iky < 0) or (ikx == 0 and iky == 0 and ikz <= 0):
                    continue
                k_radius = math.sqrt(float(ikx ** 2 + iky ** 2 + ikz ** 2))
                if k_radius > max(nmodex, nmodey, nmodez):
                    continue
                kmod = math.sqrt(kx ** 2 + ky ** 2 + kz ** 2)
                d = ((ky * bz0 - kz * by0) / kmod / B0mod,
                     (kz * bx0 - kx * bz0) / kmod / B0mod,
                     (kx * by0 - ky * bx0) / kmod / B0mod)
                idx = ((ikx + nmodex) * (2 * nmodey + 1) + iky + nmodey) * (2 * nmodez + 1) + ikz + nmodez
                ph0, ph1, ph2 = phs[idx], phs1[idx], phs2[idx]
                base = (kx * X + ky * Y) + kz * Z
                amp = 1.0 / math.sqrt(k_radius ** 3)
                c0 = np.cos(base + ph0)
                c1 = np.cos(base + ph1)
                c2 = np.cos(base + ph2)
                for i in range(3):
                    prim[1 + i] += cu * dv0 * c0 * amp * d[i]
                    prim[4 + i] += db0 * c1 * amp * d[i]
                    prim[1 + i] += correlation_vb * dv0 * c1 * amp * d[i]
                prim[0] += drho0 * c2 * amp
    return prim


# --------------------------------------------------------------------------------------
# output / restart file format (mhdoutput.f90:72-131, restart.f90:17-63)
# --------------------------------------------------------------------------------------
def write_out_file(path, time: float, uu_prim_layout: np.ndarray):
    """``uu_prim_layout`` = [8, nz, ny, nx] (rho,u,B,p when output_primitive, mhdoutput.f90:95-103).
    Header is the 12-byte Fortran record [int32 4][float32 time][int32 4] (:88-92); data at byte 12."""
    with open(path, "wb") as f:
        np.array([4], dtype="<i4").tofile(f)
        np.array([time], dtype="<f4").tofile(f)
        np.array([4], dtype="<i4").tofile(f)
        np.ascontiguousarray(uu_prim_layout, dtype="<f8").tofile(f)


def read_out_file(path, nx, ny, nz, nvar=8):
    with open(path, "rb") as f:
        head = np.fromfile(f, dtype="<i4", count=1)
        t = np.fromfile(f, dtype="<f4", count=1)
        tail = np.fromfile(f, dtype="<i4", count=1)
        assert head[0] == 4 and tail[0] == 4
        data = np.fromfile(f, dtype="<f8", count=nvar * nz * ny * nx).reshape(nvar, nz, ny, nx)
    return float(t[0]), data


def primitive_of(state: State) -> np.ndarray:
    """The array output_uu writes when output_primitive=.true. (mhdoutput.f90:95-103)."""
    out = state.uu.copy()
    out[1:4] = state.uu_prim[0:3]
    out[7] = state.uu_prim[3]
    return out
