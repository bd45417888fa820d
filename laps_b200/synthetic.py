"""Synthetic initial data for the stand-in driver (bench.py, tests): what the Fortran driver's
``background_fields_initialize`` / ``perturbation_initialize`` hooks would hand to
``laps_set_primitive``.  Host-side NumPy only; nothing here is on the hot path.

``turbulence_slab`` evaluates the same field as the reference's ``ipert=7`` mode sum
(mhdinit.f90:695-823; ``ifield=3`` background, :251-256): isotropic random-phase modes with
|k_int| <= kmax, amplitudes k^-3/2, polarised along k x B0, velocity/magnetic correlation -0.05.
The reference sums cosines point by point (O(modes x N^3)); here the modes are placed in a sparse
spectrum and synthesised with inverse FFTs, one z-slab at a time, which is the same function of
(x,y,z) up to round-off.  The phases come from seeded NumPy generators because the reference's
``random_seed`` usage is compiler specific.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.fft as sfft

PI = 3.141592653589793  # mhdinit.f90:7


def mode_table(Lx, Ly, Lz, bx0, by0, bz0, nmodex, nmodey, nmodez, seeds, db0, dv0, drho0):
    """Integer wave vectors and the complex amplitude of each of the 7 perturbed fields
    (rho, ux, uy, uz, bx, by, bz) per mode: field = sum_m Re(c[v,m] exp(i k_m.x))."""
    correlation_vb = -0.05
    nmode = (2 * nmodex + 1) * (2 * nmodey + 1) * (2 * nmodez + 1)
    phs = np.random.default_rng(seeds[0]).random(nmode) * 2 * PI
    phs1 = np.random.default_rng(seeds[1]).random(nmode) * 2 * PI
    phs2 = np.random.default_rng(seeds[2]).random(nmode) * 2 * PI
    B0mod = math.sqrt(bx0 ** 2 + by0 ** 2 + bz0 ** 2)
    cu = math.sqrt(1 - correlation_vb ** 2)
    ks, coefs = [], []
    for ikx in range(0, nmodex + 1):
        kx = ikx * 2 * PI / Lx
        for iky in range(-nmodey, nmodey + 1):
            ky = iky * 2 * PI / Ly
            for ikz in range(-nmodez, nmodez + 1):
                kz = ikz * 2 * PI / Lz
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
                e0, e1, e2 = np.exp(1j * phs[idx]), np.exp(1j * phs1[idx]), np.exp(1j * phs2[idx])
                amp = 1.0 / math.sqrt(k_radius ** 3)
                c = np.zeros(7, dtype=np.complex128)
                c[0] = drho0 * amp * e2
                for i in range(3):
                    c[1 + i] = (cu * dv0 * e0 + correlation_vb * dv0 * e1) * amp * d[i]
                    c[4 + i] = db0 * amp * e1 * d[i]
                ks.append((ikx, iky, ikz))
                coefs.append(c)
    return np.array(ks, dtype=np.int64).reshape(-1, 3), np.array(coefs).reshape(-1, 7).T.copy()


def turbulence_slab(nx, ny, nz, Lx, Ly, Lz, z_offset=0, z_size=None, bx0=1.0, by0=0.0, bz0=0.0,
                    press0=1.0, rho0=1.0, db0=0.1, dv0=0.1, drho0=0.01, kmax=8,
                    seeds=(101, 116, 132), out=None, workers=-1):
    """Primitive (rho,ux,uy,uz,bx,by,bz,p)[8, z_size, ny, nx] of the z-slab [z_offset, z_offset+z_size)."""
    z_size = nz - z_offset if z_size is None else z_size
    nxh = nx // 2 + 1
    assert kmax < nx // 2 and kmax < ny // 2 and kmax < nz // 2
    ks, coefs = mode_table(Lx, Ly, Lz, bx0, by0, bz0, kmax, kmax, kmax, seeds, db0, dv0, drho0)
    prim = out if out is not None else np.empty((8, z_size, ny, nx))
    assert prim.shape == (8, z_size, ny, nx)
    z = (np.arange(z_offset, z_offset + z_size) * (Lz / nz))
    ez = np.exp(1j * np.outer(2 * PI / Lz * ks[:, 2], z))        # [mode, z]
    # c2r doubles every kx>0 bin (adds the conjugate) and keeps Re of the kx=0 bin
    w = np.where(ks[:, 0] > 0, 0.5, 1.0)
    back = (rho0, 0.0, 0.0, 0.0, bx0, by0, bz0)
    ncol = (2 * kmax + 1)
    for v in range(7):
        # sparse spectrum after the z synthesis: [z, iky (compact), ikx (compact)]
        g = np.zeros((z_size, ncol, kmax + 1), dtype=np.complex128)
        np.add.at(g, (slice(None), ks[:, 1] + kmax, ks[:, 0]), ((coefs[v] * w)[:, None] * ez).T)
        full = np.zeros((z_size, ny, nxh), dtype=np.complex128)
        full[:, :kmax + 1, :kmax + 1] = g[:, kmax:, :]
        full[:, ny - kmax:, :kmax + 1] = g[:, :kmax, :]
        a = sfft.ifft(full, axis=1, workers=workers, norm="forward")
        prim[v] = sfft.irfft(a, n=nx, axis=2, workers=workers, norm="forward")
        prim[v] += back[v]
    prim[7] = press0
    return prim


def uniform_background(nx, ny, z_size, bx0=0.0, by0=0.0, bz0=0.0, press0=1.0, rho0=1.0, out=None):
    """``ifield = 3`` (mhdinit.f90:193-194,251-256): uniform rho, B and p; primitive [8, z_size, ny, nx]."""
    prim = out if out is not None else np.empty((8, z_size, ny, nx))
    prim[...] = 0.0
    prim[0] = rho0
    prim[4], prim[5], prim[6] = bx0, by0, bz0
    prim[7] = press0
    return prim


def add_alfven_wave(prim, nx, Lx, db0=0.1, wave_number_jet=1, cor_angle=0.0):
    """``ipert = 1`` (mhdinit.f90:328-342): circularly polarised Alfven wave along x, added in place
    to the primitive slab ``prim`` = [8, z_size, ny, nx]."""
    x = np.arange(nx) * (Lx / nx)
    kx = 2 * PI / Lx * wave_number_jet
    s = np.sin(kx * x)[None, None, :]
    c = np.cos(kx * x)[None, None, :]
    ca, sa = math.cos(cor_angle), math.sin(cor_angle)
    rs = np.sqrt(prim[0])
    prim[6] = prim[6] - db0 * s
    prim[3] = prim[3] + db0 / rs * s
    prim[1] = prim[1] + db0 / rs * c * sa
    prim[4] = prim[4] - db0 * c * sa
    prim[2] = prim[2] + db0 / rs * c * ca
    prim[5] = prim[5] - db0 * c * ca
    return prim
