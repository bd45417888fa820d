"""Shared parity checks: the CUDA path (through the C ABI) against the CPU oracle.

Used by tests/test_gpu_parity.py on a B200 (`-m gpu`) and — with the test-only kernel emulator —
by tests/test_emulated_kernels.py in the CPU-only container."""
import numpy as np

from laps_b200 import Solver
from oracle import laps_oracle as lo


def rel_l2(a, b):
    d = np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel())
    n = np.linalg.norm(np.asarray(b).ravel())
    return d / n if n > 0 else d


def make_case(nx, ny, nz, hall=True, aeb=True, corot=False, dealias=1, visc=True, resis=True,
              explicit=False, conserve_bg=False, seed=101, nmode=2):
    """Oracle parameters + primitive initial data (ifield=3 + ipert=7-style turbulence)."""
    p = lo.Params(nx=nx, ny=ny, nz=nz, Lx=24.0, Ly=24.0, Lz=24.0, adiabatic_index=1.666667,
                  if_resis=resis, resistivity=1e-4 if not explicit else 1e-3, if_resis_exp=explicit,
                  if_visc=visc, viscosity=1e-4 if not explicit else 1e-3, if_visc_exp=explicit,
                  if_conserve_background=conserve_bg, cfl=0.5, dealias_option=dealias,
                  if_AEB=aeb, radius0=30.0, Ur0=1.167 if aeb else 0.0, if_corotating=corot,
                  corotating_angle=0.3 if corot else 0.0,
                  if_hall=hall, ion_inertial_length=0.2 if hall else 0.0)
    prim = lo.ic_uniform_background(p, bx0=1.0, press0=1.0)
    prim = lo.ic_turbulence(p, prim, 1.0, 0.0, 0.0, db0=0.1, dv0=0.1, drho0=0.01,
                            nmodex=nmode, nmodey=nmode, nmodez=nmode, seeds=(seed, seed + 15, seed + 31))
    return p, prim


def make_case_2d(nx, ny, hall=True, aeb=True, z_radial=False, dealias=1, visc=True, resis=True, explicit=False,
                 conserve_bg=False, limit_dt=False, seed=3, corot=False):
    """2D tree (src_compressible/2D): oracle parameters + smooth random primitive data on (nx, ny, 1)."""
    p = lo.Params(nx=nx, ny=ny, nz=1, Lx=24.0, Ly=12.0, Lz=1.0, adiabatic_index=1.666667,
                  if_resis=resis, resistivity=1e-4 if not explicit else 1e-3, if_resis_exp=explicit,
                  if_visc=visc, viscosity=1e-4 if not explicit else 1e-3, if_visc_exp=explicit,
                  if_conserve_background=conserve_bg, cfl=0.5, dealias_option=dealias,
                  if_AEB=aeb, radius0=30.0, Ur0=1.167 if aeb else 0.0, if_z_radial=z_radial,
                  if_corotating=corot, corotating_angle=0.4 if corot else 0.0,
                  if_hall=hall, ion_inertial_length=0.2 if hall else 0.0, if_limit_dt_increase=limit_dt)
    rng = np.random.default_rng(seed)
    x = 2 * np.pi * np.arange(nx) / nx
    y = 2 * np.pi * np.arange(ny) / ny
    Y, X = np.meshgrid(y, x, indexing="ij")

    def smooth(amp):
        f = np.zeros((ny, nx))
        for kx in range(0, 3):
            for ky in range(-2, 3):
                if kx == 0 and ky <= 0:
                    continue
                f += rng.standard_normal() * np.cos(kx * X + ky * Y + rng.uniform(0, 2 * np.pi)) / (kx * kx + ky * ky)
        return amp * f

    prim = np.zeros((8, 1, ny, nx))
    prim[0, 0] = 1.0 + smooth(0.01)
    for v in (1, 2, 3):
        prim[v, 0] = smooth(0.1)
    prim[4, 0] = 1.0 + smooth(0.1)
    prim[5, 0] = smooth(0.1)
    prim[6, 0] = 0.3 + smooth(0.1)
    prim[7, 0] = 1.0 + smooth(0.02)
    return p, prim


def solver_kwargs(p: lo.Params):
    extra = dict(ndim=2, if_z_radial=p.if_z_radial, if_limit_dt_increase=p.if_limit_dt_increase) if p.nz == 1 else {}
    if p.if_external_force:
        extra["if_external_force"] = 1
    if p.incompressible:
        extra = dict(extra, incompressible=1, rho0=p.rho0)
        extra.pop("if_z_radial", None)
    return dict(**extra, **_solver_kwargs(p))


def _solver_kwargs(p: lo.Params):
    return dict(nx=p.nx, ny=p.ny, nz=p.nz, Lx=p.Lx, Ly=p.Ly, Lz=p.Lz, adiabatic_index=p.adiabatic_index,
                if_resis=p.if_resis, if_resis_exp=p.if_resis_exp, resistivity=p.resistivity,
                if_visc=p.if_visc, if_visc_exp=p.if_visc_exp, viscosity=p.viscosity,
                if_conserve_background=p.if_conserve_background, cfl=p.cfl, dealias_option=p.dealias_option,
                afx=p.afx, afy=p.afy, afz=p.afz, if_AEB=p.if_AEB, if_corotating=p.if_corotating,
                radius0=p.radius0, Ur0=p.Ur0, corotating_angle=p.corotating_angle,
                if_hall=p.if_hall, ion_inertial_length=p.ion_inertial_length)


def oracle_state(p):
    if p.incompressible:
        return lo.StateIncompressible2D(p) if p.nz == 1 else lo.StateIncompressible(p)
    return lo.State2D(p) if p.nz == 1 else lo.State(p)


def make_case_incompressible(nx, ny, nz, rho0=1.0, **kw):
    """BASELINE config 3 family (src_incompressible): the turbulence case of make_case run through the
    incompressible tree (uu(8) = pressure); drho0 = 0.01 keeps rho non-uniform, which the tree allows."""
    p, prim = make_case(nx, ny, nz, **kw)
    p.incompressible = True
    p.rho0 = rho0
    return p, prim


def make_case_incompressible_2d(nx, ny, rho0=1.0, **kw):
    """src_incompressible/2D: the smooth 2D data of make_case_2d run through the incompressible tree."""
    p, prim = make_case_2d(nx, ny, **kw)
    p.incompressible = True
    p.rho0 = rho0
    return p, prim


def run_both(p, prim, nsteps, lib_path=None, t0=0.0):
    """Drive oracle and library exactly as mhd.f90 does: [set time]; vardt; nsteps x step."""
    o = oracle_state(p)
    o.set_primitive(prim)
    g = Solver(lib_path, **solver_kwargs(p))
    g.set_primitive(prim)
    if t0:
        o.time = t0
        o.evolve_radius(t0)
        g.time = t0
        g.evolve_radius(t0)
    o.vardt()
    g.vardt()
    for _ in range(nsteps):
        o.step()
        g.step()
    return o, g


def check_fft(nx, ny, nz, lib_path=None, tol=1e-13, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((3, nz, ny, nx))
    with Solver(lib_path, nx=nx, ny=ny, nz=nz, Lx=1.0, Ly=1.0, Lz=1.0, dealias_option=0) as s:
        w = s.fft_forward(a)
        ref = lo.fft_forward(a)
        assert rel_l2(w, ref) < tol, ("forward", rel_l2(w, ref))
        # inverse of an arbitrary (not Hermitian-consistent) spectrum: c2r must drop Im(DC/Nyquist)
        spec = ref + 0.0
        spec[..., 0] += 0.25j
        spec[..., -1] -= 0.5j
        b = s.fft_inverse(spec)
        assert rel_l2(b, a) < tol, ("inverse", rel_l2(b, a))
        # single-mode known answer: cos(2 pi (2x/Lx + 3y/Ly + 1 z/Lz)) -> 1/2 at (kx,ky,kz)=(2,3,1)
        z, y, x = np.meshgrid(np.arange(nz) / nz, np.arange(ny) / ny, np.arange(nx) / nx, indexing="ij")
        m = np.cos(2 * np.pi * (2 * x + 3 * y + 1 * z))[None]
        wm = s.fft_forward(m)[0]
        assert abs(wm[1, 3, 2] - 0.5) < 1e-14
        wm[1, 3, 2] = 0
        assert np.abs(wm).max() < 1e-14


def check_state(o, g, tol_field, tol_spec=None):
    uu, prim = g.get_state()
    for v in range(8):
        assert rel_l2(uu[v], o.uu[v]) < tol_field, (v, rel_l2(uu[v], o.uu[v]))
    for v in range(3 if o.p.incompressible else 4):
        assert rel_l2(prim[v], o.uu_prim[v]) < 10 * tol_field, (v, rel_l2(prim[v], o.uu_prim[v]))
    uf = g.uu_fourier()
    for v in range(8):
        assert rel_l2(uf[v], o.uu_fourier[v]) < (tol_spec or tol_field), (v, rel_l2(uf[v], o.uu_fourier[v]))
    assert abs(g.dt - o.dt) <= 1e-12 * abs(o.dt), (g.dt, o.dt)
    assert abs(g.time - o.time) <= 1e-12 * max(abs(o.time), 1e-300)


def check_state_vectors(o, g, tol):
    """check_state for cases in which one Cartesian component of a vector field is (nearly) zero — plane-polarised
    perturbations on a strongly anisotropic grid: rho and e per field, rho u and B as vectors (relative L2 of the
    three components together), in real and in Fourier space."""
    uu, _ = g.get_state()
    uf = g.uu_fourier()
    for a, b in ((uu, o.uu), (uf, o.uu_fourier)):
        for grp in ((0,), (1, 2, 3), (4, 5, 6), (7,)):
            e = rel_l2(np.stack([a[v] for v in grp]), np.stack([b[v] for v in grp]))
            assert e < tol, (grp, e)
    assert abs(g.dt - o.dt) <= 1e-12 * abs(o.dt), (g.dt, o.dt)
    assert abs(g.time - o.time) <= 1e-12 * max(abs(o.time), 1e-300)


def check_diagnostics(o, g, tol):
    ave, rms, ru2 = g.calc_rms()
    oave, orms, oru2 = o.calc_rms()
    assert np.allclose(ave, oave, rtol=tol, atol=tol * 1e-3), (ave, oave)
    msq = np.asarray(orms) + np.asarray(oave) ** 2     # uu_rms = <u^2> - <u>^2 cancels: allowance 1e-13 <u^2>
    assert np.all(np.abs(np.asarray(rms) - np.asarray(orms)) <= tol * np.abs(orms) + 1e-13 * msq), (rms, orms)
    assert np.allclose(ru2, oru2, rtol=tol, atol=1e-18), (ru2, oru2)
    inv = g.invariants()
    oinv = o.invariants()
    assert abs(inv[0] - oinv[0]) <= tol * abs(oinv[0])
    assert abs(inv[1] - oinv[1]) <= tol * max(abs(oinv[1]), 1e-6)
    # max |k.B^| is a sum of three cancelling terms: relative tolerance + an absolute round-off allowance in their natural
    # scale k_max B_rms (without the expanding box the result itself is round-off of those terms)
    pp = o.p
    kmax = np.pi * max(pp.nx / pp.Lx, pp.ny / pp.Ly, (pp.nz / pp.Lz) if pp.nz > 1 else 0.0)
    brms = float(np.sqrt(sum(np.mean(np.asarray(o.uu[v]) ** 2) for v in (4, 5, 6))))
    assert abs(inv[2] - oinv[2]) <= tol * oinv[2] + 1e-14 * kmax * brms, (inv[2], oinv[2])
    assert abs(g.calc_max_divB() - inv[2]) == 0.0


def check_pruning_is_exact(shape, nsteps, lib_path=None, incompressible=False, **case):
    """The passes skip what the dealiasing mask removes (LAPS_TUNE_PRUNE, default on): whole kx columns and ky
    rows (rectangle), with the spherical mask also the (kx, ky) columns outside the circle (LAPS_TUNE_CIRCLE)
    and, inside a surviving column, the state / RK-history entries of masked kz (LAPS_TUNE_KZPRUNE).  The state
    must be BIT-IDENTICAL to a run that computes and moves everything, for every combination."""
    import os
    if incompressible:
        p, prim = (make_case_incompressible_2d(*shape, **case) if len(shape) == 2 else make_case_incompressible(*shape, **case))
    else:
        p, prim = (make_case_2d(*shape, **case) if len(shape) == 2 else make_case(*shape, **case))
    keys = ("LAPS_TUNE_PRUNE", "LAPS_TUNE_CIRCLE", "LAPS_TUNE_KZPRUNE")
    out = []
    for env in (dict(LAPS_TUNE_PRUNE="0"), {}, dict(LAPS_TUNE_CIRCLE="0"), dict(LAPS_TUNE_KZPRUNE="0"),
                dict(LAPS_TUNE_CIRCLE="0", LAPS_TUNE_KZPRUNE="0")):
        saved = {k: os.environ.pop(k, None) for k in keys}
        os.environ.update(env)
        try:
            g = Solver(lib_path, **solver_kwargs(p))
        finally:
            for k in keys:
                os.environ.pop(k, None)
                if saved[k] is not None:
                    os.environ[k] = saved[k]
        g.set_primitive(prim)
        g.vardt()
        for _ in range(nsteps):
            g.step()
        uu, _ = g.get_state()
        out.append((uu, g.uu_fourier(), g.dt, g.calc_max_divB(), g.pruning_counts()))
        g.close()
    for o in out[1:]:
        assert np.array_equal(out[0][0], o[0])
        assert np.array_equal(out[0][1], o[1])
        assert out[0][2] == o[2] and out[0][3] == o[3]
    assert out[1][4][1] < out[0][4][1]          # the default does skip modes
    return [o[4] for o in out]


def check_external_force(shape=(64, 32), lib_path=None, nsteps=3, tol=1e-11):
    """if_external_force of the 2D compressible tree (2D/mhdrhs.f90:129-131,216-251,370-372,480-531): the driver hands
    the field of its user routine to the library once per step (laps_set_external_force); parity with the oracle, which
    restates the shipped routine, and the force must matter (B_z differs from the unforced run far above the tolerance)."""
    p, prim = make_case_2d(*shape, hall=True, aeb=True, dealias=1)
    p.if_external_force = True
    o = oracle_state(p)
    o.set_primitive(prim)
    o.vardt()
    with Solver(lib_path, **solver_kwargs(p)) as g:
        g.set_primitive(prim)
        g.vardt()
        for _ in range(nsteps):
            assert abs(g.time - o.time) <= 1e-12 * max(abs(o.time), 1.0)
            g.set_external_force(o.calc_external_force_real())   # `time` is fixed during evolve: once per step
            o.step()
            g.step()
        check_state(o, g, tol)
        check_diagnostics(o, g, 1e-9)
        assert not g.checkNan()
        uu, _ = g.get_state()
    p0, _ = make_case_2d(*shape, hall=True, aeb=True, dealias=1)
    o0, g0 = run_both(p0, prim, nsteps, lib_path=lib_path)
    uu0, _ = g0.get_state()
    g0.close()
    assert rel_l2(uu[6], uu0[6]) > 1e-4, rel_l2(uu[6], uu0[6])
    # a handle created without the flag refuses the call; the flag exists only in the 2D compressible tree
    import pytest
    from laps_b200 import capi
    with Solver(lib_path, **solver_kwargs(p0)) as g1:
        with pytest.raises(capi.LapsError):
            g1.set_external_force(np.zeros((1,) + shape[::-1]))
    with pytest.raises(capi.LapsError):
        Solver(lib_path, nx=16, ny=16, nz=16, if_external_force=1)


def check_eight_point_lines(lib_path=None):
    """The line axis of the fused spectral pass may be 8 points long (one register-resident radix-8 stage): the 2D
    input the reference ships is 256 x 8 (src_compressible/2D/mhd.input:12-13), which also needs the half-height
    x-pass tile.  All four trees, two steps, against the oracle."""
    cases = [(make_case_2d, (256, 8), dict(hall=True, aeb=True, dealias=1)),
             (make_case_2d, (64, 8), dict(hall=True, aeb=True, z_radial=True, dealias=3)),
             (make_case_2d, (32, 8), dict(hall=False, aeb=False, dealias=2, explicit=True, conserve_bg=True, limit_dt=True)),
             (make_case_incompressible_2d, (32, 8), dict(hall=True, aeb=True, dealias=1)),
             (make_case, (16, 16, 8), dict(hall=True, aeb=True, dealias=1, nmode=1)),
             (make_case_incompressible, (16, 16, 8), dict(hall=True, aeb=True, dealias=2, nmode=1))]
    for make, shape, kw in cases:
        p, prim = make(*shape, **kw)
        o, g = run_both(p, prim, 2, lib_path=lib_path)
        check_state(o, g, 1e-11)
        check_diagnostics(o, g, 1e-9)
        g.close()
    import pytest
    from laps_b200 import capi
    with pytest.raises(capi.LapsError):      # only the line axis of the spectral pass: nx = 8 and a 3D ny = 8 are refused
        Solver(lib_path, nx=8, ny=16, nz=16)
    with pytest.raises(capi.LapsError):
        Solver(lib_path, nx=16, ny=8, nz=16)


def check_nan_detection(lib_path=None):
    """checkNan (2D/mhd.f90:563-591): clean state -> 0; one NaN anywhere in uu -> 1."""
    p, prim = make_case_2d(32, 16, hall=False, aeb=False, dealias=2)
    with Solver(lib_path, **solver_kwargs(p)) as g:
        g.set_primitive(prim)
        assert g.checkNan() is False
        bad = prim.copy()
        bad[5, 0, 7, 11] = np.nan
        g.set_primitive(bad)
        assert g.checkNan() is True


def check_hall_wave_known_answer(lib_path=None, incompressible=False, nsteps=(30, 60)):
    """Analytic known answer straight on the library (no oracle): a circularly polarised wave along B0 is an exact
    solution of Hall-MHD with  w^2 + (di B0 k^2 / rho) w - (B0 k)^2 / rho = 0.  Fixed time step (laps_rkt_init),
    two resolutions in time: the error must be the RK3 one (ratio ~ 8) and small."""
    import math
    L, di, kint, b0, B0, rho, T = 2 * np.pi, 0.5, 2, 0.05, 1.0, 1.0, 0.3
    k = float(kint)
    sig = di * B0 * k * k / rho
    w = 0.5 * (-sig + math.sqrt(sig * sig + 4 * k * k * B0 * B0 / rho))
    nx = 32
    x = (np.arange(nx) * (L / nx))[None, None, :]
    prim = np.zeros((8, 16, 16, nx))
    prim[0], prim[4], prim[7] = rho, B0, 1.0
    prim[5] += b0 * np.cos(k * x)
    prim[6] += b0 * np.sin(k * x)
    prim[2] += -(B0 * k / (rho * w)) * b0 * np.cos(k * x)
    prim[3] += -(B0 * k / (rho * w)) * b0 * np.sin(k * x)
    errs = []
    for n in nsteps:
        extra = dict(incompressible=1, rho0=1.0) if incompressible else {}
        with Solver(lib_path, nx=nx, ny=16, nz=16, Lx=L, Ly=L, Lz=L, dealias_option=1, if_hall=1, ion_inertial_length=di, **extra) as g:
            g.set_primitive(prim)
            dt = T / n
            for _ in range(n):
                g.rkt_init(dt)
                g.evolve()
            uu, _ = g.get_state()
            xs = x[0, 0]
            errs.append(max(np.abs(uu[5][3, 5, :] - b0 * np.cos(k * xs - w * T)).max(),
                            np.abs(uu[6][3, 5, :] - b0 * np.sin(k * xs - w * T)).max()) / b0)
            assert np.abs(uu[0] - rho).max() < 1e-12 and g.calc_max_divB() < 1e-13
    assert errs[0] < 2e-5 and errs[1] < 3e-6 and 6.0 < errs[0] / errs[1] < 10.0, errs
    return errs


def check_cfl_screen_is_exact(lib_path=None, shape=(32, 32, 32), nsteps=3):
    """vardt evaluates the FP64 signal speeds only where a cheap FP32 bound says they can raise a maximum
    (pointwise.cuh: cfl_may_raise): dt must be BIT-IDENTICAL to evaluating every point (LAPS_TUNE_SCREEN=0), through
    laps_vardt (k_cfl) and through laps_step (the sweep fused into calc_flux), also for a supersonic state in which the
    transverse bounds are loose, and for a uniform state (every point attains the maximum)."""
    import os
    cases = []
    p, prim = make_case(*shape, hall=True, aeb=True, dealias=1)
    cases.append((p, prim))
    fast = prim.copy()
    fast[1] += 7.0                                   # Mach ~ 5 along x
    fast[7] *= 0.05
    cases.append((p, fast))
    flat = np.zeros_like(prim)
    flat[0], flat[4], flat[7] = 1.0, 1.0, 1.0
    cases.append((p, flat))
    for p, prim in cases:
        dts = []
        for screen in ("0", "1"):
            saved = os.environ.get("LAPS_TUNE_SCREEN")
            os.environ["LAPS_TUNE_SCREEN"] = screen
            try:
                g = Solver(lib_path, **solver_kwargs(p))
            finally:
                if saved is None:
                    os.environ.pop("LAPS_TUNE_SCREEN", None)
                else:
                    os.environ["LAPS_TUNE_SCREEN"] = saved
            g.set_primitive(prim)
            d = [g.vardt()]
            for _ in range(nsteps):
                d.append(g.step())
            d.append(g.vardt())
            dts.append(d)
            g.close()
        assert dts[0] == dts[1], dts


def check_rhs_kernel_variants(lib_path=None, shape=(32, 32, 32), nsteps=2, exact=False, **case):
    """The RK-stage z pass exists in three forms that do the same arithmetic in the same order: the persistent pipelined
    kernel with two landing lines (LAPS_TUNE_RHS=1, default), with one landing line and more resident columns (=2), and
    the plain k_spec_z rows (=0).  Each against the oracle; all three agree with each other to 1e-13 (bit for bit where the
    build does not contract multiplications and additions: `exact`)."""
    import os
    p, prim = make_case(*shape, **case)
    states = []
    for v in ("1", "2", "0"):
        saved = os.environ.get("LAPS_TUNE_RHS")
        os.environ["LAPS_TUNE_RHS"] = v
        try:
            o, g = run_both(p, prim, nsteps, lib_path=lib_path)
        finally:
            if saved is None:
                os.environ.pop("LAPS_TUNE_RHS", None)
            else:
                os.environ["LAPS_TUNE_RHS"] = saved
        check_state(o, g, 1e-11)
        states.append((g.get_state()[0], g.uu_fourier()))
        g.close()
    for uu, uf in states[1:]:
        if exact:    # (the emulator build has no FMA contraction; nvcc contracts each kernel's expressions in its own way)
            assert np.array_equal(uu, states[0][0]) and np.array_equal(uf, states[0][1])
        else:
            assert rel_l2(uu, states[0][0]) < 1e-13 and rel_l2(uf, states[0][1]) < 1e-13


def check_async_output(lib_path=None, shape=(32, 32, 32)):
    """laps_get_output_async / laps_output_wait: the dump is the state at the request, whatever runs afterwards; a second
    request before the wait is served after the first (one snapshot buffer)."""
    p, prim = make_case(*shape, hall=True, aeb=True, dealias=1)
    with Solver(lib_path, **solver_kwargs(p)) as g:
        g.set_primitive(prim)
        g.vardt()
        g.step()
        want1 = g.get_output(True)
        a = np.empty_like(want1)
        g.get_output_async(a, True)
        g.step()
        g.step()
        want2 = g.get_output(False)
        b = np.empty_like(want2)
        g.get_output_async(b, False)      # first request still pending on the host side
        g.step()
        g.output_wait()
        assert np.array_equal(a, want1) and np.array_equal(b, want2)
        g.output_wait()                   # nothing pending: returns at once
