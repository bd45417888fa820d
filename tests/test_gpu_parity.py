"""Parity of the CUDA path (through the C ABI) against the CPU oracle -- runs on a B200 (-m gpu).

Tolerances are the north star's: fields within 1e-11 relative L2 after one RK step, diagnostics
within 1e-9 relative after 100 steps, dealiasing masks bit-exact; FFTs within 1e-13."""
import numpy as np
import pytest

import parity_common as pc
from laps_b200 import Solver, synthetic
from oracle import laps_oracle as lo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(16, 16, 16), (64, 64, 64), (128, 32, 64), (32, 256, 16), (16, 16, 512),
                                   (512, 16, 16), (16, 1024, 16), (2048, 16, 16), (16, 16, 2048)])
def test_fft_forward_inverse_vs_oracle(shape):
    pc.check_fft(*shape)


CASES = {
    "hall_aeb_mask": dict(hall=True, aeb=True, dealias=1),
    "hall_aeb_filter": dict(hall=True, aeb=True, dealias=2),
    "mhd_plain": dict(hall=False, aeb=False, dealias=1),
    "hall_only": dict(hall=True, aeb=False, dealias=1),
    "aeb_corot": dict(hall=True, aeb=True, corot=True, dealias=1),
    "explicit_diffusion": dict(hall=False, aeb=True, dealias=1, explicit=True, conserve_bg=True),
    "ideal": dict(hall=True, aeb=True, dealias=1, visc=False, resis=False),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_one_step_parity_64(name):
    p, prim = pc.make_case(64, 64, 64, **CASES[name])
    o, g = pc.run_both(p, prim, 1)
    pc.check_state(o, g, 1e-11)
    pc.check_diagnostics(o, g, 1e-9)
    g.close()


def test_one_step_parity_anisotropic_grid_and_late_start():
    p, prim = pc.make_case(128, 32, 64, hall=True, aeb=True, corot=True)
    o, g = pc.run_both(p, prim, 2, t0=3.0)   # restart-style: evolve_radius(t_restart) first (mhd.f90:101-103)
    pc.check_state(o, g, 1e-11)
    g.close()


@pytest.mark.parametrize("flag", ["0", "1"])
def test_fused_and_separate_flux_x_pass_parity_128(flag, monkeypatch):
    """calc_flux inside the forward x pass (k_flux_fwd_x, default for 128 <= nx <= 512) and the separate
    k_flux + k_fwd_x path, both against the oracle at 128 x 64 x 64."""
    monkeypatch.setenv("LAPS_TUNE_FUSEX", flag)
    p, prim = pc.make_case(128, 64, 64, hall=True, aeb=True, dealias=1)
    o, g = pc.run_both(p, prim, 2)
    pc.check_state(o, g, 1e-11)
    pc.check_diagnostics(o, g, 1e-9)
    g.close()


@pytest.mark.parametrize("env", [dict(LAPS_TUNE_SPEC="0"), dict(LAPS_TUNE_MASS="0"), dict(LAPS_TUNE_SYM="0"),
                                 dict(LAPS_TUNE_MASS="0", LAPS_TUNE_SYM="0", LAPS_TUNE_SPEC="0", LAPS_TUNE_PRUNE="0"),
                                 dict(LAPS_TUNE_ZCHUNK="3")])
def test_work_skipping_switched_off_parity_64(env, monkeypatch):
    """Every exact work-skipping device (continuity row from the state, symmetric flux tensor, speculative front
    half with the fused CFL sweep, mask pruning) switched off in turn: the reference's own operation count
    (19 forward transforms, separate vardt sweep) must give the same parity against the oracle."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    p, prim = pc.make_case(64, 64, 64, hall=True, aeb=True, dealias=1)
    o, g = pc.run_both(p, prim, 2)
    pc.check_state(o, g, 1e-11)
    pc.check_diagnostics(o, g, 1e-9)
    g.close()


def test_mask_pruning_is_bit_exact():
    pc.check_pruning_is_exact((64, 64, 64), 3, hall=True, aeb=True, dealias=1)
    pc.check_pruning_is_exact((128, 64), 3, hall=True, aeb=True, dealias=3)
    pc.check_pruning_is_exact((128, 64), 2, hall=True, aeb=True, dealias=1)
    pc.check_pruning_is_exact((64, 64, 64), 2, incompressible=True, hall=True, aeb=True, dealias=1)


@pytest.mark.parametrize("name,kw", [("hall_aeb", dict(hall=True, aeb=True, dealias=1)),
                                     ("z_radial_square", dict(hall=True, aeb=True, z_radial=True, dealias=3)),
                                     ("filter_explicit", dict(hall=False, aeb=False, dealias=2, explicit=True, conserve_bg=True, limit_dt=True)),
                                     ("corotating", dict(hall=True, aeb=True, corot=True, dealias=1)),
                                     ("corotating_filter_explicit", dict(hall=True, aeb=True, corot=True, dealias=2, explicit=True, conserve_bg=True)),
                                     ("corotating_nodealias", dict(hall=False, aeb=True, corot=True, dealias=0))])
def test_2d_tree_parity(name, kw):
    """BASELINE config 2 family (src_compressible/2D) at 256 x 128, two steps."""
    p, prim = pc.make_case_2d(256, 128, **kw)
    o, g = pc.run_both(p, prim, 2, t0=2.0 if kw.get("corot") else 0.0)   # (a late start: the rotated wave vectors differ from t = 0)
    pc.check_state(o, g, 1e-11)
    pc.check_diagnostics(o, g, 1e-9)
    g.close()


def test_2d_tree_2048_properties():
    """BASELINE config 2 at full size (2048^2, Hall, no expansion): k=0 mode conserved bit-exactly,
    div B conserved to round-off, forward transform of the real state reproduces the spectrum."""
    p, prim = pc.make_case_2d(2048, 2048, hall=True, aeb=False, dealias=1)
    with Solver(**pc.solver_kwargs(p)) as g:
        g.set_primitive(prim)
        s0 = g.uu_fourier()[:, 0, 0, 0].copy()
        d0 = g.calc_max_divB()     # the smooth test field is not solenoidal: div B is a conserved, non-zero quantity
        g.vardt()
        for i in range(3):
            g.step(calc_dt=(i == 2))
        assert np.array_equal(g.uu_fourier()[:, 0, 0, 0], s0)
        # dB/dt = curl E (2D/mhdrhs.f90:308-313) keeps k.B^ fixed; only the implicit resistivity damps it (~4e-7 here)
        assert abs(g.calc_max_divB() - d0) < 1e-5 * d0
        uu, _ = g.get_state()
        assert np.isfinite(uu).all()
        assert pc.rel_l2(g.fft_forward(uu[:2]), g.uu_fourier()[:2]) < 1e-13


@pytest.mark.parametrize("name,kw", [("mhd", dict(hall=False, aeb=False, dealias=1)),
                                     ("hall_aeb_mask", dict(hall=True, aeb=True, dealias=1)),
                                     ("hall_aeb_corot_filter", dict(hall=True, aeb=True, corot=True, dealias=2)),
                                     ("explicit_retransform", dict(hall=True, aeb=False, dealias=0, explicit=True, conserve_bg=True))])
def test_incompressible_tree_parity_64(name, kw):
    """BASELINE config 3 family (src_incompressible) at 64^3, two steps, against the oracle."""
    p, prim = pc.make_case_incompressible(64, 64, 64, **kw)
    o, g = pc.run_both(p, prim, 2)
    pc.check_state(o, g, 1e-11)
    pc.check_diagnostics(o, g, 1e-9)
    assert abs(g.calc_max_divV() - o.calc_max_divV()) <= 1e-9 * o.calc_max_divV()
    db, dv = g.calc_max_div_real()
    odb, odv = o.calc_max_div_real()
    assert abs(dv - odv) <= 1e-9 * odv and abs(db - odb) <= 1e-9 * max(odb, 1e-6)
    assert g.rho0 == o.rho0
    g.close()


@pytest.mark.parametrize("name,kw", [("hall_aeb_square", dict(hall=True, aeb=True, dealias=3)),
                                     ("filter_explicit", dict(hall=False, aeb=False, dealias=2, explicit=True, conserve_bg=True, limit_dt=True)),
                                     ("retransform", dict(hall=True, aeb=False, dealias=0)),
                                     ("corotating", dict(hall=True, aeb=True, corot=True, dealias=1)),
                                     ("corotating_filter_explicit", dict(hall=True, aeb=True, corot=True, dealias=2, explicit=True, conserve_bg=True))])
def test_incompressible_2d_tree_parity(name, kw):
    """src_incompressible/2D at 256 x 128, two steps, against the oracle."""
    p, prim = pc.make_case_incompressible_2d(256, 128, **kw)
    o, g = pc.run_both(p, prim, 2, t0=2.0 if kw.get("corot") else 0.0)
    pc.check_state(o, g, 1e-11)
    pc.check_diagnostics(o, g, 1e-9)
    db, dv = g.calc_max_div_real()
    odb, odv = o.calc_max_div_real()
    assert abs(dv - odv) <= 1e-9 * odv and abs(db - odb) <= 1e-9 * odb
    g.close()


def test_incompressible_tree_256_properties():
    """BASELINE config 3 at full size (256^3 decaying turbulence, no expansion, no Hall): the projection keeps
    div(rho u) fixed, the k=0 mode of rho u and B is conserved bit-exactly, div B stays at round-off."""
    n = 256
    kw = dict(nx=n, ny=n, nz=n, Lx=24.0, Ly=24.0, Lz=24.0, adiabatic_index=1.666667, if_resis=1, resistivity=1e-4,
              if_visc=1, viscosity=1e-4, cfl=0.5, dealias_option=1, incompressible=1, rho0=1.0)
    prim = synthetic.turbulence_slab(n, n, n, 24.0, 24.0, 24.0, kmax=8, drho0=0.0)
    with Solver(**kw) as g:
        g.set_primitive(prim)
        s0 = g.uu_fourier()[:7, 0, 0, 0].copy()
        dv0 = g.calc_max_divV()
        g.vardt()
        for _ in range(2):
            g.step()
        uu, _ = g.get_state()
        assert np.isfinite(uu).all()
        assert np.array_equal(g.uu_fourier()[:7, 0, 0, 0], s0)
        assert g.calc_max_divB() < 1e-13
        assert g.calc_max_divV() < max(2 * dv0, 1e-12)       # solenoidal initial velocity, uniform density
        assert abs(uu[7].mean()) < 1e-14                      # the pressure carries no k=0 mode (mhdrhs.f90:505-508)


def test_set_primitive_modes_128():
    """Initial data from a mode table (laps_set_primitive_modes) against the uploaded host field at 128^3."""
    n = 128
    ks, coefs = synthetic.mode_table(24.0, 24.0, 24.0, 1.0, 0.0, 0.0, 8, 8, 8, (101, 116, 132), 0.1, 0.1, 0.01)
    prim = synthetic.turbulence_slab(n, n, n, 24.0, 24.0, 24.0, kmax=8)
    kw = dict(nx=n, ny=n, nz=n, Lx=24.0, Ly=24.0, Lz=24.0, dealias_option=1, if_hall=1, ion_inertial_length=0.2)
    with Solver(**kw) as a, Solver(**kw) as b:
        a.set_primitive(prim)
        b.set_primitive_modes(ks, coefs, [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 1.0])
        assert np.abs(a.get_state()[0] - b.get_state()[0]).max() < 1e-13
        a.vardt(); b.vardt()
        a.step(); b.step()
        assert pc.rel_l2(b.get_state()[0], a.get_state()[0]) < 1e-12 and abs(a.dt - b.dt) < 1e-13 * a.dt


@pytest.mark.parametrize("inc", [False, True])
def test_hall_wave_known_answer(inc):
    """Independent of the oracle: the CUDA path against the exact Hall-MHD wave solution (third-order convergence
    to the whistler branch of the dispersion relation), compressible and incompressible trees."""
    pc.check_hall_wave_known_answer(incompressible=inc)


def test_dealias_mask_bit_exact():
    p, prim = pc.make_case(32, 64, 32, hall=False, aeb=False, dealias=1)
    o, g = pc.run_both(p, prim, 1)
    mask = lo.dealias_mask(p, lo.Grid(p))          # True where the mode is removed (dealiasing.f90:91-97)
    uf = g.uu_fourier()
    for v in range(8):
        assert np.array_equal(uf[v] == 0, mask | (o.uu_fourier[v] == 0))
        assert np.all(uf[v][mask] == 0)
    g.close()


def test_100_steps_diagnostics_parity():
    p, prim = pc.make_case(32, 32, 32, hall=True, aeb=True, dealias=1)
    o, g = pc.run_both(p, prim, 100)
    pc.check_diagnostics(o, g, 1e-9)
    pc.check_state(o, g, 1e-9)
    g.close()


def test_k0_mode_conserved_bit_exact_without_expansion():
    p, prim = pc.make_case(64, 64, 64, hall=True, aeb=False)
    with Solver(**pc.solver_kwargs(p)) as g:
        g.set_primitive(prim)
        k0 = g.uu_fourier()[:, 0, 0, 0].copy()
        g.vardt()
        for _ in range(5):
            g.step()
        assert np.array_equal(g.uu_fourier()[:, 0, 0, 0], k0)     # fnl(k=0) = 0 exactly
        assert g.calc_max_divB() < 1e-14


def test_transpose_yz_indexmap_matches_reference_tables():
    with Solver(nx=16, ny=32, nz=64, dealias_option=0) as g:
        m = g.transpose_yz_indexmap().reshape(g.nxh, g.ny, g.nzl, 2)
        # single rank: destination is [kx][ky][z] of rank 0
        kx, ky, z = np.meshgrid(np.arange(g.nxh), np.arange(g.ny), np.arange(g.nzl), indexing="ij")
        assert np.all(m[..., 0] == 0)
        assert np.array_equal(m[..., 1], (kx * g.ny + ky) * g.nz + z)
        m2 = g.transpose_zy_indexmap().reshape(g.nxh, g.nyl, g.nz, 2)      # the inverse direction (parallel.f90:300-324)
        kx, ky, z = np.meshgrid(np.arange(g.nxh), np.arange(g.nyl), np.arange(g.nz), indexing="ij")
        assert np.all(m2[..., 0] == 0)
        assert np.array_equal(m2[..., 1], (kx * g.ny + ky) * g.nz + z)


def test_full_size_properties_512():
    """BASELINE's full size through size-independent properties: forward/inverse round trip of the
    state, Hermitian-consistent real output, exact conservation of the k=0 mode over a step."""
    n = 512
    kw = dict(nx=n, ny=n, nz=n, Lx=24.0, Ly=24.0, Lz=24.0, adiabatic_index=1.666667, if_resis=1, resistivity=1e-4,
              if_visc=1, viscosity=1e-4, cfl=0.5, dealias_option=1, if_AEB=0, radius0=30.0, if_hall=1,
              ion_inertial_length=0.2)
    prim = synthetic.turbulence_slab(n, n, n, 24.0, 24.0, 24.0, kmax=8)
    with Solver(**kw) as g:
        g.set_primitive(prim)
        uu, pr = g.get_state()
        # prim -> cons -> forward -> (get_state reads uu as uploaded/converted): primitives come back
        assert pc.rel_l2(uu[0], prim[0]) < 1e-14
        assert pc.rel_l2(pr[0], prim[1]) < 1e-13 and pc.rel_l2(pr[3], prim[7]) < 1e-12
        s0 = g.uu_fourier()[:, 0, 0, 0].copy()
        assert abs(s0[0].real - prim[0].mean()) < 1e-13
        g.vardt()
        g.step()
        uu1, _ = g.get_state()
        assert np.isfinite(uu1).all()
        assert np.array_equal(g.uu_fourier()[:, 0, 0, 0], s0)
        assert g.calc_max_divB() < 1e-13
        # the real state after the step is the inverse transform of the (band-limited) spectrum:
        # transforming it forward again must reproduce the spectrum
        spec = g.fft_forward(uu1[:2])
        ref = g.uu_fourier()[:2]
        assert pc.rel_l2(spec, ref) < 1e-13


def test_2d_tree_external_force():
    """if_external_force (2D/mhdrhs.f90:480-531) at 256 x 128, three steps, against the oracle."""
    pc.check_external_force((256, 128))


def test_check_nan():
    pc.check_nan_detection()


def test_eight_point_lines():
    pc.check_eight_point_lines()


@pytest.mark.parametrize("name", ["hall_aeb_mask", "corot_filter_explicit", "plain_nodealias"])
def test_library_agrees_with_the_executed_reference_source(name):
    """Golden vectors made by executing the reference's own Fortran source (tests/golden/make_ref_exec_fixtures.py):
    two steps of the Principal loop, fields within 1e-11 relative L2."""
    import test_reference_source_pins as rp
    rp.check_library(name)


@pytest.mark.parametrize("name", ["incomp_hall_aeb_mask", "incomp_corot_filter_explicit", "incomp_plain_nodealias"])
def test_incompressible_library_agrees_with_the_executed_reference_source(name):
    import test_reference_source_pins as rp
    rp.check_library_incompressible(name)


@pytest.mark.parametrize("name", ["c2d_hall_aeb_mask", "c2d_zradial_square_explicit", "c2d_external_force_filter", "c2d_corotating_oracle_only"])
def test_2d_library_agrees_with_the_executed_reference_source(name):
    import test_reference_source_pins as rp
    rp.check_library_2d(name)


@pytest.mark.parametrize("name", ["i2d_hall_aeb_mask", "i2d_square_explicit_limit", "i2d_corotating"])
def test_incompressible_2d_library_agrees_with_the_executed_reference_source(name):
    import test_reference_source_pins as rp
    rp.check_library_incompressible_2d(name)


def test_100_steps_against_the_executed_reference_source():
    """North star: energy, cross helicity and div B within 1e-9 relative after 100 steps — here against golden vectors of
    the reference's own source executed for 100 steps (tests/golden/ref_exec/hall_aeb_mask_100steps.npz)."""
    import test_reference_source_pins as rp
    rp.check_library_100_steps()


@pytest.mark.parametrize("form", ["1", "2"])
def test_two_stream_schedule_single_gpu(monkeypatch, form):
    """The opt-in two-stream stage schedules (LAPS_TUNE_OVERLAP) forced on one GPU: same state as the oracle, and bit-identical
    to the one-stream schedule (the same kernels on the same data; only the order of independent launches differs)."""
    p, prim = pc.make_case(64, 64, 64, hall=True, aeb=True, dealias=1)
    monkeypatch.setenv("LAPS_TUNE_OVERLAP", form)
    o, g = pc.run_both(p, prim, 2)
    pc.check_state(o, g, 1e-11)
    uu1, uf1 = g.get_state()[0], g.uu_fourier()
    g.close()
    monkeypatch.setenv("LAPS_TUNE_OVERLAP", "0")
    o, g = pc.run_both(p, prim, 2)
    assert np.array_equal(g.get_state()[0], uu1) and np.array_equal(g.uu_fourier(), uf1)
    g.close()


def test_cfl_screen_is_exact():
    pc.check_cfl_screen_is_exact(shape=(64, 64, 64), nsteps=3)


def test_rhs_kernel_variants():
    pc.check_rhs_kernel_variants(shape=(64, 64, 128), nsteps=2, hall=True, aeb=True, dealias=1)
    pc.check_rhs_kernel_variants(shape=(32, 32, 64), nsteps=2, hall=True, aeb=True, corot=True, dealias=2, explicit=True)
    pc.check_rhs_kernel_variants(shape=(64, 32, 32), nsteps=2, hall=False, aeb=False, dealias=0)


def test_async_output():
    pc.check_async_output(shape=(64, 64, 64))
