"""The UNCHANGED kernel sources of laps_b200/csrc, compiled with g++ against the test-only
execution-model emulator (tests/emu) and driven through the same C ABI, compared with the oracle.
This validates index math, shared-memory staging and launch geometry in the CPU-only container;
the product never loads the emulator, and the real parity tests are tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import build_emu  # noqa: E402
import parity_common as pc  # noqa: E402
from laps_b200 import Solver, synthetic  # noqa: E402
from oracle import laps_oracle as lo  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    return build_emu.build()


@pytest.mark.parametrize("shape", [(16, 16, 16), (32, 16, 64), (128, 16, 16)])
def test_fft(emu, shape):
    pc.check_fft(*shape, lib_path=emu)


@pytest.mark.parametrize("kw", [dict(hall=True, aeb=True, dealias=1), dict(hall=False, aeb=True, corot=True, dealias=2),
                                dict(hall=True, aeb=False, explicit=True, conserve_bg=True)])
def test_one_step(emu, kw):
    p, prim = pc.make_case(16, 16, 16, **kw)
    o, g = pc.run_both(p, prim, 1, lib_path=emu)
    pc.check_state(o, g, 1e-11)
    pc.check_diagnostics(o, g, 1e-9)
    g.close()


@pytest.mark.parametrize("kw", [dict(hall=True, aeb=True, dealias=1), dict(hall=True, aeb=True, z_radial=True, dealias=3),
                                dict(hall=False, aeb=False, dealias=2, explicit=True, conserve_bg=True, limit_dt=True),
                                # if_corotating (2D/mhdrhs.f90:282-288): both components of the rotated wave vector vary along the line
                                dict(hall=True, aeb=True, corot=True, dealias=1),
                                dict(hall=True, aeb=True, corot=True, dealias=2, explicit=True, conserve_bg=True),
                                dict(hall=False, aeb=True, corot=True, dealias=0)])
def test_one_step_2d_tree(emu, kw):
    p, prim = pc.make_case_2d(32, 16, **kw)
    o, g = pc.run_both(p, prim, 2, lib_path=emu, t0=2.0 if kw.get("corot") else 0.0)
    pc.check_state(o, g, 1e-11)
    pc.check_diagnostics(o, g, 1e-9)
    g.close()


def test_external_force_2d_tree(emu):
    pc.check_external_force((32, 16), lib_path=emu)


def test_eight_point_lines(emu):
    pc.check_eight_point_lines(lib_path=emu)


def test_check_nan(emu):
    pc.check_nan_detection(lib_path=emu)


@pytest.mark.parametrize("kw", [dict(hall=True, aeb=True, dealias=1), dict(hall=True, aeb=True, corot=True, dealias=2),
                                dict(hall=False, aeb=False, dealias=0, explicit=True, conserve_bg=True)])
def test_one_step_incompressible_tree(emu, kw):
    """src_incompressible: projection kernel, gradient/current tasks, retransform mode (dealias_option 0)."""
    p, prim = pc.make_case_incompressible(16, 16, 16, **kw)
    o, g = pc.run_both(p, prim, 2, lib_path=emu)
    pc.check_state(o, g, 1e-11)
    pc.check_diagnostics(o, g, 1e-9)
    assert abs(g.calc_max_divV() - o.calc_max_divV()) <= 1e-9 * o.calc_max_divV()
    db, dv = g.calc_max_div_real()
    odb, odv = o.calc_max_div_real()
    assert abs(dv - odv) <= 1e-9 * odv and abs(db - odb) <= 1e-9 * max(odb, 1e-6)
    assert g.rho0 == o.rho0
    g.close()


def test_fused_flux_forward_x_pass_matches_the_separate_kernels(emu, monkeypatch):
    """k_flux_fwd_x (calc_flux inside the forward x pass) against k_flux + k_fwd_x: same expressions, same
    transform -> bit-identical state on the emulator; and against the oracle."""
    p, prim = pc.make_case(32, 16, 16, hall=True, aeb=True, dealias=1)
    out = []
    for flag in ("0", "1"):
        monkeypatch.setenv("LAPS_TUNE_FUSEX", flag)
        o, g = pc.run_both(p, prim, 2, lib_path=emu)
        pc.check_state(o, g, 1e-11)
        out.append((g.get_state()[0], g.uu_fourier()))
        g.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize("shape,kw", [((32, 16, 16), dict(hall=True, aeb=True, dealias=1)),
                                      ((32, 16), dict(hall=True, aeb=True, z_radial=True, dealias=3))])
def test_speculative_front_half_is_bit_identical(emu, monkeypatch, shape, kw):
    """laps_step runs the next step's dt-independent front half with the CFL sweep fused into calc_flux
    (k_flux<true>); the state and dt must equal evolve; set_time; vardt bit for bit, also when calls that use the
    work buffers or move the radius come in between."""
    p, prim = (pc.make_case_2d(*shape, **kw) if len(shape) == 2 else pc.make_case(*shape, **kw))
    out = []
    for flag in ("0", "1"):
        monkeypatch.setenv("LAPS_TUNE_SPEC", flag)
        o, g = pc.run_both(p, prim, 2, lib_path=emu)
        g.fft_forward(np.ones((1,) + g.real_shape))     # clobbers the work buffers: the front half must be redone
        g.step(); o.step()
        g.evolve_radius(g.time)                          # same radius again: still consistent
        g.step(); o.step()
        g.get_output(); g.calc_rms(); g.calc_max_divB()  # read-only calls keep the speculative front valid
        g.step(); o.step()
        pc.check_state(o, g, 1e-11)
        out.append((g.get_state()[0], g.uu_fourier(), g.dt))
        g.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and out[0][2] == out[1][2]


def test_get_output_is_the_array_output_uu_writes(emu):
    p, prim = pc.make_case(16, 16, 16, hall=True, aeb=True)
    o, g = pc.run_both(p, prim, 1, lib_path=emu)
    uu, pr = g.get_state()
    out = g.get_output(True)
    assert np.array_equal(out[0], uu[0]) and np.array_equal(out[1:4], pr[0:3]) and np.array_equal(out[4:7], uu[4:7])
    assert np.array_equal(out[7], pr[3]) and np.array_equal(g.get_output(False), uu)
    assert pc.rel_l2(out, lo.primitive_of(o)) < 1e-11
    g.close()


@pytest.mark.parametrize("kw", [dict(hall=True, aeb=True, dealias=3), dict(hall=False, aeb=False, dealias=2, explicit=True, conserve_bg=True, limit_dt=True),
                                dict(hall=True, aeb=False, dealias=0),
                                # if_corotating (src_incompressible/2D/mhdrhs.f90:196-201,318-323,402-407,593-598)
                                dict(hall=True, aeb=True, corot=True, dealias=1),
                                dict(hall=True, aeb=True, corot=True, dealias=2, explicit=True, conserve_bg=True)])
def test_one_step_incompressible_2d_tree(emu, kw):
    """src_incompressible/2D: kz = 0, the line axis carries ky in the projection, gradient and divergence tasks."""
    p, prim = pc.make_case_incompressible_2d(32, 16, **kw)
    o, g = pc.run_both(p, prim, 2, lib_path=emu, t0=2.0 if kw.get("corot") else 0.0)
    pc.check_state(o, g, 1e-11)
    pc.check_diagnostics(o, g, 1e-9)
    assert abs(g.calc_max_divV() - o.calc_max_divV()) <= 1e-9 * o.calc_max_divV()
    db, dv = g.calc_max_div_real()
    odb, odv = o.calc_max_div_real()
    assert abs(dv - odv) <= 1e-9 * odv and abs(db - odb) <= 1e-9 * max(odb, 1e-6)
    g.close()


def test_mask_pruning_is_bit_exact(emu):
    counts = pc.check_pruning_is_exact((32, 32, 32), 2, lib_path=emu, hall=True, aeb=True, dealias=1)
    # 32^3, spherical mask: 189 of the 231 columns of the surviving rectangle lie inside the circle,
    # 2699 of 17 * 32 * 32 modes inside the sphere
    assert counts[0] == (17 * 32, 17 * 32 * 32) and counts[1] == (189, 2699) and counts[2] == (231, 2699)
    pc.check_pruning_is_exact((32, 32), 2, lib_path=emu, hall=True, aeb=True, dealias=3)
    pc.check_pruning_is_exact((32, 16, 32), 2, lib_path=emu, incompressible=True, hall=True, aeb=True, dealias=1)


def test_set_primitive_modes_equals_the_uploaded_field(emu):
    """laps_set_primitive_modes (sparse spectrum + inverse transform on the device) against laps_set_primitive of
    the host-built field, 3D and the 2D tree; the oracle's point-by-point cosine sum pins both."""
    ks, coefs = synthetic.mode_table(24.0, 20.0, 12.0, 1.0, 0.3, 0.0, 3, 3, 3, (101, 116, 132), 0.1, 0.1, 0.01)
    back = [1.0, 0.0, 0.0, 0.0, 1.0, 0.3, 0.0, 1.0]
    p = lo.Params(nx=32, ny=16, nz=16, Lx=24.0, Ly=20.0, Lz=12.0)
    ref = lo.ic_turbulence(p, lo.ic_uniform_background(p, bx0=1.0, by0=0.3, press0=1.0), 1.0, 0.3, 0.0, nmodex=3, nmodey=3, nmodez=3)
    kw = dict(nx=32, ny=16, nz=16, Lx=24.0, Ly=20.0, Lz=12.0, dealias_option=1)
    with Solver(emu, **kw) as a, Solver(emu, **kw) as b:
        a.set_primitive(ref)
        b.set_primitive_modes(ks, coefs, back)
        assert np.abs(a.get_state()[0] - b.get_state()[0]).max() < 1e-13
        assert np.abs(a.uu_fourier() - b.uu_fourier()).max() < 1e-14
    # 2D tree: the modes with kz = 0
    sel = ks[:, 2] == 0
    with Solver(emu, nx=32, ny=16, nz=1, Lx=24.0, Ly=20.0, Lz=1.0, ndim=2, dealias_option=1) as c:
        c.set_primitive_modes(ks[sel], coefs[:, sel], back)
        uu, prim = c.get_state()
        x = np.arange(32) * (24.0 / 32)
        y = np.arange(16) * (20.0 / 16)
        Y, X = np.meshgrid(y, x, indexing="ij")
        bx = 1.0 + sum((coefs[4, m] * np.exp(1j * 2 * np.pi * (ks[m, 0] * X / 24.0 + ks[m, 1] * Y / 20.0))).real for m in np.nonzero(sel)[0])
        assert np.abs(uu[4, 0] - bx).max() < 1e-13
    with Solver(emu, **kw) as d:
        with pytest.raises(Exception, match="outside the grid"):
            d.set_primitive_modes(np.array([[17, 0, 0]]), np.zeros((7, 1), dtype=complex), back)
    # a mode on the Nyquist column kx = nx/2 (the stand-in driver's default nmodex = 8 on a 16-point grid): the
    # reference's cosine sampled there, Re(c exp(i(ky y + kz z))) (-1)^ix, at FULL amplitude (the c2r pass does not
    # double that column)
    with Solver(emu, nx=16, ny=16, nz=16, Lx=1.0, Ly=1.0, Lz=1.0, dealias_option=0) as e:
        c = np.zeros((7, 1), dtype=complex)
        c[4, 0] = 0.2 * np.exp(0.7j)
        e.set_primitive_modes(np.array([[8, 1, 0]]), c, [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 1.0])
        uu, _ = e.get_state()
        j = np.arange(16)
        want = 1.0 + (0.2 * np.exp(0.7j) * np.exp(2j * np.pi * j[:, None] / 16)).real * ((-1.0) ** j)[None, :]
        assert np.abs(uu[4, 3] - want).max() < 1e-14


@pytest.mark.parametrize("inc", [False, True])
def test_hall_wave_known_answer_on_the_kernels(emu, inc):
    pc.check_hall_wave_known_answer(emu, incompressible=inc, nsteps=(15, 30))


@pytest.mark.parametrize("shape", [(16, 16, 1024), (1024, 16, 16)])
def test_long_lines_through_every_pass(emu, shape):
    """1024-point lines (four radix stages, the line length of BASELINE config 5) through the RHS z pass and the
    x passes, not only through the unit transforms."""
    p, prim = pc.make_case(*shape, hall=True, aeb=True, dealias=1)
    o, g = pc.run_both(p, prim, 1, lib_path=emu)
    pc.check_state(o, g, 1e-11)
    g.close()


# Line lengths with an odd factor (3 * 2^k, 5 * 2^k; FFTW plans any length, fftw.f90:27-33): the composite transform of
# fft_core.cuh (P power-of-two transforms side by side + one radix-P exchange) in every pass of every tree.
@pytest.mark.parametrize("shape", [(48, 16, 16), (16, 80, 16), (16, 16, 96), (160, 48, 80)])
def test_fft_odd_factor_lines(emu, shape):
    pc.check_fft(*shape, lib_path=emu)


@pytest.mark.parametrize("shape,kw", [((48, 16, 16), dict(hall=True, aeb=True, dealias=1)),
                                      ((16, 48, 16), dict(hall=False, aeb=True, corot=True, dealias=2)),
                                      ((16, 16, 80), dict(hall=True, aeb=False, explicit=True, conserve_bg=True))])
def test_one_step_odd_factor_lines(emu, shape, kw):
    p, prim = pc.make_case(*shape, **kw)
    o, g = pc.run_both(p, prim, 1, lib_path=emu)
    pc.check_state(o, g, 1e-11)
    pc.check_diagnostics(o, g, 1e-9)
    g.close()


def test_odd_factor_lines_other_trees(emu):
    for shape, kw in [((80, 48), dict(hall=True, aeb=True, corot=True, dealias=2, explicit=True, conserve_bg=True)),
                      ((48, 80), dict(hall=True, aeb=True, z_radial=True, dealias=3))]:
        p, prim = pc.make_case_2d(*shape, **kw)
        o, g = pc.run_both(p, prim, 2, lib_path=emu, t0=2.0 if kw.get("corot") else 0.0)
        pc.check_state(o, g, 1e-11)
        pc.check_diagnostics(o, g, 1e-9)
        g.close()
    p, prim = pc.make_case_incompressible(16, 16, 48, hall=True, aeb=True, dealias=1)
    o, g = pc.run_both(p, prim, 1, lib_path=emu)
    pc.check_state(o, g, 1e-11)
    g.close()
    p, prim = pc.make_case_incompressible_2d(48, 80, hall=True, aeb=True, dealias=3)
    o, g = pc.run_both(p, prim, 2, lib_path=emu)
    pc.check_state(o, g, 1e-11)
    g.close()


def test_random_cases_fixed_seed(emu):
    """Ten cases of tools/fuzz_parity.py (random tree, grid with power-of-two / odd-factor / 8-point lines, switches, 1 - 2
    steps) with a fixed seed; the long sweeps are run by hand."""
    import random
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_parity as fz
    rng = random.Random(2)
    for _ in range(10):
        case = fz.random_case(rng, max_points=32 * 32 * 32)
        fz.run_case(case, emu)


def test_synthetic_slab_matches_the_mode_sum():
    p = lo.Params(nx=16, ny=32, nz=24, Lx=24.0, Ly=20.0, Lz=12.0)
    prim = lo.ic_uniform_background(p, bx0=1.0, by0=0.3, press0=1.0)
    prim = lo.ic_turbulence(p, prim, 1.0, 0.3, 0.0, nmodex=3, nmodey=3, nmodez=3)
    a = synthetic.turbulence_slab(16, 32, 24, 24.0, 20.0, 12.0, bx0=1.0, by0=0.3, kmax=3)
    assert np.abs(a - prim).max() < 1e-13
    b = synthetic.turbulence_slab(16, 32, 24, 24.0, 20.0, 12.0, z_offset=5, z_size=7, bx0=1.0, by0=0.3, kmax=3)
    assert np.abs(b - prim[:, 5:12]).max() < 1e-13


def test_cfl_screen_is_exact(emu):
    pc.check_cfl_screen_is_exact(emu, shape=(16, 16, 16), nsteps=2)


def test_rhs_kernel_variants(emu):
    pc.check_rhs_kernel_variants(emu, shape=(16, 16, 32), nsteps=2, exact=True, hall=True, aeb=True, dealias=1)
    pc.check_rhs_kernel_variants(emu, shape=(16, 16, 16), nsteps=1, exact=True, hall=True, aeb=True, corot=True, dealias=2, explicit=True)


def test_async_output(emu):
    pc.check_async_output(emu, shape=(16, 16, 16))
