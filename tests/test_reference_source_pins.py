"""Golden vectors produced by executing the reference's own Fortran source text (tests/golden/make_ref_exec_fixtures.py
through oracle/fortran_exec.py: every subroutine of the hot path from src_compressible/*.f90, only FFTW's 1-D
executions and one-rank mpi_allreduce supplied from outside) pin
  * the oracle restatement (oracle/laps_oracle.py), stage pieces and whole steps, here on the CPU;
  * the library on the kernel emulator here, and on the GPU in tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))
import parity_common as pc  # noqa: E402
from oracle import laps_oracle as lo  # noqa: E402

GOLD = os.path.join(HERE, "golden", "ref_exec")
# the `lines48` / `lines80` cases have line lengths with an odd factor (3 * 16, 5 * 16): FFTW plans any length (fftw.f90:27-33)
CASES = ["hall_aeb_mask", "corot_filter_explicit", "plain_nodealias", "lines48_hall_aeb_corot_mask"]
CASES_INCOMPRESSIBLE = ["incomp_hall_aeb_mask", "incomp_corot_filter_explicit", "incomp_plain_nodealias", "incomp_lines48_hall_aeb_mask"]
CASES_INCOMPRESSIBLE_2D = ["i2d_hall_aeb_mask", "i2d_square_explicit_limit", "i2d_corotating", "i2d_lines48x80_hall_aeb_mask"]
CASES_2D = ["c2d_hall_aeb_mask", "c2d_zradial_square_explicit", "c2d_external_force_filter", "c2d_lines80x48_hall_aeb_filter"]


def load_case(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    sw = {str(k): float(v) for k, v in zip(g["switch_names"], g["switches"])}
    exp = bool(sw["if_resis_exp"])
    p = lo.Params(nx=int(sw["nx"]), ny=int(sw["ny"]), nz=int(sw["nz"]), Lx=24.0, Ly=12.0, Lz=6.0, adiabatic_index=1.666667,
                  if_resis=bool(sw["if_resis"]), if_resis_exp=exp, resistivity=1e-3 if exp else 1e-4,
                  if_visc=bool(sw["if_visc"]), if_visc_exp=bool(sw["if_visc_exp"]), viscosity=1e-3 if sw["if_visc_exp"] else 1e-4,
                  if_conserve_background=bool(sw["if_conserve_background"]), cfl=0.5, dealias_option=int(sw["dealias_option"]),
                  if_AEB=bool(sw["if_aeb"]), radius0=30.0, Ur0=1.167, if_corotating=bool(sw["if_corotating"]),
                  corotating_angle=0.3 if sw["if_corotating"] else 0.0, if_hall=bool(sw["if_hall"]), ion_inertial_length=0.2)
    if name.startswith("incomp") or name.startswith("i2d"):
        p.incompressible = True
        p.rho0 = 1.0
    if p.nz == 1:          # the 2D trees (2D/mhd.f90:23,43,44)
        p.Lz = 1.0
        p.if_z_radial = bool(sw["if_z_radial"])
        p.if_limit_dt_increase = bool(sw["if_limit_dt_increase"])
        p.if_external_force = bool(sw["if_external_force"])
    return g, p


@pytest.mark.parametrize("name", CASES)
def test_oracle_agrees_with_the_executed_reference_source(name):
    g, p = load_case(name)
    o = lo.State(p)
    # grid_initialize (mhdinit.f90:58-124): wave numbers with the Nyquist kept positive, k_square
    assert np.array_equal(g["wave_numbers"], np.concatenate([o.g.wnx, o.g.wny, o.g.wnz]))
    if "k_square0" in g.files:
        assert np.allclose(g["k_square0"], np.broadcast_to(o.k_square, g["k_square0"].shape), rtol=1e-15, atol=0)
    # initial_calc_conserve_variable + transform_uu_real_to_fourier (mhd.f90:121-122)
    o.set_primitive(g["prim0"])
    if "uu_fourier0" in g.files:
        assert pc.rel_l2(o.uu_fourier, g["uu_fourier0"]) < 1e-14
    # vardt (mhd.f90:328-429)
    o.vardt()
    assert abs(o.dt - float(g["dt0"])) <= 1e-14 * o.dt
    # the pieces of the first stage: calc_current_density_real, calc_flux, transforms, calc_rhs (first case only)
    if "flux_stage1" in g.files:
        flux, expand = o.calc_flux()
        assert pc.rel_l2(flux, g["flux_stage1"]) < 1e-14
        assert pc.rel_l2(o.current_density, g["current_density_stage1"]) < 1e-13
        assert pc.rel_l2(expand, g["expand_stage1"][0]) < 1e-14
        fnl = o.calc_rhs(lo.fft_forward(flux), lo.fft_forward(expand))
        for v in range(8):
            assert pc.rel_l2(fnl[v], g["fnl_stage1"][v]) < 1e-13, v
    # two whole steps of the Principal loop (mhd.f90:244-248,285)
    for i in range(len(g["dt"])):
        o.step()
        assert abs(o.dt - g["dt"][i]) <= 1e-13 * o.dt and abs(o.time - g["time"][i]) <= 1e-14 * o.time
        assert abs(o.radius - g["radius"][i]) <= 1e-15 * o.radius
    for v in range(8):
        assert pc.rel_l2(o.uu[v], g["uu"][v]) < 1e-13, (v, pc.rel_l2(o.uu[v], g["uu"][v]))
        assert pc.rel_l2(o.uu_fourier[v], g["uu_fourier"][v]) < 1e-13, v
    for v in range(4):
        assert pc.rel_l2(o.uu_prim[v], g["uu_prim"][v]) < 1e-12, v
    if "k_square" in g.files:
        assert np.allclose(np.broadcast_to(o.k_square, g["k_square"].shape), g["k_square"], rtol=1e-14, atol=0)   # update_ksquare
    # diagnostics: calc_max_divB (mhd.f90:522-570), calc_rms (mhdrms.f90:53-126)
    assert abs(o.calc_max_divB() - float(g["max_divb"])) <= max(1e-9 * float(g["max_divb"]), 1e-14)
    ave, rms, ru2 = o.calc_rms()
    assert np.allclose(ave, g["uu_ave"], rtol=1e-13, atol=1e-14)      # means that are zero up to summation round-off
    # <u^2> - <u>^2 cancels five digits for B_x (mean 0.88, variance 1e-5): the sequential Fortran sum and NumPy's
    # pairwise sum differ at 1e-13 before the subtraction
    assert np.allclose(rms, g["uu_rms"], rtol=1e-7, atol=1e-17)
    assert np.allclose(ru2, g["rho_u2"], rtol=1e-12, atol=1e-20)


@pytest.mark.parametrize("name", CASES_INCOMPRESSIBLE)
def test_incompressible_oracle_agrees_with_the_executed_reference_source(name):
    """src_incompressible: J and grad u, the flux for the pressure, the pressure projection, E, calc_rhs, then two whole
    steps incl. update_rho_p (rho0 compounds with the radius of the evolve just done)."""
    g, p = load_case(name)
    o = lo.StateIncompressible(p)
    o.set_primitive(g["prim0"])
    if "uu_fourier0" in g.files:
        assert pc.rel_l2(o.uu_fourier, g["uu_fourier0"]) < 1e-14
    o.vardt()
    assert abs(o.dt - float(g["dt0"])) <= 1e-14 * o.dt
    if "fnl_stage1" in g.files:
        # first stage, piece by piece (src_incompressible/mhd.f90:323-350)
        o.uu_fourier = lo.fft_forward(o.uu)
        o.calc_current_density_real()
        o.calc_gradient_velocity_real()
        assert pc.rel_l2(o.current_density, g["current_density_stage1"]) < 1e-13
        assert pc.rel_l2(o.grad_velocity, g["grad_velocity_stage1"]) < 1e-13
        fp = o.calc_flux_for_pressure()
        assert pc.rel_l2(fp, g["flux_pressure_stage1"]) < 1e-13
        fpf = lo.fft_forward(fp)
        o.calc_pressure_fourier(fpf)
        assert pc.rel_l2(o.uu_fourier[7], g["pressure_fourier_stage1"]) < 1e-12
        flux = o.calc_flux()
        assert pc.rel_l2(flux, g["flux_stage1"]) < 1e-13
        fnl = o.calc_rhs(lo.fft_forward(flux), fpf)
        for v in range(8):
            ref = g["fnl_stage1"][v]
            assert pc.rel_l2(fnl[v], ref) < 1e-12 or np.abs(fnl[v] - ref).max() < 1e-15, v
    # whole steps
    o = lo.StateIncompressible(p)
    o.set_primitive(g["prim0"])
    o.vardt()
    for i in range(len(g["dt"])):
        o.step()
        assert abs(o.dt - g["dt"][i]) <= 1e-13 * o.dt and abs(o.time - g["time"][i]) <= 1e-14 * o.time
        assert abs(o.rho0 - g["rho0"][i]) <= 1e-15
    for v in range(8):
        assert pc.rel_l2(o.uu[v], g["uu"][v]) < 1e-12, (v, pc.rel_l2(o.uu[v], g["uu"][v]))
        assert pc.rel_l2(o.uu_fourier[v], g["uu_fourier"][v]) < 1e-12, v
    for v in range(3):
        assert pc.rel_l2(o.uu_prim[v], g["uu_prim"][v]) < 1e-12, v
    assert abs(o.calc_max_divV() - float(g["max_divv"])) <= 1e-9 * float(g["max_divv"])
    db, dv = o.calc_max_div_real()
    assert abs(dv - float(g["max_divv_real"])) <= 1e-9 * float(g["max_divv_real"])
    assert abs(db - float(g["max_divb_real"])) <= 1e-7 * max(float(g["max_divb_real"]), 1e-9)


def check_library_incompressible(name, lib_path=None, tol=1e-11):
    from laps_b200 import Solver
    g, p = load_case(name)
    with Solver(lib_path, **pc.solver_kwargs(p)) as s:
        s.set_primitive(g["prim0"])
        s.vardt()
        assert abs(s.dt - float(g["dt0"])) <= 1e-13 * s.dt
        for i in range(len(g["dt"])):
            s.step()
            assert abs(s.dt - g["dt"][i]) <= 1e-12 * s.dt
            assert abs(s.rho0 - g["rho0"][i]) <= 1e-15
        uu, prim = s.get_state()
        uf = s.uu_fourier()
        for v in range(8):
            assert pc.rel_l2(uu[v], g["uu"][v]) < tol, (v, pc.rel_l2(uu[v], g["uu"][v]))
            assert pc.rel_l2(uf[v], g["uu_fourier"][v]) < tol, v
        assert abs(s.calc_max_divV() - float(g["max_divv"])) <= 1e-7 * float(g["max_divv"])
        db, dv = s.calc_max_div_real()
        assert abs(dv - float(g["max_divv_real"])) <= 1e-7 * float(g["max_divv_real"])


@pytest.mark.parametrize("name", CASES_2D + ["c2d_corotating_oracle_only"])
def test_2d_oracle_agrees_with_the_executed_reference_source(name):
    """src_compressible/2D: kz = 0, if_z_radial, square truncation, if_limit_dt_increase, the external force."""
    g, p = load_case(name)
    o = lo.State2D(p)
    o.set_primitive(g["prim0"])
    if "uu_fourier0" in g.files:
        assert pc.rel_l2(o.uu_fourier, g["uu_fourier0"]) < 1e-14
    o.vardt()
    assert abs(o.dt - float(g["dt0"])) <= 1e-14 * o.dt
    if "flux_stage1" in g.files:
        flux, expand = o.calc_flux()
        # the 2D tree leaves flux(:,:,:,3:...) of the z direction formed as well: compare what both hold
        assert pc.rel_l2(flux, g["flux_stage1"]) < 1e-14
        assert pc.rel_l2(expand, g["expand_stage1"][0]) < 1e-14
        fnl = o.calc_rhs(lo.fft_forward(flux), lo.fft_forward(expand))
        if p.if_external_force:
            fnl[6] = fnl[6] + lo.fft_forward(o.calc_external_force_real())
        for v in range(8):
            assert pc.rel_l2(fnl[v], g["fnl_stage1"][v]) < 1e-13, v
    for i in range(len(g["dt"])):
        if p.if_external_force:     # calc_external_force_real (2D/mhdrhs.f90:480-531) at this step's time
            assert np.abs(o.calc_external_force_real() - g["external_force"][i]).max() < 1e-15
        o.step()
        assert abs(o.dt - g["dt"][i]) <= 1e-13 * o.dt and abs(o.time - g["time"][i]) <= 1e-14 * o.time
    for v in range(8):
        assert pc.rel_l2(o.uu[v], g["uu"][v]) < 1e-13, (v, pc.rel_l2(o.uu[v], g["uu"][v]))
        assert pc.rel_l2(o.uu_fourier[v], g["uu_fourier"][v]) < 1e-13, v
    if "k_square" in g.files:
        assert np.allclose(np.broadcast_to(o.k_square, g["k_square"].shape), g["k_square"], rtol=1e-14, atol=0)
    assert abs(o.calc_max_divB() - float(g["max_divb"])) <= max(1e-9 * float(g["max_divb"]), 1e-14)
    ave, rms, ru2 = o.calc_rms()
    assert np.allclose(ave, g["uu_ave"], rtol=1e-13, atol=1e-14)      # means that are zero up to summation round-off
    assert np.allclose(rms, g["uu_rms"], rtol=1e-7, atol=1e-17)
    assert np.allclose(ru2, g["rho_u2"], rtol=1e-12, atol=1e-20)
    assert int(g["isnanall"]) == 0


def check_library_2d(name, lib_path=None, tol=1e-11):
    from laps_b200 import Solver
    g, p = load_case(name)
    with Solver(lib_path, **pc.solver_kwargs(p)) as s:
        s.set_primitive(g["prim0"])
        s.vardt()
        assert abs(s.dt - float(g["dt0"])) <= 1e-13 * s.dt
        for i in range(len(g["dt"])):
            if p.if_external_force:     # the field the reference's user routine produced for this step
                s.set_external_force(g["external_force"][i])
            s.step()
            assert abs(s.dt - g["dt"][i]) <= 1e-12 * s.dt
        uu, prim = s.get_state()
        uf = s.uu_fourier()
        for v in range(8):
            assert pc.rel_l2(uu[v], g["uu"][v]) < tol, (v, pc.rel_l2(uu[v], g["uu"][v]))
            assert pc.rel_l2(uf[v], g["uu_fourier"][v]) < tol, v
        assert abs(s.calc_max_divB() - float(g["max_divb"])) <= 1e-9 * float(g["max_divb"]) + 1e-13
        assert s.checkNan() is bool(int(g["isnanall"]))


@pytest.mark.parametrize("name", CASES_INCOMPRESSIBLE_2D)
def test_incompressible_2d_oracle_agrees_with_the_executed_reference_source(name):
    g, p = load_case(name)
    o = lo.StateIncompressible2D(p)
    o.set_primitive(g["prim0"])
    if "uu_fourier0" in g.files:
        assert pc.rel_l2(o.uu_fourier, g["uu_fourier0"]) < 1e-14
    o.vardt()
    assert abs(o.dt - float(g["dt0"])) <= 1e-14 * o.dt
    for i in range(len(g["dt"])):
        o.step()
        assert abs(o.dt - g["dt"][i]) <= 1e-13 * o.dt and abs(o.time - g["time"][i]) <= 1e-14 * o.time
        assert abs(o.rho0 - g["rho0"][i]) <= 1e-15
    for v in range(8):
        ref = g["uu"][v]
        assert pc.rel_l2(o.uu[v], ref) < 1e-12 or np.abs(o.uu[v] - ref).max() < 1e-14, (v, pc.rel_l2(o.uu[v], ref))
        assert pc.rel_l2(o.uu_fourier[v], g["uu_fourier"][v]) < 1e-12 or np.abs(o.uu_fourier[v] - g["uu_fourier"][v]).max() < 1e-15, v
    assert abs(o.calc_max_divV() - float(g["max_divv"])) <= 1e-9 * float(g["max_divv"])
    db, dv = o.calc_max_div_real()
    assert abs(dv - float(g["max_divv_real"])) <= 1e-9 * float(g["max_divv_real"])
    assert abs(db - float(g["max_divb_real"])) <= 1e-7 * max(float(g["max_divb_real"]), 1e-9)


def check_library_incompressible_2d(name, lib_path=None, tol=1e-11):
    from laps_b200 import Solver
    g, p = load_case(name)
    with Solver(lib_path, **pc.solver_kwargs(p)) as s:
        s.set_primitive(g["prim0"])
        s.vardt()
        assert abs(s.dt - float(g["dt0"])) <= 1e-13 * s.dt
        for i in range(len(g["dt"])):
            s.step()
            assert abs(s.dt - g["dt"][i]) <= 1e-12 * s.dt
            assert abs(s.rho0 - g["rho0"][i]) <= 1e-15
        uu, prim = s.get_state()
        for v in range(8):
            ref = g["uu"][v]
            assert pc.rel_l2(uu[v], ref) < tol or np.abs(uu[v] - ref).max() < 1e-13, (v, pc.rel_l2(uu[v], ref))
        assert abs(s.calc_max_divV() - float(g["max_divv"])) <= 1e-7 * float(g["max_divv"])
        assert s.checkNan() is bool(int(g["isnanall"]))


@pytest.fixture(scope="module")
def emu():
    import build_emu
    return build_emu.build()


def check_library(name, lib_path=None, tol=1e-11):
    """The library, driven like mhd.f90, against the vectors of the executed reference source (north-star tolerances:
    fields 1e-11 relative L2, diagnostics 1e-9)."""
    from laps_b200 import Solver
    g, p = load_case(name)
    with Solver(lib_path, **pc.solver_kwargs(p)) as s:
        s.set_primitive(g["prim0"])
        if "uu_fourier0" in g.files:
            assert pc.rel_l2(s.uu_fourier(), g["uu_fourier0"]) < 1e-13
        s.vardt()
        assert abs(s.dt - float(g["dt0"])) <= 1e-13 * s.dt
        for i in range(len(g["dt"])):
            s.step()
            assert abs(s.dt - g["dt"][i]) <= 1e-12 * s.dt and abs(s.time - g["time"][i]) <= 1e-12 * s.time
        uu, prim = s.get_state()
        uf = s.uu_fourier()
        for v in range(8):
            assert pc.rel_l2(uu[v], g["uu"][v]) < tol, (v, pc.rel_l2(uu[v], g["uu"][v]))
            assert pc.rel_l2(uf[v], g["uu_fourier"][v]) < tol, v
        if p.dealias_option == 1:      # the truncation mask of dealias (dealiasing.f90:87-99), bit for bit: the same modes are zero
            assert np.array_equal(uf == 0, g["uu_fourier"] == 0)
            assert (g["uu_fourier"][0] == 0).mean() > 0.5
        for v in range(4):
            assert pc.rel_l2(prim[v], g["uu_prim"][v]) < 10 * tol, v
        ave, rms, ru2 = s.calc_rms()
        check_cancelling_diagnostics(p, s.calc_max_divB(), float(g["max_divb"]), rms, g["uu_rms"], g["uu_ave"])
        assert np.allclose(ave, g["uu_ave"], rtol=1e-9, atol=1e-12)
        assert np.allclose(ru2, g["rho_u2"], rtol=1e-9, atol=1e-18)


@pytest.mark.parametrize("name", CASES)
def test_library_on_the_emulator_agrees_with_the_executed_reference_source(emu, name):
    check_library(name, lib_path=emu)


@pytest.mark.parametrize("name", CASES_INCOMPRESSIBLE)
def test_incompressible_library_on_the_emulator_agrees_with_the_executed_reference_source(emu, name):
    check_library_incompressible(name, lib_path=emu)


@pytest.mark.parametrize("name", CASES_2D + ["c2d_corotating_oracle_only"])
def test_2d_library_on_the_emulator_agrees_with_the_executed_reference_source(emu, name):
    check_library_2d(name, lib_path=emu)


@pytest.mark.parametrize("name", CASES_INCOMPRESSIBLE_2D)
def test_incompressible_2d_library_on_the_emulator_agrees_with_the_executed_reference_source(emu, name):
    check_library_incompressible_2d(name, lib_path=emu)


# ------------------------------------------------------------------------------------------------------------------
# transpose index maps against the MPI subarray types of the executed parallel_start (parallel.f90:28-212)
# ------------------------------------------------------------------------------------------------------------------
def _par(shape, npe):
    g = np.load(os.path.join(GOLD, "parallel_start.npz"))
    nx, ny, nz = shape
    pre = f"{nx}x{ny}x{nz}_p{npe}_r"
    return [{k: g[f"{pre}{r}_{k}"] for k in ("yj_offset", "yj_size", "zj_offset", "zj_size", "xi_size", "yz_send", "yz_recv")}
            for r in range(npe)]


@pytest.mark.parametrize("shape,npe", [((16, 16, 16), 2), ((16, 16, 16), 3), ((16, 16, 16), 4), ((16, 32, 16), 8), ((32, 64, 32), 5),
                                       ((16, 48, 80), 5), ((48, 80, 48), 3)])
def test_transpose_index_maps_match_the_executed_parallel_start(emu, shape, npe):
    """transpose_yz sends, for every peer q, the MPI subarray yz_send(q) of w_yxz into the subarray yz_recv(me) of q's
    w_zxy, element by element in Fortran order (parallel.f90:185-210,273-297); transpose_zy is the reverse pairing
    (:300-324).  The types come from the reference's parallel_start executed for every rank
    (tests/golden/make_ref_exec_fixtures.py); the library (reference slabs: LAPS_TUNE_CYCLIC=0) must send every element to
    the same rank and the same local coordinates.  decompose_1d's tables (remainder on the last rank) are checked on the way."""
    from laps_b200 import Solver
    nx, ny, nz = shape
    ranks = _par(shape, npe)
    saved = os.environ.get("LAPS_TUNE_CYCLIC")
    os.environ["LAPS_TUNE_CYCLIC"] = "0"
    try:
        for me, d in enumerate(ranks):
            with Solver(emu, nx=nx, ny=ny, nz=nz, Lx=1.0, Ly=1.0, Lz=1.0, dealias_option=1, rank=me, nranks=npe) as s:
                X = int(d["xi_size"])
                assert (s.ext.z_offset, s.ext.z_size, s.ext.y_offset, s.ext.y_size) == (
                    d["zj_offset"][me], d["zj_size"][me], d["yj_offset"][me], d["yj_size"][me]) and X == s.nxh
                myz = s.transpose_yz_indexmap().reshape(s.nxh, ny, s.nzl, 2)
                mzy = s.transpose_zy_indexmap().reshape(s.nxh, s.nyl, nz, 2)
                for q, dq in enumerate(ranks):
                    # ---- transpose_yz: me -> q
                    full, sub, st = d["yz_send"][q]
                    rfull, rsub, rst = dq["yz_recv"][me]
                    assert list(sub) == list(rsub) and list(full) == [X, ny, d["zj_size"][me]] and list(rfull) == [X, dq["yj_size"][q], nz]
                    i, j, l = np.meshgrid(np.arange(sub[0]), np.arange(sub[1]), np.arange(sub[2]), indexing="ij")
                    ours = myz[st[0] + i, st[1] + j, st[2] + l]                 # source element, local coordinates in w_yxz
                    assert np.all(ours[..., 0] == q)
                    off, Yq = ours[..., 1], int(dq["yj_size"][q])
                    assert np.array_equal(off % nz, rst[2] + l)                 # destination z in q's w_zxy
                    assert np.array_equal((off // nz) % Yq, rst[1] + j)         # destination local y
                    assert np.array_equal(off // (nz * Yq), rst[0] + i)         # destination x
                    # ---- transpose_zy: me -> q sends subarray yz_recv(q) of w_zxy into subarray yz_send(me) of q's w_yxz
                    full, sub, st = d["yz_recv"][q]
                    rfull, rsub, rst = dq["yz_send"][me]
                    assert list(sub) == list(rsub)
                    i, j, l = np.meshgrid(np.arange(sub[0]), np.arange(sub[1]), np.arange(sub[2]), indexing="ij")
                    ours = mzy[st[0] + i, st[1] + j, st[2] + l]
                    assert np.all(ours[..., 0] == q)
                    off, Zq = ours[..., 1], int(dq["zj_size"][q])
                    assert np.array_equal(off % Zq, rst[2] + l)
                    assert np.array_equal((off // Zq) % ny, rst[1] + j)
                    assert np.array_equal(off // (Zq * ny), rst[0] + i)
    finally:
        if saved is None:
            os.environ.pop("LAPS_TUNE_CYCLIC", None)
        else:
            os.environ["LAPS_TUNE_CYCLIC"] = saved


def check_cancelling_diagnostics(p, divb, ref_divb, rms, ref_rms, ref_ave, what=None, tol=1e-9):
    """The two diagnostics that are differences of nearly equal numbers, held to the north star's 1e-9 RELATIVE
    tolerance plus an ABSOLUTE round-off allowance stated in the natural scale of the terms that cancel:
      max |k.B^|  (mhd.f90:541-568): |error| <= 1e-9 max|k.B^| + 1e-14 k_max B_rms     (k_x B^_x + k_y B^_y + k_z B^_z cancel)
      uu_rms = <u^2> - <u>^2 (mhdrms.f90:103-105): |error| <= 1e-9 uu_rms + 1e-13 <u^2>
    (a relative bound alone cannot hold where the result is round-off of the cancelling terms: without the expanding box
    max |k.B^| is ~1e-17 against terms of order one)."""
    ref_rms, ref_ave = np.asarray(ref_rms, dtype=float), np.asarray(ref_ave, dtype=float)
    msq = ref_rms + ref_ave ** 2
    kmax = np.pi * max(p.nx / p.Lx, p.ny / p.Ly, (p.nz / p.Lz) if p.nz > 1 else 0.0)
    brms = float(np.sqrt(msq[4:7].sum()))
    assert abs(divb - ref_divb) <= tol * ref_divb + 1e-14 * kmax * brms, (what, divb, ref_divb)
    assert np.all(np.abs(np.asarray(rms) - ref_rms) <= tol * np.abs(ref_rms) + 1e-13 * msq), (what, rms, ref_rms)


# ------------------------------------------------------------------------------------------------------------------
# 100 steps (north star: energy, cross helicity and div B within 1e-9 relative after 100 steps)
# ------------------------------------------------------------------------------------------------------------------
def check_100_steps(make_state, tol=1e-9, nsteps=100):
    """make_state(p, prim) -> object with vardt(), step(), dt, time, invariants(), calc_rms(); rows of the fixture every 10 steps:
    istep, time, dt, mean energy density, mean u.B, max |k.B^|, uu_ave(8), uu_rms(8), rho_u2(3)."""
    g, p = load_case("hall_aeb_mask_100steps")
    s = make_state(p, g["prim0"])
    s.vardt()
    rows = {int(r[0]): r for r in g["rows"]}
    for istep in range(1, nsteps + 1):
        s.step()
        if istep in rows:
            r = rows[istep]
            assert abs(s.time - r[1]) <= 1e-11 * r[1] and abs(s.dt - r[2]) <= 1e-10 * r[2], (istep, s.time, r[1], s.dt, r[2])
            inv = s.invariants()
            assert abs(inv[0] - r[3]) <= tol * abs(r[3]), (istep, inv[0], r[3])
            assert abs(inv[1] - r[4]) <= tol * max(abs(r[4]), 1e-6), (istep, inv[1], r[4])
            ave, rms, ru2 = s.calc_rms()
            check_cancelling_diagnostics(p, inv[2], r[5], rms, r[14:22], r[6:14], what=istep)
            assert np.allclose(ave, r[6:14], rtol=tol, atol=1e-13), istep
            assert np.allclose(ru2, r[22:25], rtol=1e-8, atol=1e-16), istep
    return s


def test_oracle_100_steps_against_the_executed_reference_source():
    def make(p, prim):
        o = lo.State(p)
        o.set_primitive(prim)
        return o
    o = check_100_steps(make)
    g, _ = load_case("hall_aeb_mask_100steps")
    for v in range(8):
        assert pc.rel_l2(o.uu[v], g["uu"][v]) < 1e-10, v


def check_library_100_steps(lib_path=None, nsteps=100):
    from laps_b200 import Solver
    holder = {}

    def make(p, prim):
        s = Solver(lib_path, **pc.solver_kwargs(p))
        s.set_primitive(prim)
        holder["s"] = s
        return s
    try:
        s = check_100_steps(make, nsteps=nsteps)
        g, _ = load_case("hall_aeb_mask_100steps")
        uu, _ = s.get_state()
        for v in range(8 if nsteps == 100 else 0):
            assert pc.rel_l2(uu[v], g["uu"][v]) < 1e-9, (v, pc.rel_l2(uu[v], g["uu"][v]))
    finally:
        holder["s"].close()


def test_library_40_steps_on_the_emulator_against_the_executed_reference_source(emu):
    """The emulator does the first 40 of the 100 steps (rows 10..40); the GPU test runs all of them."""
    check_library_100_steps(emu, nsteps=40)


def test_2d_corotation_where_the_reference_has_none(emu):
    """if_corotating runs in both 2D trees (fixtures c2d_corotating_oracle_only — the name dates from the round in which only the
    oracle had it — and i2d_corotating above); what stays refused is the combination with if_z_radial, at which the reference
    stops too (2D/mhd.f90:62-67)."""
    from laps_b200 import Solver, capi
    g, p = load_case("c2d_corotating_oracle_only")
    kw = pc.solver_kwargs(p)
    with pytest.raises(capi.LapsError, match="exclude each other"):
        Solver(emu, **dict(kw, if_z_radial=1))


# ------------------------------------------------------------------------------------------------------------------
# initial-condition hooks (the benchmark's synthetic turbulence, the stand-in driver's Alfven wave)
# ------------------------------------------------------------------------------------------------------------------
def test_initial_conditions_against_the_executed_reference_source():
    """background_fields_initialize case 3 + perturbation_initialize cases 7 and 1 (mhdinit.f90:183-260,300-342,695-823)
    executed from the reference's text, the three phase tables supplied by numpy.random.default_rng(ir + 100) in place of
    the compiler-specific generator.  The oracle's generators and the library-side ones (laps_b200.synthetic: bench.py's
    workload, the stand-in driver, the mode table of laps_set_primitive_modes) produce the same fields."""
    from laps_b200 import synthetic
    g = np.load(os.path.join(GOLD, "initial_conditions.npz"))
    nx, ny, nz, Lx, Ly, Lz, bx0, by0, bz0 = g["params"]
    nx, ny, nz = int(nx), int(ny), int(nz)
    p = lo.Params(nx=nx, ny=ny, nz=nz, Lx=Lx, Ly=Ly, Lz=Lz)
    # ipert = 7
    prim = lo.ic_uniform_background(p, bx0=bx0, by0=by0, bz0=bz0, press0=1.0)
    prim = lo.ic_turbulence(p, prim, bx0, by0, bz0, db0=0.1, dv0=0.1, drho0=0.01, nmodex=2, nmodey=2, nmodez=2, seeds=(101, 116, 132))
    for v in range(8):
        assert np.abs(prim[v] - g["ipert7"][v]).max() < 1e-14, v
    slab = synthetic.turbulence_slab(nx, ny, nz, Lx, Ly, Lz, bx0=bx0, by0=by0, bz0=bz0, db0=0.1, dv0=0.1, drho0=0.01, kmax=2)
    for v in range(8):
        assert np.abs(slab[v] - g["ipert7"][v]).max() < 1e-13, v
    part = synthetic.turbulence_slab(nx, ny, nz, Lx, Ly, Lz, z_offset=3, z_size=4, bx0=bx0, by0=by0, bz0=bz0, kmax=2)
    assert np.abs(part - g["ipert7"][:, 3:7]).max() < 1e-13                 # a rank's z slab
    # ipert = 1
    prim = lo.ic_uniform_background(p, bx0=bx0, by0=by0, bz0=bz0, press0=1.0)
    prim = lo.ic_alfven_wave(p, prim, db0=0.1, wave_number_jet=2, cor_angle=0.0)
    assert np.abs(prim - g["ipert1"]).max() < 1e-15
    prim2 = lo.ic_uniform_background(p, bx0=bx0, by0=by0, bz0=bz0, press0=1.0)
    synthetic.add_alfven_wave(prim2, nx, Lx, db0=0.1, wave_number_jet=2, cor_angle=0.0)
    assert np.abs(prim2 - g["ipert1"]).max() < 1e-15


def test_the_pins_are_live_a_changed_reference_source_changes_the_vectors(tmp_path):
    """Sanity of the method (needs /root/reference, skipped elsewhere): the golden vectors come from the Fortran text itself
    — re-running calc_rhs from a copy of mhdrhs.f90 in which one expanding-box coefficient is altered (3.0 -> 2.0 in the
    rho u_y row, mhdrhs.f90:239) moves fnl(3) away from the stored vector, while the unaltered text reproduces it bit for bit."""
    ref = "/root/reference/src_compressible/mhdrhs.f90"
    if not os.path.exists(ref):
        pytest.skip("/root/reference absent")
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_ref_exec_fixtures as m
    from oracle import fortran_exec as fx
    name = "corot_filter_explicit"
    g = np.load(os.path.join(GOLD, name + ".npz"))
    text = open(ref).read()
    target = "fnl(ix,iy,iz,3) = fnl(ix,iy,iz,3) - 3.0 * uu_fourier(ix,iy,iz,3) / tau_exp"
    assert text.count(target) == 1
    (tmp_path / "mhdrhs.f90").write_text(text.replace(target, target.replace("3.0", "2.0")))
    results = []
    for path in (ref, str(tmp_path / "mhdrhs.f90")):
        c = m.CASES[name]
        ns = m.build_namespace(c)
        m.load_reference(ns)
        fx.load(ns, path, ["calc_rhs"])                      # calc_rhs from this text
        st = ns["_storage"]
        ns["grid_initialize"]()
        ns["dealias_initialize"]()
        ns["aeb_calc"](ns["radius"])
        ns["cos_cor_ang"], ns["sin_cor_ang"] = float(np.cos(0.3)), float(np.sin(0.3))
        st["uu"][...] = g["prim0"]
        ns["initial_calc_conserve_variable"]()
        ns["transform_uu_real_to_fourier"]()
        ns["vardt"]()
        ns["calc_flux"]()
        ns["transform_flux_real_to_fourier"]()
        ns["calc_rhs"]()
        results.append(st["fnl"].copy())
    assert np.array_equal(results[0], g["fnl_stage1"])                                   # the reference's text: the stored vector
    assert pc.rel_l2(results[1][2], g["fnl_stage1"][2]) > 1e-3                            # the altered text: not
    assert np.array_equal(results[1][[0, 1, 3, 4, 5, 6, 7]], g["fnl_stage1"][[0, 1, 3, 4, 5, 6, 7]])
