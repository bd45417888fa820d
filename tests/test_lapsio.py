"""File formats either side of the hot path (laps_b200/lapsio.py) and the stand-in driver
(laps_b200/driver.py, the ``program mhd`` loop of mhd.f90:16-293).

The golden files under tests/golden/io were read back by the reference's own post-processing reader
when they were generated (tests/golden/make_io_fixtures.py, run in the build container where
/root/reference exists); expected.npz holds what that reader returned."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.join(HERE, "emu"))

import make_io_fixtures as fx  # noqa: E402
import parity_common as pc  # noqa: E402
from laps_b200 import lapsio  # noqa: E402
from laps_b200.driver import Driver, params_from_namelists  # noqa: E402
from oracle import laps_oracle as lo  # noqa: E402

GOLD = os.path.join(HERE, "golden", "io")


def test_writers_reproduce_the_golden_files_byte_for_byte(tmp_path):
    fx.write_all(str(tmp_path))
    for name in ("grid.dat", "parallel_info.dat", "out003.dat", "rms.dat", "EBM_info.dat"):
        with open(os.path.join(GOLD, "output", name), "rb") as a, open(tmp_path / name, "rb") as b:
            assert a.read() == b.read(), name


def test_readers_agree_with_what_the_reference_reader_returned():
    e = np.load(os.path.join(GOLD, "expected.npz"))
    out = os.path.join(GOLD, "output")
    assert lapsio.read_parallel_info(os.path.join(out, "parallel_info.dat")) == (int(e["npe"]), int(e["iproc"]), int(e["jproc"]), int(e["nvar"]))
    xg, yg, zg = lapsio.read_grid(os.path.join(out, "grid.dat"))
    assert np.array_equal(xg, e["xgrid"]) and np.array_equal(yg, e["ygrid"]) and np.array_equal(zg, e["zgrid"])
    nx, ny, nz = len(xg), len(yg), len(zg)
    f = os.path.join(out, "out003.dat")
    assert lapsio.read_out_header(f) == float(e["t"])
    full = lapsio.read_out_slab(f, nx, ny, nz)
    assert np.array_equal(full.transpose(3, 2, 1, 0), e["uu"])          # reference layout uu[ix,iy,iz,ivar]
    slab = lapsio.read_out_slab(f, nx, ny, nz, z_offset=2, z_size=3)     # a rank's block, as read_restart reads it
    assert np.array_equal(slab, full[:, 2:5])
    assert np.array_equal(full[:, 3, 1, 2], e["loc"])
    rms = np.loadtxt(os.path.join(out, "rms.dat"))
    assert np.array_equal(rms, e["rms"])


def test_2d_tree_files_against_the_reference_2d_reader(tmp_path):
    """2D/mhdoutput.f90:45-63: grid.dat = (nx, ny) + two grids, parallel_info.dat = (npe, nvar).  The golden files
    were read back by data_process/2D_Python/read_output.py when they were generated."""
    gold = os.path.join(HERE, "golden", "io2d")
    fx.write_all_2d(str(tmp_path))
    for name in ("grid.dat", "parallel_info.dat", "out007.dat", "EBM_info.dat"):
        with open(os.path.join(gold, "output", name), "rb") as a, open(tmp_path / name, "rb") as b:
            assert a.read() == b.read(), name
    e = np.load(os.path.join(gold, "expected.npz"))
    out = os.path.join(gold, "output")
    assert lapsio.read_parallel_info(os.path.join(out, "parallel_info.dat")) == (int(e["npe"]), int(e["nvar"]))
    xg, yg = lapsio.read_grid(os.path.join(out, "grid.dat"))
    assert np.array_equal(xg, e["xgrid"]) and np.array_equal(yg, e["ygrid"])
    f = os.path.join(out, "out007.dat")
    assert lapsio.read_out_header(f) == float(e["t"])
    full = lapsio.read_out_slab(f, len(xg), len(yg), 1)
    assert np.array_equal(full[:, 0].transpose(2, 1, 0), e["uu"])        # the 2D reader's layout uu[ix,iy,ivar]


def test_fortran_edit_descriptors():
    assert lapsio.fmt_1pe16_8(1.0) == "  1.00000000E+00"
    assert lapsio.fmt_1pe16_8(-3.75e5) == " -3.75000000E+05"
    assert lapsio.fmt_1pe16_8(0.0) == "  0.00000000E+00"
    assert lapsio.fmt_1pe16_8(1e-110) == "  1.00000000-110"               # three-digit exponents drop the E
    assert lapsio.rms_line(0.05, [1.0] * 8, [0.0] * 8, [2.0] * 3).startswith("    0.050000    1.00000000E+00")
    assert len(lapsio.rms_line(0.0, [0] * 8, [0] * 8, [0] * 3)) == 12 + 2 + 19 * 16
    assert [lapsio.out_name(i) for i in (0, 7, 42, 999)] == ["out000.dat", "out007.dat", "out042.dat", "out999.dat"]


INPUT = """&genr
   tmax = 100.0
   dtout = 0.5
   dtrms = 0.2   ! comment
/
&numerical
   cfl = 0.5
   dealias_option = 1
/
&prl
   ndim_parallel = 1
/
&grid
   nx = 16
   ny = 16
   nz = 16
   Lx = 24.0
   Ly = 24.0
   Lz = 24.0
/
&field
   ifield = 3
   Bx0 = 1.
   By0 = 0.
   Bz0 = 0.
   press0 = 1.0
/
&pert
   ipert = 7
   db0 = 0.1, dv0 = 0.1
   drho0 = 1d-2
   nmodex = 2
/
&phys
   adiabatic_index = 1.666667
   if_resis = T
   resistivity = 1e-4
   if_visc = .true.
   viscosity = 1e-4
/
&AEB
   if_AEB = T
   radius0 = 30.0
   Ur0 = 1.167
/
&Hall
   if_Hall = T 
   ion_inertial_length = 0.2
/
"""


def test_namelist_parser_reads_the_shipped_input_syntax(tmp_path):
    p = tmp_path / "mhd.input"
    p.write_text(INPUT)
    nl = lapsio.read_namelists(str(p))
    assert nl["genr"] == {"tmax": 100.0, "dtout": 0.5, "dtrms": 0.2}
    assert nl["pert"] == {"ipert": 7, "db0": 0.1, "dv0": 0.1, "drho0": 0.01, "nmodex": 2}
    assert nl["phys"]["if_resis"] is True and nl["phys"]["if_visc"] is True and nl["hall"]["if_hall"] is True
    kw = params_from_namelists(nl)
    assert kw["nx"] == 16 and kw["if_AEB"] is True and kw["Ur0"] == 1.167 and kw["dealias_option"] == 1
    assert kw["afx"] == 0.495 and kw["if_resis_exp"] is False          # module defaults (dealiasing.f90:10)


@pytest.fixture(scope="module")
def emu():
    import build_emu
    return build_emu.build()


def test_driver_loop_files_and_restart_on_the_emulator(emu, tmp_path):
    """The Principal loop against the oracle driven the same way, the files it leaves, and a restart from
    one of its own outNNN.dat files (restart.f90:17-63)."""
    (tmp_path / "mhd.input").write_text(INPUT)
    d = Driver(str(tmp_path / "mhd.input"), str(tmp_path), lib_path=emu)
    prim0 = d.initial_primitive()
    nsteps = d.run(max_steps=3, echo=False)
    assert nsteps == 3
    # oracle, same sequencing (mhd.f90:244-248,285)
    p = lo.Params(**{k: v for k, v in d.kw.items() if k not in ("rank", "nranks", "device")})
    o = lo.State(p)
    o.set_primitive(prim0)
    o.vardt()
    for i in range(3):
        o.evolve()
        o.time += o.dt
        o.evolve_radius(o.time)
        if i < 2:
            o.vardt()
    assert abs(d.time - o.time) < 1e-12
    # dtout = 0.5 < 3 steps of ~0.8: out000 (t=0), out001 (after step 1), out002, and the final one
    names = sorted(f for f in os.listdir(tmp_path) if f.startswith("out"))
    assert names[0] == "out000.dat" and len(names) >= 2
    last = str(tmp_path / names[-1])
    assert abs(lapsio.read_out_header(last) - np.float32(o.time)) < 1e-6
    data = lapsio.read_out_slab(last, 16, 16, 16)
    ref = lo.primitive_of(o)
    for v in range(8):
        assert pc.rel_l2(data[v], ref[v]) < 1e-10, v
    first = lapsio.read_out_slab(str(tmp_path / "out000.dat"), 16, 16, 16)
    assert pc.rel_l2(first, prim0) < 1e-13                                 # primitives round-trip through uu_prim
    rms = np.loadtxt(tmp_path / "rms.dat")
    ebm = np.loadtxt(tmp_path / "EBM_info.dat")
    assert rms.shape[1] == 20 and ebm.shape[1] == 3 and rms.shape[0] == ebm.shape[0] >= 2
    assert abs(ebm[-1, 1] - (30.0 + 1.167 * d.time)) < 1e-6 and ebm[0, 2] == 1.167
    assert os.path.exists(tmp_path / "grid.dat") and os.path.exists(tmp_path / "log")
    assert "Iterations     :       3" in (tmp_path / "log").read_text()
    d.solver.close()
    # restart from the last file
    n = int(names[-1][3:6])
    (tmp_path / "mhd.input").write_text(INPUT.replace("dtrms = 0.2", "dtrms = 0.2\n   if_restart = T\n   n_start = %d" % n))
    r = Driver(str(tmp_path / "mhd.input"), str(tmp_path), lib_path=emu)
    assert r.if_restart and r.n_start == n
    r.run(max_steps=1, echo=False)
    assert r.time > o.time and r.istep == 1
    assert os.path.exists(tmp_path / lapsio.out_name(n + 1))
    r.solver.close()


@pytest.mark.parametrize("tree", ["incompressible", "compressible2d", "incompressible2d"])
def test_driver_runs_the_other_source_trees_on_the_emulator(emu, tmp_path, tree):
    """The same mhd.input syntax drives the other three source trees (--tree): Alfven-wave data (ipert = 1),
    two steps, files in place, state equal to the oracle class of that tree."""
    text = INPUT.replace("ipert = 7", "ipert = 1").replace("Bx0 = 1.", "Bx0 = 1.\n   wave_number_jet = 2")
    (tmp_path / "mhd.input").write_text(text)
    d = Driver(str(tmp_path / "mhd.input"), str(tmp_path), lib_path=emu, tree=tree)
    prim0 = d.initial_primitive()
    assert prim0.shape == (8, 1 if tree.endswith("2d") else 16, 16, 16)
    assert d.run(max_steps=2, echo=False) == 2
    kw = {k: v for k, v in d.kw.items() if k not in ("rank", "nranks", "device", "ndim", "incompressible")}
    p = lo.Params(incompressible=tree.startswith("incompressible"), **kw)
    o = pc.oracle_state(p)
    o.set_primitive(prim0)
    o.vardt()
    for i in range(2):
        o.evolve()
        o.time += o.dt
        o.evolve_radius(o.time)
        if i == 0 and not tree.endswith("2d"):       # the 2D drivers call vardt every 20 steps only
            o.vardt()
    assert abs(d.time - o.time) < 1e-12
    if tree.endswith("2d"):      # 2D/mhdoutput.f90:45-63
        assert len(lapsio.read_grid(str(tmp_path / "grid.dat"))) == 2
        assert lapsio.read_parallel_info(str(tmp_path / "parallel_info.dat")) == (1, 8)
    else:
        assert len(lapsio.read_grid(str(tmp_path / "grid.dat"))) == 3
        assert lapsio.read_parallel_info(str(tmp_path / "parallel_info.dat")) == (1, 1, 1, 8)
    names = sorted(f for f in os.listdir(tmp_path) if f.startswith("out"))
    data = lapsio.read_out_slab(str(tmp_path / names[-1]), 16, 16, 1 if tree.endswith("2d") else 16)
    ref = o.uu.copy()
    ref[1:4] = o.uu_prim[0:3]
    if not tree.startswith("incompressible"):
        ref[7] = o.uu_prim[3]
    for v in range(8):
        assert pc.rel_l2(data[v], ref[v]) < 1e-10 or np.abs(data[v] - ref[v]).max() < 1e-13, v
    d.solver.close()


def test_driver_external_force_and_checknan_on_the_emulator(emu, tmp_path):
    """&pert if_external_force = T in the 2D compressible tree: the stand-in driver evaluates the shipped user routine
    (2D/mhdrhs.f90:480-531) once per step and hands the field to the library; same state as the oracle.  checkNan
    (2D/mhd.f90:242-252) runs at its cadence and does not stop a healthy run."""
    text = INPUT.replace("ipert = 7", "ipert = 1\n   if_external_force = T").replace("Bx0 = 1.", "Bx0 = 1.\n   wave_number_jet = 2")
    (tmp_path / "mhd.input").write_text(text)
    d = Driver(str(tmp_path / "mhd.input"), str(tmp_path), lib_path=emu, tree="compressible2d")
    assert d.kw["if_external_force"] is True
    d.dstep_checknan = 2
    prim0 = d.initial_primitive()
    assert d.run(max_steps=3, echo=False) == 3 and not d.stopped_on_nan
    kw = {k: v for k, v in d.kw.items() if k not in ("rank", "nranks", "device", "ndim", "incompressible")}
    o = lo.State2D(lo.Params(**kw))
    o.set_primitive(prim0)
    o.vardt()
    for _ in range(3):
        o.step(calc_dt=False)
    assert abs(d.time - o.time) < 1e-12
    uu, _ = d.solver.get_state()
    for v in range(8):
        assert pc.rel_l2(uu[v], o.uu[v]) < 1e-11 or np.abs(uu[v] - o.uu[v]).max() < 1e-13, v
    assert np.abs(o.uu[6]).max() > 1e-3          # the forcing has built up a B_z
    d.solver.close()


def test_driver_runs_the_shipped_2d_input_grid(emu, tmp_path):
    """&grid of the input the reference ships for the 2D compressible tree: nx = 256, ny = 8
    (src_compressible/2D/mhd.input:12-13) — an 8-point line axis and the half-height x-pass tile."""
    text = (INPUT.replace("ipert = 7", "ipert = 1").replace("Bx0 = 1.", "Bx0 = 1.\n   wave_number_jet = 2")
            .replace("nx = 16", "nx = 256").replace("ny = 16", "ny = 8"))
    assert "nx = 256" in text and "ny = 8" in text
    (tmp_path / "mhd.input").write_text(text)
    d = Driver(str(tmp_path / "mhd.input"), str(tmp_path), lib_path=emu, tree="compressible2d")
    prim0 = d.initial_primitive()
    assert prim0.shape == (8, 1, 8, 256)
    assert d.run(max_steps=2, echo=False) == 2
    kw = {k: v for k, v in d.kw.items() if k not in ("rank", "nranks", "device", "ndim", "incompressible")}
    o = lo.State2D(lo.Params(**kw))
    o.set_primitive(prim0)
    o.vardt()
    for _ in range(2):
        o.step(calc_dt=False)
    uu, _ = d.solver.get_state()
    for v in range(8):
        assert pc.rel_l2(uu[v], o.uu[v]) < 1e-11 or np.abs(uu[v] - o.uu[v]).max() < 1e-13, v
    d.solver.close()


@pytest.mark.parametrize("path,tree", [("src_compressible/mhd.input", "compressible"), ("src_compressible/2D/mhd.input", "compressible2d"),
                                       ("src_incompressible/mhd.input", "incompressible"), ("src_incompressible/2D/mhd.input", "incompressible2d")])
def test_the_four_shipped_inputs_are_accepted(emu, path, tree):
    """The mhd.input files the reference ships (read where they lie; skipped where /root/reference does not exist):
    the parser takes their syntax (one has a stray '/' line), and laps_create accepts their physics switches — on the
    shipped grid where the emulator can hold it (256 x 8), else on a 16-point grid."""
    full = os.path.join("/root/reference", path)
    if not os.path.exists(full):
        pytest.skip("/root/reference absent")
    nl = lapsio.read_namelists(full)
    kw = params_from_namelists(nl, 0, 1, 0, tree)
    shipped = {"compressible": (512, 512, 512), "compressible2d": (256, 8, 1), "incompressible": (512, 512, 512),
               "incompressible2d": (256, 1024, 1)}[tree]
    assert (kw["nx"], kw["ny"], kw["nz"]) == shipped
    assert kw["dealias_option"] == 1
    if tree != "compressible2d":
        kw.update(nx=16, ny=16, nz=1 if tree.endswith("2d") else 16)
    from laps_b200 import Solver
    with Solver(emu, **kw) as g:
        assert (g.nx, g.ny) == (kw["nx"], kw["ny"])
