"""The reference's own C reader of its output files (data_process/3D_C/iofunctions.c), compiled where it lies by
oracle/Makefile into oracle/_ref/ and EXECUTED here, reads what laps_b200.lapsio and the stand-in driver write:
grid.dat, parallel_info.dat, EBM_info.dat and outNNN.dat (mhdoutput.f90:51-131, AEBmod.f90:75-85).  This is the
one piece of reference code the toolchain of this image can build."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))
sys.path.insert(0, os.path.join(HERE, "golden"))

from laps_b200 import lapsio  # noqa: E402
from laps_b200.driver import Driver  # noqa: E402
from oracle import ref_io  # noqa: E402


@pytest.fixture(scope="module")
def ref():
    path = ref_io.build()
    if path is None:
        pytest.skip("oracle/_ref/libref_io.so not built and /root/reference absent")
    return ref_io.ReferenceReader(path)


def test_reference_reader_reads_lapsio_files(ref, tmp_path):
    import make_io_fixtures as mk
    nx, ny, nz, nranks = 6, 4, 5, 2
    mk.write_all(str(tmp_path), nx, ny, nz, nranks)
    gx, gy, gz, x, y, z = ref.read_grid(str(tmp_path / "grid.dat"))
    assert (gx, gy, gz) == (nx, ny, nz)
    assert np.array_equal(x, (np.arange(nx) * (24.0 / nx)).astype(np.float32))
    assert np.array_equal(y, (np.arange(ny) * (12.0 / ny)).astype(np.float32))
    assert np.array_equal(z, (np.arange(nz) * (6.0 / nz)).astype(np.float32))
    assert ref.read_parallel_info(str(tmp_path / "parallel_info.dat")) == (nranks, 1, nranks, 8)
    t, uu = ref.read_output(str(tmp_path / lapsio.out_name(3)), nx, ny, nz)
    assert t == 1.25
    want = mk.sample_fields(nx, ny, nz)                    # [v, z, y, x], slabs written by two "ranks" in reverse order
    assert np.array_equal(uu, want.transpose(0, 3, 2, 1))  # the C reader stores [ivar][ix][iy][iz]
    te, radius, ur = ref.read_EBM(str(tmp_path / "EBM_info.dat"))
    ours = np.loadtxt(tmp_path / "EBM_info.dat", ndmin=2)
    assert len(te) == ours.shape[0] >= 1
    assert np.array_equal(te, ours[:, 0]) and np.array_equal(radius, ours[:, 1]) and np.array_equal(ur, ours[:, 2])


def test_reference_reader_reads_a_driver_run(ref, tmp_path):
    """program mhd's stand-in on the kernel emulator leaves an output directory the reference's post-processing
    program can open: every file of main.c's start-up sequence, and the last outNNN.dat equal to the solver state."""
    import build_emu
    from test_lapsio import INPUT
    emu = build_emu.build()
    (tmp_path / "mhd.input").write_text(INPUT)
    d = Driver(str(tmp_path / "mhd.input"), str(tmp_path), lib_path=emu)
    d.run(max_steps=2, echo=False)
    nx, ny, nz, _, _, _ = ref.read_grid(str(tmp_path / "grid.dat"))
    assert (nx, ny, nz) == (d.nx, d.ny, d.nz)
    npe, iproc, jproc, nvar = ref.read_parallel_info(str(tmp_path / "parallel_info.dat"))
    assert (npe, iproc, jproc, nvar) == (1, 1, 1, 8)
    te, radius, ur = ref.read_EBM(str(tmp_path / "EBM_info.dat"))
    assert radius[0] == 30.0 and ur[0] == 1.167 and abs(radius[-1] - (30.0 + 1.167 * te[-1])) < 1e-3
    names = sorted(f for f in os.listdir(tmp_path) if f.startswith("out"))
    t, uu = ref.read_output(str(tmp_path / names[-1]), nx, ny, nz, nvar)
    assert abs(t - np.float32(d.time)) < 1e-6
    state = d.solver.get_output(primitive=True)            # what output_uu writes: rho, u, B, p
    assert np.array_equal(uu, state.transpose(0, 3, 2, 1))
    d.solver.close()
