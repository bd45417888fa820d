"""integration/mhd_main.cpp — the compiled stand-in for `program mhd` over the C ABI — against the Python stand-in
(laps_b200/driver.py) on the same mhd.input: same files, same numbers.  Here the program is linked with the test-only kernel
emulator; tests/test_gpu_zz_cpp_driver.py links it with the real library on the GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "emu"))
import build_emu  # noqa: E402
from laps_b200 import lapsio  # noqa: E402
from laps_b200.driver import Driver  # noqa: E402
from test_lapsio import INPUT  # noqa: E402


def build_cpp_driver(lib, out):
    """g++ against `lib` (the emulator's .so here, laps_b200/_lib/liblaps_b200.so on the GPU box)."""
    src = os.path.join(ROOT, "integration", "mhd_main.cpp")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        d = os.path.dirname(os.path.abspath(lib))
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), src, "-L" + d,
                               "-l:" + os.path.basename(lib), "-Wl,-rpath," + d, "-o", out])
    return out


def compare_runs(exe, lib_for_python, tmp_path, text, steps, tree="compressible"):
    """Run both stand-ins on the same input in two directories and compare every file they leave."""
    a, b = tmp_path / "cpp", tmp_path / "py"
    for d in (a, b):
        d.mkdir()
        (d / "mhd.input").write_text(text)
    out = subprocess.run([exe, "--input", str(a / "mhd.input"), "--outdir", str(a), "--max-steps", str(steps), "--tree", tree],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:]
    assert "OUTPUT RMS at time:" in out.stdout
    d = Driver(str(b / "mhd.input"), str(b), lib_path=lib_for_python, tree=tree)
    assert d.run(max_steps=steps, echo=False) == steps
    d.close()
    names = sorted(os.listdir(b))
    assert sorted(os.listdir(a)) == names and "out000.dat" in names and "rms.dat" in names
    for n in ("grid.dat", "parallel_info.dat"):                       # byte for byte
        assert (a / n).read_bytes() == (b / n).read_bytes(), n
    nx = ny = 16
    nz = 1 if tree.endswith("2d") else 16
    for n in names:
        if n.startswith("out"):
            assert lapsio.read_out_header(str(a / n)) == lapsio.read_out_header(str(b / n))
            x, y = lapsio.read_out_slab(str(a / n), nx, ny, nz), lapsio.read_out_slab(str(b / n), nx, ny, nz)
            # the two programs evaluate the initial sine wave with different libm's: round-off apart, not bit-equal
            assert np.abs(x - y).max() <= 1e-12 * np.abs(y).max(), n
    ra, rb = np.loadtxt(a / "rms.dat", ndmin=2), np.loadtxt(b / "rms.dat", ndmin=2)
    assert ra.shape == rb.shape and ra.shape[1] == 20 and np.array_equal(ra[:, 0], rb[:, 0])
    assert np.allclose(ra, rb, rtol=1e-7, atol=1e-12)
    assert (a / "EBM_info.dat").read_text() == (b / "EBM_info.dat").read_text()
    assert "Iterations     :%8d" % steps in (a / "log").read_text()
    return a, b


ALFVEN = INPUT.replace("ipert = 7", "ipert = 1").replace("Bx0 = 1.", "Bx0 = 1.\n   wave_number_jet = 2")


@pytest.fixture(scope="module")
def emu():
    return build_emu.build()


def test_compiled_driver_writes_the_files_the_python_stand_in_writes(emu, tmp_path):
    exe = build_cpp_driver(emu, os.path.join(HERE, "_build", "mhd_main_emu"))
    a, _ = compare_runs(exe, emu, tmp_path, ALFVEN, 3)
    # restart from its own last file (restart.f90:17-63): one more step, the next file appears
    last = sorted(f for f in os.listdir(a) if f.startswith("out"))[-1]
    n = int(last[3:6])
    (a / "mhd.input").write_text(ALFVEN.replace("dtrms = 0.2", "dtrms = 0.2\n   if_restart = T\n   n_start = %d" % n))
    out = subprocess.run([exe, "--input", str(a / "mhd.input"), "--outdir", str(a), "--max-steps", "1", "--quiet"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:]
    assert os.path.exists(a / lapsio.out_name(n + 1))
    assert lapsio.read_out_header(str(a / lapsio.out_name(n + 1))) > lapsio.read_out_header(str(a / last))


@pytest.mark.parametrize("tree", ["incompressible", "compressible2d", "incompressible2d"])
def test_compiled_driver_runs_the_other_source_trees(emu, tmp_path, tree):
    exe = build_cpp_driver(emu, os.path.join(HERE, "_build", "mhd_main_emu"))
    compare_runs(exe, emu, tmp_path, ALFVEN, 2, tree=tree)


def test_compiled_driver_external_force_of_the_2d_tree(emu, tmp_path):
    """The user routine calc_external_force_real as shipped (2D/mhdrhs.f90:480-531), evaluated by the driver once per step."""
    exe = build_cpp_driver(emu, os.path.join(HERE, "_build", "mhd_main_emu"))
    text = ALFVEN.replace("ipert = 1", "ipert = 1\n   if_external_force = T")
    a, _ = compare_runs(exe, emu, tmp_path, text, 2, tree="compressible2d")
    x0 = lapsio.read_out_slab(str(a / "out000.dat"), 16, 16, 1)
    x1 = lapsio.read_out_slab(str(a / sorted(f for f in os.listdir(a) if f.startswith("out"))[-1]), 16, 16, 1)
    assert np.abs(x1[6] - x0[6]).max() > 1e-3          # B_z was driven


def test_compiled_driver_reports_library_errors(emu, tmp_path):
    exe = build_cpp_driver(emu, os.path.join(HERE, "_build", "mhd_main_emu"))
    (tmp_path / "mhd.input").write_text(ALFVEN.replace("nx = 16", "nx = 24"))
    out = subprocess.run([exe, "--input", str(tmp_path / "mhd.input"), "--outdir", str(tmp_path)], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=60)
    assert out.returncode == 1 and "laps_create" in out.stdout and "2^k" in out.stdout
