"""integration/mhd_main.cpp linked with the real library, on the GPU: the compiled stand-in for `program mhd` and the Python
one leave the same files (tests/test_cpp_driver.py does the same on the kernel emulator).  Sorted late on purpose: written
after the last GPU visit of round 2 — the host code is the one the emulator test runs, the library the one every other
GPU test runs."""
import os

import pytest

import test_cpp_driver as cd

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tree", ["compressible", "incompressible", "compressible2d", "incompressible2d"])
def test_compiled_driver_on_the_gpu(tmp_path, tree):
    from laps_b200 import capi
    lib = capi.DEFAULT_LIB
    exe = cd.build_cpp_driver(lib, os.path.join(cd.HERE, "_build", "mhd_main"))
    cd.compare_runs(exe, None, tmp_path, cd.ALFVEN, 3, tree=tree)
