"""Builds the TEST-ONLY CPU emulation of the CUDA kernels (tests/_build/liblaps_emu.so).

The same kernel sources the product compiles with nvcc are compiled here with g++ against
tests/emu/cuda_emu.h so that index math can be checked in a container without a GPU.  Never
imported by the package."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "tests", "_build", "liblaps_emu.so")
SRC = [os.path.join(ROOT, "laps_b200", "csrc", "solver.cu"), os.path.join(ROOT, "tests", "emu", "cuda_emu.cpp")]
CSRC = os.path.join(ROOT, "laps_b200", "csrc")
DEPS = SRC + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [
    os.path.join(ROOT, "tests", "emu", "cuda_emu.h"), os.path.join(ROOT, "include", "laps_b200.h")]


def build(force=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-DLAPS_EMU_BUILD",
           "-I" + os.path.join(ROOT, "tests", "emu"), "-I" + os.path.join(ROOT, "laps_b200", "csrc"),
           "-x", "c++"] + SRC + ["-o", OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
