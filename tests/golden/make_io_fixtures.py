#!/usr/bin/env python
"""Generates tests/golden/io/: small files in the reference's formats written by laps_b200.lapsio, read back
HERE by the reference's own post-processing reader (/root/reference/data_process/3D_Python/read_output.py:
read_parallel_info, read_grid, read_uu, read_output_location, read_output_slice, read_EBM — the function
definitions only; the module's top-level plotting code is not executed) and, for rms.dat, by the recipe of
read_rms.py (np.loadtxt + column slicing).  What the reference reader returned is stored in expected.npz.

Run in the build container (needs /root/reference):  python tests/golden/make_io_fixtures.py
tests/test_lapsio.py then checks, without the reference, that lapsio still writes these bytes and reads them
back to the stored arrays."""
import ast
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from laps_b200 import lapsio  # noqa: E402

REF = "/root/reference/data_process/3D_Python/read_output.py"
REF_2D = "/root/reference/data_process/2D_Python/read_output.py"


def reference_reader_functions(path=REF):
    tree = ast.parse(open(path).read())
    funcs = [n for n in tree.body if isinstance(n, ast.FunctionDef)]
    ns = {"np": np, "struct": struct}
    exec(compile(ast.Module(body=funcs, type_ignores=[]), path, "exec"), ns)
    return ns


def write_all_2d(outdir, nx=6, ny=4, nranks=1):
    """The files of the 2D trees (2D/mhdoutput.f90:45-63: grid.dat holds nx, ny and two grids, parallel_info.dat
    npe and nvar; outNNN.dat is the 3D layout with nz = 1)."""
    os.makedirs(outdir, exist_ok=True)
    grids = [np.arange(n) * (l / n) for n, l in zip((nx, ny), (24.0, 12.0))]
    lapsio.write_grid(os.path.join(outdir, "grid.dat"), *grids)
    lapsio.write_parallel_info(os.path.join(outdir, "parallel_info.dat"), nranks)
    a = sample_fields(nx, ny, 1)
    path = os.path.join(outdir, lapsio.out_name(7))
    lapsio.write_out_header(path, 0.375)
    with open(path, "r+b") as f:
        f.truncate(lapsio.OUT_DISPLACEMENT + 8 * a.size)
    lapsio.write_out_slab(path, a, 1, 0)
    with open(os.path.join(outdir, "EBM_info.dat"), "w") as f:
        f.write(lapsio.ebm_line(0.0, 30.0, 1.167) + "\n")      # a single line: the reader's reshape branch
    return a, grids


def sample_fields(nx, ny, nz, nvar=8):
    """Deterministic, position-encoding values (the idea of ipert = 999, mhdinit.f90:1021-1030)."""
    v, z, y, x = np.meshgrid(np.arange(nvar), np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    return (1000.0 * v + 100.0 * z + 10.0 * y + x + 0.5 * np.sin(1.0 + x + 2 * y + 3 * z + 5 * v)).astype(np.float64)


def write_all(outdir, nx=6, ny=4, nz=5, nranks=2):
    os.makedirs(outdir, exist_ok=True)
    L = (24.0, 12.0, 6.0)
    grids = [np.arange(n) * (l / n) for n, l in zip((nx, ny, nz), L)]
    lapsio.write_grid(os.path.join(outdir, "grid.dat"), *grids)
    lapsio.write_parallel_info(os.path.join(outdir, "parallel_info.dat"), nranks, 1, nranks, 8)
    a = sample_fields(nx, ny, nz)
    path = os.path.join(outdir, lapsio.out_name(3))
    lapsio.write_out_header(path, 1.25)
    with open(path, "r+b") as f:
        f.truncate(lapsio.OUT_DISPLACEMENT + 8 * a.size)
    q = nz // nranks                      # decompose_1d: nz/P each, the remainder on the last rank
    for r in reversed(range(nranks)):     # any order: every rank writes only its own slab
        zo = r * q
        zn = q if r < nranks - 1 else nz - zo
        lapsio.write_out_slab(path, a[:, zo:zo + zn], nz, zo)
    with open(os.path.join(outdir, "rms.dat"), "w") as f:
        for k in range(3):
            t = 0.05 * k
            f.write(lapsio.rms_line(t, 1.0 + 0.1 * np.arange(8) * (k + 1), 1e-3 * np.arange(8) ** 2, [1e-10 * (k + 1), 2.5e-3, -3.75e5]) + "\n")
    with open(os.path.join(outdir, "EBM_info.dat"), "w") as f:
        for k in range(3):
            f.write(lapsio.ebm_line(0.05 * k, 30.0 + 1.167 * 0.05 * k, 1.167) + "\n")
    return a, grids


def main():
    base = os.path.join(HERE, "io")
    outdir = os.path.join(base, "output")            # the reference reader opens './output/...'
    a, grids = write_all(outdir)
    ns = reference_reader_functions()
    cwd = os.getcwd()
    os.chdir(base)
    try:
        npe, iproc, jproc, nvar = ns["read_parallel_info"]()
        xg, yg, zg = ns["read_grid"]()
        nx, ny, nz = len(xg), len(yg), len(zg)
        t, uu = ns["read_uu"]("./output/out003.dat", nx, ny, nz, nvar)
        t1, loc = ns["read_output_location"]("./output/out003.dat", 2, 1, 3, nvar, nx, ny, nz)
        t2, sl = ns["read_output_slice"]("./output/out003.dat", [-1, -1, 2], nvar, nx, ny, nz)
        t_ebm, radius, ur = ns["read_EBM"]()
        rms = np.loadtxt("./output/rms.dat")          # read_rms.py:28
    finally:
        os.chdir(cwd)
    # the reference reader returns uu[ix,iy,iz,ivar]; lapsio's layout is [ivar,iz,iy,ix]
    assert (npe, iproc, jproc, nvar) == (2, 1, 2, 8)
    assert np.array_equal(uu, a.transpose(3, 2, 1, 0)) and t == np.float32(1.25)
    assert np.array_equal(loc, a[:, 3, 1, 2]) and np.array_equal(sl, a[:, 2].transpose(0, 2, 1))
    assert np.allclose(xg, grids[0], rtol=1e-7) and np.allclose(zg, grids[2], rtol=1e-7)
    assert rms.shape == (3, 20) and np.allclose(t_ebm, [0.0, 0.05, 0.1])
    np.savez(os.path.join(base, "expected.npz"), npe=npe, iproc=iproc, jproc=jproc, nvar=nvar, xgrid=xg, ygrid=yg, zgrid=zg,
             t=t, uu=uu, loc=loc, slice_xy=sl, t_ebm=t_ebm, radius=radius, ur=ur, rms=rms)
    print("fixtures written to", base)
    # ---- the 2D trees, read back by data_process/2D_Python/read_output.py
    base2 = os.path.join(HERE, "io2d")
    a2, grids2 = write_all_2d(os.path.join(base2, "output"))
    ns = reference_reader_functions(REF_2D)
    os.chdir(base2)
    try:
        npe, nvar = ns["read_parallel_info"]()
        xg, yg = ns["read_grid"]()
        t, uu = ns["read_uu"]("./output/out007.dat", len(xg), len(yg), nvar)
        t_ebm, radius, ur = ns["read_EBM"]()
    finally:
        os.chdir(cwd)
    assert (npe, nvar) == (1, 8) and t == np.float32(0.375)
    assert np.array_equal(uu, a2[:, 0].transpose(2, 1, 0))       # the 2D reader returns uu[ix,iy,ivar]
    assert np.allclose(xg, grids2[0], rtol=1e-7) and np.allclose(yg, grids2[1], rtol=1e-7)
    np.savez(os.path.join(base2, "expected.npz"), npe=npe, nvar=nvar, xgrid=xg, ygrid=yg, t=t, uu=uu, t_ebm=t_ebm, radius=radius, ur=ur)
    print("fixtures written to", base2)


if __name__ == "__main__":
    main()
