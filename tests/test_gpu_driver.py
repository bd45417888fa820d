"""The stand-in for `program mhd` (laps_b200/driver.py) on the GPU: the Principal loop against the oracle driven the same
way, the files it leaves read back by the reference's own C reader (data_process/3D_C/iofunctions.c, compiled into
oracle/_ref/ in the build container; the .so travels to the GPU box), and a restart cycle from one of its own
outNNN.dat files (restart.f90:17-63) that continues exactly as a solver started from the file's contents."""
import os

import numpy as np
import pytest

import parity_common as pc
from laps_b200 import Solver, lapsio
from laps_b200.driver import Driver
from oracle import laps_oracle as lo
from oracle import ref_io
from test_lapsio import INPUT

pytestmark = pytest.mark.gpu


def test_driver_restart_cycle_and_reference_reader(tmp_path):
    n = 64
    text = INPUT.replace("nx = 16", f"nx = {n}").replace("ny = 16", f"ny = {n}").replace("nz = 16", f"nz = {n}")
    assert f"nx = {n}" in text
    (tmp_path / "mhd.input").write_text(text)
    d = Driver(str(tmp_path / "mhd.input"), str(tmp_path))
    prim0 = d.initial_primitive()
    assert d.run(max_steps=4, echo=False) == 4
    # oracle, same sequencing (mhd.f90:244-248,285)
    p = lo.Params(**{k: v for k, v in d.kw.items() if k not in ("rank", "nranks", "device")})
    o = lo.State(p)
    o.set_primitive(prim0)
    o.vardt()
    for i in range(4):
        o.evolve()
        o.time += o.dt
        o.evolve_radius(o.time)
        if i < 3:
            o.vardt()
    assert abs(d.time - o.time) < 1e-12
    names = sorted(f for f in os.listdir(tmp_path) if f.startswith("out"))
    last = str(tmp_path / names[-1])
    data = lapsio.read_out_slab(last, n, n, n)
    ref = lo.primitive_of(o)
    for v in range(8):
        assert pc.rel_l2(data[v], ref[v]) < 1e-10, v
    rms = np.loadtxt(tmp_path / "rms.dat")
    ebm = np.loadtxt(tmp_path / "EBM_info.dat")
    assert rms.shape[1] == 20 and ebm.shape[1] == 3 and rms.shape[0] == ebm.shape[0] >= 2
    # the reference's own reader opens every file of its post-processing start-up sequence
    path = ref_io.build()
    if path is not None:
        rd = ref_io.ReferenceReader(path)
        gx, gy, gz, _, _, _ = rd.read_grid(str(tmp_path / "grid.dat"))
        assert (gx, gy, gz) == (n, n, n)
        assert rd.read_parallel_info(str(tmp_path / "parallel_info.dat")) == (1, 1, 1, 8)
        te, radius, ur = rd.read_EBM(str(tmp_path / "EBM_info.dat"))
        assert radius[0] == 30.0 and abs(radius[-1] - (30.0 + 1.167 * te[-1])) < 1e-3
        t, uu = rd.read_output(last, n, n, n, 8)
        assert abs(t - np.float32(d.time)) < 1e-6
        assert np.array_equal(uu, d.solver.get_output(primitive=True).transpose(0, 3, 2, 1))
    d.close()
    # restart from the last file (restart.f90:17-63: time from the header, primitives from the body)
    k = int(names[-1][3:6])
    (tmp_path / "mhd.input").write_text(text.replace("dtrms = 0.2", "dtrms = 0.2\n   if_restart = T\n   n_start = %d" % k))
    r = Driver(str(tmp_path / "mhd.input"), str(tmp_path))
    assert r.if_restart and r.n_start == k
    r.run(max_steps=2, echo=False)
    assert r.istep == 2 and os.path.exists(tmp_path / lapsio.out_name(k + 1))
    # ... continues exactly as a solver started by hand from the file's contents
    t0 = float(lapsio.read_out_header(last))
    with Solver(**{kk: v for kk, v in d.kw.items()}) as g:
        g.time = t0
        g.evolve_radius(t0)
        g.set_primitive(data)
        g.vardt()
        for i in range(2):          # the driver's own call sequence (mhd.f90:245-248,285): evolve; time; evolve_radius; vardt
            g.step(calc_dt=False)
            if i == 0:
                g.vardt()
        assert g.time == r.time
        assert np.array_equal(g.get_state()[0], r.solver.get_state()[0])
    r.close()
