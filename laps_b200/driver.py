"""Stand-in for the reference's ``program mhd`` (src_compressible/mhd.f90:16-293) around the C ABI:
reads ``mhd.input``, builds the initial data (or reads a restart file), runs the Principal loop and
writes ``grid.dat``, ``parallel_info.dat``, ``outNNN.dat``, ``rms.dat``, ``EBM_info.dat`` and ``log``
in the reference's formats (``lapsio.py``).  This is the caller on either side of the hot path
(SURVEY.md 8(f) rank 1); a Fortran driver patched as in INTEGRATION.md makes the same calls.

    python -m laps_b200.driver [--input mhd.input] [--outdir .] [--max-steps N]
                               [--tree compressible|compressible2d|incompressible|incompressible2d]
    torchrun --nproc-per-node P -m laps_b200.driver ...          (one rank per GPU, slab decomposition)

Initial conditions built in: ``ifield = 3`` (uniform background) with ``ipert`` 0 (none), 1 (Alfven wave,
mhdinit.f90:328-342) or 7 (random-phase turbulence, :695-823, seeded NumPy phases — the reference's
own phases are compiler specific); anything else must come through a restart file (``if_restart``),
exactly as the reference restarts (restart.f90:17-63).  All compute is the CUDA library's.
"""
from __future__ import annotations

import argparse
import math
import os
import sys
import time as _time

import numpy as np

from . import lapsio, synthetic
from .solver import Solver


def _get(groups, group, key, default):
    return groups.get(group, {}).get(key.lower(), default)


TREES = ("compressible", "compressible2d", "incompressible", "incompressible2d")


def params_from_namelists(nl, rank=0, nranks=1, device=0, tree="compressible"):
    """laps_params fields from the namelists of mhd.f90:30-53 (module defaults where a key is absent:
    mhdinit.f90:5-54, dealiasing.f90:9-10, AEBmod.f90:10-12, mhd.f90:21).  ``tree`` selects which of the
    reference's four source trees the input file belongs to (they share the namelist syntax; the 2D trees add
    if_limit_dt_increase and, compressible only, if_z_radial; the incompressible trees use rho0 = 1,
    src_incompressible/mhdinit.f90:15)."""
    if tree not in TREES:
        raise ValueError(f"tree must be one of {TREES}")
    g = lambda grp, key, d: _get(nl, grp, key, d)  # noqa: E731
    extra = {}
    if tree.endswith("2d"):
        extra.update(ndim=2, if_limit_dt_increase=bool(g("numerical", "if_limit_dt_increase", False)))
        if tree == "compressible2d":
            extra.update(if_z_radial=bool(g("aeb", "if_z_radial", False)),
                         if_external_force=bool(g("pert", "if_external_force", False)))   # 2D/mhd.f90:43
    if tree.startswith("incompressible"):
        extra.update(incompressible=1, rho0=1.0)
    # &prl ndim_parallel (parallel.f90:56-70): the shipped 3D inputs ask for pencils (ndim_parallel = 2).  The library
    # decomposes in slabs whatever the namelist says — results and outNNN.dat files do not depend on the decomposition,
    # and parallel_info.dat records what was used (npe, iproc = 1, jproc = npe).
    return dict(
        nx=int(g("grid", "nx", 128)), ny=int(g("grid", "ny", 128)), nz=1 if tree.endswith("2d") else int(g("grid", "nz", 64)),
        Lx=float(g("grid", "Lx", 1.0)), Ly=float(g("grid", "Ly", 1.0)), Lz=float(g("grid", "Lz", 1.0)),
        adiabatic_index=float(g("phys", "adiabatic_index", 5.0 / 3.0)),
        if_resis=bool(g("phys", "if_resis", False)), resistivity=float(g("phys", "resistivity", 0.0)),
        if_visc=bool(g("phys", "if_visc", False)), viscosity=float(g("phys", "viscosity", 0.0)),
        if_resis_exp=bool(g("numerical", "if_resis_exp", False)), if_visc_exp=bool(g("numerical", "if_visc_exp", False)),
        if_conserve_background=bool(g("numerical", "if_conserve_background", False)),
        cfl=float(g("numerical", "cfl", 0.5)), dealias_option=int(g("numerical", "dealias_option", 2)),
        afx=float(g("numerical", "afx", 0.495)), afy=float(g("numerical", "afy", 0.495)), afz=float(g("numerical", "afz", 0.495)),
        if_AEB=bool(g("aeb", "if_AEB", False)), if_corotating=bool(g("aeb", "if_corotating", False)),
        radius0=float(g("aeb", "radius0", 30.0)), Ur0=float(g("aeb", "Ur0", 0.0)),
        corotating_angle=float(g("aeb", "corotating_angle", 0.0)),
        if_hall=bool(g("hall", "if_Hall", False)), ion_inertial_length=float(g("hall", "ion_inertial_length", 0.0)),
        rank=rank, nranks=nranks, device=device, **extra)


class Driver:
    def __init__(self, input_path="mhd.input", outdir=".", rank=0, nranks=1, device=0, lib_path=None, barrier=None,
                 connect=None, tree="compressible", bcast=None):
        self.nl = lapsio.read_namelists(input_path)
        self.tree = tree
        self.two_d = tree.endswith("2d")
        self.dstep_calcdt = 20 if self.two_d else 1          # 2D/mhd.f90:22,237-240: vardt every 20 steps
        self.dstep_checknan = 200 if self.two_d else 0       # 2D/mhd.f90:25,242-252: checkNan every 200 steps
        self.dstep_checksave = 40                            # mhd.f90:20: look at the wall clock every 40 steps ...
        self.delta_clocktime_output = 60.0 * 50.0            # mhd.f90:16: ... and dump out999.dat every 50 minutes of it
        self.stopped_on_nan = False
        self.outdir = outdir
        self.rank, self.nranks = rank, nranks
        self.barrier = barrier or (lambda: None)
        self.bcast = bcast                                   # rank 0's integer to every rank (MPI_Bcast of mhd.f90:208)
        kw = params_from_namelists(self.nl, rank, nranks, device, tree)
        self.kw = kw
        g = lambda grp, key, d: _get(self.nl, grp, key, d)  # noqa: E731
        self.tmax = float(g("genr", "tmax", 1.0))
        self.dtout = float(g("genr", "dtout", 1.0))
        self.dtrms = float(g("genr", "dtrms", 1.0))
        self.output_primitive = bool(g("genr", "output_primitive", True))
        self.if_restart = bool(g("genr", "if_restart", False))
        self.n_start = int(g("genr", "n_start", 0))
        self.solver = Solver(lib_path, **kw)
        if connect is not None and nranks > 1:
            connect(self.solver)
        s = self.solver
        self.nx, self.ny, self.nz = s.nx, s.ny, s.nz
        self.zo, self.zn = s.ext.z_offset, s.ext.z_size
        self.Ur = kw["Ur0"] if kw["if_AEB"] else 0.0          # mhd.f90:88-90
        self.radius = kw["radius0"]
        self.time = 0.0
        self.istep = 0
        self.clock0 = _time.perf_counter()
        os.makedirs(outdir, exist_ok=True)

    def path(self, name):
        return os.path.join(self.outdir, name)

    def external_force(self):
        """The user routine calc_external_force_real as the reference ships it (2D/mhdrhs.f90:480-531): a Gaussian
        forcing of B_z centred at x = Lx/2 whose y centre moves at speed 0.3, with its two periodic images."""
        Lx, Ly = self.kw["Lx"], self.kw["Ly"]
        x = np.arange(self.nx) * (Lx / self.nx)
        y = np.arange(self.ny) * (Ly / self.ny)
        dBdt, xc, w = 0.2, 0.5 * Lx, 0.05 * Ly
        yc = math.fmod(0.2 * Ly + 0.3 * self.time, Ly)
        fx = np.exp(-((x - xc) / w) ** 2)[None, :]
        f = dBdt * fx * np.exp(-((y[:, None] - yc) / w) ** 2)
        f = f + dBdt * fx * np.exp(-((y[:, None] - (yc + Ly)) / w) ** 2)
        f = f + dBdt * fx * np.exp(-((y[:, None] - (yc - Ly)) / w) ** 2)
        return f[None]

    # ------------------------------------------------------------------ initial data
    def initial_primitive(self):
        g = lambda grp, key, d: _get(self.nl, grp, key, d)  # noqa: E731
        kw = self.kw
        if self.if_restart:                                   # mhd.f90:101-107, restart.f90:17-63
            fname = self.path(lapsio.out_name(self.n_start))
            self.time = lapsio.read_out_header(fname)
            return lapsio.read_out_slab(fname, self.nx, self.ny, self.nz, self.zo, self.zn)
        ifield, ipert = int(g("field", "ifield", 3)), int(g("pert", "ipert", 0))
        if ifield != 3:
            raise NotImplementedError("built-in initial data: ifield = 3 only (others: restart from an outNNN.dat)")
        bx0, by0, bz0 = float(g("field", "Bx0", 0.0)), float(g("field", "By0", 0.0)), float(g("field", "Bz0", 0.0))
        press0 = float(g("field", "press0", 1.0))
        if ipert == 7 and not self.two_d:
            n = int(g("pert", "nmodex", 8))
            return synthetic.turbulence_slab(self.nx, self.ny, self.nz, kw["Lx"], kw["Ly"], kw["Lz"], z_offset=self.zo,
                                             z_size=self.zn, bx0=bx0, by0=by0, bz0=bz0, press0=press0,
                                             db0=float(g("pert", "db0", 0.1)), dv0=float(g("pert", "dv0", 0.0)),
                                             drho0=float(g("pert", "drho0", 0.0)), kmax=n)
        prim = synthetic.uniform_background(self.nx, self.ny, self.zn, bx0, by0, bz0, press0)
        if ipert == 7 and self.two_d:
            raise NotImplementedError("built-in turbulence (ipert = 7) is the 3D mode table; use a restart file in 2D")
        if ipert == 1:
            ang = kw["corotating_angle"] if kw["if_corotating"] else 0.0
            synthetic.add_alfven_wave(prim, self.nx, kw["Lx"], db0=float(g("pert", "db0", 0.1)),
                                      wave_number_jet=int(g("pert", "wave_number_jet", 1)), cor_angle=ang)
        elif ipert != 0:
            raise NotImplementedError("built-in perturbations: ipert = 0, 1, 7 (others: restart from an outNNN.dat)")
        return prim

    # ------------------------------------------------------------------ output (mhdoutput.f90, mhdrms.f90, AEBmod.f90)
    def output_uu(self, iout):
        s = self.solver
        data = s.get_output(self.output_primitive)
        fname = self.path(lapsio.out_name(iout))
        if self.rank == 0:
            lapsio.write_out_header(fname, self.time)
            with open(fname, "r+b") as f:                     # size the file once, ranks then fill their slabs
                f.truncate(lapsio.OUT_DISPLACEMENT + 8 * 8 * self.nx * self.ny * self.nz)
        self.barrier()
        lapsio.write_out_slab(fname, data, self.nz, self.zo)
        self.barrier()

    def output_rms(self):
        ave, rms, ru2 = self.solver.calc_rms()                # collective
        if self.rank == 0:
            with open(self.path("rms.dat"), "a") as f:
                f.write(lapsio.rms_line(self.time, ave, rms, ru2) + "\n")

    def output_aeb(self):
        if self.rank == 0:
            with open(self.path("EBM_info.dat"), "a") as f:
                f.write(lapsio.ebm_line(self.time, self.radius, self.Ur) + "\n")

    def write_log(self, dt):                                  # mhd.f90:431-457
        if self.rank != 0:
            return
        clock = _time.perf_counter() - self.clock0
        hour = math.floor(clock / 3600.0)
        minute = math.floor((clock / 3600.0 - hour) * 60)
        second = math.floor(((clock / 3600.0 - hour) * 60 - minute) * 60)
        with open(self.path("log"), "w") as f:
            f.write("   Simulation time:%8.4f\n" % self.time)
            f.write(" dt:  %r\n" % dt)
            f.write("   Real time (sec):%15.2f\n" % clock)
            f.write("   Real time (hh,mm,ss):%3dh%3dm%3ds\n" % (hour, minute, second))
            f.write("   Iterations     :%8d\n" % self.istep)
            f.write(" tasks: %12d\n" % self.nranks)

    # ------------------------------------------------------------------ program mhd
    def run(self, max_steps=None, echo=True):
        s, kw = self.solver, self.kw
        g = lambda grp, key, d: _get(self.nl, grp, key, d)  # noqa: E731
        if (not self.if_restart and not self.two_d and int(g("field", "ifield", 3)) == 3 and int(g("pert", "ipert", 0)) == 7):
            # the mode table goes to the device as it is (laps_set_primitive_modes): no N^3 host array, no upload
            bx0, by0, bz0 = float(g("field", "Bx0", 0.0)), float(g("field", "By0", 0.0)), float(g("field", "Bz0", 0.0))
            n = int(g("pert", "nmodex", 8))
            ks, coefs = synthetic.mode_table(kw["Lx"], kw["Ly"], kw["Lz"], bx0, by0, bz0, n, n, n, (101, 116, 132),
                                             float(g("pert", "db0", 0.1)), float(g("pert", "dv0", 0.0)), float(g("pert", "drho0", 0.0)))
            s.set_primitive_modes(ks, coefs, [1.0, 0.0, 0.0, 0.0, bx0, by0, bz0, float(g("field", "press0", 1.0))])
        else:
            prim = self.initial_primitive()
            if self.if_restart:
                s.time = self.time
                s.evolve_radius(self.time)                    # mhd.f90:101-103
                self.radius = kw["radius0"] + self.Ur * self.time
            s.set_primitive(prim)                             # mhd.f90:121-122
        s.dt = 0.0
        dt = s.vardt()                                        # mhd.f90:135-136
        if self.rank == 0:
            x = np.arange(self.nx) * (kw["Lx"] / self.nx)
            y = np.arange(self.ny) * (kw["Ly"] / self.ny)
            z = np.arange(self.nz) * (kw["Lz"] / self.nz)
            if self.two_d:                                    # 2D/mhdoutput.f90:45-63: nx, ny / npe, nvar only
                lapsio.write_grid(self.path("grid.dat"), x, y)
                lapsio.write_parallel_info(self.path("parallel_info.dat"), self.nranks)
            else:
                lapsio.write_grid(self.path("grid.dat"), x, y, z)                                  # mhd.f90:139
                lapsio.write_parallel_info(self.path("parallel_info.dat"), self.nranks, 1, self.nranks)   # :140
            # rms_initialize / AEB_initialize (mhdrms.f90:29-36, AEBmod.f90:34-41): an existing file is opened for
            # APPEND — a fresh run in a used directory adds to the old rows, as the reference does — else created
            open(self.path("rms.dat"), "a").close()
            open(self.path("EBM_info.dat"), "a").close()
        dtlog = min(self.dtout, self.dtrms) / 10.0            # mhd.f90:142-149
        iout = self.n_start
        tout, toutrms, tlog = self.time + self.dtout, self.time + self.dtrms, self.time + dtlog
        self.output_uu(iout)                                  # mhd.f90:154-163
        iout += 1
        max_divb = s.calc_max_divB()
        if echo and self.rank == 0:
            print("      OUTPUT RMS at time:  %10.4f, max(div B) = %10.2E, dt = %12.4E" % (self.time, max_divb, dt))
        self.output_rms()
        self.output_aeb()
        clocktime_output = self.delta_clocktime_output        # mhd.f90:165-166
        while True:                                           # Principal, mhd.f90:169-287
            if self.time >= tout:
                self.output_uu(iout)
                iout += 1
                tout += self.dtout
            if self.time >= toutrms:
                max_divb = s.calc_max_divB()
                if echo and self.rank == 0:
                    print("      OUTPUT RMS at time:  %10.4f, max(div B) = %10.2E, dt = %12.4E" % (self.time, max_divb, dt))
                self.output_rms()
                self.output_aeb()
                toutrms += self.dtrms
            if self.istep > 0 and self.istep % self.dstep_checksave == 0:   # wall-clock backup dump, mhd.f90:194-214
                save = (_time.perf_counter() - self.clock0) >= clocktime_output   # rank 0's clock decides (MPI_Bcast)
                if self.bcast is not None:
                    save = bool(self.bcast(int(save)))
                if save:
                    if echo and self.rank == 0:
                        print("   OUTPUT for backup at real time (sec):%15.2f  , time =   %10.4f" % (_time.perf_counter() - self.clock0, self.time))
                    self.output_uu(999)
                    clocktime_output += self.delta_clocktime_output
            if dt < 1e-8:                                     # mhd.f90:205-228
                self.output_uu(iout)
                self.output_rms()
                self.output_aeb()
                break
            s.time = self.time
            if kw.get("if_external_force"):                   # calc_external_force_real, called from calc_flux with the
                s.set_external_force(self.external_force())   # step's `time` (2D/mhdrhs.f90:129-131,480-531)
            s.evolve()                                        # mhd.f90:245
            self.time = self.time + dt
            self.istep += 1
            s.time = self.time
            s.evolve_radius(self.time)                        # :248
            self.radius = kw["radius0"] + self.Ur * self.time
            if self.time >= self.tmax or (max_steps is not None and self.istep >= max_steps):   # :250-276
                self.output_uu(iout)
                s.calc_max_divB()
                self.output_rms()
                self.output_aeb()
                break
            if self.time >= tlog:
                self.write_log(dt)
                tlog += dtlog
            if self.istep % self.dstep_calcdt == 0:
                dt = s.vardt()                                # :285 (2D trees: every dstep_calcdt steps)
            if self.dstep_checknan and self.istep % self.dstep_checknan == 0 and s.checkNan():   # 2D/mhd.f90:242-252
                if echo and self.rank == 0:
                    print(" NaN encountered!!! Exit the program at t = %10.4f" % self.time)
                self.stopped_on_nan = True
                break
        self.write_log(dt)
        return self.istep

    def close(self):
        """parallel_end (mhd.f90:291-293): no rank may free its exchange buffers while a peer can still store into them."""
        self.solver.sync()
        self.barrier()
        self.solver.close()


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--input", default="mhd.input")
    ap.add_argument("--outdir", default=".")
    ap.add_argument("--max-steps", type=int, default=None)
    ap.add_argument("--tree", default="compressible", choices=TREES)
    args = ap.parse_args(argv)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    barrier, connect, bcast = None, None, None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

        def barrier():
            dist.barrier()

        def bcast(v):
            t = torch.tensor([int(v)], device="cuda")
            dist.broadcast(t, 0)
            return int(t.item())

        def connect(g):
            blob = torch.from_numpy(np.frombuffer(g.export_peer_blob(), dtype=np.uint8).copy()).cuda()
            blobs = [torch.empty_like(blob) for _ in range(world)]
            dist.all_gather(blobs, blob)
            g.import_peer_blobs(b"".join(bytes(b.cpu().numpy().tobytes()) for b in blobs))
            dist.barrier()
    d = Driver(args.input, args.outdir, rank, world, local, barrier=barrier, connect=connect, tree=args.tree, bcast=bcast)
    n = d.run(args.max_steps)
    if rank == 0:
        print(f"{n} steps, time = {d.time:.6f}")
    d.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
