#!/usr/bin/env python
"""Device time per RK step of the other BASELINE.json configurations on one GPU (the headline
configuration is bench.py's): config 1 (64^3 compressible MHD), config 2 (2048^2 2D compressible
Hall-MHD), config 3 (256^3 incompressible MHD).  Prints one JSON line per configuration.
State resident in HBM, CUDA events on the library's stream (laps_last_step_ms), 3 warm-up steps."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laps_b200 import Solver, synthetic  # noqa: E402

COMMON = dict(Lx=24.0, Ly=24.0, Lz=24.0, adiabatic_index=1.666667, if_resis=1, resistivity=1e-4, if_visc=1,
              viscosity=1e-4, cfl=0.5, dealias_option=1, radius0=30.0)


def run(name, kw, prim, steps=10, warm=3):
    with Solver(**kw) as g:
        g.set_primitive(prim)
        g.vardt()
        for _ in range(warm):
            g.step()
        ms = []
        for _ in range(steps):
            g.step()
            ms.append(g.last_step_ms()[0])
        uu, _ = g.get_state()
        npts = g.nx * g.ny * g.nz
        t = float(np.median(ms))
        print(json.dumps({"config": name, "ms_per_step_evolve": t, "grid_point_steps_per_s": npts / (t * 1e-3),
                          "launches_per_step": g.last_step_ms()[1], "finite": bool(np.isfinite(uu).all()), "dt": g.dt}))


def main():
    n = 64
    run("config 1: 3D compressible MHD 64^3", dict(COMMON, nx=n, ny=n, nz=n),
        synthetic.turbulence_slab(n, n, n, 24.0, 24.0, 24.0, kmax=8))
    n = 2048
    prim3 = synthetic.turbulence_slab(n, n, 32, 24.0, 24.0, 24.0, kmax=8, z_size=1)
    run("config 2: 2D compressible Hall-MHD 2048^2", dict(COMMON, nx=n, ny=n, nz=1, ndim=2, if_hall=1, ion_inertial_length=0.2), prim3)
    n = 256
    run("config 3: 3D incompressible MHD 256^3", dict(COMMON, nx=n, ny=n, nz=n, incompressible=1, rho0=1.0),
        synthetic.turbulence_slab(n, n, n, 24.0, 24.0, 24.0, kmax=8, drho0=0.0))
    run("config 3 + Hall + expanding box 256^3", dict(COMMON, nx=n, ny=n, nz=n, incompressible=1, rho0=1.0, if_hall=1,
                                                       ion_inertial_length=0.2, if_AEB=1, Ur0=1.167),
        synthetic.turbulence_slab(n, n, n, 24.0, 24.0, 24.0, kmax=8, drho0=0.0))


if __name__ == "__main__":
    main()
