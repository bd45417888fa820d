"""Field-level parity at the line lengths and grid sizes of the BASELINE.json configurations (the one-step parity
tests of tests/test_gpu_parity.py run at 64^3 ... 256 x 128; the full-size tests there check invariants only).

* 512-, 1024- and 2048-point lines along each axis through a whole RK step (k_fwd_x/k_inv_x, k_fwd_y/k_inv_y,
  k_rhs_z/k_spec_z at the instantiations configs 2, 4 and 5 run) against the NumPy oracle;
* 256^3 compressible Hall-MHD + expanding box, two steps, against the C restatement oracle/laps_cpu.c (itself held to
  the NumPy oracle and to the executed reference source at 1e-11, tests/test_cpu_port.py);
* config 2 at full size (2048^2 2D Hall-MHD) and config 3 at full size (256^3 incompressible), one step, against the
  NumPy oracle.
Tolerance: the north star's 1e-11 relative L2 per field after one RK step."""
import pytest

import parity_common as pc
from laps_b200 import Solver, synthetic
from oracle import laps_oracle as lo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(512, 16, 16), (16, 512, 16), (16, 16, 512), (1024, 16, 16), (16, 1024, 16), (16, 16, 1024),
                                   (2048, 16, 16), (16, 16, 2048)])
def test_long_lines_one_step_parity(shape):
    p, prim = pc.make_case(*shape, hall=True, aeb=True, dealias=1, nmode=1)
    o, g = pc.run_both(p, prim, 1)
    pc.check_state_vectors(o, g, 1e-11)   # (the single-mode perturbation has no u_x: vectors as a whole, see there)
    g.close()


def _params(shape, hall, aeb, **kw):
    nx, ny, nz = shape
    return lo.Params(nx=nx, ny=ny, nz=nz, Lx=24.0, Ly=24.0, Lz=24.0, adiabatic_index=1.666667, if_resis=True, resistivity=1e-4,
                     if_visc=True, viscosity=1e-4, cfl=0.5, dealias_option=1, if_AEB=aeb, radius0=30.0, Ur0=1.167 if aeb else 0.0,
                     if_hall=hall, ion_inertial_length=0.2 if hall else 0.0, **kw)


@pytest.mark.parametrize("shape", [(64, 512, 64), (512, 64, 64), (64, 64, 512)])
def test_long_lines_many_columns(shape):
    """512-point lines with thousands of lines per pass (many CTA waves, the persistent z pass looping over items)."""
    p = _params(shape, hall=True, aeb=True)
    prim = synthetic.turbulence_slab(*shape, p.Lx, p.Ly, p.Lz, kmax=6)
    o, g = pc.run_both(p, prim, 1)
    pc.check_state(o, g, 1e-11)
    g.close()


def test_256_cubed_two_steps_against_the_c_restatement():
    """The bench physics (Hall + expanding box + spherical mask) at 256^3: fields after two steps within 1e-11 of
    oracle/laps_cpu.c; dt identical to 1e-12."""
    from oracle import cpu_port
    n = 256
    p = _params((n, n, n), hall=True, aeb=True)
    prim = synthetic.turbulence_slab(n, n, n, p.Lx, p.Ly, p.Lz, kmax=8)
    c = cpu_port.CpuPort(p)
    c.set_primitive(prim)
    c.vardt()
    with Solver(**pc.solver_kwargs(p)) as g:
        g.set_primitive(prim)
        g.vardt()
        assert abs(g.dt - c.dt) <= 1e-12 * c.dt
        for _ in range(2):
            c.step()
            g.step()
        uu, prim_g = g.get_state()
        cu, cprim = c.get_state()
        for v in range(8):
            assert pc.rel_l2(uu[v], cu[v]) < 1e-11, (v, pc.rel_l2(uu[v], cu[v]))
        for v in range(4):
            assert pc.rel_l2(prim_g[v], cprim[v]) < 1e-10, (v, pc.rel_l2(prim_g[v], cprim[v]))
        assert abs(g.dt - c.dt) <= 1e-12 * c.dt and abs(g.time - c.time) <= 1e-12 * c.time
    c.close()


def test_config2_2048_squared_one_step_against_the_oracle():
    """BASELINE config 2 at full size: 2D compressible Hall-MHD 2048^2 (src_compressible/2D), one step."""
    p, prim = pc.make_case_2d(2048, 2048, hall=True, aeb=False, dealias=1)
    o, g = pc.run_both(p, prim, 1)
    pc.check_state(o, g, 1e-11)
    pc.check_diagnostics(o, g, 1e-9)
    g.close()


def test_config3_256_cubed_incompressible_one_step_against_the_oracle():
    """BASELINE config 3 at full size: 3D incompressible MHD 256^3 decaying turbulence (src_incompressible), one step."""
    n = 256
    p = _params((n, n, n), hall=False, aeb=False)
    p.incompressible = True
    p.rho0 = 1.0
    prim = synthetic.turbulence_slab(n, n, n, p.Lx, p.Ly, p.Lz, kmax=8, drho0=0.0)
    o, g = pc.run_both(p, prim, 1)
    pc.check_state(o, g, 1e-11)
    # max |k.(rho u)^| / rho0 is round-off of three cancelling terms here (the projection keeps the field solenoidal):
    # relative tolerance + the absolute allowance 1e-14 k_max |rho u|_rms, as for max |k.B^| (parity_common.check_diagnostics)
    import numpy as np
    kmax = np.pi * n / p.Lx
    scale = float(np.sqrt(sum(np.mean(o.uu[v] ** 2) for v in (1, 2, 3))))
    assert abs(g.calc_max_divV() - o.calc_max_divV()) <= 1e-9 * o.calc_max_divV() + 1e-14 * kmax * scale
    g.close()
