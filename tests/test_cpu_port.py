"""oracle/laps_cpu.c (the C + OpenMP restatement timed as the CPU baseline) against the NumPy oracle."""
import numpy as np
import pytest

import parity_common as pc
from oracle import cpu_port
from oracle import laps_oracle as lo


@pytest.mark.parametrize("kw", [dict(hall=True, aeb=True, dealias=1), dict(hall=False, aeb=True, corot=True, dealias=2),
                                dict(hall=True, aeb=False, explicit=True, conserve_bg=True), dict(hall=False, aeb=False, dealias=0, visc=False, resis=False)])
def test_c_port_matches_the_numpy_oracle(kw):
    p, prim = pc.make_case(32, 16, 16, **kw)
    o = lo.State(p)
    o.set_primitive(prim)
    c = cpu_port.CpuPort(p)
    c.set_primitive(prim)
    assert abs(c.vardt() - o.vardt()) <= 1e-13 * o.dt
    for _ in range(3):
        o.step()
        c.step()
    uu, prim_c = c.get_state()
    for v in range(8):
        assert pc.rel_l2(uu[v], o.uu[v]) < 1e-11, (v, pc.rel_l2(uu[v], o.uu[v]))
    for v in range(4):
        assert pc.rel_l2(prim_c[v], o.uu_prim[v]) < 1e-10
    assert abs(c.dt - o.dt) <= 1e-12 * o.dt and abs(c.time - o.time) <= 1e-12 * o.time
    assert c.threads >= 1
    c.close()


@pytest.mark.parametrize("name", ["hall_aeb_mask", "corot_filter_explicit"])
def test_cpu_port_against_the_executed_reference_source(name):
    """The C + OpenMP port that bench.py times as the CPU baseline, against the golden vectors made by executing the
    reference's own Fortran source (tests/golden/make_ref_exec_fixtures.py): two steps of the Principal loop."""
    import test_reference_source_pins as rp
    g, p = rp.load_case(name)
    c = cpu_port.CpuPort(p)
    c.set_primitive(g["prim0"])
    assert abs(c.vardt() - float(g["dt0"])) <= 1e-13 * float(g["dt0"])
    for i in range(len(g["dt"])):
        c.step()
        assert abs(c.dt - g["dt"][i]) <= 1e-12 * c.dt
    uu, prim = c.get_state()
    for v in range(8):
        assert pc.rel_l2(uu[v], g["uu"][v]) < 1e-11, (v, pc.rel_l2(uu[v], g["uu"][v]))
    for v in range(4):
        assert pc.rel_l2(prim[v], g["uu_prim"][v]) < 1e-10, v
    c.close()
