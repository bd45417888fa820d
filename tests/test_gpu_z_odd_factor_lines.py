"""Line lengths with an odd factor (3 * 2^k and 5 * 2^k, 48 ... 1536) on a B200, through the C ABI, against the oracle
and against the executed-reference vectors — FFTW plans any length (fftw.f90:27-33); the library's composite transform
(fft_core.cuh: P power-of-two transforms side by side and one radix-P exchange) covers the lengths a 2/3-rule grid is
usually given.  Same tolerances as tests/test_gpu_parity.py.

The file sorts after test_gpu_parity.py on purpose: these kernels were added late in round 2 (see DESIGN.md section 6 for
what was run where), and `pytest -x` should reach them after the power-of-two suite."""
import numpy as np
import pytest

import parity_common as pc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(48, 16, 16), (16, 80, 16), (16, 16, 96), (160, 192, 16), (16, 320, 384), (640, 16, 16),
                                   (16, 768, 16), (16, 16, 1280), (1536, 16, 16), (16, 16, 1536)])
def test_fft_forward_inverse_vs_oracle(shape):
    pc.check_fft(*shape)


@pytest.mark.parametrize("shape,kw", [((48, 80, 96), dict(hall=True, aeb=True, dealias=1)),
                                      ((96, 48, 160), dict(hall=True, aeb=True, corot=True, dealias=2, explicit=True, conserve_bg=True)),
                                      ((16, 16, 384), dict(hall=True, aeb=True, dealias=1)),
                                      ((384, 16, 16), dict(hall=False, aeb=False, dealias=0)),
                                      ((16, 640, 16), dict(hall=True, aeb=True, dealias=1))])
def test_one_step_parity(shape, kw):
    p, prim = pc.make_case(*shape, **kw)
    o, g = pc.run_both(p, prim, 1)
    pc.check_state(o, g, 1e-11)
    pc.check_state_vectors(o, g, 1e-11)
    pc.check_diagnostics(o, g, 1e-9)
    g.close()


def test_two_steps_parity_192x96x160():
    """2.9 million points, every axis with an odd factor, two steps through laps_step (dt from the fused CFL pass)."""
    p, prim = pc.make_case(192, 96, 160, hall=True, aeb=True, dealias=1)
    o, g = pc.run_both(p, prim, 2)
    pc.check_state(o, g, 1e-11)
    pc.check_diagnostics(o, g, 1e-9)
    g.close()


@pytest.mark.parametrize("shape,kw", [((384, 320), dict(hall=True, aeb=True, corot=True, dealias=1)),
                                      ((80, 48), dict(hall=True, aeb=True, z_radial=True, dealias=3)),
                                      ((1536, 96), dict(hall=True, aeb=True, dealias=2))])
def test_2d_tree_parity(shape, kw):
    p, prim = pc.make_case_2d(*shape, **kw)
    o, g = pc.run_both(p, prim, 2, t0=2.0 if kw.get("corot") else 0.0)
    pc.check_state(o, g, 1e-11)
    pc.check_diagnostics(o, g, 1e-9)
    g.close()


def test_incompressible_trees_parity():
    p, prim = pc.make_case_incompressible(48, 96, 80, hall=True, aeb=True, dealias=1)
    o, g = pc.run_both(p, prim, 2)
    pc.check_state(o, g, 1e-11)
    g.close()
    p, prim = pc.make_case_incompressible_2d(160, 192, hall=True, aeb=True, corot=True, dealias=1)
    o, g = pc.run_both(p, prim, 2, t0=2.0)
    pc.check_state(o, g, 1e-11)
    g.close()


def test_library_agrees_with_the_executed_reference_source():
    """tests/golden/ref_exec/*lines48* / *lines80*: the reference's own Fortran text executed at 48- and 80-point lines."""
    import test_reference_source_pins as rp
    rp.check_library("lines48_hall_aeb_corot_mask")
    rp.check_library_incompressible("incomp_lines48_hall_aeb_mask")
    rp.check_library_2d("c2d_lines80x48_hall_aeb_filter")
    rp.check_library_incompressible_2d("i2d_lines48x80_hall_aeb_mask")


def test_mask_pruning_is_bit_exact():
    """The work skipping (columns and rows the 1/3 mask removes) at line lengths where n/3 is exact: same bits with it off."""
    pc.check_pruning_is_exact((48, 96, 80), 2, hall=True, aeb=True, dealias=1)


def test_odd_factor_lines_across_ranks_sharing_the_gpu():
    """48- and 80-point lines on the exchanged axes with 4, 3 and 5 ranks as threads of one process (laps_connect_local; the ranks
    share the device, tests/test_gpu_multirank.py): round-robin ky rows, the reference's slabs, remainder planes on the last rank."""
    from test_gpu_multirank import _local
    hall = dict(hall=True, aeb=True, dealias=1)
    _local([dict(world=4, shape=(48, 48, 48), case=hall, steps=2, expect_stride=4),
            dict(world=3, shape=(32, 48, 48), case=hall, steps=2, env=dict(LAPS_TUNE_CYCLIC="0"), expect_stride=1),
            dict(world=5, shape=(32, 80, 48), case=hall, steps=1, expect_stride=5),
            dict(world=2, shape=(48, 80, 32), case=hall, steps=1, incompressible=True)])


def test_unsupported_lengths_are_refused_with_a_message():
    from laps_b200 import LapsError, Solver
    for n in (24, 144, 100):
        p, _ = pc.make_case(64, 64, 64, hall=True, aeb=True, dealias=1)
        kw = pc.solver_kwargs(p)
        kw["nz"] = n
        with pytest.raises(LapsError, match="3 \\* 2\\^k or 5 \\* 2\\^k"):
            Solver(**kw)
