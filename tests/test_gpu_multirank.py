"""Slab decomposition over 2 (and 4) real GPUs, one process per GPU wired through NCCL like
bench.py: peer stores over NVLink + device-side flag barriers, checked against the single-grid
oracle (tests/mp_worker.py).  Skipped on boxes with a single GPU."""
import pytest

from test_multirank_gloo import run_ranks

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_one_step_parity(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    run_ranks(world, dict(backend="nccl", shape=(64, 64, 64), case=dict(hall=True, aeb=True, dealias=1), steps=2))


@pytest.mark.parametrize("world", [4, 8])
def test_multi_gpu_reference_slabs(world):
    """The reference's contiguous ky slabs at 4 and 8 GPUs (the default there deals the rows round-robin)."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    run_ranks(world, dict(backend="nccl", shape=(64, 64, 64), case=dict(hall=True, aeb=True, dealias=1), steps=2,
                          env=dict(LAPS_TUNE_CYCLIC="0")))


def test_three_gpus_remainder_on_last_rank():
    if _ngpu() < 3:
        pytest.skip("needs 3 GPUs")
    run_ranks(3, dict(backend="nccl", shape=(32, 64, 32), case=dict(hall=True, aeb=True, corot=True, dealias=2), steps=1))


def test_two_gpus_incompressible_tree():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    run_ranks(2, dict(backend="nccl", shape=(64, 64, 64), incompressible=True, case=dict(hall=True, aeb=True, dealias=1), steps=2))


# ---------------------------------------------------------------------------------------------------------------
# The same decomposition with every rank in ONE process (laps_connect_local, one host thread per handle).  The ranks
# may share a device, so these run — and check the 4- and 8-rank ownership forms against the oracle — on a 1-GPU box.
# ---------------------------------------------------------------------------------------------------------------
def _devices(world):
    n = _ngpu()
    return [r % n for r in range(world)]


@pytest.mark.parametrize("world,stride", [(2, 1), (3, 1), (4, 4), (8, 8)])
def test_connect_local_threads_parity(world, stride):
    """Default ownership: decompose_1d slabs up to 3 ranks, round-robin ky rows from 4 ranks on."""
    from local_ranks import run_local
    shape = (64, 64, 64) if world != 3 else (32, 64, 32)
    run_local(world, shape, dict(hall=True, aeb=True, dealias=1), steps=2, devices=_devices(world), expect_stride=stride)


@pytest.mark.parametrize("world", [4, 8])
def test_connect_local_threads_reference_slabs(world):
    from local_ranks import run_local
    run_local(world, (64, 64, 64), dict(hall=True, aeb=True, dealias=1), steps=2, devices=_devices(world),
              env=dict(LAPS_TUNE_CYCLIC="0"), expect_stride=1)


def test_connect_local_threads_other_physics():
    """Corotation + filter dealiasing (no pruning) on 3 ranks with a remainder slab; incompressible tree on 2."""
    from local_ranks import run_local
    run_local(3, (32, 64, 32), dict(hall=True, aeb=True, corot=True, dealias=2), steps=1, devices=_devices(3))
    run_local(2, (64, 64, 64), dict(hall=True, aeb=True, dealias=1), steps=2, devices=_devices(2), incompressible=True)


def test_exchange_wait_is_bounded():
    """A rank that never makes the matching collective call must not wedge its peers (exchange.cuh): the waiting rank's
    flag kernel gives up after LAPS_XCHG_TIMEOUT_S, the failure reaches the host through laps_last_error, and the abort
    word it raises fails the other rank's next call as well."""
    import time
    import parity_common as pc
    from laps_b200 import capi
    from local_ranks import make_solvers
    p, prim = pc.make_case(32, 32, 32, hall=True, aeb=True)
    gs = make_solvers(2, p, _devices(2), env=dict(LAPS_XCHG_TIMEOUT_S="1.0"))
    try:
        t0 = time.perf_counter()
        with pytest.raises(capi.LapsError, match="ran out of its budget"):
            gs[0].set_primitive(prim[:, :16])       # collective; rank 1 never calls it
        assert time.perf_counter() - t0 < 20.0
        with pytest.raises(capi.LapsError, match="abort"):
            gs[0].vardt()                           # the handle stays dead
        with pytest.raises(capi.LapsError, match="abort"):
            gs[1].set_primitive(prim[:, 16:])       # and the peer learns about it at its next wait
            gs[1].sync()
    finally:
        for g in gs:
            g.close()
