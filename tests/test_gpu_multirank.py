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
# Each case runs tests/local_ranks.py in a child process with CUDA_DEVICE_MAX_CONNECTIONS=32 (see there).
# ---------------------------------------------------------------------------------------------------------------
def _local(cfg, timeout=900):
    import json
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    # EAGER module loading: the first launch of a kernel otherwise loads it lazily, which synchronises the device — behind the
    # flag kernel of the rank that is waiting for exactly this launch when the ranks share one GPU
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32", CUDA_MODULE_LOADING="EAGER", LAPS_ORACLE_WORKERS="4")
    out = subprocess.run([sys.executable, os.path.join(here, "local_ranks.py"), json.dumps(cfg)], env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=timeout)
    assert out.returncode == 0 and "local ranks ok" in out.stdout, out.stdout[-4000:]


_HALL = dict(hall=True, aeb=True, dealias=1)


def test_connect_local_threads_parity():
    """Default ownership and schedule: decompose_1d slabs up to 3 ranks, round-robin ky rows from 4 ranks on (2, 3, 4, 8 ranks);
    the reference's slabs forced at 4 and 8 ranks."""
    _local([dict(world=2, shape=(64, 64, 64), case=_HALL, steps=2, expect_stride=1),
            dict(world=4, shape=(64, 64, 64), case=_HALL, steps=2, expect_stride=4),
            dict(world=8, shape=(64, 64, 64), case=_HALL, steps=2, expect_stride=8),
            dict(world=4, shape=(64, 64, 64), case=_HALL, steps=2, env=dict(LAPS_TUNE_CYCLIC="0"), expect_stride=1),
            dict(world=8, shape=(64, 64, 64), case=_HALL, steps=2, env=dict(LAPS_TUNE_CYCLIC="0"), expect_stride=1),
            dict(world=3, shape=(32, 64, 32), case=_HALL, steps=2, expect_stride=1)])


def test_connect_local_threads_stage_schedules():
    """LAPS_TUNE_OVERLAP=0 / 1 / 2 on real streams: one stream (forced at 8 ranks, where form 1 is the default), form 1 (forced at
    2 ranks, default at 8 above), form 2 at 2, 3 and 8 ranks (use_overlap in csrc/solver.cu)."""
    _local([dict(world=8, shape=(64, 64, 64), case=_HALL, steps=2, env=dict(LAPS_TUNE_OVERLAP="0")),
            dict(world=2, shape=(64, 64, 64), case=_HALL, steps=2, env=dict(LAPS_TUNE_OVERLAP="1")),
            dict(world=2, shape=(64, 64, 64), case=_HALL, steps=2, env=dict(LAPS_TUNE_OVERLAP="2")),
            dict(world=8, shape=(64, 64, 64), case=_HALL, steps=2, env=dict(LAPS_TUNE_OVERLAP="2")),
            dict(world=3, shape=(32, 64, 32), case=_HALL, steps=2, env=dict(LAPS_TUNE_OVERLAP="2"))])


def test_connect_local_threads_other_physics():
    """Corotation + filter dealiasing (no pruning) on 3 ranks with a remainder slab; incompressible tree on 2."""
    _local([dict(world=3, shape=(32, 64, 32), case=dict(hall=True, aeb=True, corot=True, dealias=2), steps=1),
            dict(world=2, shape=(64, 64, 64), case=_HALL, steps=2, incompressible=True)])


def test_exchange_wait_is_bounded():
    _local(dict(mode="bounded_wait"))


@pytest.mark.parametrize("form", ["1", "2"])
def test_two_gpus_two_stream_schedules(form):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    run_ranks(2, dict(backend="nccl", shape=(64, 64, 64), case=dict(hall=True, aeb=True, dealias=1), steps=2, env=dict(LAPS_TUNE_OVERLAP=form)))
