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
