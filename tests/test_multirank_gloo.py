"""World-size 2 and 3 runs of the slab-decomposed path on CPU: one process per rank, wired through
torch.distributed (gloo) exactly as bench.py wires GPUs through NCCL, running the UNCHANGED kernel
sources on the test-only emulator (whose "device memory" is POSIX shared memory, so the ranks
really store into each other's exchange buffers and synchronise with the device-side flags).
Checks decompose_1d tables incl. the remainder-on-last-rank rule, the transpose_yz index map,
distributed FFT parity, one-step parity against the single-grid oracle, and the allreduces."""
import json
import os
import socket
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import build_emu  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_ranks(world, cfg, timeout=600):
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1", LAPS_ORACLE_WORKERS="2", **cfg.get("env", {}))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "mp_worker.py"), json.dumps(cfg)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    try:
        for p in procs:
            out, _ = p.communicate(timeout=timeout)
            outs.append(out)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{out[-3000:]}"
        assert f"rank {r}/{world} ok" in out


@pytest.fixture(scope="module")
def emu():
    return build_emu.build()


def test_two_ranks_hall_aeb(emu):
    run_ranks(2, dict(lib=emu, shape=(16, 16, 16), case=dict(hall=True, aeb=True, dealias=1), steps=1))


def test_three_ranks_remainder_on_last_rank(emu):
    # ny = nz = 16 over 3 ranks: slabs of 5, 5, 6 (parallel.f90:326-349)
    run_ranks(3, dict(lib=emu, shape=(16, 16, 16), case=dict(hall=True, aeb=True, corot=True, dealias=2), steps=1))


def test_two_ranks_incompressible_tree(emu):
    # src_incompressible: pressure projection, 12 inverse + 6 forward transforms per stage through the same exchange
    run_ranks(2, dict(lib=emu, shape=(16, 16, 16), incompressible=True, case=dict(hall=True, aeb=True, dealias=1), steps=1))


def test_eight_ranks_two_of_them_without_surviving_columns(emu):
    # ny = 32 over 8 ranks with the 1/3 mask (|ky| <= 10 survives): ranks 3 and 4 own rows 12..19 only, i.e. no
    # column the z pass has to visit — the situation of the 512^3 benchmark on 8 GPUs
    # (LAPS_TUNE_CYCLIC=0 pins the reference's contiguous slabs; from 4 ranks on the default deals the rows round-robin)
    run_ranks(8, dict(lib=emu, shape=(16, 32, 16), case=dict(hall=True, aeb=True, dealias=1), steps=2, env=dict(LAPS_TUNE_CYCLIC="0")))


def test_four_ranks_default_ownership_is_cyclic(emu):
    # no switch set: with a masked dealiasing option and >= 4 ranks the library deals the ky rows round-robin
    # (mp_worker checks ky_rows == rank, rank + P, ... and the matching transpose index map when y_stride > 1)
    run_ranks(4, dict(lib=emu, shape=(16, 16, 16), case=dict(hall=True, aeb=True, dealias=1), steps=1, expect_stride=4))


def test_four_ranks_filter_dealiasing_keeps_reference_slabs(emu):
    # dealias option 2 prunes nothing, so the reference's decompose_1d slabs stay (y_stride = 1)
    run_ranks(4, dict(lib=emu, shape=(16, 16, 16), case=dict(hall=True, aeb=True, dealias=2), steps=1, expect_stride=1))


@pytest.mark.parametrize("world,shape", [(2, (16, 16, 16)), (3, (16, 16, 16)), (8, (16, 32, 16))])
def test_cyclic_ky_ownership(emu, world, shape):
    """LAPS_TUNE_CYCLIC=1: Fourier rows dealt round-robin to the ranks (the default from 4 ranks on).  Same results as the slab
    ownership — the oracle parity, the distributed FFT and the allreduces of mp_worker hold unchanged."""
    run_ranks(world, dict(lib=emu, shape=shape, case=dict(hall=True, aeb=True, dealias=1), steps=2, env=dict(LAPS_TUNE_CYCLIC="1")))


def test_cyclic_ky_ownership_incompressible(emu):
    run_ranks(2, dict(lib=emu, shape=(16, 16, 16), incompressible=True, case=dict(hall=True, aeb=True, dealias=2), steps=1,
                      env=dict(LAPS_TUNE_CYCLIC="1")))


def test_four_ranks_incompressible_default_ownership(emu):
    # the incompressible tree at 4 ranks with the spherical mask: round-robin rows by default (12 inverse + 6 forward
    # transforms per stage and the projection kernel over the strided rows)
    run_ranks(4, dict(lib=emu, shape=(16, 16, 16), incompressible=True, case=dict(hall=True, aeb=True, dealias=1), steps=1,
                      expect_stride=4))


def test_eight_ranks_default_path_whole_z_tiles(emu):
    # the 8-GPU benchmark situation in small: 64^3 over 8 ranks (8 planes and 8 round-robin rows per rank, i.e. whole
    # 8-line z tiles in the y passes), Hall + expanding box + spherical mask, two steps through laps_step
    run_ranks(8, dict(lib=emu, shape=(64, 64, 64), case=dict(hall=True, aeb=True, dealias=1), steps=2, expect_stride=8), timeout=1200)


def test_five_ranks_uneven_round_robin_rows(emu):
    # 64 rows over 5 ranks: 13, 13, 13, 13, 12 round-robin rows; z planes 12, 12, 12, 12, 16 (decompose_1d remainder)
    run_ranks(5, dict(lib=emu, shape=(32, 64, 32), case=dict(hall=True, aeb=True, dealias=1), steps=2, expect_stride=5), timeout=1200)


@pytest.mark.parametrize("world,shape,env", [(4, (16, 48, 16), {}),                        # 48 round-robin ky rows, 12 per rank
                                             (3, (16, 16, 48), dict(LAPS_TUNE_CYCLIC="0")),   # 48-point z lines stored across 3 slabs of 16 planes
                                             (5, (16, 80, 48), {})])                          # 16 rows per rank; planes 9, 9, 9, 9, 12
def test_odd_factor_lines_across_ranks(emu, world, shape, env):
    """Line lengths with an odd factor (3 * 16, 5 * 16) on the exchanged axes: the composite transforms of fft_core.cuh store
    through the same peer tables as the power-of-two ones."""
    run_ranks(world, dict(lib=emu, shape=shape, case=dict(hall=True, aeb=True, dealias=1), steps=1, env=env), timeout=1200)


def test_absent_rank_does_not_wedge_the_others(emu):
    # rank 1 never calls the collective: ranks 0 and 2 must come back with an error inside the time budget
    run_ranks(3, dict(lib=emu, shape=(16, 16, 16), case=dict(hall=True, aeb=True, dealias=1), absent_rank=1,
                      env=dict(LAPS_XCHG_TIMEOUT_S="1.0")), timeout=120)


@pytest.mark.parametrize("world,form", [(1, "1"), (2, "1"), (4, "1"), (1, "2"), (2, "2"), (3, "2"), (8, "2")])
def test_two_stream_schedule_in_sequence(emu, world, form):
    """The two-stream stage schedules (LAPS_TUNE_OVERLAP; opt-in, see use_overlap in csrc/solver.cu) give the same state as
    the oracle.  Form 1: the y pass and the z passes store into the peers' buffers from the exchange stream, in field
    chunks / row groups, with grid-capped launches.  Form 2: the y pass stores into local staging blocks and a copy
    kernel on the exchange stream (k_xchg_push) does transpose_yz, chunk by chunk, with the flag barriers as events.
    The emulator runs launches synchronously, so this checks the index math, the chunk / group bookkeeping and the
    barrier pairing, not the stream ordering (tests/test_gpu_multirank.py does that on hardware)."""
    shape = (16, 16, 16) if world != 3 else (16, 32, 16)
    run_ranks(world, dict(lib=emu, shape=shape, case=dict(hall=True, aeb=True, dealias=1), steps=2, env=dict(LAPS_TUNE_OVERLAP=form)))
    if world == 2:   # no Hall term, no dealiasing mask (the continuity row is then an ordinary RHS row), 2 chunks
        run_ranks(world, dict(lib=emu, shape=(16, 16, 16), case=dict(hall=False, aeb=False, dealias=0), steps=2,
                              env=dict(LAPS_TUNE_OVERLAP=form, LAPS_TUNE_OVL_CHUNKS="2")))
