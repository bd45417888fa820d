"""The executable model of the in-shared-memory FFT index math (tools/fft_model.py mirrors
laps_b200/csrc/fft_core.cuh) against numpy.fft, for every supported line length."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import fft_model as fm  # noqa: E402


@pytest.mark.parametrize("N", [16, 32, 64, 128, 256, 512, 1024, 2048])
def test_staged_fft_matches_numpy(N):
    rng = np.random.default_rng(N)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    assert np.abs(fm.fft_model(x, -1) - np.fft.fft(x)).max() < 1e-12 * N
    assert np.abs(fm.fft_model(x, +1) - np.fft.ifft(x) * N).max() < 1e-12 * N


def test_smem_padding_is_conflict_free_for_tile_mappings():
    for N, TL in ((512, 8), (512, 4), (64, 8)):
        res = fm.report_conflicts(N, TL)
        # mapping A (threads walk a line) must be conflict free in every stage
        assert all(avg <= 1.0001 for (s, m), (avg, worst) in res.items() if m == "A"), res


@pytest.mark.parametrize("N", [48, 80, 96, 160, 192, 320, 384, 640, 768, 1280, 1536])
def test_composite_fft_matches_numpy(N):
    """Line lengths with an odd factor: the register identity (thread u holds the stage-0 inputs of lane u // P of class u % P),
    the disjoint parts of the padded line and the radix-P exchange, as Fft<N, DIR, P> does them."""
    rng = np.random.default_rng(N)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    assert np.abs(fm.fft_model_odd(x, -1) - np.fft.fft(x)).max() < 1e-12 * N
    assert np.abs(fm.fft_model_odd(x, +1) - np.fft.ifft(x) * N).max() < 1e-12 * N
