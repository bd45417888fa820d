#!/usr/bin/env python
"""cuFFT as a comparison point (BASELINE.json north_star: "cuFFT is timed alongside only as a comparison
point"): torch.fft.rfftn / irfftn (cuFFT D2Z / Z2D, FP64) of one n^3 field on the GPU, against this library's
own unit transforms (laps_fft_forward / laps_fft_inverse time the same passes the RK stage uses, without the
host copies: measured through the per-launch profile).  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laps_b200 import Solver  # noqa: E402


def main(n=512, reps=10):
    dev = torch.device("cuda", 0)
    a = torch.randn(n, n, n, dtype=torch.float64, device=dev)
    for _ in range(2):
        w = torch.fft.rfftn(a)
        b = torch.fft.irfftn(w, s=(n, n, n))
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(reps):
        w = torch.fft.rfftn(a)
    e[1].record()
    for _ in range(reps):
        b = torch.fft.irfftn(w, s=(n, n, n))
    e[2].record()
    torch.cuda.synchronize()
    cufft_fwd, cufft_inv = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
    err = float((b - a).abs().max())
    del a, b, w
    torch.cuda.empty_cache()
    # this library: 8 fields per call, unpruned (dealias_option 0), device time of the passes only
    with Solver(nx=n, ny=n, nz=n, Lx=1.0, Ly=1.0, Lz=1.0, dealias_option=0) as s:
        x = np.random.default_rng(0).standard_normal((8,) + s.real_shape)
        s.set_profiling(True)
        spec = s.fft_forward(x)
        first = s.get_profile()
        prof_f = dict()
        for name, ms in first:
            prof_f[name] = prof_f.get(name, 0.0) + ms
        y = s.fft_inverse(spec)
        prof_i = dict()
        for name, ms in s.get_profile()[len(first):]:       # the profile accumulates until the next laps_evolve
            prof_i[name] = prof_i.get(name, 0.0) + ms
        rt = float(np.abs(y - x).max())
    ours_fwd = sum(v for k, v in prof_f.items() if k.startswith(("fwd_x", "fwd_y", "fwd_z"))) / 8
    ours_inv = sum(v for k, v in prof_i.items() if k.startswith(("inv_z", "inv_y", "inv_x"))) / 8 - 0.0
    print(json.dumps({"n": n, "cufft_d2z_ms_per_field": cufft_fwd, "cufft_z2d_ms_per_field": cufft_inv, "cufft_roundtrip_err": err,
                      "laps_forward_ms_per_field": ours_fwd, "laps_inverse_ms_per_field": ours_inv, "laps_roundtrip_err": rt,
                      "laps_forward_profile": prof_f, "laps_inverse_profile": prof_i,
                      "note": "unpruned transforms of a full spectrum; inside an RK stage the passes skip the dealiased modes and are fused with the pointwise work"}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 512)
