#!/usr/bin/env python
"""A/B of the equivalent kernels / launch shapes (laps_set_tune) on ONE resident state: for every variant a few steps
are timed with CUDA events on the library's stream, then one instrumented step gives the per-kernel times.
Usage: python tools/ab_tune.py [--n 512] [--steps 3] [--variants name=k:v,k:v ...]   (one JSON line per variant)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from laps_b200 import Solver, synthetic  # noqa: E402

DEFAULT = [
    "default=",
    "rhs0=rhs:0",
    "rhs0_cgz1=rhs:0,cgz:1",
    "rhs0_cgz4=rhs:0,cgz:4",
    "rcg2=rcg:2",
    "z0=z:0", "z1=z:1", "z3=z:3", "z4=z:4", "z7=z:7",
    "nospec=spec:0",
    "overlap=overlap:1",
    "overlap_c2=overlap:1,ovl_chunks:2",
    "overlap_c4=overlap:1,ovl_chunks:4",
]
RESET = dict(rhs=1, cgz=0, rcg=0, z=3, spec=1, overlap=-1, ovl_chunks=3, ovl_y=16, ovl_z=8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--variants", nargs="*", default=DEFAULT)
    ap.add_argument("--reset", default="", help="extra k:v defaults to restore between variants")
    args = ap.parse_args()
    import torch
    n = args.n
    kw = bench.workload_params(n)
    g = Solver(**kw)
    stream = torch.cuda.ExternalStream(g.cuda_stream())
    prim = synthetic.turbulence_slab(n, n, n, kw["Lx"], kw["Ly"], kw["Lz"], kmax=min(8, n // 2 - 1))
    g.set_primitive(prim)
    g.vardt()
    for _ in range(3):
        g.step()
    reset = dict(RESET)
    for kv in filter(None, args.reset.split(",")):
        k, v = kv.split(":")
        reset[k] = int(v)
    for var in args.variants:
        name, _, spec = var.partition("=")
        for k, v in reset.items():
            try:
                g.set_tune(k, v)
            except Exception:
                pass
        for kv in filter(None, spec.split(",")):
            k, v = kv.split(":")
            g.set_tune(k, int(v))
        g.step()                                   # settle (the speculative front half of the previous variant is discarded)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            g.step()
        e1.record(stream)
        g.sync()
        ms = e0.elapsed_time(e1) / args.steps
        g.set_profiling(True)
        g.step()
        prof = {}
        for nm, t, by in g.get_profile(with_bytes=True):
            a = prof.setdefault(nm, [0.0, 0, 0.0])
            a[0] += t; a[1] += 1; a[2] += by
        g.set_profiling(False)
        uu0 = g.calc_rms()[0]
        print(json.dumps({"variant": name, "tune": spec, "ms_per_step": round(ms, 3),
                          "kernels_ms_per_launch": {k: round(v[0] / v[1], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
                          "kernels_GBps": {k: round(v[2] / v[1] / (v[0] / v[1] * 1e-3) / 1e9) for k, v in prof.items() if v[2] > 0},
                          "finite": bool(np.isfinite(uu0).all())}), flush=True)
    g.close()


if __name__ == "__main__":
    main()
