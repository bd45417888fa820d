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
    "rhs2=rhs:2",
    "noscreen=screen:0",
    "nospec=spec:0",
    "overlap=overlap:1",
]
RESET = dict(rhs=1, cgz=0, rcg=0, z=3, spec=1, overlap=-1, ovl_chunks=3, ovl_push=2, ovl_y=16, ovl_z=12, screen=1, tly=0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--rounds", type=int, default=3, help="passes over the whole variant list (the variants are interleaved: clocks drift under the power cap)")
    ap.add_argument("--variants", nargs="*", default=DEFAULT)
    ap.add_argument("--reset", default="", help="extra k:v defaults to restore between variants")
    args = ap.parse_args()
    import torch
    n = args.n
    kw = bench.workload_params(n)
    args.gpus = int(os.environ.get("WORLD_SIZE", "1"))
    R = bench.Ranks(args)                              # one process per GPU under torchrun, like bench.py
    g = Solver(rank=R.rank, nranks=R.world, device=R.local, **kw)
    R.connect(g)
    R.barrier()
    stream = torch.cuda.ExternalStream(g.cuda_stream(), device=torch.device("cuda", R.local))
    prim = synthetic.turbulence_slab(n, n, n, kw["Lx"], kw["Ly"], kw["Lz"], z_offset=g.ext.z_offset, z_size=g.ext.z_size,
                                     kmax=min(8, n // 2 - 1))
    g.set_primitive(prim)
    g.vardt()
    for _ in range(3):
        g.step()
    reset = dict(RESET)
    for kv in filter(None, args.reset.split(",")):
        k, v = kv.split(":")
        reset[k] = int(v)
    results = {}
    for rnd in range(args.rounds):
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
            R.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                g.step()
            e1.record(stream)
            g.sync()
            R.barrier()
            r = results.setdefault(name, {"tune": spec, "ms": [], "prof": {}})
            r["ms"].append(R.reduce(e0.elapsed_time(e1), "max") / args.steps)
            if rnd == args.rounds - 1:
                g.set_profiling(True)
                g.step()
                for nm, t, by in g.get_profile(with_bytes=True):
                    a = r["prof"].setdefault(nm, [0.0, 0, 0.0])
                    a[0] += t; a[1] += 1; a[2] += by
                g.set_profiling(False)
    finite = bool(np.isfinite(g.calc_rms()[0]).all())
    R.barrier()
    g.close()
    if R.rank != 0:
        R.close()
        return
    for name, r in results.items():
        prof = r["prof"]
        print(json.dumps({"variant": name, "tune": r["tune"], "ms_per_step_median": round(float(np.median(r["ms"])), 3),
                          "ms_per_step_min": round(min(r["ms"]), 3), "ms_per_step_all": [round(x, 2) for x in r["ms"]],
                          "kernels_ms_per_launch": {k: round(v[0] / v[1], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
                          "kernels_GBps": {k: round(v[2] / v[1] / (v[0] / v[1] * 1e-3) / 1e9) for k, v in prof.items() if v[2] > 0},
                          "finite": finite, "n_gpus": R.world}), flush=True)
    R.close()


if __name__ == "__main__":
    main()
