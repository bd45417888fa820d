"""All ranks of a slab-decomposed run inside ONE process, one host thread per handle, wired with
laps_connect_local (include/laps_b200.h) — the single-process multi-GPU mode of the boundary.  `devices` may name the
same GPU for every rank: the peer stores then stay on one device, but the decomposition tables, the transpose index
maps (contiguous ky slabs or round-robin rows), the device-side flag barriers and the allreduce are exactly those of
an N-GPU run, so a box with a single GPU checks the multi-rank path against the single-grid oracle too.
GPU only (the test-only kernel emulator is not thread-safe).

Ranks that SHARE a device need every stream of every handle in a hardware work queue of its own: with the default of 8
queues (CUDA_DEVICE_MAX_CONNECTIONS) the 2 x N streams alias, and a flag kernel that spins in one stream then holds up the
kernels of the rank it is waiting for — a deadlock that one-GPU-per-rank runs cannot have.  tests/test_gpu_multirank.py
therefore runs this module in a child process with CUDA_DEVICE_MAX_CONNECTIONS=32 (it must be set before the CUDA
context exists) and CUDA_MODULE_LOADING=EAGER (a lazily loaded kernel synchronises the device at its first launch — behind the
flag kernel of the rank that waits for that very launch):   python tests/local_ranks.py '<json>'."""
import os
import sys
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.dirname(_HERE), _HERE):      # run as a script by tests/test_gpu_multirank.py
    if _p not in sys.path:
        sys.path.insert(0, _p)

import parity_common as pc  # noqa: E402
from laps_b200 import Solver  # noqa: E402


class _Env:
    def __init__(self, env):
        self.env, self.saved = dict(env or {}), {}

    def __enter__(self):
        for k, v in self.env.items():
            self.saved[k] = os.environ.get(k)
            os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def make_solvers(world, p, devices=None, lib_path=None, env=None):
    devices = list(devices) if devices is not None else [0] * world
    base = {"LAPS_XCHG_TIMEOUT_S": os.environ.get("LAPS_XCHG_TIMEOUT_S", "30")}
    base.update(env or {})
    with _Env(base):
        gs = [Solver(lib_path, rank=r, nranks=world, device=devices[r], **pc.solver_kwargs(p)) for r in range(world)]
    Solver.connect_local(gs)
    return gs


def run_threads(gs, work):
    """work(rank, solver) on one thread per handle; returns the list of results, re-raises the first failure."""
    out, err = [None] * len(gs), [None] * len(gs)

    def body(r):
        try:
            out[r] = work(r, gs[r])
        except BaseException as e:  # noqa: BLE001 - reported below
            err[r] = e

    ts = [threading.Thread(target=body, args=(r,)) for r in range(len(gs))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out


_ORACLE = {}    # the single-grid oracle run of a case, shared by the decompositions checked against it in one process


def run_local(world, shape, case, steps=2, devices=None, lib_path=None, env=None, incompressible=False, tol=1e-11,
              expect_stride=None):
    """`steps` Principal-loop steps on `world` ranks in this process against the single-grid oracle: fields and spectra
    of every rank's slabs within `tol` (relative L2), dt and diagnostics identical on every rank."""
    key = (tuple(shape), tuple(sorted(case.items())), steps, incompressible)
    if key not in _ORACLE:
        make = pc.make_case_incompressible if incompressible else pc.make_case
        p, prim = make(*shape, **case)
        o = pc.oracle_state(p)
        o.set_primitive(prim)
        o.vardt()
        dt0 = o.dt
        for _ in range(steps):
            o.step()
        _ORACLE[key] = (p, prim, o, dt0)
    p, prim, o, dt0 = _ORACLE[key]
    gs = make_solvers(world, p, devices, lib_path, env)
    if expect_stride is not None:
        assert all(g.ext.y_stride == expect_stride for g in gs), [g.ext.y_stride for g in gs]

    def work(r, g):
        zo, zn = g.ext.z_offset, g.ext.z_size
        g.set_primitive(prim[:, zo:zo + zn])
        g.vardt()
        d0 = g.dt
        for _ in range(steps):
            g.step()
        uu, _ = g.get_state()
        return dict(dt0=d0, dt=g.dt, uu=uu, uf=g.uu_fourier(), rms=np.concatenate(g.calc_rms()), inv=g.invariants(),
                    nan=g.checkNan())

    try:
        res = run_threads(gs, work)
        worst = 0.0
        for r, (g, q) in enumerate(zip(gs, res)):
            zo, zn, rows = g.ext.z_offset, g.ext.z_size, g.ky_rows
            assert abs(q["dt0"] - dt0) <= 1e-13 * dt0 and abs(q["dt"] - o.dt) <= 1e-12 * o.dt, (q["dt0"], dt0, q["dt"], o.dt)
            for v in range(8):
                e = pc.rel_l2(q["uu"][v], o.uu[v, zo:zo + zn])
                worst = max(worst, e)
                assert e < tol, (r, v, e)
                ref = o.uu_fourier[v][:, rows, :]
                scale = max(1.0, np.linalg.norm(o.uu_fourier[v]) / max(np.linalg.norm(ref), 1e-300))
                assert pc.rel_l2(q["uf"][v], ref) < tol * scale, (r, v)
            assert not q["nan"]
            assert np.array_equal(q["rms"], res[0]["rms"]) and np.array_equal(q["inv"], res[0]["inv"]) and q["dt"] == res[0]["dt"]
        oave, orms, oru2 = o.calc_rms()
        assert np.allclose(res[0]["rms"][:8], oave, rtol=1e-9, atol=1e-12)
        assert np.allclose(res[0]["rms"][8:16], orms, rtol=1e-9, atol=1e-15)
        assert abs(res[0]["inv"][0] - o.invariants()[0]) <= 1e-9 * abs(o.invariants()[0])
        return worst
    finally:
        for g in gs:
            g.close()


def check_bounded_wait(devices):
    """A rank that never makes the matching collective call must not wedge its peers (exchange.cuh): the waiting rank's
    flag kernel gives up after LAPS_XCHG_TIMEOUT_S, the failure reaches the host through laps_last_error, and the abort
    word it raises fails the other rank's next call as well."""
    import time
    from laps_b200 import capi
    p, prim = pc.make_case(32, 32, 32, hall=True, aeb=True)
    gs = make_solvers(2, p, devices, env=dict(LAPS_XCHG_TIMEOUT_S="1.0"))

    def fails(fn, pattern):
        try:
            fn()
        except capi.LapsError as e:
            assert pattern in str(e), str(e)
            return
        raise AssertionError("did not raise a LapsError")

    try:
        t0 = time.perf_counter()
        fails(lambda: gs[0].set_primitive(prim[:, :16]), "ran out of its budget")     # collective; rank 1 never calls it
        assert time.perf_counter() - t0 < 20.0
        fails(lambda: gs[0].vardt(), "abort")                                         # the handle stays dead

        def other():                                                                  # and the peer learns about it at its next wait
            gs[1].set_primitive(prim[:, 16:])
            gs[1].sync()
        fails(other, "abort")
    finally:
        for g in gs:
            g.close()


if __name__ == "__main__":
    import json
    import sys
    cfgs = json.loads(sys.argv[1])
    if isinstance(cfgs, dict):
        cfgs = [cfgs]
    import torch
    ndev = torch.cuda.device_count()
    for cfg in cfgs:
        if cfg.get("mode") == "bounded_wait":
            check_bounded_wait([r % ndev for r in range(2)])
            print("bounded wait ok", flush=True)
            continue
        world = cfg["world"]
        worst = run_local(world, tuple(cfg["shape"]), cfg.get("case", {}), steps=cfg.get("steps", 2), devices=[r % ndev for r in range(world)],
                          env=cfg.get("env"), incompressible=cfg.get("incompressible", False), expect_stride=cfg.get("expect_stride"))
        print(f"{world} ranks {cfg.get('env') or ''}: max rel L2 {worst:.3e}", flush=True)
    print("local ranks ok")
