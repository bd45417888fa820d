"""Randomised parity sweep: random source tree, grid (power-of-two and odd-factor line lengths, 8-point line axis), physics
switches and step count; the library against the oracle at 1e-11 relative L2 per field.  By default on the test-only kernel
emulator (no GPU needed); with --gpu through the real library.  Development tool; tests/test_emulated_kernels.py runs a short
fixed-seed sweep of it.

    python tools/fuzz_parity.py --seed 7 --seconds 1500        # round 2: 471 cases, 0 failures
    python tools/fuzz_parity.py --ranks --seed 3 --seconds 1200   # 2 - 8 ranks over gloo on the emulator: 108 cases, 0 failures
"""
import argparse
import os
import random
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu")):
    if d not in sys.path:
        sys.path.insert(0, d)

SIZES = [16, 32, 48, 64, 80, 96]


def random_case(rng, max_points=48 * 48 * 32):
    tree = rng.choice(["c3", "c3", "c2", "i3", "i2"])
    kw = dict(hall=rng.random() < 0.7, aeb=rng.random() < 0.7)
    kw["corot"] = kw["aeb"] and rng.random() < 0.4
    kw["dealias"] = rng.choice([0, 1, 2] if tree.endswith("3") else [0, 1, 2, 3])
    if rng.random() < 0.3:
        kw.update(explicit=True)
    if rng.random() < 0.3:
        kw.update(conserve_bg=True)
    if tree == "c2" and not kw["corot"] and kw["aeb"] and rng.random() < 0.3:
        kw["z_radial"] = True
    if tree.endswith("3"):
        while True:
            shape = (rng.choice(SIZES), rng.choice(SIZES), rng.choice(SIZES + [8]))
            if shape[0] * shape[1] * shape[2] <= max_points:
                break
    else:
        shape = (rng.choice(SIZES), rng.choice(SIZES + [8]))
    return tree, shape, kw, rng.choice([1, 2]), rng.choice([0.0, 2.0])


def run_case(case, lib_path, tol=1e-11):
    import parity_common as pc
    tree, shape, kw, nsteps, t0 = case
    mk = {"c3": pc.make_case, "i3": pc.make_case_incompressible, "c2": pc.make_case_2d, "i2": pc.make_case_incompressible_2d}[tree]
    p, prim = mk(*shape, **kw)
    o, g = pc.run_both(p, prim, nsteps, lib_path=lib_path, t0=t0)
    try:
        pc.check_state(o, g, tol)
    finally:
        g.close()


def random_multirank_case(rng):
    """World size 2 - 8, 3D trees, both ownership forms and the three stage schedules (tests/test_multirank_gloo.py runs the ranks)."""
    while True:
        world = rng.choice([2, 3, 4, 5, 8])
        sizes = [16, 32, 48]
        shape = (rng.choice(sizes), rng.choice([n for n in sizes + [80] if n >= world]), rng.choice([n for n in sizes if n >= world]))
        if shape[0] * shape[1] * shape[2] <= 32 * 48 * 48:
            break
    case = dict(hall=rng.random() < 0.7, aeb=rng.random() < 0.7, dealias=rng.choice([0, 1, 2]))
    if case["aeb"] and rng.random() < 0.4:
        case["corot"] = True
    env = {}
    r = rng.random()
    if r < 0.5:
        env["LAPS_TUNE_CYCLIC"] = "0" if r < 0.25 else "1"
    r = rng.random()
    if r < 0.6:
        env["LAPS_TUNE_OVERLAP"] = "1" if r < 0.3 else ("2" if r < 0.5 else "0")
    return world, dict(shape=shape, case=case, steps=rng.choice([1, 2]), env=env, incompressible=rng.random() < 0.25)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--seconds", type=float, default=600.0)
    ap.add_argument("--gpu", action="store_true", help="the real library on cuda:0 instead of the kernel emulator")
    ap.add_argument("--ranks", action="store_true", help="multi-rank cases (separate processes over gloo on the emulator)")
    a = ap.parse_args()
    lib = None
    if not a.gpu:
        import build_emu
        lib = build_emu.build()
    rng = random.Random(a.seed)
    t0, n, bad = time.time(), 0, 0
    while time.time() - t0 < a.seconds:
        n += 1
        try:
            if a.ranks:
                import test_multirank_gloo as mg
                world, cfg = random_multirank_case(rng)
                case = (world, cfg)
                mg.run_ranks(world, dict(cfg, lib=lib), timeout=900)
            else:
                case = random_case(rng)
                run_case(case, lib)
            print(n, *case, "ok", flush=True)
        except (Exception, AssertionError) as e:   # noqa: BLE001
            bad += 1
            print(n, *case, "FAIL", repr(e)[:300], flush=True)
            traceback.print_exc()
    print("done:", n, "cases,", bad, "failures")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
