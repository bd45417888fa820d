#!/usr/bin/env python
"""Benchmark of the LAPS RK-step hot path (BASELINE.json: 512^3 compressible Hall-MHD + expanding
box, grid-point-steps/s) on N B200s of one node, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 512] [--impl reference]

A "step" is one pass of the driver's Principal loop body (mhd.f90:245-248,285):
evolve (3 RK stages) + evolve_radius + vardt, through the C ABI (include/laps_b200.h).

value  : K steps with the state resident in HBM, CUDA events on the library's stream, max over ranks.
e2e    : a driver session through the same C ABI with HOST buffers inside the timed region:
         laps_set_primitive from pinned host memory (H2D + conversion + 8 forward FFTs), K steps
         (each reads dt back to the host), laps_get_output into pinned host memory (D2H of the 8-field
         array output_uu writes) -- the traffic a LAPS driver generates between two outNNN.dat dumps.
roofline: dominant kernel by device time (CUDA events around every launch, laps_set_profiling),
         algorithmic bytes per launch as stated in DESIGN.md, against MEASURED_PEAKS.json.
cpu_baseline / --impl reference: the CPU port (oracle/laps_cpu.c, a C + OpenMP restatement with the
         reference's structure, kind "port": the Fortran+MPI+FFTW reference cannot be built in this
         image) on all host cores, on a 256^3 grid of the same physics (bounded sample), in
         grid-point-steps/s.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "grid_point_steps_per_s"
UNIT = "grid-point-steps/s"


def workload_params(n):
    """BASELINE config 4 / SURVEY 8(d): src_compressible/mhd.input with Hall + expanding box on."""
    return dict(nx=n, ny=n, nz=n, Lx=24.0, Ly=24.0, Lz=24.0, adiabatic_index=1.666667,
                if_resis=1, if_resis_exp=0, resistivity=1e-4, if_visc=1, if_visc_exp=0, viscosity=1e-4,
                if_conserve_background=0, cfl=0.5, dealias_option=1, afx=0.495, afy=0.495, afz=0.495,
                if_AEB=1, if_corotating=0, radius0=30.0, Ur0=1.167, corotating_angle=0.0,
                if_hall=1, ion_inertial_length=0.2)


def workload_name(n):
    return f"3D compressible Hall-MHD + expanding box {n}^3, FP64, RK3 step (src_compressible/mhd.input physics)"


# ------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for t, line in self.rows:
            if t0 is not None and not (t0 <= t <= t1):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# algorithmic bytes per launch of each kernel; R, C in bytes per field
# ------------------------------------------------------------------------------------------
def kernel_bytes(name, R, C, nf, ni, hall, fx=1.0, fcol=1.0, fmode=1.0, mass=False):
    """Algorithmic HBM bytes of one launch.  Pass launches carry their field
    count in the name (fwd_x13, inv_y11, inv_y3 ...).  fx = fraction of the kx columns that survive
    the dealiasing mask, fcol = fraction of this rank's (kx, ky) columns, fmode = fraction of its
    (kx, ky, kz) modes (laps_get_pruning, laps_get_pruning_counts): the passes skip the rest exactly.
    mass: the continuity row takes its fluxes from the state (laps_get_field_counts) and runs inside
    the curl_b_inv_z launch instead of the spec_z one."""
    import re
    rows = 7 if mass else 8
    m = re.fullmatch(r"(fwd_x|fwd_y|inv_y|inv_x)(\d+)", name)
    if m:
        k, n = m.group(1), int(m.group(2))
        return {"fwd_x": n * R + n * C * fx, "fwd_y": n * C * fx + n * C * fcol,
                "inv_y": n * C * fcol + n * C * fx, "inv_x": n * C * fx + n * R}[k]
    flux = (8 + (3 if hall else 0)) * R + nf * R
    table = {
        "flux": flux,
        "flux+cfl": flux,
        # calc_flux fused into the forward x pass: reads uu (8R) + J (3R), writes nf half spectra
        "flux_fwd_x": (8 + (3 if hall else 0)) * R + nf * C * fx,
        # reads nf flux lines and writes `rows` inverse-z lines of every surviving column; reads u + fnl_rk (stages
        # 2,3) and writes u + fnl_rk (stages 1,2) of the surviving modes only: averaged over the three stages
        "spec_z": (nf + rows) * C * fcol + (rows + rows * 2 / 3 + rows + rows * 2 / 3) * C * fmode,
        # J^ = ik x B^ (3 state reads, 3 lines out) + the continuity row (reads rho u, rho, fnl_rk; writes rho, fnl_rk, 1 line out)
        "curl_b_inv_z": 3 * C * fmode + 3 * C * fcol + (((4 + 2 / 3 + 1 + 2 / 3) * C * fmode + C * fcol) if mass else 0.0),
        "fwd_z": 2 * 8 * C,
        "cfl": 8 * R,
    }
    return table.get(name)


def nvlink_model(prof, live_cols, total_cols, rows, mass, n, nzl, steps, ms_step):
    """Bytes one rank stores into its peers' buffers (the reference's transpose_yz / transpose_zy, parallel.f90:273-324,
    fused into the passes) per launch of each exchanging kernel, over that launch's mean duration.  prof: name ->
    [total ms, launches]; live_cols / total_cols: surviving (kx, ky) columns owned by this rank / by all ranks."""
    import re
    per_kernel = {}
    for k, (tms, cnt) in prof.items():
        m = re.fullmatch(r"fwd_y(\d+)", k)
        if m:      # every surviving (kx, ky) column of this rank's z slab goes to the owner of ky (transpose_yz)
            remote = int(m.group(1)) * 16.0 * nzl * (total_cols - live_cols)
        elif k == "spec_z":       # the inverse-z lines of this rank's columns go to the owners of z (transpose_zy)
            remote = rows * 16.0 * live_cols * (n - nzl)
        elif k == "curl_b_inv_z":  # J (3 lines) + the continuity row when it is taken from the state
            remote = (3 + (1 if mass else 0)) * 16.0 * live_cols * (n - nzl)
        else:
            continue
        per_kernel[k] = {"remote_bytes_per_launch": remote, "avg_launch_ms": tms / cnt,
                         "egress_GBps": remote / (max(tms / cnt, 1e-9) * 1e-3) / 1e9}
    sent = sum(v["remote_bytes_per_launch"] * prof[k][1] for k, v in per_kernel.items()) / steps
    return {"what": "bytes this rank stores into its peers' buffers over NVLink inside the y pass (transpose_yz) and the z passes "
                    "(transpose_zy), per launch, over the launch's own duration (the kernel does its HBM work in the same time)",
            "peak": 900.0, "unit": "GB/s per direction (nominal NVLink 5)", "per_kernel": per_kernel,
            "egress_bytes_per_step": sent, "egress_GBps_over_the_step": sent / (ms_step * 1e-3) / 1e9}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
# CPU oracle leg
# ------------------------------------------------------------------------------------------
def cpu_oracle_run(n, steps, warmup):
    """The CPU port's Principal-loop step on an n^3 grid with the workload's physics; returns
    (grid-point-steps/s, seconds per step, threads, description).  The port is oracle/laps_cpu.c (C + OpenMP, all
    host threads; the reference's structure: one field at a time, line-at-a-time transforms, separate pointwise
    sweeps); if it cannot be built on this host, the NumPy/SciPy oracle."""
    from oracle import laps_oracle as lo
    from laps_b200 import synthetic
    kw = workload_params(n)
    p = lo.Params(**{k: (bool(v) if k.startswith("if_") else v) for k, v in kw.items()})
    prim = synthetic.turbulence_slab(n, n, n, p.Lx, p.Ly, p.Lz, kmax=min(8, n // 2 - 1))
    try:
        from oracle import cpu_port
        s = cpu_port.CpuPort(p)
        what, cores = "oracle/laps_cpu.c (C + OpenMP restatement)", s.threads
    except Exception as e:  # no compiler on this host
        s = lo.State(p)
        what, cores = f"oracle/laps_oracle.py (NumPy/SciPy restatement; C port unavailable: {type(e).__name__})", lo._WORKERS
    s.set_primitive(prim)
    s.vardt()
    for _ in range(warmup):
        s.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        s.step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n ** 3 / dt, dt, cores, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.cpu_n
    steps = max(1, min(args.steps, 3))
    warm = 1
    v, sec, cores, what = cpu_oracle_run(n, steps, warm)
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.n), "sample": f"{n}^3 grid of the same physics"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{what}, {n}^3 grid, {steps} RK steps after {warm} warm-up, {cores} threads"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from laps_b200 import Solver, synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.n
    kw = workload_params(n)
    g = Solver(rank=rank, nranks=world, device=local, **kw)
    if world > 1:
        blob = torch.from_numpy(np.frombuffer(g.export_peer_blob(), dtype=np.uint8).copy()).cuda()
        blobs = [torch.empty_like(blob) for _ in range(world)]
        dist.all_gather(blobs, blob)
        g.import_peer_blobs(b"".join(bytes(b.cpu().numpy().tobytes()) for b in blobs))
    stream = torch.cuda.ExternalStream(g.cuda_stream(), device=torch.device("cuda", local))

    # synthetic turbulence on this rank's z-slab, in pinned host memory (the driver's uu array)
    shape = (8,) + g.real_shape
    host_in = torch.empty(shape, dtype=torch.float64).pin_memory()
    prim = host_in.numpy()
    synthetic.turbulence_slab(n, n, n, kw["Lx"], kw["Ly"], kw["Lz"], z_offset=g.ext.z_offset, z_size=g.ext.z_size,
                              kmax=min(8, n // 2 - 1), out=prim)
    host_uu = torch.empty(shape, dtype=torch.float64).pin_memory()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # ---------------- warm-up ----------------
    g.set_primitive(prim)
    g.vardt()
    for _ in range(args.warmup):
        g.step()
    g.sync()

    # ---------------- timed: resident state ----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    g.set_profiling(True)
    barrier()
    e0, e1 = ev(), ev()
    tw0 = time.perf_counter()
    e0.record(stream)
    launches = 0
    prof = {}
    for _ in range(args.steps):
        g.step()
        _, nl = g.last_step_ms()
        launches += nl
        for name, ms in g.get_profile():
            a = prof.setdefault(name, [0.0, 0])
            a[0] += ms; a[1] += 1
    e1.record(stream)
    barrier()
    tw1 = time.perf_counter()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    g.set_profiling(False)
    clocks = sampler.stop(tw0, tw1) if rank == 0 else None
    ms_step = ms_total / args.steps
    value = float(n) ** 3 / (ms_step * 1e-3)

    # ---------------- timed: end to end through the C ABI with host buffers ----------------
    barrier()
    f0, f1 = ev(), ev()
    f0.record(stream)
    g.time = 0.0
    g.evolve_radius(0.0)
    g.set_primitive(prim)                      # H2D from pinned memory
    g.dt = 0.0
    g.vardt()
    for _ in range(args.steps):
        g.step()                               # reads dt back every step
    g.get_output(True, out=host_uu.numpy())   # D2H of the array output_uu writes (rho, u, B, p) into pinned memory
    f1.record(stream)
    barrier()
    e2e_ms = max_over_ranks(f0.elapsed_time(f1)) / args.steps
    h2d = world * prim.nbytes / args.steps
    d2h = world * host_uu.numel() * 8 / args.steps + 8
    finite = bool(np.isfinite(host_uu.numpy()).all())

    # ---------------- roofline of the dominant kernel ----------------
    R = 8.0 * n * n * g.nzl
    C = 16.0 * g.nxh * g.nyl * n
    nf, ni, rows = g.field_counts()
    hall = bool(kw["if_hall"])
    nkx, kymax, nkyl = g.pruning()
    live_cols, live_modes = g.pruning_counts()
    fx = nkx / g.nxh
    fcol = live_cols / float(g.nxh * g.nyl)
    fmode = live_modes / float(g.nxh * g.nyl * n)
    mass = rows < 8
    kb = lambda k: kernel_bytes(k, R, C, nf, ni, hall, fx, fcol, fmode, mass)  # noqa: E731
    peak, peak_src = peaks()
    # the dominant kernel among those that move data (the one-CTA flag kernels have no byte model)
    top = max((kv for kv in prof.items() if kb(kv[0])), key=lambda kv: kv[1][0], default=None)
    roofline = None
    shares = {}
    # DRAM bytes per launch measured by ncu --set full for this exact configuration (committed capture), or None
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r01i_traffic.json")) as f:
            tj = json.load(f)
        c = tj["config"]
        if (c["n"], c["n_gpus"], c["if_hall"], c["if_AEB"], c["dealias_option"]) == (n, world, kw["if_hall"], kw["if_AEB"], kw["dealias_option"]):
            traffic, traffic_src = {k: v["dram_bytes_per_launch"] for k, v in tj["kernels"].items()}, tj["source"]
    except Exception:
        pass
    if top:
        tot = sum(v[0] for v in prof.values()) or 1e-12
        shares = {k: round(v[0] / tot, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
        name, (tms, cnt) = top
        b = kb(name)
        avg_ms = max(tms / cnt, 1e-9)
        ach = b / (avg_ms * 1e-3) / 1e9 if b else None
        roofline = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": (ach / peak) if ach else None, "traffic": (traffic or {}).get(name), "traffic_source": traffic_src if (traffic or {}).get(name) else None,
                    "peak_source": peak_src,
                    "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": b,
                    "per_kernel_GBps": {k: (kb(k) or 0) / (max(v[0] / v[1], 1e-9) * 1e-3) / 1e9 for k, v in prof.items() if kb(k)},
                    "pruning": {"nkx": nkx, "nxh": g.nxh, "kymax": kymax, "ny": n, "nky_local": nkyl, "live_column_fraction": fcol, "live_mode_fraction": fmode,
                                "what": "columns removed by the dealiasing mask are skipped exactly (bit-identical state)"},
                    "time_share": shares}

    # ---------------- NVLink: the slab transposes are the remote stores of the y pass and the z pass ----------------
    nvlink = None
    if world > 1:
        t = torch.tensor([float(live_cols)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        nvlink = nvlink_model(prof, live_cols, float(t.item()), rows, mass, n, g.nzl, args.steps, ms_step)

    barrier()          # no rank may free its exchange buffers while a peer can still store into them
    g.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, sec, cores, what = cpu_oracle_run(args.cpu_n, 2, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{what}, {args.cpu_n}^3 grid of the same physics, 2 RK steps after 1 warm-up, {cores} threads",
               "ms_per_step": sec * 1e3}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n), "decomposition": f"slab over {world} GPU(s), z in real space / ky in Fourier space"
                                    + (f" (ky rows dealt round-robin, y_stride {g.ext.y_stride})" if g.ext.y_stride > 1 else " (decompose_1d slabs)"),
                   "l2": "inputs larger than L2 (every pass streams >= 8 GB per GPU at 512^3/8 and above); no flush",
                   "ic": "ifield=3 uniform B0=(1,0,0) + ipert=7-style random-phase modes |k|<=8, k^-3/2"},
        "clocks": clocks,
        "e2e": {"value": float(n) ** 3 / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "what": "laps_set_primitive(host) + K x laps_step + laps_get_output(host), per step"},
        "gpu_launches": launches,
        "nvlink": nvlink,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "state_finite": finite,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=512, help="grid size (512 = BASELINE config 4)")
    ap.add_argument("--cpu-n", type=int, default=256, help="grid of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
