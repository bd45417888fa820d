#!/usr/bin/env python
"""Benchmark of the LAPS RK-step hot path (BASELINE.json: 512^3 compressible Hall-MHD + expanding
box, grid-point-steps/s) on N B200s of one node, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--n 512] [--impl reference]

A "step" is one pass of the driver's Principal loop body (mhd.f90:245-248,285):
evolve (3 RK stages) + evolve_radius + vardt, through the C ABI (include/laps_b200.h).

parity : BEFORE anything is timed, every rank runs two steps of a 64^3 Hall-MHD + expanding-box case through the same
         decomposition (N ranks, both Fourier-row ownership forms when N > 1) and compares its slabs with the CPU oracle
         (relative L2 <= 1e-11); a failure ends the run with a non-zero exit code and no bench line.
value  : K steps with the state resident in HBM, CUDA events on the library's stream, max over ranks; no per-launch
         instrumentation inside this region.
e2e    : a driver session through the same C ABI with HOST buffers inside the timed region:
         laps_set_primitive from pinned host memory (H2D + conversion + 8 forward FFTs), K steps
         (each reads dt back to the host), laps_get_output into pinned host memory (D2H of the 8-field
         array output_uu writes) -- the traffic a LAPS driver generates between two outNNN.dat dumps; the
         h2d/d2h byte counts are those two copies spread over the K steps, not a per-step copy.
roofline: a separate instrumented pass (CUDA events around every launch, laps_set_profiling) after the timed one:
         dominant kernel by device time, algorithmic bytes per launch as the library states them
         (laps_get_profile_bytes) against MEASURED_PEAKS.json; step_frac = the same for the
         whole step (sum of algorithmic bytes / step time / peak).
cpu_baseline / --impl reference: the CPU port (oracle/laps_cpu.c, a C + OpenMP restatement with the
         reference's structure, kind "port": the Fortran+MPI+FFTW reference cannot be built in this
         image) on ALL host cores (whatever OMP_NUM_THREADS the launcher exported), on the workload's own grid when
         the host has the memory for it (512^3 needs about 90 GB), else on a 256^3 sample, in grid-point-steps/s.
"""
from __future__ import annotations

import argparse
import json
import os
import re
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
PARITY_TOL = 1e-11


def workload_params(n):
    """BASELINE config 4 / SURVEY 8(d): src_compressible/mhd.input with Hall + expanding box on."""
    return dict(nx=n, ny=n, nz=n, Lx=24.0, Ly=24.0, Lz=24.0, adiabatic_index=1.666667,
                if_resis=1, if_resis_exp=0, resistivity=1e-4, if_visc=1, if_visc_exp=0, viscosity=1e-4,
                if_conserve_background=0, cfl=0.5, dealias_option=1, afx=0.495, afy=0.495, afz=0.495,
                if_AEB=1, if_corotating=0, radius0=30.0, Ur0=1.167, corotating_angle=0.0,
                if_hall=1, ion_inertial_length=0.2)


def workload_name(n):
    return f"3D compressible Hall-MHD + expanding box {n}^3, FP64, RK3 step (src_compressible/mhd.input physics)"


def config_spec(config, n_override=None):
    """The five BASELINE.json configurations: (name, solver parameters, shape of one real field [nz, ny, nx], kwargs of
    synthetic.turbulence_slab, number of grid points)."""
    base = workload_params(64)
    plain = dict(base, if_AEB=0, if_hall=0, ion_inertial_length=0.0, Ur0=0.0)
    if config == 1:
        n = n_override or 64
        kw = dict(plain, nx=n, ny=n, nz=n)
        return dict(name=f"config 1: 3D compressible MHD {n}^3 (src_compressible/mhd.input physics)", kw=kw, grid=(n, n, n), turb={})
    if config == 2:
        n = n_override or 2048
        kw = dict(plain, nx=n, ny=n, nz=1, ndim=2, if_hall=1, ion_inertial_length=0.2)
        return dict(name=f"config 2: 2D compressible Hall-MHD {n}^2 (src_compressible/2D)", kw=kw, grid=(n, n, 1), turb={})
    if config == 3:
        n = n_override or 256
        kw = dict(plain, nx=n, ny=n, nz=n, incompressible=1, rho0=1.0)
        return dict(name=f"config 3: 3D incompressible MHD {n}^3 decaying turbulence (src_incompressible)", kw=kw, grid=(n, n, n),
                    turb=dict(drho0=0.0))
    if config in (4, 5):
        n = n_override or (512 if config == 4 else 1024)
        return dict(name=workload_name(n), kw=workload_params(n), grid=(n, n, n), turb={})
    raise SystemExit(f"unknown --config {config}")


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
# algorithmic bytes per launch of each kernel; R, C in bytes per field.
# The library reports the same figures per launch (laps_get_profile_bytes); this host-side model is what
# tests/test_bench_model.py holds them to.
# ------------------------------------------------------------------------------------------
def kernel_bytes(name, R, C, nf, ni, hall, fx=1.0, fcol=1.0, fmode=1.0, mass=False, fcol_all=None):
    """Algorithmic HBM bytes of one launch.  Pass launches carry their field
    count in the name (fwd_x13, inv_y11, inv_y3 ...).  fx = fraction of the kx columns that survive
    the dealiasing mask, fcol = fraction of this rank's (kx, ky) columns, fmode = fraction of its
    (kx, ky, kz) modes (laps_get_pruning, laps_get_pruning_counts): the passes skip the rest exactly.
    fcol_all = surviving fraction of the (kx, ky) columns of ALL ranks (what the y passes of this rank's z slab
    visit; equal to fcol on one rank).  mass: the continuity row takes its fluxes from the state
    (laps_get_field_counts) and runs inside the curl_b_inv_z launch instead of the spec_z one."""
    rows = 7 if mass else 8
    fy = fcol if fcol_all is None else fcol_all
    m = re.fullmatch(r"(fwd_x|fwd_y|inv_y|inv_x)(\d+)", name)
    if m:
        k, n = m.group(1), int(m.group(2))
        return {"fwd_x": n * R + n * C * fx, "fwd_y": n * C * fx + n * C * fy,
                "inv_y": n * C * fy + n * C * fx, "inv_x": n * C * fx + n * R}[k]
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
    per_kernel = {}
    for k, v in prof.items():
        tms, cnt = v[0], v[1]
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


def all_cores():
    """Every CPU this process may run on after lifting the launcher's / the NUMA pinning's restrictions."""
    n = os.cpu_count() or 1
    try:
        os.sched_setaffinity(0, range(n))
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return n


def pin_to_gpu_numa(index):
    """Run this rank's host thread (and allocate its pinned buffers) on the CPUs next to its GPU: with 8 ranks the
    H2D / D2H copies of the end-to-end session otherwise cross the socket interconnect for half of the GPUs."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return f"nvmlDeviceSetCpuAffinity(gpu {index}): {len(os.sched_getaffinity(0))} CPUs"
    except Exception as e:
        return f"not pinned ({type(e).__name__})"


# ------------------------------------------------------------------------------------------
# CPU oracle leg
# ------------------------------------------------------------------------------------------
def cpu_sample_size(n):
    """The workload's own grid when the host has the memory for the port's arrays (34 real + 46 spectral fields +
    k_square, the reference's own layout: 81.5 x 8 n^3 bytes), else 256^3."""
    if n & (n - 1):      # the port's 1-D transforms are radix-2/4 only: an odd-factor grid (--n 384) is sampled on the power of two below
        n = 1 << (n.bit_length() - 1)
    need = 81.5 * 8 * float(n) ** 3 * 1.25 + 8e9
    try:
        import psutil
        if psutil.virtual_memory().available > need:
            return n
    except Exception:
        pass
    return min(n, 256)


def cpu_oracle_run(n, steps, warmup):
    """The CPU port's Principal-loop step on an n^3 grid with the workload's physics; returns
    (grid-point-steps/s, seconds per step, threads, description).  The port is oracle/laps_cpu.c (C + OpenMP, all
    host threads; the reference's structure: one field at a time, 19 + 11 three-dimensional transforms per stage,
    separate pointwise sweeps); if it cannot be built on this host, the NumPy/SciPy oracle."""
    from oracle import laps_oracle as lo
    from laps_b200 import synthetic
    kw = workload_params(n)
    p = lo.Params(**{k: (bool(v) if k.startswith("if_") else v) for k, v in kw.items()})
    cores = all_cores()
    prim = synthetic.turbulence_slab(n, n, n, p.Lx, p.Ly, p.Lz, kmax=min(8, n // 2 - 1), workers=cores)
    try:
        from oracle import cpu_port
        s = cpu_port.CpuPort(p)
        s.set_threads(cores)       # torchrun exports OMP_NUM_THREADS=1: use every core regardless
        what, cores = "oracle/laps_cpu.c (C + OpenMP restatement, batched SIMD transforms)", s.threads
    except Exception as e:  # no compiler on this host
        s = lo.State(p)
        what, cores = f"oracle/laps_oracle.py (NumPy/SciPy restatement; C port unavailable: {type(e).__name__})", lo._WORKERS
    s.set_primitive(prim)
    del prim
    s.vardt()
    for _ in range(warmup):
        s.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        s.step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    if hasattr(s, "close"):
        s.close()
    return n ** 3 / dt, dt, cores, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    spec = config_spec(args.config, args.n)
    n = args.cpu_n or cpu_sample_size(spec["grid"][0])
    steps = max(1, min(args.steps, 2 if n > 256 else 3))
    warm = 1
    v, sec, cores, what = cpu_oracle_run(n, steps, warm)
    same = n == spec["grid"][0] and args.config in (4, 5)
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(spec["grid"][0]) if args.config in (4, 5) else spec["name"],
                   "sample": ("the workload's own grid" if same else f"{n}^3 grid of the 3D compressible Hall + expanding-box physics")},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{what}, {n}^3 grid, {steps} RK steps after {warm} warm-up, {cores} threads"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------------------
# multi-rank plumbing shared by the parity gate and the timed run
# ------------------------------------------------------------------------------------------
class Ranks:
    def __init__(self, args):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x, op="max"):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN}[op])
        return float(t.item())

    def connect(self, g):
        if not self.dist:
            return
        torch = self.torch
        blob = torch.from_numpy(np.frombuffer(g.export_peer_blob(), dtype=np.uint8).copy()).cuda()
        blobs = [torch.empty_like(blob) for _ in range(self.world)]
        self.dist.all_gather(blobs, blob)
        g.import_peer_blobs(b"".join(bytes(b.cpu().numpy().tobytes()) for b in blobs))

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def parity_gate(R):
    """Two Principal-loop steps of a 64^3 Hall-MHD + expanding-box case on this run's ranks against the CPU oracle
    (oracle/laps_oracle.py State, the restatement pinned by the executed reference source): every rank's real-space and
    Fourier-space slabs, relative L2 per field.  With N > 1 both Fourier-row ownership forms are run: the reference's
    decompose_1d slabs and the round-robin rows (the default from 4 ranks on)."""
    from laps_b200 import Solver
    from oracle import laps_oracle as lo
    n = int(os.environ.get("LAPS_BENCH_PARITY_N", "64"))   # tests/bench_on_emulator.py runs the gate at 16^3 (the kernel emulator is slow)
    kw = workload_params(n)
    p = lo.Params(**{k: (bool(v) if k.startswith("if_") else v) for k, v in kw.items()})
    prim = lo.ic_uniform_background(p, bx0=1.0, press0=1.0)
    prim = lo.ic_turbulence(p, prim, 1.0, 0.0, 0.0, db0=0.1, dv0=0.1, drho0=0.01, nmodex=2, nmodey=2, nmodez=2, seeds=(101, 116, 132))
    o = lo.State(p)
    o.set_primitive(prim)
    o.vardt()
    for _ in range(2):
        o.step()
    forms = [None] if R.world == 1 else ["0", "1"]
    cases, ok = [], True
    for form in forms:
        saved = os.environ.get("LAPS_TUNE_CYCLIC")
        if form is not None:
            os.environ["LAPS_TUNE_CYCLIC"] = form
        try:
            g = Solver(rank=R.rank, nranks=R.world, device=R.local, **kw)
        finally:
            if form is not None:
                if saved is None:
                    os.environ.pop("LAPS_TUNE_CYCLIC", None)
                else:
                    os.environ["LAPS_TUNE_CYCLIC"] = saved
        R.connect(g)
        R.barrier()
        zo, zn, rows = g.ext.z_offset, g.ext.z_size, g.ky_rows
        g.set_primitive(prim[:, zo:zo + zn])
        g.vardt()
        for _ in range(2):
            g.step()
        uu, _ = g.get_state()
        uf = g.uu_fourier()
        worst = 0.0
        for v in range(8):
            a, b = uu[v], o.uu[v, zo:zo + zn]
            worst = max(worst, float(np.linalg.norm(a - b) / np.linalg.norm(b)))
            # this rank's Fourier rows against the norm of the whole field (a rank may own only near-empty rows)
            a, b = uf[v], o.uu_fourier[v][:, rows, :]
            worst = max(worst, float(np.linalg.norm(a - b) / np.linalg.norm(o.uu_fourier[v]) * np.sqrt(R.world)))
        dt_err = abs(g.dt - o.dt) / o.dt
        worst_all = R.reduce(worst, "max")
        dt_all = R.reduce(dt_err, "max")
        cases.append({"y_stride": int(g.ext.y_stride), "max_rel_l2": worst_all, "dt_rel": dt_all})
        ok = ok and worst_all <= PARITY_TOL and dt_all <= 1e-12
        R.barrier()
        g.close()
        R.barrier()
    return {"ranks": R.world, "case": f"{n}^3 Hall-MHD + expanding box, 2 steps, vs oracle/laps_oracle.py State", "tol": PARITY_TOL,
            "fields": "uu(1:8) of every rank's z slab and uu_fourier(1:8) of its ky rows, relative L2", "cases": cases, "ok": bool(ok)}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_gpu(args):
    R = Ranks(args)
    torch = R.torch
    from laps_b200 import Solver, synthetic
    rank, world, local = R.rank, R.world, R.local
    affinity = pin_to_gpu_numa(local)

    parity = None
    if not args.no_parity:
        parity = parity_gate(R)
        if not parity["ok"]:
            if rank == 0:
                sys.stderr.write("PARITY GATE FAILED: " + json.dumps(parity) + "\n")
            R.close()
            return 3

    spec = config_spec(args.config, args.n)
    kw = spec["kw"]
    gx, gy, gz = spec["grid"]
    npoints = float(gx) * gy * gz
    g = Solver(rank=rank, nranks=world, device=local, **kw)
    R.connect(g)
    stream = torch.cuda.ExternalStream(g.cuda_stream(), device=torch.device("cuda", local))
    footprint = R.reduce(float(g.footprint()), "max")

    # synthetic turbulence on this rank's z-slab, in pinned host memory (the driver's uu array)
    shape = (8,) + g.real_shape
    host_in = torch.empty(shape, dtype=torch.float64).pin_memory()
    prim = host_in.numpy()
    if kw.get("ndim") == 2:     # the z = 0 plane of a 3D turbulence field
        prim[...] = synthetic.turbulence_slab(gx, gy, 32, kw["Lx"], kw["Ly"], kw["Lz"], kmax=8, z_size=1, **spec["turb"])
    else:
        synthetic.turbulence_slab(gx, gy, gz, kw["Lx"], kw["Ly"], kw["Lz"], z_offset=g.ext.z_offset, z_size=g.ext.z_size,
                                  kmax=min(8, gx // 2 - 1), out=prim, **spec["turb"])
    host_uu = torch.empty(shape, dtype=torch.float64).pin_memory()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # ---------------- warm-up ----------------
    g.set_primitive(prim)
    rho_mean0 = float(g.calc_rms()[0][0])     # the k = 0 mode of rho (collective; outside every timed region)
    g.vardt()
    history = []                               # (time, dt) at the start of every step since set_primitive
    for _ in range(args.warmup):
        history.append((g.time, g.dt))
        g.step()
    g.sync()

    # ---------------- timed: resident state, no instrumentation ----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # Working sets that fit the 126 MB L2 (configs 1 and 2: 64^3, 2048^2) are timed step by step with an L2 flush (a
    # 256 MB write on the same stream) between the steps, outside the timed intervals; everything else streams >= 8 GB
    # per pass and is timed as one region.
    small = footprint < 4e9
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small else None
    R.barrier()
    e0, e1 = ev(), ev()
    tw0 = time.perf_counter()
    e0.record(stream)
    launches = 0
    per_step = []
    for _ in range(args.steps):
        history.append((g.time, g.dt))
        if small:
            with torch.cuda.stream(stream):
                flush.zero_()
            a, b = ev(), ev()
            a.record(stream)
            g.step()
            b.record(stream)
            per_step.append((a, b))
        else:
            g.step()
        launches += g.last_step_ms()[1]
    e1.record(stream)
    R.barrier()
    tw1 = time.perf_counter()
    ms_total = R.reduce(sum(a.elapsed_time(b) for a, b in per_step) if small else e0.elapsed_time(e1), "max")
    clocks = sampler.stop(tw0, tw1) if rank == 0 else None
    ms_step = ms_total / args.steps
    value = npoints / (ms_step * 1e-3)
    # Size-independent property at the full grid size: the k = 0 mode of rho.  The continuity row has no flux at k = 0,
    # so d<rho>/dt = -(2 / tau) <rho> with tau = R(t) / U_r frozen during a step (mhdrhs.f90:235-237): every RK3 step
    # multiplies it by 1 + z + z^2/2 + z^3/6, z = -2 dt / tau (rktmod.f90:17-61 applied to a linear term) — and by
    # exactly 1 without the expanding box.
    rho_mean1 = float(g.calc_rms()[0][0])
    expect = 1.0
    if kw.get("if_AEB") and not kw.get("incompressible"):
        for t, dt in history:
            z = -2.0 * dt / ((kw["radius0"] + kw["Ur0"] * t) / kw["Ur0"])
            expect *= 1.0 + z + z * z / 2.0 + z * z * z / 6.0
    k0_mode = {"what": "mean rho (the k = 0 mode) after warm-up + K steps over its initial value, against the RK3 polynomial of the "
                       "expanding-box term (exactly 1 without the box)", "measured": rho_mean1 / rho_mean0, "expected": expect,
               "rel_err": abs(rho_mean1 / rho_mean0 / expect - 1.0), "ok": abs(rho_mean1 / rho_mean0 / expect - 1.0) < 1e-11}

    # ---------------- instrumented pass: per-launch CUDA events (outside the timed region) ----------------
    psteps = max(1, min(args.steps, 4))
    if world > 1:
        g.set_tune("overlap", 0)     # per-kernel times on one stream (the two-stream schedule runs passes side by side)
    g.set_profiling(True)
    R.barrier()
    p0, p1 = ev(), ev()
    p0.record(stream)
    prof = {}     # name -> [total ms, launches, total algorithmic bytes]
    for _ in range(psteps):
        g.step()
        for name, ms, by in g.get_profile(with_bytes=True):
            a = prof.setdefault(name, [0.0, 0, 0.0])
            a[0] += ms; a[1] += 1; a[2] += by
    p1.record(stream)
    R.barrier()
    ms_prof_step = R.reduce(p0.elapsed_time(p1), "max") / psteps
    g.set_profiling(False)
    if world > 1:
        g.set_tune("overlap", -1)

    # ---------------- timed: end to end through the C ABI with host buffers ----------------
    # One driver session between two dumps: the state comes from the host (laps_set_primitive: H2D + conversion + 8 forward
    # transforms), K steps run (each reads dt back), and ONE dump of the 8-field output array goes back to the host.  The
    # dump is requested after step ceil(K/2) through laps_get_output_async — packed on the device, copied on a copy stream
    # while the remaining steps run — and waited for at the end of the session.  (blocking: the same session with the dump
    # taken by the blocking laps_get_output after the last step, as round 1 measured it.)
    def session(blocking):
        R.barrier()
        f0, f1 = ev(), ev()
        f0.record(stream)
        g.time = 0.0
        g.evolve_radius(0.0)
        g.set_primitive(prim)                      # H2D from pinned memory
        g.dt = 0.0
        g.vardt()
        for i in range(args.steps):
            g.step()                               # reads dt back every step
            if not blocking and i + 1 == (args.steps + 1) // 2:
                g.get_output_async(host_uu.numpy(), True)
        if blocking:
            g.get_output(True, out=host_uu.numpy())   # D2H of the array output_uu writes (rho, u, B, p) into pinned memory
        else:
            g.output_wait()
        f1.record(stream)
        g.sync()
        R.barrier()
        return R.reduce(f0.elapsed_time(f1), "max") / args.steps
    e2e_blocking_ms = session(True)
    e2e_ms = session(False)
    h2d = world * prim.nbytes / args.steps
    d2h = world * host_uu.numel() * 8 / args.steps + 8
    finite = bool(np.isfinite(host_uu.numpy()).all())

    # ---------------- roofline ----------------
    nf, ni, rows = g.field_counts()
    nkx, kymax, nkyl = g.pruning()
    live_cols, live_modes = g.pruning_counts()
    nline = g.nz if kw.get("ndim") != 2 else g.ny
    fcol = live_cols / float(g.nxh * max(g.nyl, 1)) if kw.get("ndim") != 2 else live_cols / float(g.nxh)
    fmode = live_modes / float(g.nxh * max(g.nyl, 1) * nline) if kw.get("ndim") != 2 else live_modes / float(g.nxh * nline)
    mass = rows < 8
    peak, peak_src = peaks()
    timed = {k: v for k, v in prof.items() if v[2] > 0}
    top = max(timed.items(), key=lambda kv: kv[1][0], default=None)
    roofline = None
    # DRAM bytes per launch measured by ncu --set full for this exact configuration (committed capture), or None
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        c = tj["config"]
        if (c["config"], c["n"], c["n_gpus"]) == (args.config, gx, world):
            traffic, traffic_src = {k: v["dram_bytes_per_launch"] for k, v in tj["kernels"].items()}, tj["source"]
    except Exception:
        pass
    if top:
        tot = sum(v[0] for v in prof.values()) or 1e-12
        shares = {k: round(v[0] / tot, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
        name, (tms, cnt, tby) = top
        b = tby / cnt
        avg_ms = max(tms / cnt, 1e-9)
        ach = b / (avg_ms * 1e-3) / 1e9
        step_bytes = sum(v[2] for v in prof.values()) / psteps
        step_ach = step_bytes / (ms_step * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": (traffic or {}).get(name), "traffic_source": traffic_src if (traffic or {}).get(name) else None,
                    "peak_source": peak_src,
                    "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": b,
                    "step_frac": step_ach / peak, "step_achieved": step_ach, "step_algorithmic_bytes": step_bytes,
                    "step_what": "sum of the algorithmic bytes of every launch of one step (laps_get_profile_bytes; exactly skipped "
                                 "columns and modes left out) / ms_per_step of the un-instrumented timed region / peak",
                    "per_kernel_GBps": {k: (v[2] / v[1]) / (max(v[0] / v[1], 1e-9) * 1e-3) / 1e9 for k, v in timed.items()},
                    "per_kernel_frac": {k: round((v[2] / v[1]) / (max(v[0] / v[1], 1e-9) * 1e-3) / 1e9 / peak, 4) for k, v in timed.items()},
                    "pruning": {"nkx": nkx, "nxh": g.nxh, "kymax": kymax, "ny": g.ny, "nky_local": nkyl, "live_column_fraction": fcol, "live_mode_fraction": fmode,
                                "what": "columns removed by the dealiasing mask are skipped exactly (bit-identical state)"},
                    "profiled_steps": psteps, "ms_per_step_instrumented": ms_prof_step,
                    "time_share": shares}
        if roofline["traffic"] is None and world > 1:     # no ncu capture at N > 1: the key is left out rather than printed as null
            roofline.pop("traffic"); roofline.pop("traffic_source")

    # ---------------- NVLink: the slab transposes are the remote stores of the y pass and the z pass ----------------
    nvlink = None
    if world > 1 and not kw.get("incompressible"):
        total_cols = R.reduce(float(live_cols), "sum")
        nvlink = nvlink_model(prof, live_cols, total_cols, rows, mass, gz, g.nzl, psteps, ms_prof_step)
        nvlink["egress_GBps_over_the_step"] = nvlink["egress_bytes_per_step"] / (ms_step * 1e-3) / 1e9   # un-instrumented step time

    R.barrier()          # no rank may free its exchange buffers while a peer can still store into them
    g.close()
    if rank != 0:
        R.close()
        return 0

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cn = args.cpu_n or cpu_sample_size(gx if args.config in (4, 5) else 256)
        v, sec, cores, what = cpu_oracle_run(cn, 2, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{what}, {cn}^3 grid of the 3D compressible Hall + expanding-box physics, 2 RK steps after 1 warm-up, {cores} threads",
               "ms_per_step": sec * 1e3}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["name"], "decomposition": f"slab over {world} GPU(s), z in real space / ky in Fourier space"
                                    + (f" (ky rows dealt round-robin, y_stride {g.ext.y_stride})" if g.ext.y_stride > 1 else " (decompose_1d slabs)"),
                   "l2": ("working set fits L2: every step timed on its own between CUDA events, 256 MB written to flush L2 before each" if small else
                          "inputs larger than L2 (every pass streams >= 8 GB per GPU at 512^3/8 and above); no flush"),
                   "ic": "ifield=3 uniform B0=(1,0,0) + ipert=7-style random-phase modes |k|<=8, k^-3/2",
                   "device_bytes_per_gpu": footprint, "host_affinity": affinity},
        "parity": parity,
        "clocks": clocks,
        "e2e": {"value": npoints / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "ms_per_step_blocking_output": e2e_blocking_ms,
                "what": "one driver session between two dumps, per step: laps_set_primitive(host) + K x laps_step + one dump of the output "
                        "array requested after step ceil(K/2) with laps_get_output_async (device snapshot + copy stream, overlapping the "
                        "remaining steps) and waited for at the end; the two copies happen once per session, their bytes are spread over "
                        "the K steps.  ms_per_step_blocking_output: the same with the blocking laps_get_output after the last step"},
        "gpu_launches": launches,
        "nvlink": nvlink,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "state_finite": finite,
    }
    out["k0_mode"] = k0_mode
    print(json.dumps(out))
    R.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=4, help="BASELINE.json configuration 1..5 (4 = the headline 512^3; 5 = 1024^3 over 8 GPUs)")
    ap.add_argument("--n", type=int, default=None, help="override the configuration's grid size")
    ap.add_argument("--cpu-n", type=int, default=None, help="grid of the CPU-baseline sample (default: the workload's own when the host has the memory, else 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the pre-timing parity gate (profiler runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
