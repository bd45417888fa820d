"""The byte models bench.py reports: algorithmic HBM bytes per launch and the NVLink egress of
the fused transposes.  Pure host arithmetic — importing bench.py needs neither a GPU nor the CUDA library."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_stage_bytes_add_up_to_the_survey_figure():
    """with nothing pruned and the reference's 19 + 11 fields the passes of one stage move
    B_stage = 2 ni R + (4 nf + 4 ni + 40) C."""
    b = _bench()
    R, C, nf, ni = 3.0, 5.0, 19, 11
    kb = lambda k: b.kernel_bytes(k, R, C, nf, ni, True)  # noqa: E731
    total = (kb("flux") + kb(f"fwd_x{nf}") + kb(f"fwd_y{nf}") + kb("spec_z") + kb("curl_b_inv_z")
             + kb(f"inv_y{ni}") + kb(f"inv_x{ni}"))
    # against that fused model: the separate calc_flux sweep writes the nf fluxes and the x pass reads them back
    # (+2 nf R); the RK history is not read in stage 1 and not written in stage 3 (-16/3 C on the stage average); the
    # inverse z transform starts from registers (-8 C); the current re-reads B^ (+3 C)
    fused = 2 * ni * R + (4 * nf + 4 * ni + 40) * C
    extra = 2 * nf * R + (-16.0 / 3 - 8 + 3) * C
    assert abs(total - (fused + extra)) < 1e-9 * total, (total, fused + extra)


def test_pruned_pass_bytes_scale_with_the_surviving_fractions():
    b = _bench()
    full = b.kernel_bytes("fwd_y13", 1.0, 1.0, 13, 11, True)
    half = b.kernel_bytes("fwd_y13", 1.0, 1.0, 13, 11, True, fx=0.5, fcol=0.25)
    assert full == 26.0 and half == 13 * 0.5 + 13 * 0.25
    assert b.kernel_bytes("no_such_kernel", 1.0, 1.0, 13, 11, True) is None


def test_nvlink_model_counts_the_remote_share():
    b = _bench()
    prof = {"fwd_y13": [20.0, 10], "spec_z": [30.0, 30], "curl_b_inv_z": [6.0, 30], "flux": [9.0, 30]}
    n, nzl, mine, total = 512, 128, 1000, 4000.0
    out = b.nvlink_model(prof, mine, total, rows=7, mass=True, n=n, nzl=nzl, steps=10, ms_step=25.0)
    k = out["per_kernel"]
    assert set(k) == {"fwd_y13", "spec_z", "curl_b_inv_z"}
    assert k["fwd_y13"]["remote_bytes_per_launch"] == 13 * 16 * nzl * 3000
    assert k["spec_z"]["remote_bytes_per_launch"] == 7 * 16 * mine * (n - nzl)
    assert k["curl_b_inv_z"]["remote_bytes_per_launch"] == 4 * 16 * mine * (n - nzl)
    assert abs(k["fwd_y13"]["egress_GBps"] - 13 * 16 * nzl * 3000 / 2e-3 / 1e9) < 1e-9
    per_step = (13 * 16 * nzl * 3000 * 10 + 11 * 16 * mine * (n - nzl) * 30) / 10
    assert out["egress_bytes_per_step"] == per_step


def test_bench_line_is_assembled_on_the_emulator():
    """bench.py's single-rank measurement code, run on the CPU kernel emulator with the few torch.cuda calls stubbed
    (tests/bench_on_emulator.py): every key of the driver's contract is present and well-formed.  The numbers are
    meaningless here; this guards the code that the driver runs once per round on a real B200."""
    import json
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    emu = build_emu.build()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_on_emulator.py"), emu, "16"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "nvlink"):
        assert k in d, k
    assert d["metric"] == "grid_point_steps_per_s" and d["dtype"] == "f64" and d["n_gpus"] == 1 and d["steps"] == 2
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["vs_baseline"] is None and d["nvlink"] is None
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * 16 ** 3 * 8 / 2
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "step_frac", "per_kernel_frac"}
    assert d["parity"]["ok"] is True and d["parity"]["ranks"] == 1 and d["parity"]["cases"][0]["max_rel_l2"] < 1e-11
    assert d["k0_mode"]["ok"] is True and d["k0_mode"]["expected"] < 1.0
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    assert "workload" in d["config"] and "decompose_1d slabs" in d["config"]["decomposition"]
    assert d["state_finite"] is True


def test_bench_line_multi_rank_on_the_emulator():
    """The same with 4 ranks wired through gloo (the helper swaps NCCL for gloo and device tensors for host ones):
    blob exchange, max over ranks, the round-robin default ownership in config.decomposition, the NVLink section."""
    import json
    import socket
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    emu = build_emu.build()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 4
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "bench_on_emulator.py"), emu, "16"],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    for r, (p, (so, se)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r}:\n{se[-3000:]}"
    assert all(not so.strip() for so, _ in outs[1:]), "only rank 0 prints"
    d = json.loads(outs[0][0].strip().splitlines()[-1])
    assert d["n_gpus"] == world and d["scaling"] == "strong" and d["cpu_baseline"] is None
    assert d["parity"]["ok"] is True and d["parity"]["ranks"] == world
    assert sorted(c["y_stride"] for c in d["parity"]["cases"]) == [1, world] and d["k0_mode"]["ok"] is True
    assert "y_stride 4" in d["config"]["decomposition"]
    nv = d["nvlink"]
    assert set(nv["per_kernel"]) == {"fwd_y13", "spec_z", "curl_b_inv_z"} and nv["egress_bytes_per_step"] > 0
    assert d["roofline"]["kernel"] != "xchg_barrier" and d["roofline"]["algorithmic_bytes_per_launch"] > 0


def test_library_reported_bytes_equal_the_host_model():
    """laps_get_profile_bytes (what bench.py's roofline now uses) against kernel_bytes, the host-side statement of
    DESIGN.md section 4, for the headline physics on the emulator: per launch for the passes, summed over the three
    stages for the z passes (the model averages the RK-history traffic over the stages)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    from laps_b200 import Solver
    b = _bench()
    emu = build_emu.build()
    n = 32
    kw = b.workload_params(n)
    rng = __import__("numpy").random.default_rng(0)
    prim = __import__("numpy").ones((8, n, n, n))
    prim[1:4] = 0.01 * rng.standard_normal((3, n, n, n))
    prim[4:7] += 0.01 * rng.standard_normal((3, n, n, n))
    with Solver(emu, **kw) as g:
        g.set_primitive(prim)
        g.vardt()
        g.step()
        g.set_profiling(True)
        g.step()
        prof = g.get_profile(with_bytes=True)
        nf, ni, rows = g.field_counts()
        nkx, kymax, nkyl = g.pruning()
        cols, modes = g.pruning_counts()
        R, C = 8.0 * n ** 3, 16.0 * g.nxh * n * n
        kb = lambda k: b.kernel_bytes(k, R, C, nf, ni, True, nkx / g.nxh, cols / (g.nxh * n), modes / (g.nxh * n * n), rows < 8)  # noqa: E731
        sums = {}
        for name, ms, by in prof:
            sums.setdefault(name, []).append(by)
        for name, vals in sums.items():
            model = kb(name)
            if name in ("spec_z", "curl_b_inv_z"):
                # (curl_b_inv_z of the third stage carries no J tasks in the expanding box: compare spec_z only)
                if name == "spec_z":
                    assert len(vals) == 3 and abs(sum(vals) - 3 * model) < 1e-9 * sum(vals), (name, vals, model)
            elif model is not None:
                assert all(abs(v - model) < 1e-9 * model for v in vals), (name, vals, model)
        assert {"flux", "flux+cfl", "fwd_x13", "fwd_y13", "inv_y11", "inv_x11", "spec_z"} <= set(sums)


def test_reference_arm_prints_the_contract_line_and_only_rank_0_runs():
    """`bench.py --impl reference` (the CPU restatement oracle/laps_cpu.c on all host cores) needs no GPU: the JSON line of the
    contract, a cpu_baseline block that describes the run, zero copy bytes; under a launcher every rank but 0 exits silently."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cpu-n", "32", "--steps", "2", "--warmup", "1"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run(cmd, env=dict(env, OMP_NUM_THREADS="1"), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "grid_point_steps_per_s" and d["unit"] == "grid-point-steps/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["value"] - 32 ** 3 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "32^3" in cb["sample"]
    # the launcher's OMP_NUM_THREADS=1 is ignored: every host core is used and the count is printed
    assert cb["cores"] == (os.cpu_count() or 1) and f"{cb['cores']} threads" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "512^3" in d["config"]["workload"] and "32^3" in d["config"]["sample"]
    out = subprocess.run(cmd + ["--gpus", "2"], env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip() == ""
