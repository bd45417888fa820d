#!/usr/bin/env python
"""Per-kernel DRAM traffic of one RK stage from an `ncu --set full` capture (raw page as CSV) -> profiles/traffic.json, the
file bench.py reads `roofline.traffic` from, plus a markdown table of the counters that explain each kernel.
Usage: python tools/ncu_traffic.py <raw.csv> <config> <n> <n_gpus> <source label> [out.json] [out.md]"""
import csv
import json
import re
import sys


def launch_name(kernel, grid):
    gy = int(grid.strip("() ").split(",")[1])
    if "k_fwd_x" in kernel: return f"fwd_x{gy}"
    if "k_fwd_y" in kernel: return f"fwd_y{gy}"
    if "k_inv_y" in kernel: return f"inv_y{gy}"
    if "k_inv_x" in kernel: return f"inv_x{gy}"
    if "k_flux<(bool)1>" in kernel or "k_flux<true>" in kernel or "k_flux<1>" in kernel: return "flux+cfl"
    if "k_flux" in kernel: return "flux"
    if "k_rhs_z" in kernel: return "spec_z"
    if "k_spec_z" in kernel: return {4: "curl_b_inv_z", 3: "curl_b_inv_z3", 1: "mass_inv_z"}.get(gy, f"spec_z_rows{gy}")
    if "k_cfl" in kernel: return "cfl"
    return re.sub(r"<.*", "", kernel.split("::")[-1])


def main():
    raw, config, n, ngpu, label = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    out_json = sys.argv[6] if len(sys.argv) > 6 else "profiles/traffic.json"
    out_md = sys.argv[7] if len(sys.argv) > 7 else None
    rows = list(csv.reader(open(raw)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, key):
        try:
            return float(r[col[key]].replace(",", ""))
        except Exception:
            return float("nan")
    units = rows[1]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    kernels, table = {}, []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = launch_name(r[col["Kernel Name"]], r[col["Grid Size"]])
        rd = val(r, "dram__bytes_read.sum") * scale.get(units[col["dram__bytes_read.sum"]], 1.0)
        wr = val(r, "dram__bytes_write.sum") * scale.get(units[col["dram__bytes_write.sum"]], 1.0)
        t_ms = val(r, "gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[col["gpu__time_duration.sum"]], 1e-6)
        k = kernels.setdefault(name, {"dram_bytes_per_launch": 0.0, "launches": 0, "ms": 0.0})
        k["dram_bytes_per_launch"] += rd + wr; k["launches"] += 1; k["ms"] += t_ms
        table.append((name, r[col["Kernel Name"]][:38], r[col["Grid Size"]], t_ms, rd / 1e9, wr / 1e9,
                      val(r, "launch__registers_per_thread"), val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                      val(r, "sm__inst_issued.avg.pct_of_peak_sustained_active") if "sm__inst_issued.avg.pct_of_peak_sustained_active" in col else float("nan"),
                      val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                      val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                      (rd + wr) / 1e9 / (t_ms * 1e-3) if t_ms > 0 else float("nan")))
    for k in kernels.values():
        k["dram_bytes_per_launch"] /= k["launches"]; k["ms"] /= k["launches"]
    json.dump({"config": {"config": config, "n": n, "n_gpus": ngpu}, "source": label, "kernels": kernels}, open(out_json, "w"), indent=1)
    if out_md:
        with open(out_md, "w") as f:
            f.write("| launch | kernel | grid | time ms | dram read GB | dram write GB | regs | warps active % | issue % | fp64 pipe % | shared-memory wavefronts % | DRAM GB/s |\n|" + "---|" * 12 + "\n")
            for t in table:
                f.write("| `%s` | `%s` | %s | %.3f | %.3f | %.3f | %.0f | %.1f | %.1f | %.1f | %.1f | %.0f |\n" % t)
    print(json.dumps(kernels, indent=1))


if __name__ == "__main__":
    main()
