#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into a small JSON + text table for profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_name [frames_per_launch]
Needs the `ncu` CLI (no GPU). Keeps only per-kernel headline metrics; the .ncu-rep itself stays in gpurun_out/ (scratch).
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return x


def main():
    rep, out = sys.argv[1], sys.argv[2]
    frames = int(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = {"value": num(r[i]), "unit": units[i]}
        if frames:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
            tr = sum(d[k]["value"] * scale.get(d[k]["unit"], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in d)
            d["frames_per_launch"] = frames
            d["dram_bytes_per_frame"] = tr / frames
            if "smsp__inst_executed.sum" in d:
                d["warp_instructions_per_frame"] = d["smsp__inst_executed.sum"]["value"] / frames
            t = d["gpu__time_duration.sum"]
            secs = t["value"] * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(t["unit"], 1e-9)
            d["frames_per_s_under_profiler"] = frames / secs
        res.append(d)
    json.dump(res, open(out + ".json", "w"), indent=1)
    with open(out + ".txt", "w") as f:
        for d in res:
            f.write(f"== {d['kernel']}\n")
            for k, v in d.items():
                if isinstance(v, dict):
                    f.write(f"  {k:90s} {v['value']} {v['unit']}\n")
                elif k != "kernel":
                    f.write(f"  {k:90s} {v}\n")
    print(open(out + ".txt").read())


if __name__ == "__main__":
    main()
