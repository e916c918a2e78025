#!/usr/bin/env python
"""Key counters of one-kernel ncu reports as JSON (cache hit rates, DRAM traffic, issue rate,
divergence). Usage: ncu_scene_summary.py name=report.ncu-rep [name=report ...]"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "kernel_ms",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_inst",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed.avg.per_cycle_active": "ipc_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "l1tex__t_bytes.sum": "l1_bytes",
    "lts__t_bytes.sum": "l2_bytes",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__icc_request_hit_rate.pct": "icache_hit_pct",
    "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed": "gpc_instruction_fetch_pct_of_peak",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction_per_issue",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier_per_issue",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait_per_issue",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard_per_issue",
}
out = {}
for arg in sys.argv[1:]:
    name, rep = arg.split("=", 1)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")]}
    for k, label in WANT.items():
        if k in hdr:
            i = hdr.index(k)
            try:
                v = float(vals[i].replace(",", ""))
            except ValueError:
                continue
            d[label] = v
            if units[i] and units[i] not in ("%", ""):
                d[label + "_unit"] = units[i]
    if "dram_read" in d and "kernel_ms" in d:
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
        rd = d["dram_read"] * scale.get(d.get("dram_read_unit", "byte"), 1.0)
        wr = d["dram_write"] * scale.get(d.get("dram_write_unit", "byte"), 1.0)
        ms = d["kernel_ms"] * (1e-3 if d.get("kernel_ms_unit") == "us" else 1.0) * (1e3 if d.get("kernel_ms_unit") == "s" else 1.0)
        d["dram_GBps"] = (rd + wr) / (ms * 1e-3) / 1e9
    out[name] = d
json.dump(out, sys.stdout, indent=1)
print()
