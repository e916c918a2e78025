#!/usr/bin/env python
"""profiles/ncu_summary.json from one `ncu --set full` capture of the dominant render kernel
(raw page as CSV: `ncu -i report.ncu-rep --page raw --csv > raw.csv`).

Usage: ncu_summary.py <raw.csv> <paths traced by the captured launch> "<how it was captured>" [config+variant, e.g. 4A] [fast|parity]

The summary records the fingerprint of the CUDA sources it was captured from (the same sha1 bench.py
computes at run time): bench.py applies the per-path constants only to a build with that fingerprint.

bench.py reads the result for `roofline.traffic` (DRAM bytes per launch -- the framebuffer
traffic of a launch does not depend on spp) and for the issue-slot figures (warp instructions per
path x measured paths/s against the SM's 4 warp instructions per clock)."""
import csv
import hashlib
import json
import os
import sys


def source_fingerprint():
    h = hashlib.sha1()
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tiny-path-tracer_b200", "csrc")
    for fn in sorted(os.listdir(d)):
        if fn.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(d, fn), "rb").read())
    return h.hexdigest()[:16]

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def main():
    raw, paths, how = sys.argv[1], float(sys.argv[2]), sys.argv[3]
    config = sys.argv[4] if len(sys.argv) > 4 else "4A"
    mode = sys.argv[5] if len(sys.argv) > 5 else "fast"
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]

    def get(name, scale=None):
        i = hdr.index(name)
        v = float(vals[i].replace(",", ""))
        return v * (scale or {}).get(units[i], 1.0)

    name = vals[hdr.index("Kernel Name")]
    key = "render_wave_kernel" if "render_wave" in name else "render_mega_kernel"
    rd, wr = get("dram__bytes_read.sum", SCALE), get("dram__bytes_write.sum", SCALE)
    inst = get("smsp__inst_executed.sum")
    out = {key: {
        "source": f"{os.path.relpath(raw)} ({how})",
        "kernel": name.strip(),
        "source_fingerprint": source_fingerprint(),
        "file": os.path.relpath(raw),
        "config": config,
        "mode": mode,
        "paths": paths,
        "dram_bytes_read": rd,
        "dram_bytes_write": wr,
        "dram_bytes_per_launch": rd + wr,
        "duration_ms": get("gpu__time_duration.sum", TIME),
        "registers_per_thread": int(get("launch__registers_per_thread")),
        "avg_active_threads_per_inst": get("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "issue_active_pct": get("sm__inst_issued.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "l1_hit_pct": get("l1tex__t_sector_hit_rate.pct"),
        "l2_hit_pct": get("lts__t_sector_hit_rate.pct"),
        "pipe_alu_pct": get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "pipe_fma_pct": get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "inst_executed": inst,
        "warp_inst_per_path": inst / paths,
    }}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
