#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: warp instructions executed, thread
instructions, average active threads and stall samples.  Usage: ncu_lines.py report.ncu-rep [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
cur_file = ""
lines = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        d = dict(zip(hdr, r))
        try:
            inst = int(d["Instructions Executed"])
            tinst = int(d["Thread Instructions Executed"])
            samples = int(d["# Samples"])
        except ValueError:
            continue
        lines.append((inst, tinst, samples, cur_file, r[0], r[1].strip()[:90]))
tot_i = sum(l[0] for l in lines)
tot_t = sum(l[1] for l in lines)
tot_s = sum(l[2] for l in lines)
print(f"total warp-inst {tot_i:.4g}  thread-inst {tot_t:.4g}  avg active {tot_t / max(tot_i, 1):.2f}  samples {tot_s}")
# ---- per function (nearest preceding definition line in the same file) ----
import bisect, collections, os, re
here = os.path.dirname(os.path.abspath(__file__))
src_dir = os.path.join(here, "..", "..", "tiny-path-tracer_b200", "csrc")
defs = {}
for fn in os.listdir(src_dir):
    if fn.endswith((".cuh", ".cu", ".h")):
        L = open(os.path.join(src_dir, fn)).read().split("\n")
        marks = []
        for i, l in enumerate(L):
            m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:TPT_DEV|__global__|__device__)[^(]*?([A-Za-z_0-9]+)\s*\(", l)
            if m and not l.startswith("TPT_DEV V3 operator") and not l.startswith("TPT_DEV float dot") and "{ return" not in l:
                marks.append((i + 1, m.group(1)))
        defs[fn] = marks
agg = collections.defaultdict(lambda: [0, 0, 0])
for inst, tinst, samples, f, ln, src in lines:
    name = f
    if f in defs and defs[f]:
        idx = bisect.bisect_right([m[0] for m in defs[f]], int(ln)) - 1
        if idx >= 0:
            name = defs[f][idx][1]
    a = agg[name]
    a[0] += inst; a[1] += tinst; a[2] += samples
print(f"{'warp-inst%':>10} {'samples%':>9} {'active':>6}  function")
for name, (inst, tinst, samples) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{100 * inst / tot_i:10.2f} {100 * samples / max(tot_s, 1):9.2f} {tinst / max(inst, 1):6.1f}  {name}")
print()
print(f"{'warp-inst%':>10} {'samples%':>9} {'active':>6}  file:line  source")
for inst, tinst, samples, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{100 * inst / tot_i:10.2f} {100 * samples / max(tot_s, 1):9.2f} {tinst / max(inst, 1):6.1f}  {f}:{ln}  {src}")
