#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(out.splitlines()))
fname, hdr, lines, seen_kernel = None, None, {}, 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) >= 2 and r[0] == "Function Name":
        pass
    elif len(r) > 10 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = dict(zip(hdr, r))
        key = (fname, int(r[0]))
        if key in lines:
            continue  # the report repeats sections per launch: keep the first
        try:
            lines[key] = (int(d["Instructions Executed"]), int(d["# Samples"]), r[1].strip(), {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0})
        except (ValueError, KeyError):
            pass
tot_i = sum(v[0] for v in lines.values()) or 1
tot_s = sum(v[1] for v in lines.values()) or 1
print("kernel %s: %d source lines, %.1f M warp instructions, %d stall samples" % (kern, len(lines), tot_i / 1e6, tot_s))
print("--- by instructions executed")
for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100 * v[0] / tot_i, 100 * v[1] / tot_s, f, l, v[2][:90]))
print("--- by stall samples")
for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    st = sorted(v[3].items(), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s   %s" % (100 * v[1] / tot_s, 100 * v[0] / tot_i, f, l, v[2][:70], st))
