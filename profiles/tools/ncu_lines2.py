#!/usr/bin/env python
"""Per-source-line table of one kernel from an .ncu-rep (needs -lineinfo + --import-source on): warp instructions, stall samples,
shared-memory wavefronts, global L1 tag requests and L2 sectors — the LSU data-pipe pressure per line.
usage: ncu_lines2.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(out.splitlines()))
fname, hdr, lines = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = dict(zip(hdr, r))
        key = (fname, int(r[0]))
        if key in lines:
            continue
        def gi(k):
            try:
                return int(d.get(k, "0") or 0)
            except ValueError:
                return 0
        lines[key] = dict(inst=gi("Instructions Executed"), smp=gi("# Samples"), src=r[1].strip(), shw=gi("L1 Wavefronts Shared"),
                          shi=gi("L1 Wavefronts Shared Ideal"), tag=gi("L1 Tag Requests Global"), sec=gi("L2 Theoretical Sectors Global"),
                          seci=gi("L2 Theoretical Sectors Global Ideal"))
T = {k: sum(v[k] for v in lines.values()) or 1 for k in ("inst", "smp", "shw", "tag", "sec", "shi", "seci")}
print("kernel %s: %.1f M warp instr, %d samples, shared wavefronts %.1f M (ideal %.1f M), global tag requests %.1f M, L2 sectors %.1f M (ideal %.1f M)"
      % (kern, T["inst"] / 1e6, T["smp"], T["shw"] / 1e6, T["shi"] / 1e6, T["tag"] / 1e6, T["sec"] / 1e6, T["seci"] / 1e6))
for name, k in (("instructions", "inst"), ("stall samples", "smp"), ("shared wavefronts", "shw"), ("global tag requests", "tag")):
    print("--- by " + name)
    for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][k])[:top]:
        if v[k] == 0:
            break
        print("%5.1f%% %-5s | inst %4.1f%% smp %4.1f%% shw %4.1f%% tag %4.1f%% | %s:%d  %s" % (100 * v[k] / T[k], k, 100 * v["inst"] / T["inst"], 100 * v["smp"] / T["smp"],
              100 * v["shw"] / T["shw"], 100 * v["tag"] / T["tag"], f, l, v["src"][:100]))
