#!/usr/bin/env python
"""Kernel-wide warp stall reasons (sampled) from an .ncu-rep source page. usage: ncu_stalls.py report.ncu-rep kernel_regex"""
import csv, subprocess, sys
from collections import Counter
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(out.splitlines()))
hdr, tot, ninst, first = None, Counter(), 0, True
for r in rows:
    if len(r) > 10 and r[0] in ("Address", "Line No", "#"):
        if hdr is not None: break        # first launch only
        hdr = r
    elif hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v.isdigit(): tot[k] += int(v)
        if d.get("Instructions Executed", "").isdigit(): ninst += int(d["Instructions Executed"])
s = sum(tot.values()) or 1
print("kernel %s: %.1f M warp instructions, %d stall samples" % (kern, ninst / 1e6, s))
for k, v in tot.most_common(12): print("%6.1f%%  %s" % (100.0 * v / s, k))
