#!/usr/bin/env python
"""Static SASS instructions per source file / line bucket of one kernel, split into executed and never-executed
(instruction-cache footprint). usage: ncu_codesize.py report.ncu-rep kernel_regex"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(out.splitlines()))
hdr = None
tot = ex = 0
for r in rows:
    if len(r) > 5 and r[0] == "Address":
        if hdr is not None:
            break           # first launch only
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        tot += 1
        try:
            if int(d["Instructions Executed"]) > 0:
                ex += 1
        except ValueError:
            pass
print("%s: %d SASS instructions (%.1f KB), executed at least once: %d (%.1f KB)" % (kern, tot, tot * 16 / 1024, ex, ex * 16 / 1024))
