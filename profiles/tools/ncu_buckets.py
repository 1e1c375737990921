#!/usr/bin/env python
"""Warp instructions / stall samples / LSU pressure of one kernel summed over source-line buckets.
usage: ncu_buckets.py report.ncu-rep kernel_regex file:lo-hi=name ...   (unmatched lines go to 'other')"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
buckets = []
for a in sys.argv[3:]:
    spec, name = a.split("=")
    f, rng = spec.split(":")
    lo, hi = rng.split("-")
    buckets.append((f, int(lo), int(hi), name))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(out.splitlines()))
fname, hdr, seen, acc = None, None, set(), {}
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = dict(zip(hdr, r))
        key = (fname, int(r[0]))
        if key in seen:
            continue
        seen.add(key)
        name = "other:" + fname
        for f, lo, hi, n in buckets:
            if f == fname and lo <= key[1] <= hi:
                name = n
                break
        def gi(k):
            try:
                return int(d.get(k, "0") or 0)
            except ValueError:
                return 0
        a = acc.setdefault(name, [0, 0, 0, 0])
        a[0] += gi("Instructions Executed"); a[1] += gi("# Samples"); a[2] += gi("L1 Wavefronts Shared"); a[3] += gi("L1 Tag Requests Global")
T = [sum(a[i] for a in acc.values()) or 1 for i in range(4)]
print("%-28s %10s %7s %7s %9s %9s" % ("bucket", "Minstr", "inst%", "smp%", "shw M", "tag M"))
for n, a in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %10.1f %6.1f%% %6.1f%% %9.1f %9.1f" % (n, a[0] / 1e6, 100 * a[0] / T[0], 100 * a[1] / T[1], a[2] / 1e6, a[3] / 1e6))
