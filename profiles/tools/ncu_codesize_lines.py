#!/usr/bin/env python
"""Executed static SASS instructions per source line (instruction-cache footprint). usage: ncu_codesize_lines.py rep kernel [top]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(out.splitlines()))
fname, hdr, cur, cnt, src, seen = None, None, None, collections.Counter(), {}, set()
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr):
        if r[0] not in ("", "Line No"):
            cur = (fname, int(r[0])); src[cur] = r[1].strip()
        elif cur and r[2].startswith("0x"):
            d = dict(zip(hdr, r))
            if (r[2]) in seen:
                continue
            seen.add(r[2])
            try:
                if int(d["Instructions Executed"]) > 0:
                    cnt[cur] += 1
            except ValueError:
                pass
byfile = collections.Counter()
for (f, l), c in cnt.items():
    byfile[f] += c
print("executed static SASS by file:", dict(byfile))
for (f, l), c in cnt.most_common(top):
    print("%5d  %s:%d  %s" % (c, f, l, src[(f, l)][:110]))
