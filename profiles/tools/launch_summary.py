#!/usr/bin/env python
"""Markdown table of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total, average, share per kernel."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ci, cv, cu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    try:
        v = float(r[cv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[cu], 1.0)
    k = r[ci].split("(")[0].replace("void ", "")
    tot[k] += v; cnt[k] += 1
s = sum(tot.values())
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k in sorted(tot, key=lambda k: -tot[k]):
    print("| %s | %d | %.1f | %.1f | %.1f%% |" % (k, cnt[k], tot[k], tot[k] / cnt[k], 100 * tot[k] / s))
