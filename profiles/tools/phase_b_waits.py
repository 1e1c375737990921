#!/usr/bin/env python
"""Where the phase-B wavefront waits for the LEFT macroblock (fh264_debug_timeline slots 10/11) on one 1080p picture."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import h264_fer_b200 as fh
from h264_fer_b200 import synth

W, H = 1920, 1080
c = synth.SynthClip(W, H, 100)
fr = [tuple(synth.crop16(p, chroma=(i > 0)) for i, p in enumerate(c.frame(t))) for t in range(3)]
Hc = fr[0][0].shape[0]
with fh.Session(W, Hc) as s:
    s.upload_recon(0, *fr[0])
    s.debug_timeline(0, read=False)
    for t in (1, 2):
        s.upload_source(0, *fr[t])
        rec = s.encode_p(28, 32, 3)[0]
    tl = s.debug_timeline(0)
    print("last_timings", s.last_timings())
skip = rec["mb_type"] == 31
n = len(tl)
where = tl[:, 10].astype(np.int64)
names = ["skip: B==C, one candidate passes", "skip: B!=C", "partition 0 (up q2 != up q3)", "partition 2 (q0 != q1)", "merge: left q1", "merge: left q3"]
print("MBs %d, skip %d" % (n, skip.sum()))
for i, nm in enumerate(names):
    cnt = ((where >> (8 * i)) & 255) > 0
    print("  left fetched at %-34s %5d MBs (%.1f%%)" % (nm, cnt.sum(), 100.0 * cnt.mean()))
crit = (((where >> 0) & 255) + ((where >> 8) & 255) + ((where >> 16) & 255)) > 0
print("MBs whose first half depends on the left MB (serial chain): %.1f%%" % (100.0 * crit.mean()))
print("time spent waiting for / fetching the left MB per MB: mean %.0f ns, median %.0f, p90 %.0f" % (tl[:, 11].mean(), np.median(tl[:, 11]), np.percentile(tl[:, 11], 90)))
d = np.diff(tl[:, :10], axis=1).astype(float)
seg = ["prefetch issue", "wait row above", "nb mv load", "skip test", "part0", "part1", "part2", "part3", "merge+publish"]
ns = d[~skip]
for i, nm in enumerate(seg):
    print("  %-16s mean %8.0f  median %8.0f  p90 %8.0f" % (nm, ns[:, i].mean(), np.median(ns[:, i]), np.percentile(ns[:, i], 90)))
sub = np.diff(np.concatenate([tl[~skip][:, 5:6], tl[~skip][:, 12:18], tl[~skip][:, 6:7]], axis=1), axis=1).astype(float)
for i, nm in enumerate(["predictor", "issue feature loads + stage-2/3 keys", "stage-1 keys + barrier", "stage-1 select", "SAD loads issue + lazy stage 2", "SAD reduce + block min", "decode + publish"]):
    print("  part1 %-38s mean %7.0f median %7.0f" % (nm, sub[:, i].mean(), np.median(sub[:, i])))
print("picture wavefront: first start -> last publish = %.3f ms" % ((tl[:, 9].max() - tl[:, 0].min()) / 1e6))
Wmb = W // 16
T = tl.reshape(-1, Wmb, tl.shape[1])
rows_t = (T[:, :, 9].max(1) - T[:, :, 0].min(1)) / 1e3
print("per MB row: first start -> last end, median %.0f us; row-to-row start lag median %.1f us" % (np.median(rows_t), np.median(np.diff(T[:, 0, 2])) / 1e3))
