#!/usr/bin/env python
"""Per-macroblock clock64() timeline of the phase-B wavefront (fh264_debug_timeline) on one 1080p picture."""
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
    print(s.last_timings())
skip = rec["mb_type"] == 31
names = ["prefetch issue", "wait deps", "nb mv load", "skip test", "part0", "part1", "part2", "part3", "merge+publish"]
d = np.diff(tl[:, :10], axis=1).astype(float)
ns = d[~skip]
print("non-skip MBs: %d, skip: %d  (globaltimer ns)" % ((~skip).sum(), skip.sum()))
for i, n in enumerate(names):
    print("  %-16s mean %8.0f  median %8.0f  p90 %8.0f" % (n, ns[:, i].mean(), np.median(ns[:, i]), np.percentile(ns[:, i], 90)))
print("  busy (after deps) mean %.0f ns" % (ns[:, 2:].sum(1).mean()))
sub = np.diff(np.concatenate([tl[~skip][:, 4:5], tl[~skip][:, 10:18], tl[~skip][:, 5:6]], axis=1), axis=1).astype(float)
for i, n in enumerate(["mvp + issue feature loads", "stage-2 keys + barrier", "stage-2 select", "stage-2/3 evaluate", "stage-1 keys + barrier",
                       "stage-1 select", "stage-1 SAD", "warp min + barrier", "decode"]):
    print("  part0 %-26s mean %7.0f median %7.0f" % (n, sub[:, i].mean(), np.median(sub[:, i])))
sk = d[skip]
if len(sk):
    print("skip MBs: after-deps mean %.0f cycles" % (tl[skip, 4] - tl[skip, 2]).mean())

# critical-path view: when were the dependencies published vs when did the MB notice and start?
Wmb, Hmb = W // 16, fr[0][0].shape[0] // 16
T = tl.reshape(Hmb, Wmb, -1)
lag, starve = [], []
for y in range(Hmb):
    for x in range(Wmb):
        ready = 0
        if x > 0: ready = max(ready, T[y, x - 1, 6] if T[y, x - 1, 6] else T[y, x - 1, 4])       # left: quadrant 1 published (or skip decided)
        if y > 0:
            if x < Wmb - 1: ready = max(ready, T[y - 1, x + 1, 7] if T[y - 1, x + 1, 7] else T[y - 1, x + 1, 4])
            ready = max(ready, T[y - 1, x, 8] if T[y - 1, x, 8] else T[y - 1, x, 4])
        if ready:
            lag.append(T[y, x, 2] - ready)            # > 0: noticed after the deps were ready
            starve.append(T[y, x, 1] - ready)         # > 0: the CTA had not even finished prefetching when the deps were ready
lag, starve = np.array(lag, float), np.array(starve, float)
print("dependency ready -> MB proceeds: median %.0f ns, mean %.0f, p90 %.0f" % (np.median(lag), lag.mean(), np.percentile(lag, 90)))
print("MBs whose CTA was still fetching its ticket/prefetch when the deps were ready: %.1f%% (median lateness %.0f ns)" % (100 * (starve > 0).mean(), np.median(starve[starve > 0]) if (starve > 0).any() else 0))
print("picture wavefront: first start -> last publish = %.3f ms" % ((tl[:, 9].max() - tl[:, 0].min()) / 1e6))
