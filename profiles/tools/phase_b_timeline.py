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
print("non-skip MBs: %d, skip: %d  (cycles @ SM clock; 1965 MHz => 1000 cyc = 0.51 us)" % ((~skip).sum(), skip.sum()))
for i, n in enumerate(names):
    print("  %-16s mean %8.0f  median %8.0f  p90 %8.0f" % (n, ns[:, i].mean(), np.median(ns[:, i]), np.percentile(ns[:, i], 90)))
print("  busy (after deps) mean %.0f cycles = %.1f us" % (ns[:, 2:].sum(1).mean(), ns[:, 2:].sum(1).mean() / 1965.0))
sub = np.diff(np.concatenate([tl[~skip][:, 4:5], tl[~skip][:, 10:18], tl[~skip][:, 5:6]], axis=1), axis=1).astype(float)
for i, n in enumerate(["mvp + issue feature loads", "stage-2 keys + barrier", "stage-2 select", "stage-2/3 evaluate", "stage-1 keys + barrier",
                       "stage-1 select", "stage-1 SAD", "warp min + barrier", "decode"]):
    print("  part0 %-26s mean %7.0f median %7.0f" % (n, sub[:, i].mean(), np.median(sub[:, i])))
sk = d[skip]
if len(sk):
    print("skip MBs: after-deps mean %.0f cycles" % (tl[skip, 4] - tl[skip, 2]).mean())
