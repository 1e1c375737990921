#!/usr/bin/env python
"""Per-partition clock64() duration of k_stage2 on one 1080p picture (debug timeline slots 3..5 of each partition)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import h264_fer_b200 as fh
from h264_fer_b200 import synth

W, H = 1920, 1080
c = synth.SynthClip(W, H, 100)
fr = [tuple(synth.crop16(p, chroma=(i > 0)) for i, p in enumerate(c.frame(t))) for t in range(3)]
with fh.Session(W, fr[0][0].shape[0]) as s:
    s.upload_recon(0, *fr[0])
    s.debug_timeline(0, read=False)
    for t in (1, 2):
        s.upload_source(0, *fr[t])
        s.encode_p(28, 32, 3)
    tl = s.debug_timeline(0).reshape(-1, 6)
    print(s.last_timings())
d, ns, n2 = tl[:, 3].astype(float), tl[:, 4], tl[:, 5]
print("k_stage2 per-partition cycles: mean %.0f median %.0f p99 %.0f max %.0f (%.1f us)" % (d.mean(), np.median(d), np.percentile(d, 99), d.max(), d.max() / 1965))
print("gated survivors ns: mean %.0f max %d; kept n2: mean %.0f max %d" % (ns.mean(), ns.max(), n2.mean(), n2.max()))
i = np.argsort(-d)[:8]
print("slowest:", [(int(k), int(d[k]), int(ns[k]), int(n2[k])) for k in i])
