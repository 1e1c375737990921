"""Three fh264_encode_i calls on 8 x 1080p pictures (for an ncu capture of k_intra: skip the first launch)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import h264_fer_b200 as fh  # noqa: E402
from h264_fer_b200 import synth  # noqa: E402

nseq = 8
frames = [synth.SynthClip(1920, 1080, 100 + b).frame(0) for b in range(nseq)]
frames = [(synth.crop16(y), synth.crop16(u, True), synth.crop16(v, True)) for y, u, v in frames]
with fh.Session(1920, 1072, batch=nseq) as s:
    for it in range(3):
        for b in range(nseq):
            s.upload_source(b, *frames[b])
        s.encode_i(28)
        print("k_intra %.3f ms" % s.last_intra_ms())
