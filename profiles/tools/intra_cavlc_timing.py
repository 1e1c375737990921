"""Times fh264_cavlc_i (device CAVLC of I slices) after fh264_encode_i on 8 x 1080p pictures: host wall clock of the synchronous
call (one coder kernel, scan, pack, D2H of the slice data). Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import h264_fer_b200 as fh  # noqa: E402
from h264_fer_b200 import synth  # noqa: E402

nseq = 8
frames = [synth.SynthClip(1920, 1080, 100 + b).frame(0) for b in range(nseq)]
frames = [(synth.crop16(y), synth.crop16(u, True), synth.crop16(v, True)) for y, u, v in frames]
with fh.Session(1920, 1072, batch=nseq) as s:
    for b in range(nseq):
        s.upload_source(b, *frames[b])
    s.encode_i(28)
    t = []
    for it in range(4):
        t0 = time.perf_counter()
        res = s.cavlc_i(first_bit=5)
        t.append((time.perf_counter() - t0) * 1e3)
print(json.dumps({"workload": "8 x 1080p I pictures, qp 28", "cavlc_i_call_ms": round(min(t[1:]), 3), "slice_bytes_per_picture": [int((n + 7) // 8) for _, n in res], "record_bytes_per_picture": 8040 * 832}))
