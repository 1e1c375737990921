"""Times fh264_encode_i (device I pictures) at 1080p: CUDA-event time of the wavefront kernel and wall
clock of the whole synchronous call (kernel + record D2H + dpb swap + phase R), for 1 and 8 sequences per launch and both lane
modes. Prints one JSON line. Usage: python profiles/tools/intra_timing.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import h264_fer_b200 as fh  # noqa: E402
from h264_fer_b200 import synth  # noqa: E402

W, H, QP = 1920, 1072, 28
out = {"workload": "1080p (coded 1920x1072) synthetic I pictures, qp 28", "runs": []}
for nseq in (1, 8):
    frames = [synth.SynthClip(1920, 1080, 100 + b).frame(0) for b in range(nseq)]
    frames = [(synth.crop16(y), synth.crop16(u, True), synth.crop16(v, True)) for y, u, v in frames]
    with fh.Session(W, H, batch=nseq) as s:
        for lanes in (32, 1):
            os.environ["FH264_INTRA_LANES"] = str(lanes)
            ker, wall = [], []
            for it in range(4):
                for b in range(nseq):
                    s.upload_source(b, *frames[b])
                s.sync()
                t0 = time.perf_counter()
                rec = s.encode_i(QP)
                wall.append((time.perf_counter() - t0) * 1e3)
                ker.append(s.last_intra_ms())
            k, w_ = min(ker[1:]), min(wall[1:])
            out["runs"].append({"seqs": nseq, "lanes": lanes, "kernel_ms": round(k, 3), "call_ms": round(w_, 3),
                                "i_pictures_per_s_kernel": round(nseq / k * 1e3, 1), "i_pictures_per_s_call": round(nseq / w_ * 1e3, 1),
                                "intra4x4_share": float((rec["mb_type"] == 0).mean())})
os.environ.pop("FH264_INTRA_LANES", None)
print(json.dumps(out))
