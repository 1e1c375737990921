#!/usr/bin/env python
"""Frames per second of the drop-in binaries (the UNMODIFIED reference host code linked against libfh264_b200.so)
on a synthetic 1080p clip: wall clock of the whole process (Y4M parsing, NAL writing and CUDA start-up included) for N and for 2
pictures; the difference / (N - 2) is the steady-state time per picture. usage: dropin_fps.py [frames]"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from h264_fer_b200 import synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 22
tmp = tempfile.mkdtemp(prefix="fh264_dropin_")
y4m = os.path.join(tmp, "in.y4m")
synth.write_y4m(y4m, 1920, 1080, 100, N)
out = {}
for name in ("fh264_encoder_b200", "fh264_encoder_b200_cavlc", "fh264_encoder_b200_intra", "fh264_encoder_b200_all"):
    exe = os.path.join(ROOT, "integration", "_build", name)
    if not os.path.isfile(exe):
        continue
    t = {}
    for n in (2, N):
        cmd = [exe, y4m, os.path.join(tmp, name + ".264"), "-", str(n), "28", "0", "32", "3", "1000", "0", "-1"]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        t[n] = time.perf_counter() - t0
        assert r.returncode == 0, r.stderr.decode()[-500:]
    per = (t[N] - t[2]) / (N - 2)
    out[name] = {"seconds_%d_pictures" % N: round(t[N], 3), "seconds_2_pictures": round(t[2], 3), "ms_per_p_picture": round(1000 * per, 2), "p_pictures_per_s": round(1.0 / per, 1)}
print(json.dumps({"clip": "synthetic 1080p, seed 100, %d pictures (1 I + P), QP 28, WindowSize 32, MAXDIFF 3; one sequence, batch 1, synchronous host loop of the reference" % N, "binaries": out}, indent=1))
