#!/usr/bin/env python
"""Frames per second of the drop-in binaries (the UNMODIFIED reference host code linked against libfh264_b200.so)
on a synthetic 1080p clip, from the harness' own per-picture clock (steady_clock around selectNALUnitType + RBSP_encode + writeNAL,
oracle/ref_harness/driver.cpp): median over the P pictures after the first. usage: dropin_fps.py [frames]"""
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
    cmd = [exe, y4m, os.path.join(tmp, name + ".264"), "-", str(N), "28", "0", "32", "3", "1000", "0", "-1"]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    wall = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr.decode()[-500:]
    summ = json.loads([l for l in r.stdout.decode().splitlines() if l.startswith("{")][-1])     # the harness' own per-picture clock (oracle/ref_harness/driver.cpp)
    tp = sorted(summ["t_picture"][2:])                      # P pictures after the first (selectNALUnitType + RBSP_encode + writeNAL, steady_clock)
    med = tp[len(tp) // 2]
    out[name] = {"process_wall_s": round(wall, 2), "i_picture_s": round(summ["t_picture"][0], 3), "p_picture_ms_median": round(1000 * med, 2),
                 "p_picture_ms_min": round(1000 * tp[0], 2), "p_pictures_per_s": round(1.0 / med, 1)}
print(json.dumps({"clip": "synthetic 1080p, seed 100, %d pictures (1 I + P), QP 28, WindowSize 32, MAXDIFF 3; one sequence, batch 1, synchronous host loop of the reference" % N, "binaries": out}, indent=1))
