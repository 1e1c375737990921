"""ORACLE / TEST INFRASTRUCTURE: run oracle/_ref/ref_encoder (the unmodified reference, built by oracle/Makefile)
and parse its chunked dump (format defined in oracle/ref_harness/driver.cpp)."""
from __future__ import annotations

import json
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ENCODER = os.path.join(HERE, "_ref", "ref_encoder")
REC_INTS = 405

D_MBREC, D_RECON, D_SOURCE, D_PHASE_R, D_TQIO, D_INTRA16, D_SLICE, D_TABLES, D_IMBREC = 1, 2, 4, 8, 16, 32, 64, 128, 256
IREC_INTS = 439


def have_ref_encoder() -> bool:
    return os.path.isfile(REF_ENCODER) and os.access(REF_ENCODER, os.X_OK)


def run_reference(y4m_path, frames, qp=28, basic=0, window=16, maxdiff=3, intra_every=1000, dumpmask=0, planes_pic=-1,
                  out_264=None, dump_path=None, timeout=3600):
    """Returns (summary dict, dump path or None, .264 path)."""
    tmpdir = tempfile.mkdtemp(prefix="fh264_ref_")
    out_264 = out_264 or os.path.join(tmpdir, "out.264")
    if dumpmask and dump_path is None:
        dump_path = os.path.join(tmpdir, "dump.bin")
    cmd = [REF_ENCODER, y4m_path, out_264, dump_path if dumpmask else "-", str(frames), str(qp), str(basic), str(window),
           str(maxdiff), str(intra_every), str(dumpmask), str(planes_pic)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout, cwd=tmpdir)
    if res.returncode != 0:
        raise RuntimeError("ref_encoder failed (%d): %s" % (res.returncode, res.stderr.decode()[-2000:]))
    line = [l for l in res.stdout.decode().splitlines() if l.startswith("{")][-1]
    return json.loads(line), (dump_path if dumpmask else None), out_264


def parse_dump(path):
    """-> list of pictures; each a dict: hdr (nal_type, bytes, w, h, counts[5], qp) and the arrays present."""
    pics = {}
    with open(path, "rb") as f:
        data = f.read()
    off = 0
    while off < len(data):
        tag = data[off:off + 4].decode()
        pic, n = struct.unpack_from("<II", data, off + 4)
        payload = data[off + 12:off + 12 + n]
        off += 12 + n
        p = pics.setdefault(pic, {})
        if tag == "PICH":
            h = struct.unpack("<10i", payload)
            p["nal_type"], p["bytes"], p["w"], p["h"], p["counts"], p["qp"] = h[0], h[1], h[2], h[3], list(h[4:9]), h[9]
        elif tag == "MBRC":
            p["mbrec"] = np.frombuffer(payload, dtype=np.int32).reshape(-1, REC_INTS).copy()
        elif tag == "IMBR":
            p["imbrec"] = np.frombuffer(payload, dtype=np.int32).reshape(-1, IREC_INTS).copy()
        elif tag == "I16M":
            p["i16"] = np.frombuffer(payload, dtype=np.int16).reshape(-1, 1024).copy()
        elif tag == "TQIO":
            p["tqio"] = np.frombuffer(payload, dtype=np.uint8).reshape(-1, 768).copy()
        elif tag in ("RECY", "RECU", "RECV", "SRCY", "SRCU", "SRCV"):
            p[tag] = np.frombuffer(payload, dtype=np.uint8).copy()
        elif tag == "PLNE":
            p.setdefault("planes", []).append(np.frombuffer(payload, dtype=np.uint8).copy())
        elif tag == "KARF":
            p.setdefault("kar", []).append(np.frombuffer(payload, dtype=np.uint16).copy())  # order: f-major, then k
        elif tag == "SORT":
            p.setdefault("sorted", []).append(np.frombuffer(payload, dtype=np.int32).copy())
        elif tag == "KOLI":
            p["koliko"] = np.frombuffer(payload, dtype=np.int32).copy()
        elif tag == "SLDT":         # P picture: bit position of the first slice_data bit, then the whole slice RBSP
            p["slice_bit0"] = struct.unpack_from("<i", payload, 0)[0]
            p["rbsp"] = np.frombuffer(payload[4:], dtype=np.uint8).copy()
        elif tag == "CVTB":         # the reference's CAVLC coder tables (see driver.cpp for the order)
            p["cavlc_tables"] = np.frombuffer(payload, dtype=np.int32).copy()
    out = []
    for k in sorted(pics):
        d = pics[k]
        if "w" in d:
            w, h = d["w"], d["h"]
            for t in ("RECY", "SRCY"):
                if t in d:
                    d[t] = d[t].reshape(h, w)
            for t in ("RECU", "RECV", "SRCU", "SRCV"):
                if t in d:
                    d[t] = d[t].reshape(h // 2, w // 2)
        out.append(d)
    return out
