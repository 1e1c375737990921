"""TEST INFRASTRUCTURE (helper of tests/test_intra_host.py, run as a subprocess so that it can be started under ThreadSanitizer):
runs the I-picture core as 32 host threads per macroblock (tests/intra_host_lanes.cpp) on the named golden fixtures and compares
every record and reconstruction with the reference. usage: python intra_lanes_check.py lib.so fixture [fixture ...]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_intra_host as T  # noqa: E402
from h264_fer_b200 import native as fh  # noqa: E402


def main():
    lib = C.CDLL(sys.argv[1])
    vp = C.c_void_p
    for name in sys.argv[2:]:
        params, pics = T.golden_pictures(name)
        prev = None
        for n, p in enumerate(pics):
            if "imbrec" not in p:
                prev = p["mbrec"][:, 0].astype(np.int32)
                continue
            h, w = p["SRCY"].shape
            out = np.zeros((w // 16) * (h // 16), fh.MB_RESULT_I_DTYPE)
            src = [np.ascontiguousarray(p[t]) for t in ("SRCY", "SRCU", "SRCV")]
            rec = [np.zeros_like(a) for a in src]
            pt = None if prev is None else np.ascontiguousarray(prev, np.int32)
            lib.intra_host_picture_lanes(*[a.ctypes.data_as(vp) for a in src], *[a.ctypes.data_as(vp) for a in rec], w, h, int(params[4]),
                                         None if pt is None else pt.ctypes.data_as(vp), out.ctypes.data_as(vp))
            T.compare_i_records(fh.i_records_to_ints(out), p["imbrec"], "%s picture %d (32 lanes)" % (name, n))
            for r, t in zip(rec, ("RECY", "RECU", "RECV")):
                assert np.array_equal(r, p[t]), "%s picture %d: %s differs" % (name, n, t)
            prev = p["imbrec"][:, 0].astype(np.int32)
        print("LANES-OK", name)


if __name__ == "__main__":
    main()
