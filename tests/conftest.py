import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


class Golden:
    """A committed fixture produced by tests/golden/make_golden.py from the compiled reference."""

    def __init__(self, path):
        self.name = os.path.splitext(os.path.basename(path))[0]
        z = np.load(path)
        self.w_in, self.h_in, self.seed, self.frames, self.qp, self.window, self.maxdiff, self.basic = [int(v) for v in z["params"]]
        self.types = [int(t) for t in z["types"]]
        self.w, self.h = z["RECY_0"].shape[1], z["RECY_0"].shape[0]
        self.z = z

    def src(self, n):
        return self.z["SRCY_%d" % n], self.z["SRCU_%d" % n], self.z["SRCV_%d" % n]

    def rec(self, n):
        return self.z["RECY_%d" % n], self.z["RECU_%d" % n], self.z["RECV_%d" % n]

    def mbrec(self, n):
        return self.z["mbrec_%d" % n].astype(np.int32)

    def tqio(self, n):
        return self.z["tqio_%d" % n]

    def counts(self, n):
        return [int(c) for c in self.z["counts_%d" % n]]

    def i16(self):
        """Intra16x16 luma records of the I pictures: [n, 1024] int16 = 256 src, 256 pred, 16 dc, 240 ac, 256 recon."""
        out = [self.z[k] for k in self.z.files if k.startswith("i16_")]
        return np.concatenate(out) if out else np.zeros((0, 1024), np.int16)

    def records(self, n):
        """P picture n as fh264_mb_result records (the layout the device CAVLC and the decoder path read)."""
        from h264_fer_b200 import native as fh
        r = self.mbrec(n)
        out = np.zeros(len(r), fh.MB_RESULT_DTYPE)
        out["mb_type"] = r[:, 0]
        nparts = {0: 1, 1: 2, 2: 2, 4: 4, 31: 0}
        out["num_parts"] = [nparts[int(t)] for t in r[:, 0]]
        out["mv"] = r[:, 1:9].reshape(-1, 4, 2)
        out["mvd"] = r[:, 9:17].reshape(-1, 4, 2)
        out["sad"] = r[:, 17:21]
        out["luma"] = r[:, 21:277].reshape(-1, 16, 16)
        out["chroma_dc"] = r[:, 277:285].reshape(-1, 2, 4)
        out["chroma_ac"] = r[:, 285:405].reshape(-1, 2, 4, 15)
        return out

    def slice_rbsp(self, n):
        """P picture n: (RBSP bytes of the slice NAL, bit position of the first slice_data bit) as written by the reference."""
        return self.z["rbsp_%d" % n], int(self.z["slbit0_%d" % n][0])

    def p_pictures(self):
        return [n for n, t in enumerate(self.types) if t == 1]


def golden_paths():
    # P-path clip fixtures only (cavlc_tables.npz holds the reference's CAVLC coder tables, see tests/test_cavlc_host.py; the
    # intra_*.npz fixtures have their own layout and tests: tests/test_intra_host.py, tests/test_gpu_intra.py)
    return sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if os.path.basename(p) != "cavlc_tables.npz" and not os.path.basename(p).startswith("intra_"))


@pytest.fixture(params=golden_paths(), ids=lambda p: os.path.splitext(os.path.basename(p))[0])
def golden(request):
    return Golden(request.param)


@pytest.fixture(scope="session")
def gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
