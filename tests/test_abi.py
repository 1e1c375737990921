"""CPU suite: the C-ABI library builds, loads and exports every symbol include/fh264_b200.h declares; the record layout
matches; and without a GPU the product path fails loudly (no CPU fallback, no route into oracle/)."""
import ctypes
import os
import re

import numpy as np
import pytest

import h264_fer_b200 as fh
from h264_fer_b200 import build, native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fh264_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fh264_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.isfile(path)
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    assert sorted(native.EXPORTS) == syms
    assert lib.fh264_abi_version() == 1


def test_record_layout_matches_header():
    assert native.MB_RESULT_DTYPE.itemsize == 832
    f = native.MB_RESULT_DTYPE.fields
    assert f["mv"][1] == 4 and f["mvd"][1] == 20 and f["sad"][1] == 36 and f["luma"][1] == 44
    assert f["chroma_dc"][1] == 556 and f["chroma_ac"][1] == 572 and f["reserved"][1] == 812
    r = np.zeros(2, native.MB_RESULT_DTYPE)
    r["mb_type"] = [4, 31]; r["luma"][0, 3, 5] = -7; r["chroma_ac"][1, 1, 2, 14] = 9; r["sad"][0] = [1, 2, 3, 4]
    ints = fh.records_to_ints(r)
    assert ints.shape == (2, 405) and ints[0, 0] == 4 and ints[1, 0] == 31
    assert ints[0, 21 + 3 * 16 + 5] == -7 and ints[1, 285 + 60 + 2 * 15 + 14] == 9 and ints[0, 17:21].tolist() == [1, 2, 3, 4]


def test_argument_validation_needs_no_device():
    lib = fh.load_library()
    h = ctypes.c_void_p()
    assert lib.fh264_open(100, 144, 1, 0, ctypes.byref(h)) == -1          # not a multiple of 16
    assert b"multiples of 16" in lib.fh264_last_error()
    assert lib.fh264_open(176 * 16, 144 * 16, 1, 0, ctypes.byref(h)) == -1  # > 10000 macroblocks
    assert lib.fh264_open(176, 144, 0, 0, ctypes.byref(h)) == -1
    assert lib.fh264_close(None) == 0


def test_no_cpu_fallback_without_gpu(gpu_available):
    if gpu_available:
        pytest.skip("a GPU is present")
    with pytest.raises(fh.Fh264Error) as e:
        fh.Session(176, 144)
    assert e.value.code == -3


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "h264_fer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("oracle/", "").lower() or f == "synth.py" or "import oracle" not in text
                assert "from oracle" not in text and "import oracle" not in text and "fh264_oracle" not in text, f
