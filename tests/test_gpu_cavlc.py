"""GPU parity of the device CAVLC (fh264_cavlc_p), through the C ABI: the slice data of every P
picture of the golden clips against the RBSP the unmodified reference wrote, and at 720p / batch 2 against the host build
of the same coder core fed with the device's own records (parallel assembly: skip runs, bit offsets, packing)."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import h264_fer_b200 as fh
from h264_fer_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bits(a):
    return np.unpackbits(np.asarray(a, np.uint8))


def test_slice_data_matches_reference_golden(golden):
    if not golden.p_pictures():
        pytest.skip("I-only fixture")
    checked = 0
    with fh.Session(golden.w, golden.h) as s:
        for n, t in enumerate(golden.types):
            if t == 5:
                s.upload_recon(0, *golden.rec(n))
                continue
            s.upload_source(0, *golden.src(n))
            s.encode_p(golden.qp, golden.window, golden.maxdiff, golden.basic)
            rbsp, bit0 = golden.slice_rbsp(n)
            data, nbits = s.cavlc_p(first_bit=bit0 % 8)[0]
            nd = nbits - bit0 % 8
            ref = _bits(rbsp)
            mine = _bits(data)
            assert not mine[:bit0 % 8].any() and not mine[nbits:].any(), "bits outside the slice data must be zero"
            assert bit0 + nd + 1 <= len(ref)
            assert np.array_equal(mine[bit0 % 8:nbits], ref[bit0:bit0 + nd]), "%s picture %d: slice data differs" % (golden.name, n)
            tail = ref[bit0 + nd:]
            assert tail[0] == 1 and not tail[1:].any() and len(tail) <= 8        # rbsp_trailing_bits follow in the reference
            checked += 1
    assert checked == len(golden.p_pictures())


@pytest.fixture(scope="module")
def host_core():
    out = os.path.join(tempfile.mkdtemp(prefix="fh264_cavlc_"), "libcavlc_host.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "cavlc_host.cpp")], check=True)
    return C.CDLL(out)


def test_720p_batch_against_host_core(host_core):
    w, h, nseq = 1280, 720, 2
    clips = [synth.SynthClip(w, h, 40 + b) for b in range(nseq)]
    fr = [[tuple(synth.crop16(p, chroma=(i > 0)) for i, p in enumerate(c.frame(t))) for t in range(3)] for c in clips]
    with fh.Session(w, h, batch=nseq) as s:
        for b in range(nseq):
            s.upload_recon(b, *fr[b][0])
        for t in (1, 2):
            for b in range(nseq):
                s.upload_source(b, *fr[b][t])
            rec = s.encode_p(28, 32, 3, 0)
            for first_bit in (0, 5):
                got = s.cavlc_p(first_bit=first_bit)
                for b in range(nseq):
                    r = np.ascontiguousarray(rec[b])
                    out = np.zeros(600000, np.uint8)
                    nbits, bad = C.c_int(0), C.c_int(0)
                    rc = host_core.cavlc_host_slice(r.ctypes.data_as(C.c_void_p), len(r), w // 16, first_bit, out.ctypes.data_as(C.c_void_p), len(out),
                                                    C.byref(nbits), C.byref(bad))
                    assert rc == 0 and bad.value == 0
                    data, nb = got[b]
                    assert nb == nbits.value, (nb, nbits.value)
                    assert np.array_equal(data, out[:(nb + 7) // 8]), "sequence %d picture %d first_bit %d" % (b, t, first_bit)
                    assert nb > 10000            # a real slice, not an all-skip picture


def test_all_skip_picture_is_one_skip_run():
    """A picture identical to its reference: every macroblock is P_Skip, slice_data is the single mb_skip_run ue(v)."""
    w, h = 176, 144
    c = synth.SynthClip(w, h, 5)
    y, u, v = (synth.crop16(p, chroma=(i > 0)) for i, p in enumerate(c.frame(0)))
    with fh.Session(w, h) as s:
        s.upload_recon(0, y, u, v)
        s.upload_source(0, y, u, v)
        rec = s.encode_p(28, 16, 3, 0)[0]
        assert (rec["mb_type"] == 31).all()
        data, nbits = s.cavlc_p()[0]
        # ue(99): 99 + 1 = 0b1100100 -> 6 zeros, then 1100100
        assert nbits == 13 and np.array_equal(_bits(data)[:13], np.array([0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 1, 0, 0], np.uint8))


def test_cavlc_error_paths():
    with fh.Session(176, 144) as s:
        with pytest.raises(fh.Fh264Error) as e:
            s.cavlc_p()
        assert e.value.code == -4                     # FH264_E_STATE: nothing coded yet
        with pytest.raises(fh.Fh264Error) as e:
            s.cavlc_p(first_bit=8)
        assert e.value.code == -1
