"""GPU parity of the decoder inverse path (fh264_decode_p): from the per-macroblock records the
reference encoder produced (golden vectors) it must rebuild exactly the reconstruction the reference holds — the picture any
conforming decoder outputs — and, chained over pictures and sequences, exactly what fh264_encode_p reconstructed."""
import numpy as np
import pytest

import h264_fer_b200 as fh
from h264_fer_b200 import synth

pytestmark = pytest.mark.gpu


def test_decode_p_rebuilds_the_reference_reconstruction(golden):
    if not golden.p_pictures():
        pytest.skip("I-only fixture")
    with fh.Session(golden.w, golden.h) as s:
        for n, t in enumerate(golden.types):
            if t == 5:
                s.upload_recon(0, *golden.rec(n))          # I pictures: host path
                continue
            s.decode_p(golden.records(n), golden.qp)
            for got, want in zip(s.download_recon(0), golden.rec(n)):
                assert np.array_equal(got, want), "%s picture %d" % (golden.name, n)


def test_decode_follows_encode_on_a_batch():
    w, h, nseq, qp = 640, 480, 2, 26
    clips = [synth.SynthClip(w, h, 60 + b) for b in range(nseq)]
    fr = [[c.frame(t) for t in range(4)] for c in clips]
    with fh.Session(w, h, batch=nseq) as enc, fh.Session(w, h, batch=nseq) as dec:
        for b in range(nseq):
            enc.upload_recon(b, *fr[b][0])
            dec.upload_recon(b, *fr[b][0])
        for t in (1, 2, 3):
            for b in range(nseq):
                enc.upload_source(b, *fr[b][t])
            rec = enc.encode_p(qp, 32, -1, 0)
            dec.decode_p(rec, qp)
            for b in range(nseq):
                for a, c in zip(enc.download_recon(b), dec.download_recon(b)):
                    assert np.array_equal(a, c), "sequence %d picture %d" % (b, t)


def test_decode_error_paths():
    with fh.Session(176, 144) as s:
        rec = np.zeros((1, 99), fh.MB_RESULT_DTYPE)
        with pytest.raises(fh.Fh264Error) as e:
            s.decode_p(rec, 28)
        assert e.value.code == -4                     # no reference picture yet
