"""GPU parity of the device I-picture path (fh264_encode_i) through the C ABI: per-macroblock records
(final mb_type, prediction modes, both bit-cost trials, CodedBlockPattern, levels) and the reconstruction against the compiled
reference — the committed fixtures tests/golden/intra_*.npz and, at 1080p, a live run of oracle/_ref/ref_encoder. The P pictures
between the I pictures go through fh264_encode_p, so the chain I -> P -> I also proves that the device reconstruction of an I
picture is a correct reference picture and that the P_Skip state the next I picture's first trial reads is the session's own."""
import os
import tempfile
import time

import numpy as np
import pytest

import h264_fer_b200 as fh
from h264_fer_b200 import synth
from test_intra_host import INTRA_GOLDENS, check_slice_bits, compare_i_records, golden_pictures

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[32, 1], ids=["warp", "lane0"])
def lanes(request):
    old = os.environ.get("FH264_INTRA_LANES")
    os.environ["FH264_INTRA_LANES"] = str(request.param)
    yield request.param
    if old is None:
        os.environ.pop("FH264_INTRA_LANES", None)
    else:
        os.environ["FH264_INTRA_LANES"] = old


def run_clip(s, seq, pics, qp, window, maxdiff, what):
    for n, p in enumerate(pics):
        s.upload_source(seq, p["SRCY"], p["SRCU"], p["SRCV"])
        if "imbrec" in p:
            rec = s.encode_i(qp, seq0=seq, nseq=1)[0]
            compare_i_records(fh.i_records_to_ints(rec), p["imbrec"], "%s picture %d (I)" % (what, n))
            if "rbsp" in p:          # the I slice's slice_data() entropy-coded on the device against the reference's RBSP
                data, nbits = s.cavlc_i(first_bit=p["slice_bit0"] % 8, seq0=seq, nseq=1)[0]
                check_slice_bits(data, nbits, p["rbsp"], p["slice_bit0"], "%s picture %d (I slice data)" % (what, n))
        else:
            rec = s.encode_p(qp, window, maxdiff, 0, seq0=seq, nseq=1)[0]
            mine, ref = fh.records_to_ints(rec), p["mbrec"]
            assert np.array_equal(mine[:, :17], ref[:, :17]), "%s picture %d (P): types / MVs differ" % (what, n)
            assert np.array_equal(mine[:, 21:], ref[:, 21:]), "%s picture %d (P): levels differ" % (what, n)
        for got, t in zip(s.download_recon(seq), ("RECY", "RECU", "RECV")):
            assert np.array_equal(got, p[t]), "%s picture %d: %s differs" % (what, n, t)


@pytest.mark.parametrize("name", INTRA_GOLDENS)
def test_encode_i_matches_the_reference_on_the_golden_clips(name, lanes):
    params, pics = golden_pictures(name)
    h, w = pics[0]["SRCY"].shape
    with fh.Session(w, h) as s:
        run_clip(s, 0, pics, int(params[4]), int(params[5]), int(params[6]), name)


def test_batch_of_sequences_equals_single_sequences():
    """Three sequences in one launch (interleaved tickets) give what each gives alone."""
    w, h, qp = 320, 240, 27
    clips = [synth.SynthClip(w, h, 70 + b, contrast=(1.0, 0.1, 0.4)[b]) for b in range(3)]
    frames = [c.frame(0) for c in clips]
    single = []
    for b in range(3):
        with fh.Session(w, h) as s:
            s.upload_source(0, *frames[b])
            single.append((s.encode_i(qp)[0], s.download_recon(0)))
    with fh.Session(w, h, batch=3) as s:
        for b in range(3):
            s.upload_source(b, *frames[b])
        rec = s.encode_i(qp)
        for b in range(3):
            assert np.array_equal(fh.i_records_to_ints(rec[b]), fh.i_records_to_ints(single[b][0])), "sequence %d records" % b
            for a, c in zip(s.download_recon(b), single[b][1]):
                assert np.array_equal(a, c), "sequence %d reconstruction" % b
    kinds = set(np.unique(np.minimum(np.concatenate([r[0]["mb_type"] for r in single]), 1)))
    assert kinds == {0, 1}


def test_1080p_i_picture_matches_a_live_reference_run():
    from oracle import refdump
    if not refdump.have_ref_encoder():
        pytest.skip("oracle/_ref/ref_encoder not built")
    qp = 28
    y4m = os.path.join(tempfile.mkdtemp(prefix="fh264_intra_hd_"), "in.y4m")
    synth.write_y4m(y4m, 1920, 1080, 4, 1)
    summ, dump, _ = refdump.run_reference(y4m, 1, qp=qp, window=32, dumpmask=refdump.D_RECON | refdump.D_SOURCE | refdump.D_IMBREC)
    p = refdump.parse_dump(dump)[0]
    with fh.Session(p["w"], p["h"]) as s:
        s.upload_source(0, p["SRCY"], p["SRCU"], p["SRCV"])
        rec = s.encode_i(qp)[0]
        compare_i_records(fh.i_records_to_ints(rec), p["imbrec"], "1080p")
        for got, t in zip(s.download_recon(0), ("RECY", "RECU", "RECV")):
            assert np.array_equal(got, p[t]), "1080p: %s differs" % t
        # timing note for the log (not an assertion): the same picture again, wall clock around the synchronous call
        s.upload_source(0, p["SRCY"], p["SRCU"], p["SRCV"])
        s.sync()
        t0 = time.perf_counter()
        s.encode_i(qp)
        dt = time.perf_counter() - t0
    print("\n1080p I picture: device %.1f ms (call incl. record D2H and phase R), reference %.0f ms on one host core" % (dt * 1e3, summ["t_picture"][0] * 1e3))


@pytest.mark.parametrize("w,h,seed", [(16, 16, 51), (176, 16, 52), (16, 144, 53)], ids=["one_mb", "one_row", "one_column"])
def test_degenerate_picture_shapes_match_a_live_reference_run(w, h, seed):
    """One macroblock, one macroblock row, one macroblock column: the wavefront's neighbour waits at every picture edge."""
    from oracle import refdump
    if not refdump.have_ref_encoder():
        pytest.skip("oracle/_ref/ref_encoder not built")
    qp = 28
    y4m = os.path.join(tempfile.mkdtemp(prefix="fh264_intra_edge_"), "in.y4m")
    synth.write_y4m(y4m, w, h, seed, 2)
    _, dump, _ = refdump.run_reference(y4m, 2, qp=qp, intra_every=1, dumpmask=refdump.D_RECON | refdump.D_SOURCE | refdump.D_IMBREC | refdump.D_SLICE)
    with fh.Session(w, h) as s:
        run_clip(s, 0, refdump.parse_dump(dump), qp, 16, 3, "%dx%d" % (w, h))


def test_cavlc_i_of_a_batch_and_its_error_path():
    w, h, qp = 320, 240, 27
    frames = [synth.SynthClip(w, h, 70 + b, contrast=(1.0, 0.1)[b]).frame(0) for b in range(2)]
    single = []
    for b in range(2):
        with fh.Session(w, h) as s:
            s.upload_source(0, *frames[b])
            s.encode_i(qp)
            single.append(s.cavlc_i(first_bit=3)[0])
    with fh.Session(w, h, batch=2) as s:
        with pytest.raises(fh.Fh264Error) as e:
            s.cavlc_i()
        assert e.value.code == -4                     # no I picture coded yet
        for b in range(2):
            s.upload_source(b, *frames[b])
        s.encode_i(qp)
        for b, (data, nbits) in enumerate(s.cavlc_i(first_bit=3)):
            assert nbits == single[b][1] and np.array_equal(data, single[b][0]), "sequence %d" % b
            assert nbits > 3 and not (data[0] & 0xE0)          # the first three bits stay clear for the slice header's tail


def test_encode_i_error_paths():
    with fh.Session(176, 144) as s:
        with pytest.raises(fh.Fh264Error) as e:
            s.encode_i(52)
        assert e.value.code == -1
        with pytest.raises(fh.Fh264Error) as e:
            s.encode_i(28, seq0=1, nseq=1)
        assert e.value.code == -1
