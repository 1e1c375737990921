"""GPU suite (-m gpu): the streaming step (fh264_encode_p_stream / fh264_upload_source_batch, include/fh264_b200.h) against the
plain call sequence it replaces — fh264_scene_sad + fh264_encode_p + fh264_cavlc_p, the host-driven mirror of selectNALUnitType
(ref_frames.cpp:185-234) and RBSP_encode's P branch (rbsp_encoding.cpp:139-323), which the other GPU tests pin against the oracle
and the reference's golden vectors. Bit-exact: records, slice data, side information, reconstructions, scene SADs, and the
IDR decision taken on the device."""
import numpy as np
import pytest

import h264_fer_b200 as fh
from h264_fer_b200 import native, synth
from oracle import port

pytestmark = pytest.mark.gpu

W, H, QP, WINDOW, MAXDIFF = 352, 288, 28, 32, 3


def _clips(nseq, npic, seed0=40):
    return [[synth.SynthClip(W, H, seed0 + b).frame(t) for t in range(npic)] for b in range(nseq)]


def _block(pics):
    """One pinned block holding Y | Cb | Cr of every sequence back to back (what fh264_upload_source_batch takes)."""
    n = len(pics)
    pic = W * H * 3 // 2
    blk = native.PinnedArray((n, pic), np.uint8)
    for b, (y, cb, cr) in enumerate(pics):
        blk.array[b, :W * H] = y.ravel()
        blk.array[b, W * H:W * H + W * H // 4] = cb.ravel()
        blk.array[b, W * H + W * H // 4:] = cr.ravel()
    return blk, pic


def _plain_run(clips, npic, idr_at=None):
    """The host-driven call sequence: per picture scene SAD -> IDR decision on the host -> encode_p + cavlc_p, or encode_i."""
    nseq = len(clips)
    out = []
    with fh.Session(W, H, batch=nseq) as s:
        for b in range(nseq):
            s.upload_recon(b, *clips[b][0])
        for t in range(1, npic):
            for b in range(nseq):
                s.upload_source(b, *clips[b][t])
            sads = s.scene_sad_batch()
            step = {"sad": sads, "rec": [None] * nseq, "slice": [None] * nseq, "info": [None] * nseq, "irec": [None] * nseq, "recon": []}
            for b in range(nseq):
                if sads[b] > (s.nmb << 12):
                    step["irec"][b] = s.encode_i(QP, seq0=b, nseq=1)[0].copy()
                else:
                    step["rec"][b] = s.encode_p(QP, WINDOW, MAXDIFF, seq0=b, nseq=1)[0].copy()
                    (sl,), info = s.cavlc_p(first_bit=3, seq0=b, nseq=1, mb_info=True)
                    step["slice"][b] = sl
                    step["info"][b] = info[0].copy()
            step["recon"] = [s.download_recon(b) for b in range(nseq)]
            out.append(step)
    return out


def _assert_step(got_out, want, b, what):
    assert np.array_equal(got_out.records.array[b], want["rec"][b]), "%s: records of sequence %d" % (what, b)
    data, nbits = got_out.slices()[b]
    wdata, wbits = want["slice"][b]
    assert nbits == wbits and np.array_equal(data, wdata), "%s: slice data of sequence %d" % (what, b)
    assert int(got_out.slice_stat.array[b, 0]) == 0
    assert np.array_equal(got_out.mb_info.array[b], want["info"][b]), "%s: side information of sequence %d" % (what, b)


@pytest.mark.parametrize("piped", [0, 1])
@pytest.mark.parametrize("sync_every_step", [True, False])
def test_stream_steps_equal_the_plain_call_sequence(sync_every_step, piped):
    nseq, npic = 5, 5                                        # odd batch: the two pipeline halves differ in size
    clips = _clips(nseq, npic)
    want = _plain_run(clips, npic)
    with fh.Session(W, H, batch=nseq) as s:
        s.set_pipeline(piped)
        for b in range(nseq):
            s.upload_recon(b, *clips[b][0])
        outs, blocks = [], []
        for t in range(1, npic):
            blk, pic = _block([clips[b][t] for b in range(nseq)])
            blocks.append(blk)
            o = native.StreamOut(nseq, s.nmb, records=True, slice_bytes=65536, mb_info=True, first_bit=3)
            outs.append(o)
            s.upload_source_batch(blk.ptr, pic)
            s.encode_p_stream(QP, WINDOW, MAXDIFF, scene_gate=True, out=o)
            if sync_every_step:
                s.sync()
        s.sync()
        for t, o in enumerate(outs):
            assert o.coded() == [True] * nseq
            assert o.scene_sad() == want[t]["sad"]
            for b in range(nseq):
                _assert_step(o, want[t], b, "picture %d" % (t + 1))
        for b in range(nseq):
            s.picture_status(b)
            for a, w_ in zip(s.download_recon(b), want[-1]["recon"][b]):
                assert np.array_equal(a, w_)


def test_stream_step_against_the_oracle():
    clips = _clips(2, 2, seed0=77)
    with fh.Session(W, H, batch=2) as s:
        for b in range(2):
            s.upload_recon(b, *clips[b][0])
        blk, pic = _block([clips[b][1] for b in range(2)])
        o = native.StreamOut(2, s.nmb, records=True)
        s.upload_source_batch(blk.ptr, pic)
        s.encode_p_stream(QP, WINDOW, MAXDIFF, scene_gate=True, out=o)
        s.sync()
        for b in range(2):
            orc = port.Oracle(W, H)
            assert not orc.phase_r(clips[b][0][0])
            want_rec, want_recon = orc.encode_p(clips[b][1], clips[b][0], QP, WINDOW, MAXDIFF)
            assert np.array_equal(fh.records_to_ints(o.records.array[b]), want_rec)
            for a, w_ in zip(s.download_recon(b), want_recon):
                assert np.array_equal(a, w_)
            assert o.scene_sad()[b] == port.scene_sad(clips[b][1][0], clips[b][0][0])


@pytest.mark.parametrize("piped", [0, 1])
def test_scene_gate_stops_a_cut_and_the_sequence_goes_on_with_an_idr_picture(piped):
    nseq, npic = 3, 5
    clips = _clips(nseq, npic)
    other = synth.SynthClip(W, H, 999, pan=(0, 0), contrast=1.0)
    for t in range(2, npic):                                 # sequence 1 cuts to different content at picture 2
        clips[1][t] = other.frame(t)
    want = _plain_run(clips, npic)
    assert want[1]["irec"][1] is not None and want[1]["rec"][0] is not None, "the fixture must contain exactly the cut it is about"
    with fh.Session(W, H, batch=nseq) as s:
        s.set_pipeline(piped)
        for b in range(nseq):
            s.upload_recon(b, *clips[b][0])
        for t in range(1, npic):
            blk, pic = _block([clips[b][t] for b in range(nseq)])
            o = native.StreamOut(nseq, s.nmb, records=True, slice_bytes=65536, mb_info=True, first_bit=3)
            s.upload_source_batch(blk.ptr, pic)
            s.encode_p_stream(QP, WINDOW, MAXDIFF, scene_gate=True, out=o)
            s.sync()
            coded = o.coded()
            assert coded == [want[t - 1]["rec"][b] is not None for b in range(nseq)], "picture %d: IDR decisions" % t
            assert o.scene_sad() == want[t - 1]["sad"]
            for b in range(nseq):
                if coded[b]:
                    _assert_step(o, want[t - 1], b, "picture %d" % t)
                else:
                    assert int(o.slice_stat.array[b, 1]) == 3          # no slice data: only the caller's first_bit offset
                    irec = s.encode_i(QP, seq0=b, nseq=1)[0]           # the source picture is still current
                    assert np.array_equal(irec, want[t - 1]["irec"][b]), "picture %d: I records of sequence %d" % (t, b)
            for b in range(nseq):
                for a, w_ in zip(s.download_recon(b), want[t - 1]["recon"][b]):
                    assert np.array_equal(a, w_), "picture %d: reconstruction of sequence %d" % (t, b)
        assert int(o.status.array[1, native.ST_GATED_TOTAL]) == 1 and int(o.status.array[0, native.ST_GATED_TOTAL]) == 0


def test_pipelined_encode_p_async_equals_the_plain_schedule():
    nseq, npic = 4, 4
    clips = _clips(nseq, npic, seed0=60)
    res = {}
    for piped in (0, 1):
        with fh.Session(W, H, batch=nseq) as s:
            s.set_pipeline(piped)
            for b in range(nseq):
                s.upload_recon(b, *clips[b][0])
            outs = []
            for t in range(1, npic):
                for b in range(nseq):
                    s.upload_source(b, *clips[b][t])
                outs.append(native.PinnedArray((nseq, s.nmb), fh.MB_RESULT_DTYPE))
                s.encode_p(QP, WINDOW, MAXDIFF, out=outs[-1].array, sync=False)
            s.sync()
            for b in range(nseq):
                s.picture_status(b)
            res[piped] = ([o.array.copy() for o in outs], [s.download_recon(b) for b in range(nseq)])
    for a, b_ in zip(res[0][0], res[1][0]):
        assert np.array_equal(a, b_)
    for ra, rb in zip(res[0][1], res[1][1]):
        for a, b_ in zip(ra, rb):
            assert np.array_equal(a, b_)


def test_stream_error_paths():
    with fh.Session(W, H, batch=2) as s:
        with pytest.raises(fh.Fh264Error) as e:
            s.encode_p_stream(QP, WINDOW, MAXDIFF)                       # no reference picture yet
        assert e.value.code == -4
        y, cb, cr = synth.SynthClip(W, H, 1).frame(0)
        for b in range(2):
            s.upload_recon(b, y, cb, cr)
        o = native.StreamOut(2, s.nmb, slice_bytes=1024)
        o.struct.slice_stat = None
        with pytest.raises(fh.Fh264Error) as e:
            s.encode_p_stream(QP, WINDOW, MAXDIFF, out=o)
        assert e.value.code == -1
        with pytest.raises(fh.Fh264Error) as e:
            s.upload_source_batch(0, 10)
        assert e.value.code == -1


def test_batch_encoder_follows_the_reference_picture_type_rules():
    """BatchEncoder (host mirror of selectNALUnitType + RBSP_encode for a batch, h264_fer_b200/encoder.py) against one
    SequenceEncoder per sequence driven the reference's way: first picture and every IntraEvery-th picture IDR (ref_frames.cpp:191),
    scene cut IDR (:210-224), everything else P — same picture types and the same slice data, picture by picture."""
    nseq, npic, intra_every = 3, 7, 5
    clips = _clips(nseq, npic, seed0=80)
    other = synth.SynthClip(W, H, 555, pan=(0, 0))
    for t in range(3, npic):                                 # sequence 2 cuts at picture 3
        clips[2][t] = other.frame(t)
    want = []
    with fh.Session(W, H, batch=nseq) as s:
        encs = [fh.SequenceEncoder(s, b, qp=QP, window=WINDOW, maxdiff_set=MAXDIFF, intra_every=intra_every) for b in range(nseq)]
        for t in range(npic):
            row = []
            for b in range(nseq):
                nal, _ = encs[b].encode_picture(*clips[b][t])
                sl = s.cavlc_i(first_bit=5, seq0=b, nseq=1)[0] if nal == fh.NAL_IDR else s.cavlc_p(first_bit=5, seq0=b, nseq=1)[0]
                row.append((nal,) + sl)
            want.append(row)
    assert [r[2][0] for r in want] == [5, 1, 1, 5, 1, 5, 1], "sequence 2 must see the cut at picture 3 besides the periodic IDRs"
    with fh.Session(W, H, batch=nseq) as s:
        be = fh.BatchEncoder(s, qp=QP, window=WINDOW, maxdiff_set=MAXDIFF, intra_every=intra_every, first_bit=5)
        for t in range(npic):
            got = be.encode_pictures([clips[b][t] for b in range(nseq)])
            for b in range(nseq):
                assert got[b][0] == want[t][b][0], "picture %d sequence %d: nal_unit_type" % (t, b)
                assert got[b][2] == want[t][b][2] and np.array_equal(got[b][1], want[t][b][1]), "picture %d sequence %d: slice data" % (t, b)
