"""GPU suite, BASELINE.json's configurations at their stated LENGTH (long chains are where bit-exactness drifts if it can):
byte-identical .264 through the drop-in binary vs a live run of the unmodified reference encoder.
  * CIF 352x288, 300 frames, WindowSize 16 and 32, MAXDIFF 3 and adaptive (config 2)
  * QCIF, 520 frames with one IDR: crosses the reference's 9-bit frame_num wrap (headers_and_parameter_sets.cpp:195,319)
  * 720p, 16 pictures, WindowSize 32 (config 3; the reference needs ~10 s per picture)
  * 1080p, one P picture vs the oracle at QP 24, 26, 30 (config 5's QP sweep; QP 28 is in test_gpu_parity.py)
The reference runs take minutes of host time, so these tests only run with FH264_LONG=1 (the builder runs them through gpurun
and commits the log under profiles/); the reference encoders of all cases run concurrently first."""
import os
import subprocess

import numpy as np
import pytest

import h264_fer_b200 as fh
from h264_fer_b200 import synth
from oracle import port, refdump

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("FH264_LONG") != "1", reason="long runs: set FH264_LONG=1")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENCODER = os.path.join(ROOT, "integration", "_build", "fh264_encoder_b200")
ENCODER_ALL = ENCODER + "_all"

# name: (w, h, seed, frames, qp, window, maxdiff, intra_every)
CASES = {
    "cif300_w16_md3": (352, 288, 2, 300, 28, 16, 3, 1000),
    "cif300_w32_md3": (352, 288, 2, 300, 28, 32, 3, 1000),
    "cif300_w16_adaptive": (352, 288, 2, 300, 28, 16, -1, 1000),
    "cif300_w32_adaptive": (352, 288, 2, 300, 28, 32, -1, 1000),
    "qcif520_frame_num_wrap": (176, 144, 1, 520, 28, 16, 3, 1000),
    "720p16_w32": (1280, 720, 3, 16, 28, 32, 3, 1000),
}


@pytest.fixture(scope="module")
def reference_runs(tmp_path_factory):
    """All reference encoders at once (one host core each); returns {name: (y4m, ref264)}."""
    if not refdump.have_ref_encoder() or not os.path.isfile(ENCODER):
        pytest.skip("compiled reference / integration binary not present")
    d = tmp_path_factory.mktemp("long")
    procs, out = {}, {}
    for name, (w, h, seed, frames, qp, window, maxdiff, ie) in CASES.items():
        y4m, ref264 = str(d / (name + ".y4m")), str(d / (name + ".ref.264"))
        synth.write_y4m(y4m, w, h, seed, frames)
        cmd = [refdump.REF_ENCODER, y4m, ref264, "-", str(frames), str(qp), "0", str(window), str(maxdiff), str(ie), "0"]
        procs[name] = subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=str(d))
        out[name] = (y4m, ref264)
    for name, p in procs.items():
        assert p.wait(timeout=3000) == 0, name
    return out


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("binary", ["host_syntax", "all_on_device"])
def test_long_sequence_bitstream_is_byte_identical(reference_runs, tmp_path, name, binary):
    enc = ENCODER if binary == "host_syntax" else ENCODER_ALL
    if not os.path.isfile(enc):
        pytest.skip("binary not built")
    w, h, seed, frames, qp, window, maxdiff, ie = CASES[name]
    y4m, ref264 = reference_runs[name]
    out = str(tmp_path / "b200.264")
    cmd = [enc, y4m, out, "-", str(frames), str(qp), "0", str(window), str(maxdiff), str(ie), "0", "-1"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=3000)
    assert res.returncode == 0, res.stderr.decode()[-2000:]
    a, b = open(out, "rb").read(), open(ref264, "rb").read()
    assert len(a) == len(b) and a == b, "%s: bitstreams differ (%d vs %d bytes)" % (name, len(a), len(b))
    print("%s/%s: %d pictures, %d bytes, byte-identical" % (name, binary, frames, len(a)))


@pytest.mark.parametrize("qp", [24, 26, 30])
def test_1080p_picture_against_oracle_over_the_qp_sweep(qp):
    clip = synth.SynthClip(1920, 1080, 101)
    fr = [tuple(synth.crop16(p, chroma=(i > 0)) for i, p in enumerate(clip.frame(t))) for t in range(2)]
    h, w = fr[0][0].shape
    o = port.Oracle(w, h)
    assert not o.phase_r(fr[0][0])
    want, want_recon = o.encode_p(fr[1], fr[0], qp, 32, 3)
    with fh.Session(w, h) as s:
        s.upload_recon(0, *fr[0])
        s.upload_source(0, *fr[1])
        got = fh.records_to_ints(s.encode_p(qp, 32, 3)[0])
        recon = s.download_recon(0)
    assert np.array_equal(got, want), np.argwhere(got != want)[:6]
    assert all(np.array_equal(a, b) for a, b in zip(recon, want_recon))


@pytest.mark.parametrize("piped", [0, 1])
def test_streaming_step_soak_against_the_plain_call_sequence(piped):
    """The asynchronous step (fh264_upload_source_batch + fh264_encode_p_stream, nothing synchronised between pictures) over a long
    chain with periodic scene cuts, against the host-driven sequence fh264_scene_sad + fh264_encode_p / fh264_encode_i picture by
    picture: any ordering bug between the coding stream, the copy stream and the upload stream shows as a drifting reconstruction."""
    from h264_fer_b200 import native
    w, h, nseq, npic, qp, window, maxdiff = 352, 288, 4, 48, 28, 32, 3
    clips = [[synth.SynthClip(w, h, 300 + b + 17 * (t // 16 if b == 1 else 0)).frame(t) for t in range(npic)] for b in range(nseq)]   # sequence 1 cuts every 16 pictures
    pic = w * h * 3 // 2
    want_rec, want_recon, want_idr = [], None, 0
    with fh.Session(w, h, batch=nseq) as s:
        for b in range(nseq):
            s.upload_recon(b, *clips[b][0])
        for t in range(1, npic):
            for b in range(nseq):
                s.upload_source(b, *clips[b][t])
            sads = s.scene_sad_batch()
            row = []
            for b in range(nseq):
                if sads[b] > (s.nmb << 12):
                    s.encode_i(qp, seq0=b, nseq=1); row.append(None); want_idr += 1
                else:
                    row.append(s.encode_p(qp, window, maxdiff, seq0=b, nseq=1)[0].copy())
            want_rec.append(row)
        want_recon = [s.download_recon(b) for b in range(nseq)]
    assert want_idr == 2
    with fh.Session(w, h, batch=nseq) as s:
        s.set_pipeline(piped)
        for b in range(nseq):
            s.upload_recon(b, *clips[b][0])
        blocks, outs = [], []
        for t in range(1, npic):
            blk = native.PinnedArray((nseq, pic), np.uint8)
            for b in range(nseq):
                blk.array[b] = np.concatenate([p.ravel() for p in clips[b][t]])
            so = native.StreamOut(nseq, s.nmb, records=True)
            blocks.append(blk); outs.append(so)
            s.upload_source_batch(blk.ptr, pic)
            s.encode_p_stream(qp, window, maxdiff, scene_gate=1, out=so)
            if any(r is None for r in want_rec[t - 1]):          # the host of a real encoder learns about a cut by looking: sync, then the IDR picture
                s.sync()
                coded = so.coded()
                assert coded == [r is not None for r in want_rec[t - 1]], "picture %d" % t
                for b in range(nseq):
                    if not coded[b]:
                        s.encode_i(qp, seq0=b, nseq=1)
        s.sync()
        for t, so in enumerate(outs):
            for b in range(nseq):
                if want_rec[t][b] is not None:
                    assert np.array_equal(so.records.array[b], want_rec[t][b]), "picture %d sequence %d" % (t + 1, b)
        for b in range(nseq):
            s.picture_status(b)
            assert all(np.array_equal(a, c) for a, c in zip(s.download_recon(b), want_recon[b])), "final reconstruction of sequence %d" % b
