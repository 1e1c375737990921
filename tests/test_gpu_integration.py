"""GPU suite: the drop-in claim end to end. integration/_build/fh264_encoder_b200 is the UNMODIFIED reference host code
(CAVLC, NAL, headers, intra pictures, Y4M reader) linked against libfh264_b200.so through integration/fh264_ref_shim.cpp.
Its Annex-B bitstream must be byte-identical to the reference encoder's, and its reconstruction too."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from h264_fer_b200 import synth
from oracle import refdump

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENCODER = os.path.join(ROOT, "integration", "_build", "fh264_encoder_b200")
ENCODER_CAVLC = ENCODER + "_cavlc"          # same, with the P-slice entropy coding on the device too
ENCODER_INTRA = ENCODER + "_intra"          # same as the first, with the I pictures coded on the device too
ENCODER_ALL = ENCODER + "_all"              # both picture types coded AND entropy-coded on the device (fh264_cavlc_p + fh264_cavlc_i)

needs_binary = pytest.mark.skipif(not (os.path.isfile(ENCODER) and os.path.isfile(ENCODER_CAVLC)), reason="integration binaries not built (make -C integration)")
both_encoders = pytest.mark.parametrize("encoder", [ENCODER, ENCODER_CAVLC], ids=["host_cavlc", "device_cavlc"])


def run_b200_encoder(y4m, out264, dump, frames, qp, basic, window, maxdiff, intra_every=1000, dumpmask=0, encoder=ENCODER):
    cmd = [encoder, y4m, out264, dump if dumpmask else "-", str(frames), str(qp), str(basic), str(window), str(maxdiff), str(intra_every),
           str(dumpmask), "-1"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert res.returncode == 0, res.stderr.decode()[-2000:]
    return res.stdout.decode()


def md5(path):
    return hashlib.md5(open(path, "rb").read()).digest()


@needs_binary
@both_encoders
def test_bitstream_and_reconstruction_match_golden(golden, tmp_path, encoder):
    y4m = str(tmp_path / "in.y4m")
    z = golden.z
    synth.write_y4m(y4m, golden.w_in, golden.h_in, golden.seed, golden.frames, noise=float(z["noise"][0]), square=bool(int(z["square"][0])),
                    contrast=float(z["contrast"][0]))
    assert md5(y4m) == bytes(z["y4m_md5"]), "synthetic clip generator drifted from the one that made the golden vectors"
    out, dump = str(tmp_path / "out.264"), str(tmp_path / "dump.bin")
    run_b200_encoder(y4m, out, dump, golden.frames, golden.qp, golden.basic, golden.window, golden.maxdiff, dumpmask=refdump.D_RECON, encoder=encoder)
    assert md5(out) == bytes(z["bitstream_md5"]), "%s: .264 differs from the reference encoder's" % golden.name
    pics = refdump.parse_dump(dump)
    assert [p["nal_type"] for p in pics] == golden.types
    for n, p in enumerate(pics):
        ey, eu, ev = golden.rec(n)
        assert np.array_equal(p["RECY"], ey) and np.array_equal(p["RECU"], eu) and np.array_equal(p["RECV"], ev), "picture %d" % n


@needs_binary
@pytest.mark.skipif(not refdump.have_ref_encoder(), reason="compiled reference not present")
@pytest.mark.parametrize("w,h,seed,frames,qp,window,maxdiff,intra_every", [(176, 144, 1, 30, 28, 16, 3, 1000),     # BASELINE config 1 (QCIF, 30 frames)
                                                                            (352, 288, 2, 12, 28, 32, -1, 5),      # CIF, adaptive MAXDIFF, periodic IDR
                                                                            (176, 144, 7, 5, 12, 16, 3, 1000)])    # QP 12: large levels (CAVLC escape codes, suffixLength up to 6)
@both_encoders
def test_bitstream_matches_live_reference(tmp_path, w, h, seed, frames, qp, window, maxdiff, intra_every, encoder):
    y4m = str(tmp_path / "in.y4m")
    synth.write_y4m(y4m, w, h, seed, frames)
    ref264 = str(tmp_path / "ref.264")
    summ, _, _ = refdump.run_reference(y4m, frames, qp=qp, window=window, maxdiff=maxdiff, intra_every=intra_every, out_264=ref264)
    out = str(tmp_path / "b200.264")
    run_b200_encoder(y4m, out, "-", frames, qp, 0, window, maxdiff, intra_every=intra_every, encoder=encoder)
    assert "P" in summ["types"]
    assert open(out, "rb").read() == open(ref264, "rb").read(), "bitstreams differ (%s)" % summ["types"]


@needs_binary
@pytest.mark.skipif(not refdump.have_ref_encoder(), reason="compiled reference not present")
@both_encoders
def test_720p_window32_bitstream_matches_live_reference(tmp_path, encoder):
    """BASELINE config 3 geometry (1280x720, +-16 search): 1 I + 2 P pictures, byte-identical bitstream."""
    y4m = str(tmp_path / "in.y4m")
    synth.write_y4m(y4m, 1280, 720, 3, 3)
    ref264 = str(tmp_path / "ref.264")
    summ, _, _ = refdump.run_reference(y4m, 3, qp=28, window=32, maxdiff=3, out_264=ref264)
    out = str(tmp_path / "b200.264")
    run_b200_encoder(y4m, out, "-", 3, 28, 0, 32, 3, encoder=encoder)
    assert summ["types"] == "IPP"
    assert open(out, "rb").read() == open(ref264, "rb").read()


needs_intra_binary = pytest.mark.skipif(not os.path.isfile(ENCODER_INTRA), reason="integration binary with device I pictures not built (make -C integration)")


@needs_intra_binary
@pytest.mark.skipif(not refdump.have_ref_encoder(), reason="compiled reference not present")
@pytest.mark.parametrize("w,h,seed,frames,qp,window,maxdiff,intra_every,kw", [
    (176, 144, 1, 30, 28, 16, 3, 1000, {}),                                   # BASELINE config 1: first picture + a scene-change IDR
    (352, 288, 2, 9, 28, 32, -1, 3, {}),                                      # CIF, periodic IDR every 3 pictures
    (176, 144, 6, 6, 30, 16, 3, 2, {"contrast": 0.05, "noise": 0.3}),         # low contrast: Intra16x16 macroblocks, flat P pictures
    (176, 144, 31, 5, 12, 16, 3, 2, {"pan": (0, 0), "noise": 0.0, "square": False}),   # static: all-P_Skip pictures before the IDRs
    (200, 120, 9, 4, 44, 16, 3, 2, {}),                                       # cropped input, coarse quantiser
])
def test_bitstream_with_device_i_pictures_matches_live_reference(tmp_path, w, h, seed, frames, qp, window, maxdiff, intra_every, kw):
    """The unmodified reference host code with BOTH picture types coded on the device (fh264_encode_i + fh264_encode_p): the
    reference's own CAVLC writes what the device decided, and the Annex-B stream must be byte-identical to the reference's."""
    y4m = str(tmp_path / "in.y4m")
    synth.write_y4m(y4m, w, h, seed, frames, **kw)
    ref264, refdumpf = str(tmp_path / "ref.264"), str(tmp_path / "ref.bin")
    summ, rd, _ = refdump.run_reference(y4m, frames, qp=qp, window=window, maxdiff=maxdiff, intra_every=intra_every, out_264=ref264,
                                        dumpmask=refdump.D_RECON, dump_path=refdumpf)
    out, dump = str(tmp_path / "b200.264"), str(tmp_path / "b200.bin")
    run_b200_encoder(y4m, out, dump, frames, qp, 0, window, maxdiff, intra_every=intra_every, dumpmask=refdump.D_RECON, encoder=ENCODER_INTRA)
    assert summ["types"].count("I") >= 1 and "P" in summ["types"]
    for n, (a, b) in enumerate(zip(refdump.parse_dump(dump), refdump.parse_dump(rd))):
        assert a["nal_type"] == b["nal_type"], "picture %d type" % n
        assert np.array_equal(a["RECY"], b["RECY"]) and np.array_equal(a["RECU"], b["RECU"]) and np.array_equal(a["RECV"], b["RECV"]), "picture %d reconstruction (%s)" % (n, summ["types"])
    assert open(out, "rb").read() == open(ref264, "rb").read(), "bitstreams differ (%s)" % summ["types"]


@pytest.mark.skipif(not os.path.isfile(ENCODER_ALL), reason="integration binary with everything on the device not built (make -C integration)")
@pytest.mark.skipif(not refdump.have_ref_encoder(), reason="compiled reference not present")
@pytest.mark.parametrize("w,h,seed,frames,qp,window,maxdiff,intra_every,kw", [
    (176, 144, 1, 30, 28, 16, 3, 1000, {}),                                   # BASELINE config 1: first picture + a scene-change IDR
    (352, 288, 2, 7, 28, 32, -1, 3, {}),                                      # CIF, adaptive MAXDIFF, periodic IDR every 3 pictures
])
def test_bitstream_with_everything_on_the_device_matches_live_reference(tmp_path, w, h, seed, frames, qp, window, maxdiff, intra_every, kw):
    """fh264_encoder_b200_all: slice_data() of BOTH picture types comes from the device (fh264_encode_i + fh264_cavlc_i,
    fh264_encode_p + fh264_cavlc_p); the reference host code only writes parameter sets, slice headers and NAL framing."""
    y4m = str(tmp_path / "in.y4m")
    synth.write_y4m(y4m, w, h, seed, frames, **kw)
    ref264 = str(tmp_path / "ref.264")
    summ, rd, _ = refdump.run_reference(y4m, frames, qp=qp, window=window, maxdiff=maxdiff, intra_every=intra_every, out_264=ref264,
                                        dumpmask=refdump.D_RECON, dump_path=str(tmp_path / "ref.bin"))
    out, dump = str(tmp_path / "b200.264"), str(tmp_path / "b200.bin")
    run_b200_encoder(y4m, out, dump, frames, qp, 0, window, maxdiff, intra_every=intra_every, dumpmask=refdump.D_RECON, encoder=ENCODER_ALL)
    assert summ["types"].count("I") >= 1 and "P" in summ["types"]
    for n, (a, b) in enumerate(zip(refdump.parse_dump(dump), refdump.parse_dump(rd))):
        assert a["nal_type"] == b["nal_type"], "picture %d type" % n
        assert np.array_equal(a["RECY"], b["RECY"]) and np.array_equal(a["RECU"], b["RECU"]) and np.array_equal(a["RECV"], b["RECV"]), "picture %d reconstruction (%s)" % (n, summ["types"])
    assert open(out, "rb").read() == open(ref264, "rb").read(), "bitstreams differ (%s)" % summ["types"]
