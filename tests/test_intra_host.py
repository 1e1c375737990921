"""CPU checks of the device I-picture core (h264_fer_b200/csrc/intra_core.h, compiled for the host by g++) against I pictures of
the compiled reference: the committed fixtures tests/golden/intra_*.npz (made by tests/golden/make_golden.py from
oracle/_ref/ref_encoder, chunk IMBR) and, when the reference binary is present, live runs. Every field the reference's
intraPredictionEncoding + quantizationTransform leave behind is compared bit for bit: final mb_type, prediction modes, the
bits of both coded_mb_size() trials, CodedBlockPattern, all levels and the reconstruction. SURVEY.md §8(f) rank 2."""
import ctypes as C
import glob
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
from h264_fer_b200 import native as fh  # noqa: E402

INTRA_GOLDENS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "intra_*.npz")))


@pytest.fixture(scope="module")
def host_lib():
    out = os.path.join(tempfile.mkdtemp(prefix="fh264_intra_"), "libintra_host.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "intra_host.cpp")], check=True)
    return C.CDLL(out)


def host_picture(lib, sy, su, sv, qp, prev_types):
    h, w = sy.shape
    out = np.zeros((w // 16) * (h // 16), fh.MB_RESULT_I_DTYPE)
    rec = [np.zeros_like(np.ascontiguousarray(p)) for p in (sy, su, sv)]
    pt = None if prev_types is None else np.ascontiguousarray(prev_types, np.int32)
    vp = C.c_void_p
    lib.intra_host_picture(*[np.ascontiguousarray(p).ctypes.data_as(vp) for p in (sy, su, sv)], *[r.ctypes.data_as(vp) for r in rec], w, h, qp,
                           None if pt is None else pt.ctypes.data_as(vp), out.ctypes.data_as(vp))
    return out, rec


def compare_i_records(mine, ref, what):
    """mine: [nmb, 439] from i_records_to_ints; ref: the reference dump's IMBR records."""
    ref = np.asarray(ref, np.int32)
    names = [("mb_type", 0, 1), ("Intra16x16PredMode", 1, 2), ("intra_chroma_pred_mode", 2, 3), ("bits of the Intra16x16 trial", 3, 4),
             ("bits of the Intra4x4 trial", 4, 5), ("CodedBlockPatternLuma", 5, 6), ("CodedBlockPatternChroma", 6, 7), ("Intra4x4PredMode", 7, 23),
             ("prev_intra4x4_pred_mode_flag", 23, 39), ("luma levels", 55, 311), ("chroma DC levels", 311, 319), ("chroma AC levels", 319, 439)]
    for name, a, b in names:
        bad = np.nonzero((mine[:, a:b] != ref[:, a:b]).any(axis=1))[0]
        assert len(bad) == 0, "%s: %s differs in %d macroblocks, first %d: mine %s ref %s" % (what, name, len(bad), bad[0], mine[bad[0], a:b], ref[bad[0], a:b])
    coded = ref[:, 23:39] == 0          # rem_intra4x4_pred_mode is only meaningful where the flag is 0
    assert np.array_equal(mine[:, 39:55][coded], ref[:, 39:55][coded]), what + ": rem_intra4x4_pred_mode differs"


def check_clip(lib, pics, qp, what):
    """pics: list of dicts with SRCY/SRCU/SRCV, RECY/RECU/RECV and imbrec (I pictures) or mbrec (P pictures)."""
    prev, n_i = None, 0
    for n, p in enumerate(pics):
        if "imbrec" not in p:
            prev = np.asarray(p["mbrec"])[:, 0].astype(np.int32)
            continue
        out, rec = host_picture(lib, p["SRCY"], p["SRCU"], p["SRCV"], qp, prev)
        compare_i_records(fh.i_records_to_ints(out), p["imbrec"], "%s picture %d" % (what, n))
        for r, t in zip(rec, ("RECY", "RECU", "RECV")):
            assert np.array_equal(r, p[t]), "%s picture %d: %s differs" % (what, n, t)
        prev = np.asarray(p["imbrec"])[:, 0].astype(np.int32)
        n_i += 1
    assert n_i > 0


def golden_pictures(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    pics = []
    for n in range(len(g["types"])):
        p = {t: g["%s_%d" % (t, n)] for t in ("SRCY", "SRCU", "SRCV", "RECY", "RECU", "RECV")}
        for k in ("imbrec", "mbrec"):
            if "%s_%d" % (k, n) in g:
                p[k] = g["%s_%d" % (k, n)].astype(np.int32)
        if "rbsp_%d" % n in g:
            p["rbsp"], p["slice_bit0"] = g["rbsp_%d" % n], int(g["slbit0_%d" % n][0])
        pics.append(p)
    return g["params"], pics


def test_fixtures_exist():
    assert len(INTRA_GOLDENS) >= 5


@pytest.mark.parametrize("name", INTRA_GOLDENS)
def test_core_matches_the_reference_on_the_golden_i_pictures(host_lib, name):
    params, pics = golden_pictures(name)
    check_clip(host_lib, pics, int(params[4]), name)


def test_fixtures_cover_both_macroblock_kinds_and_the_p_skip_state():
    kinds, skipped_before_i = set(), 0
    for name in INTRA_GOLDENS:
        _, pics = golden_pictures(name)
        for n, p in enumerate(pics):
            if "imbrec" in p:
                kinds |= set(np.unique(np.minimum(p["imbrec"][:, 0], 1)))
                if n and "mbrec" in pics[n - 1]:
                    skipped_before_i += int((pics[n - 1]["mbrec"][:, 0] == 31).sum())
    assert kinds == {0, 1}, "fixtures must contain Intra4x4 and Intra16x16 macroblocks"
    assert skipped_before_i > 0, "fixtures must contain an I picture that follows P_Skip macroblocks"


def test_the_previous_pictures_p_skip_state_is_part_of_the_result(host_lib):
    """The first bit-cost trial reads mb_type_array[CurrMbAddr] of the PREVIOUS picture (intra.cpp:1008-1012): with that state
    dropped, the Intra16x16 trial bits of some macroblock after P_Skip macroblocks must change."""
    params, pics = golden_pictures("intra_static_qp12_ipi")
    p = pics[2]
    prev = pics[1]["mbrec"][:, 0].astype(np.int32)
    with_state, _ = host_picture(host_lib, p["SRCY"], p["SRCU"], p["SRCV"], int(params[4]), prev)
    without, _ = host_picture(host_lib, p["SRCY"], p["SRCU"], p["SRCV"], int(params[4]), None)
    assert np.array_equal(with_state["bits_intra16x16"], p["imbrec"][:, 3])
    assert (with_state["bits_intra16x16"] != without["bits_intra16x16"]).sum() > 50
    assert (with_state["mb_type"] != without["mb_type"]).any()


@pytest.mark.parametrize("w,h,seed,frames,qp,intra_every,kw", [
    (176, 144, 21, 3, 24, 2, {}),
    (96, 80, 22, 3, 36, 1, {"contrast": 0.2}),
    (352, 288, 23, 2, 28, 1000, {}),
    (128, 64, 24, 4, 51, 2, {"contrast": 0.6}),
    (64, 64, 25, 2, 0, 1, {}),
    (16, 16, 51, 3, 28, 1, {}),             # one macroblock
    (176, 16, 52, 3, 30, 2, {}),            # one macroblock row
    (16, 144, 53, 3, 26, 2, {}),            # one macroblock column
])
def test_core_matches_live_runs_of_the_reference(host_lib, w, h, seed, frames, qp, intra_every, kw):
    from h264_fer_b200 import synth
    from oracle import refdump
    if not refdump.have_ref_encoder():
        pytest.skip("oracle/_ref/ref_encoder not built")
    y4m = os.path.join(tempfile.mkdtemp(prefix="fh264_intra_live_"), "in.y4m")
    synth.write_y4m(y4m, w, h, seed, frames, **kw)
    _, dump, _ = refdump.run_reference(y4m, frames, qp=qp, intra_every=intra_every, dumpmask=refdump.D_MBREC | refdump.D_RECON | refdump.D_SOURCE | refdump.D_IMBREC)
    check_clip(host_lib, refdump.parse_dump(dump), qp, "live %dx%d qp %d" % (w, h, qp))


def build_lanes_lib(tsan):
    out = os.path.join(tempfile.mkdtemp(prefix="fh264_intra_lanes_"), "libintra_lanes.so")
    cmd = ["g++", "-O1", "-g", "-std=c++20", "-pthread", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "intra_host_lanes.cpp")]
    if tsan:
        cmd.insert(1, "-fsanitize=thread")
    subprocess.run(cmd, check=True)
    return out


def test_core_run_as_32_lanes_matches_the_reference():
    """The device runs a macroblock on the 32 lanes of a warp. tests/intra_host_lanes.cpp runs the same source as 32 host threads
    sharing one IcCtx (ic_sync = barrier, reductions through a shared array): the cross-lane structure gives the reference's results."""
    lib = build_lanes_lib(False)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "intra_lanes_check.py"), lib] + INTRA_GOLDENS,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert res.returncode == 0 and res.stdout.decode().count("LANES-OK") == len(INTRA_GOLDENS), res.stdout.decode()[-3000:]


def test_core_run_as_32_lanes_has_no_data_race():
    """The same under ThreadSanitizer: every access to the state the lanes share is ordered by a barrier (on the device: __syncwarp)."""
    tsan = subprocess.run(["gcc", "-print-file-name=libtsan.so"], stdout=subprocess.PIPE).stdout.decode().strip()
    if not os.path.isabs(tsan) or not os.path.isfile(tsan):
        pytest.skip("libtsan not available")
    lib = build_lanes_lib(True)
    env = dict(os.environ, LD_PRELOAD=tsan, TSAN_OPTIONS="halt_on_error=0 exitcode=0")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "intra_lanes_check.py"), lib, "intra_small_qp12_iii", "intra_lowcontrast_qp30_ipi"],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    text = res.stdout.decode()
    if "Running under ThreadSanitizer" not in text and "LANES-OK" not in text:
        pytest.skip("ThreadSanitizer could not be started here: " + text[-300:])
    assert "WARNING: ThreadSanitizer" not in text, text[text.index("WARNING: ThreadSanitizer"):][:3000]
    assert text.count("LANES-OK") == 2, text[-3000:]


def i_records_from_ints(r):
    """reference dump records [nmb, 439] -> fh264_mb_result_i array (what fh264_encode_i returns)."""
    r = np.asarray(r, np.int32)
    out = np.zeros(len(r), fh.MB_RESULT_I_DTYPE)
    out["mb_type"], out["intra16x16_pred_mode"], out["intra_chroma_pred_mode"] = r[:, 0], r[:, 1], r[:, 2]
    out["bits_intra16x16"], out["bits_intra4x4"], out["cbp_luma"], out["cbp_chroma"] = r[:, 3], r[:, 4], r[:, 5], r[:, 6]
    out["intra4x4_pred_mode"], out["prev_intra4x4_pred_mode_flag"] = r[:, 7:23], r[:, 23:39]
    out["rem_intra4x4_pred_mode"] = np.where(r[:, 23:39] == 0, r[:, 39:55], 0)
    out["luma"] = r[:, 55:311].reshape(-1, 16, 16)
    out["chroma_dc"] = r[:, 311:319].reshape(-1, 2, 4)
    out["chroma_ac"] = r[:, 319:439].reshape(-1, 2, 4, 15)
    return out


def check_slice_bits(data, nbits, rbsp, bit0, what):
    """slice data in bits [bit0 % 8, nbits) of data == the reference RBSP from its first slice_data bit, up to the trailing bits."""
    ref = np.unpackbits(np.asarray(rbsp, np.uint8))
    mine = np.unpackbits(np.asarray(data, np.uint8))[bit0 % 8:nbits]
    want = ref[bit0:bit0 + len(mine)]
    assert len(want) == len(mine), what + ": slice data longer than the reference RBSP"
    bad = np.nonzero(mine != want)[0]
    assert len(bad) == 0, "%s: slice data differs from bit %d of %d" % (what, bad[0], len(mine))
    tail = ref[bit0 + len(mine):]
    assert tail[0] == 1 and not tail[1:].any() and len(tail) <= 8, what + ": the reference's slice data does not end where ours does (rbsp_trailing_bits expected)"


@pytest.mark.parametrize("name", INTRA_GOLDENS)
def test_i_slice_data_matches_the_reference_rbsp(host_lib, name):
    """macroblock_layer() of every I macroblock (ic_write_macroblock) from the reference's records against the reference's RBSP."""
    params, pics = golden_pictures(name)
    done = 0
    for n, p in enumerate(pics):
        if "imbrec" not in p:
            continue
        h, w = p["SRCY"].shape
        rec = i_records_from_ints(p["imbrec"])
        out = np.zeros(len(p["rbsp"]) + 64, np.uint8)
        nbits, bad = C.c_int(0), C.c_int(0)
        bit0 = p["slice_bit0"]
        rc = host_lib.intra_host_slice(rec.ctypes.data_as(C.c_void_p), len(rec), w // 16, bit0 % 8, out.ctypes.data_as(C.c_void_p), len(out), C.byref(nbits), C.byref(bad))
        assert rc == 0 and bad.value == 0
        check_slice_bits(out, nbits.value, p["rbsp"], bit0, "%s picture %d" % (name, n))
        done += 1
    assert done > 0


SHIM_MOCK = os.path.join(ROOT, "integration", "_build", "shimtest_all_on_cpu_mock")


@pytest.mark.skipif(not os.path.isfile(SHIM_MOCK), reason="integration/_build/shimtest_all_on_cpu_mock not built (make -C integration, needs the reference sources)")
@pytest.mark.parametrize("w,h,seed,frames,qp,kw", [(176, 144, 41, 3, 28, {}), (96, 80, 42, 3, 36, {"contrast": 0.1}), (200, 120, 43, 2, 12, {})])
def test_reference_side_binding_for_device_idr_slices_on_a_mock_device(tmp_path, w, h, seed, frames, qp, kw):
    """Host logic of the all-on-device integration variant (integration/fh264_ref_shim.cpp, encode_idr_on_device: slice header by the
    reference, the device's slice data appended without shifting, trailing bits, idr_pic_id sequence, frame / dpb bookkeeping),
    linked against tests/mock_fh264_intra.cpp — a TEST-ONLY stand-in for the C ABI built from the host compile of the same
    I-picture core. All-IDR clips (IntraEvery 1): the Annex-B stream and the reconstructions must equal the reference's. The real
    library runs the same binding on the B200 in tests/test_gpu_integration.py."""
    from h264_fer_b200 import synth
    from oracle import refdump
    if not refdump.have_ref_encoder():
        pytest.skip("oracle/_ref/ref_encoder not built")
    y4m = str(tmp_path / "in.y4m")
    synth.write_y4m(y4m, w, h, seed, frames, **kw)
    summ, rd, ref264 = refdump.run_reference(y4m, frames, qp=qp, intra_every=1, dumpmask=refdump.D_RECON, out_264=str(tmp_path / "ref.264"),
                                             dump_path=str(tmp_path / "ref.bin"))
    out, dump = str(tmp_path / "mock.264"), str(tmp_path / "mock.bin")
    res = subprocess.run([SHIM_MOCK, y4m, out, dump, str(frames), str(qp), "0", "16", "3", "1", str(refdump.D_RECON), "-1"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert res.returncode == 0, res.stderr.decode()[-2000:]
    assert summ["types"] == "I" * frames
    assert open(out, "rb").read() == open(ref264, "rb").read(), "bitstreams differ"
    for n, (a, b) in enumerate(zip(refdump.parse_dump(dump), refdump.parse_dump(rd))):
        assert np.array_equal(a["RECY"], b["RECY"]) and np.array_equal(a["RECU"], b["RECU"]) and np.array_equal(a["RECV"], b["RECV"]), "picture %d" % n
