"""CPU suite: pins the oracle (oracle/fh264_oracle.c) against the committed golden vectors, which were produced by the
unmodified reference (tests/golden/make_golden.py), and — when the compiled reference binary is present
(oracle/_ref/ref_encoder travels with the snapshot) — against a fresh run of the reference itself."""
import os

import numpy as np
import pytest

from h264_fer_b200 import synth
from oracle import port, refdump


def test_oracle_matches_golden_records_and_recon(golden):
    o = port.Oracle(golden.w, golden.h)
    for n in golden.p_pictures():
        ref = golden.rec(n - 1)
        assert not o.phase_r(ref[0]), "golden input must stay clear of the reference's undefined behaviour"
        rec, recon = o.encode_p(golden.src(n), ref, golden.qp, golden.window, golden.maxdiff, golden.basic)
        want = golden.mbrec(n)
        assert np.array_equal(rec, want), "picture %d: records differ at %s" % (n, np.argwhere(rec != want)[:5].tolist())
        for got, exp, nm in zip(recon, golden.rec(n), "Y Cb Cr".split()):
            assert np.array_equal(got, exp), "picture %d: %s reconstruction differs" % (n, nm)


def test_oracle_tq_matches_golden_tqio(golden):
    """Transform/quant/reconstruction in isolation: reference's own (snapped source, prediction) -> levels, recon."""
    for n in golden.p_pictures():
        io, want = golden.tqio(n), golden.mbrec(n)
        ry, ru, rv = golden.rec(n)
        wmb = golden.w // 16
        for mb in range(0, io.shape[0], 7):
            if want[mb, 0] == 31:
                continue
            lv, rc = port.tq_mb(io[mb, :384], io[mb, 384:], golden.qp)
            assert np.array_equal(lv, want[mb, 21:]), (n, mb)
            x, y = (mb % wmb) * 16, (mb // wmb) * 16
            assert np.array_equal(rc[:256].reshape(16, 16), ry[y:y + 16, x:x + 16])
            assert np.array_equal(rc[256:320].reshape(8, 8), ru[y // 2:y // 2 + 8, x // 2:x // 2 + 8])
            assert np.array_equal(rc[320:].reshape(8, 8), rv[y // 2:y // 2 + 8, x // 2:x // 2 + 8])


def test_mode_coverage_of_goldens():
    """The committed vectors exercise P_Skip and all four inter partitionings, and both quantiser branches."""
    from conftest import Golden, golden_paths
    seen, qps = set(), set()
    for p in golden_paths():
        g = Golden(p)
        qps.add(g.qp < 24)
        for n in g.p_pictures():
            seen |= set(np.unique(g.mbrec(n)[:, 0]).tolist())
    assert {0, 1, 2, 4, 31} <= seen, seen
    assert qps == {True, False}


def test_oracle_intra16_luma_dc_path_matches_reference(golden):
    """a13: per-block transform with DC kept + 4x4 Hadamard of the DCs, against records tapped from the reference's I pictures
    (quantizationTransform for Intra16x16 MBs; both quantiser branches qP < 36 and >= 36 are in the goldens)."""
    rec = golden.i16()
    for r in rec:
        src, pred = r[:256].astype(np.uint8), r[256:512].astype(np.uint8)
        dc, ac, recon = port.tq_luma_intra16(src, pred, golden.qp)
        assert np.array_equal(dc, r[512:528]) and np.array_equal(ac.ravel(), r[528:768]) and np.array_equal(recon, r[768:].astype(np.uint8))


def test_intra16_goldens_cover_both_dc_quantiser_branches():
    from conftest import Golden, golden_paths
    qps = [Golden(p).qp for p in golden_paths() if Golden(p).i16().shape[0] >= 50]
    assert any(q < 36 for q in qps) and any(q >= 36 for q in qps)


def test_scene_sad_oracle():
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, 5000, dtype=np.uint8)
    b = rng.integers(0, 256, 5000, dtype=np.uint8)
    assert port.scene_sad(a, b) == int(np.abs(a.astype(int) - b.astype(int)).sum())
    assert port.scene_sad(a, a) == 0


def test_intra16_dc_path_roundtrip_properties():
    """Intra16x16 luma-DC Hadamard path (a13): zero residual -> zero levels and recon == pred; a flat residual only
    touches the DC list."""
    pred = np.full(256, 100, np.uint8)
    dc, ac, rc = port.tq_luma_intra16(pred, pred, 28)
    assert not dc.any() and not ac.any() and np.array_equal(rc, pred)
    src = np.full(256, 140, np.uint8)
    dc, ac, rc = port.tq_luma_intra16(src, pred, 28)
    assert dc[0] != 0 and not dc[1:].any() and not ac.any()
    assert abs(int(rc[0]) - 140) <= 6 and len(set(rc.tolist())) == 1


@pytest.mark.skipif(not refdump.have_ref_encoder(), reason="compiled reference (oracle/_ref/ref_encoder) not present")
@pytest.mark.parametrize("w,h,seed,frames,qp,window,maxdiff,basic", [(176, 144, 21, 3, 24, 16, 3, 0), (112, 96, 22, 3, 28, 32, -1, 0),
                                                                      (96, 96, 23, 3, 12, 8, 3, 1)])
def test_oracle_matches_live_reference(tmp_path, w, h, seed, frames, qp, window, maxdiff, basic):
    y4m = str(tmp_path / "in.y4m")
    synth.write_y4m(y4m, w, h, seed, frames, square=(w > 128))
    summ, dump, _ = refdump.run_reference(y4m, frames, qp=qp, basic=basic, window=window, maxdiff=maxdiff,
                                          dumpmask=refdump.D_MBREC | refdump.D_RECON | refdump.D_SOURCE | refdump.D_PHASE_R, planes_pic=0)
    pics = refdump.parse_dump(dump)
    W, H = pics[0]["w"], pics[0]["h"]
    o = port.Oracle(W, H)
    o.phase_r(pics[0]["RECY"])
    for f in range(16):
        assert np.array_equal(o.plane(f).ravel(), pics[0]["planes"][f]), "plane %d" % f
    i = 0
    for f in range(16):
        for k in range(5):
            assert np.array_equal(o.kar(k, f).ravel(), pics[0]["kar"][i]), "feature %d of plane %d" % (k, f)
            i += 1
    for a in range(5):
        assert np.array_equal(o.sorted(a), pics[0]["sorted"][a])
    assert np.array_equal(o.bucket_start()[:16384], pics[0]["koliko"])
    npic = 0
    for n in range(1, len(pics)):
        if pics[n]["nal_type"] != 1:
            continue
        ref = (pics[n - 1]["RECY"], pics[n - 1]["RECU"], pics[n - 1]["RECV"])
        o.phase_r(ref[0])
        rec, recon = o.encode_p((pics[n]["SRCY"], pics[n]["SRCU"], pics[n]["SRCV"]), ref, qp, window, maxdiff, basic)
        assert np.array_equal(rec, pics[n]["mbrec"])
        assert np.array_equal(recon[0], pics[n]["RECY"]) and np.array_equal(recon[1], pics[n]["RECU"]) and np.array_equal(recon[2], pics[n]["RECV"])
        npic += 1
    assert npic >= 1


def test_synth_is_deterministic_and_crops_like_reference():
    a = synth.SynthClip(200, 120, 5).frame(3)
    b = synth.SynthClip(200, 120, 5).frame(3)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    y = synth.crop16(a[0]); c = synth.crop16(a[1], chroma=True)
    assert y.shape == (112, 192) and c.shape == (56, 96)
    assert np.array_equal(y, a[0][4:116, 4:196]) and np.array_equal(c, a[1][2:58, 2:98])
    assert a[0].min() >= 16 and a[0].max() <= 235
