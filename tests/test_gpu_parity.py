"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the oracle on seeded inputs and against the
committed golden vectors from the unmodified reference. Bit-exact everywhere (integer/byte arithmetic)."""
import numpy as np
import pytest

import h264_fer_b200 as fh
from h264_fer_b200 import synth
from oracle import port

pytestmark = pytest.mark.gpu


def _loaded_native():
    import os
    maps = open("/proc/self/maps").read()
    return os.path.basename(fh.lib_path()) in maps          # (FH264_B200_LIB: an alternative BUILD of the same library, e.g. the -DFH_BOUNDS one)


def test_native_library_is_what_runs():
    s = fh.Session(64, 48)
    s.close()
    assert _loaded_native()


@pytest.mark.parametrize("w,h,seed", [(176, 144, 1), (208, 112, 2), (64, 48, 3)])
def test_phase_r_planes_and_features(w, h, seed):
    y, cb, cr = synth.SynthClip(w, h, seed).frame(0)
    o = port.Oracle(w, h)
    o.phase_r(y)
    with fh.Session(w, h) as s:
        s.upload_recon(0, y, cb, cr)
        for f in range(16):
            assert np.array_equal(s.debug_plane(0, f), o.plane(f)), "plane %d" % f
            for k in range(5):
                assert np.array_equal(s.debug_feature(0, k, f), o.kar(k, f)), "feature %d plane %d" % (k, f)


def test_phase_r_random_noise_edges():
    rng = np.random.default_rng(5)
    w, h = 80, 64
    y = rng.integers(16, 236, (h, w), dtype=np.uint8)
    c = rng.integers(16, 240, (h // 2, w // 2), dtype=np.uint8)
    o = port.Oracle(w, h)
    o.phase_r(y)
    with fh.Session(w, h) as s:
        s.upload_recon(0, y, c, c)
        for f in range(16):
            assert np.array_equal(s.debug_plane(0, f), o.plane(f))
            assert np.array_equal(s.debug_feature(0, 0, f), o.kar(0, f))
            assert np.array_equal(s.debug_feature(0, 4, f), o.kar(4, f))


@pytest.mark.parametrize("qp", [0, 12, 23, 24, 28, 37, 51])
def test_fused_tq_against_oracle(qp):
    rng = np.random.default_rng(qp)
    n = 300
    src = rng.integers(0, 256, (n, 384), dtype=np.uint8)
    pred = np.clip(src.astype(int) + rng.integers(-40, 41, (n, 384)), 0, 255).astype(np.uint8)
    pred[:20] = src[:20]                       # zero residual
    pred[20:40] = rng.integers(0, 256, (20, 384), dtype=np.uint8)   # large residual
    with fh.Session(64, 48) as s:
        lv, rc = s.tq_macroblocks(src, pred, qp)
    for i in range(n):
        elv, erc = port.tq_mb(src[i], pred[i], qp)
        assert np.array_equal(lv[i].astype(np.int32), elv), (qp, i)
        assert np.array_equal(rc[i], erc), (qp, i)


@pytest.mark.parametrize("qp", [10, 28, 36, 44])
def test_intra16_luma_dc_path_against_oracle(qp):
    rng = np.random.default_rng(100 + qp)
    n = 64
    src = rng.integers(0, 256, (n, 256), dtype=np.uint8)
    pred = np.clip(src.astype(int) + rng.integers(-30, 31, (n, 256)), 0, 255).astype(np.uint8)
    with fh.Session(64, 48) as s:
        dc, ac, rc = s.tq_luma_intra16(src, pred, qp)
    for i in range(n):
        edc, eac, erc = port.tq_luma_intra16(src[i], pred[i], qp)
        assert np.array_equal(dc[i].astype(np.int32), edc) and np.array_equal(ac[i].astype(np.int32), eac) and np.array_equal(rc[i], erc), (qp, i)


def test_intra16_luma_dc_path_against_reference_golden(golden):
    rec = golden.i16()
    if rec.shape[0] == 0:
        pytest.skip("no Intra16x16 macroblocks in this fixture")
    with fh.Session(64, 48) as s:
        dc, ac, recon = s.tq_luma_intra16(rec[:, :256].astype(np.uint8), rec[:, 256:512].astype(np.uint8), golden.qp)
    assert np.array_equal(dc, rec[:, 512:528]) and np.array_equal(ac.reshape(-1, 240), rec[:, 528:768])
    assert np.array_equal(recon, rec[:, 768:].astype(np.uint8))


def test_fused_tq_against_golden_tqio(golden):
    with fh.Session(64, 48) as s:
        for n in golden.p_pictures():
            io, want = golden.tqio(n), golden.mbrec(n)
            sel = np.where(want[:, 0] != 31)[0]
            lv, rc = s.tq_macroblocks(io[sel, :384], io[sel, 384:], golden.qp)
            assert np.array_equal(lv.astype(np.int32), want[sel, 21:])


def test_motion_compensation_including_picture_edges():
    w, h = 96, 80
    y, cb, cr = synth.SynthClip(w, h, 9).frame(0)
    rng = np.random.default_rng(1)
    nmb = (w // 16) * (h // 16)
    qmv = rng.integers(-80, 81, (nmb, 4, 2)).astype(np.int32)     # quarter-pel, up to 20 px: many cross the border
    qmv[0] = [[-400, -400]] * 4
    qmv[1] = [[3, 3], [2, 2], [1, 3], [3, 1]]
    o = port.Oracle(w, h)
    o.phase_r(y)
    want = o.mc_picture((y, cb, cr), qmv.reshape(nmb, 8))
    with fh.Session(w, h) as s:
        s.upload_recon(0, y, cb, cr)
        got = s.motion_compensate(0, qmv)
    assert np.array_equal(got, want), np.argwhere(got != want)[:5]


def test_scene_sad():
    w, h = 176, 144
    c = synth.SynthClip(w, h, 4)
    (y0, u0, v0), (y1, u1, v1) = c.frame(0), c.frame(1)
    with fh.Session(w, h) as s:
        s.upload_recon(0, y0, u0, v0)
        s.upload_source(0, y1, u1, v1)
        assert s.scene_sad(0) == port.scene_sad(y1, y0)
        s.upload_source(0, y0, u0, v0)
        assert s.scene_sad(0) == 0


def _check_picture(s, seq, golden, n, got_rec):
    want = golden.mbrec(n)
    got = fh.records_to_ints(got_rec)
    if not np.array_equal(got, want):
        d = np.argwhere(got != want)
        raise AssertionError("%s picture %d: %d record fields differ, first %s got %s want %s" % (
            golden.name, n, len(d), d[:6].tolist(), [int(got[a, b]) for a, b in d[:6]], [int(want[a, b]) for a, b in d[:6]]))
    ry, ru, rv = s.download_recon(seq)
    ey, eu, ev = golden.rec(n)
    assert np.array_equal(ry, ey) and np.array_equal(ru, eu) and np.array_equal(rv, ev), "%s picture %d recon" % (golden.name, n)


def test_encode_p_matches_reference_golden(golden):
    """Whole P pictures: MVs, mvds, SADs, mb_types, quantised levels and the reconstruction, against vectors dumped from
    the unmodified reference; I pictures are taken from the reference's reconstruction (host path)."""
    with fh.Session(golden.w, golden.h) as s:
        for n, t in enumerate(golden.types):
            if t == 5:
                s.upload_recon(0, *golden.rec(n))
                continue
            s.upload_source(0, *golden.src(n))
            rec = s.encode_p(golden.qp, golden.window, golden.maxdiff, golden.basic)[0]
            _check_picture(s, 0, golden, n, rec)
            counts = s.mode_counts(0)
            want_counts = golden.counts(n)
            if not golden.basic:        # with BasicInterEncoding the reference double-counts (moestimation.cpp:326,353,422)
                assert counts == want_counts, (counts, want_counts)


def test_encode_p_against_oracle_on_seeded_clip():
    """Chained P pictures (each predicts from the GPU's own reconstruction) against the oracle, WindowSize 32."""
    w, h, qp, window, maxdiff = 208, 160, 26, 32, 3
    clip = synth.SynthClip(w, h, 31)
    o = port.Oracle(w, h)
    ref = clip.frame(0)
    with fh.Session(w, h) as s:
        s.upload_recon(0, *ref)
        for t in range(1, 4):
            cur = clip.frame(t)
            assert not o.phase_r(ref[0])
            erec, erecon = o.encode_p(cur, ref, qp, window, maxdiff)
            s.upload_source(0, *cur)
            assert s.scene_sad(0) == port.scene_sad(cur[0], ref[0])
            got = fh.records_to_ints(s.encode_p(qp, window, maxdiff)[0])
            assert np.array_equal(got, erec), np.argwhere(got != erec)[:6]
            ry, ru, rv = s.download_recon(0)
            assert np.array_equal(ry, erecon[0]) and np.array_equal(ru, erecon[1]) and np.array_equal(rv, erecon[2])
            ref = erecon


def test_batch_of_sequences_equals_single_sequence_runs():
    """Sequences of a batch are independent: a batch-of-3 call equals three batch-of-1 sessions."""
    w, h, qp, window, maxdiff = 112, 96, 28, 16, -1
    clips = [synth.SynthClip(w, h, 40 + i, square=False) for i in range(3)]
    singles = []
    for c in clips:
        with fh.Session(w, h) as s:
            s.upload_recon(0, *c.frame(0))
            s.upload_source(0, *c.frame(1))
            singles.append((s.encode_p(qp, window, maxdiff)[0].copy(), s.download_recon(0)))
    with fh.Session(w, h, batch=3) as s:
        for i, c in enumerate(clips):
            s.upload_recon(i, *c.frame(0))
            s.upload_source(i, *c.frame(1))
        out = s.encode_p(qp, window, maxdiff)
        for i in range(3):
            assert out[i].tobytes() == singles[i][0].tobytes()
            for a, b in zip(s.download_recon(i), singles[i][1]):
                assert np.array_equal(a, b)


def test_error_paths():
    with fh.Session(64, 48) as s:
        with pytest.raises(fh.Fh264Error) as e:
            s.encode_p(28, 16, 3)
        assert e.value.code == -4          # no reference picture yet
        y = np.full((48, 64), 120, np.uint8); c = np.full((24, 32), 128, np.uint8)
        s.upload_recon(0, y, c, c)
        s.upload_source(0, y, c, c)
        with pytest.raises(fh.Fh264Error) as e:
            s.encode_p(28, 128, 3)
        assert e.value.code == -7          # WindowSize > 64
        with pytest.raises(fh.Fh264Error) as e:
            s.encode_p(77, 16, 3)
        assert e.value.code == -1


def test_reference_undefined_input_is_reported_not_emulated():
    """An all-zero reference has 8x8 sums of 0: the reference's counting sort is undefined there (moestimation.cpp:153-158)."""
    w, h = 64, 48
    z = np.zeros((h, w), np.uint8); c = np.full((h // 2, w // 2), 128, np.uint8)
    with fh.Session(w, h) as s:
        s.upload_recon(0, z, c, c)
        s.upload_source(0, z, c, c)
        with pytest.raises(fh.Fh264Error) as e:
            s.encode_p(28, 16, 3)
        assert e.value.code == -5


def test_wavefront_mix_of_skip_and_coded_macroblocks_is_exact_and_deterministic():
    """640x480 with adaptive MAXDIFF gives a mix of P_Skip and coded MBs: stresses the fine-grained wavefront dependencies
    (a P_Skip MB finishes early). Checked against the oracle and repeated for run-to-run determinism."""
    w, h, qp, window = 640, 480, 30, 32
    clip = synth.SynthClip(w, h, 77)
    ref, cur = clip.frame(0), clip.frame(1)
    o = port.Oracle(w, h)
    assert not o.phase_r(ref[0])
    want, want_recon = o.encode_p(cur, ref, qp, window, -1)
    types = set(np.unique(want[:, 0]).tolist())
    assert 31 in types and len(types) >= 3
    first = None
    for rep in range(4):
        with fh.Session(w, h) as s:
            s.upload_recon(0, *ref)
            s.upload_source(0, *cur)
            got = s.encode_p(qp, window, -1)[0]
            recon = s.download_recon(0)
        if first is None:
            first = got.tobytes()
            ints = fh.records_to_ints(got)
            assert np.array_equal(ints, want), np.argwhere(ints != want)[:6]
            assert all(np.array_equal(a, b) for a, b in zip(recon, want_recon))
        else:
            assert got.tobytes() == first, "run %d differs from run 0" % rep


def test_1080p_picture_against_oracle():
    """BASELINE's full size (1920x1080 input -> coded 1920x1072, WindowSize 32): one P picture, every record and the
    reconstruction against the oracle (the reference itself needs ~22 s per 1080p picture; the pinned port ~10 s)."""
    clip = synth.SynthClip(1920, 1080, 100)
    fr = [tuple(synth.crop16(p, chroma=(i > 0)) for i, p in enumerate(clip.frame(t))) for t in range(2)]
    h, w = fr[0][0].shape
    assert (w, h) == (1920, 1072)
    o = port.Oracle(w, h)
    assert not o.phase_r(fr[0][0])
    want, want_recon = o.encode_p(fr[1], fr[0], 28, 32, 3)
    with fh.Session(w, h) as s:
        s.upload_recon(0, *fr[0])
        s.upload_source(0, *fr[1])
        assert s.scene_sad(0) == port.scene_sad(fr[1][0], fr[0][0])
        got = fh.records_to_ints(s.encode_p(28, 32, 3)[0])
        recon = s.download_recon(0)
        counts = s.mode_counts(0)
    assert np.array_equal(got, want), np.argwhere(got != want)[:6]
    assert all(np.array_equal(a, b) for a, b in zip(recon, want_recon))
    assert sum(counts) == got.shape[0] == 8040
    # size-independent property: P_Skip macroblocks carry no levels and reconstruct to the prediction-only picture
    skip = got[:, 0] == 31
    assert not got[skip, 9:].any()


@pytest.mark.parametrize("contrast,noise", [(0.06, 0.5), (0.15, 1.0)])
def test_low_contrast_content_many_gated_candidates(contrast, noise):
    """Low-contrast pictures gate tens of thousands of stage-2 positions per partition (dense 8x8 sums); only those up to
    j_stop are candidates. Exercises the running j_stop bound / compaction of k_stage2 against the oracle."""
    w, h, qp, window = 320, 240, 28, 16
    clip = synth.SynthClip(w, h, 5, contrast=contrast, noise=noise, square=False)
    ref, cur = clip.frame(0), clip.frame(1)
    o = port.Oracle(w, h)
    assert not o.phase_r(ref[0])
    want, want_recon = o.encode_p(cur, ref, qp, window, 3)
    with fh.Session(w, h) as s:
        s.upload_recon(0, *ref)
        s.upload_source(0, *cur)
        got = fh.records_to_ints(s.encode_p(qp, window, 3)[0])       # raises on FH264_E_CAPACITY
        recon = s.download_recon(0)
        st = s.debug_status(0)
    assert st[0] == 0, "status flags %d" % st[0]
    assert np.array_equal(got, want), np.argwhere(got != want)[:6]
    assert all(np.array_equal(a, b) for a, b in zip(recon, want_recon))


def test_y4m_frame_ingest_crops_like_the_reference():
    """fh264_upload_source_frame: the centre crop of ReadFromY4M (fileIO.cpp:286-337) done by the strided H2D copy gives the
    same pictures — hence the same records and reconstruction — as cropping on the host (200x120 -> 192x112, odd crop offsets)."""
    w_in, h_in = 200, 120
    c = synth.SynthClip(w_in, h_in, 11)
    raw = [c.frame(t) for t in range(3)]
    crop = [tuple(synth.crop16(p, chroma=(i > 0)) for i, p in enumerate(f)) for f in raw]
    h, w = crop[0][0].shape
    out = []
    for mode in ("host_crop", "frame"):
        with fh.Session(w, h) as s:
            s.upload_recon(0, *crop[0])
            for t in (1, 2):
                if mode == "frame":
                    s.upload_source_frame(0, np.concatenate([p.ravel() for p in raw[t]]), w_in, h_in)
                else:
                    s.upload_source(0, *crop[t])
                sad = s.scene_sad(0)
                rec = s.encode_p(24, 32, 3, 0)[0].copy()
            out.append((sad, rec, s.download_recon(0)))
    assert out[0][0] == out[1][0]
    assert np.array_equal(out[0][1], out[1][1])
    for a, b in zip(out[0][2], out[1][2]):
        assert np.array_equal(a, b)
    with fh.Session(w, h) as s:
        with pytest.raises(fh.Fh264Error):
            s.upload_source_frame(0, np.zeros(176 * 144 * 3 // 2, np.uint8), 176, 144)      # does not crop to 192x112


@pytest.mark.parametrize("contrast,noise", [(0.0, 1.0), (0.0, 0.0), (0.02, 0.5)])
def test_flat_content_takes_the_slow_stage2_path_and_stays_exact(contrast, noise):
    """Flat / noisy-flat pictures: thousands of positions share one 8x8 sum, the stage-2 candidate set up to j_stop no longer
    fits the phase-A buffers, and phase B enumerates it itself (stage2_slow). Records and reconstruction against the oracle."""
    from oracle import port
    w, h, qp, window, maxdiff = 176, 144, 28, 16, 3
    clip = synth.SynthClip(w, h, 21, noise=noise, contrast=contrast)
    ref, cur = clip.frame(0), clip.frame(1)
    o = port.Oracle(w, h)
    ub = o.phase_r(ref[0])
    want_rec, want_recon = o.encode_p(cur, ref, qp, window, maxdiff)
    with fh.Session(w, h) as s:
        s.upload_recon(0, *ref)
        s.upload_source(0, *cur)
        if ub:
            with pytest.raises(fh.Fh264Error):
                s.encode_p(qp, window, maxdiff)
            return
        got = fh.records_to_ints(s.encode_p(qp, window, maxdiff)[0])
        recon = s.download_recon(0)
    assert np.array_equal(got, want_rec), "records differ: MBs %s" % np.nonzero((got != want_rec).any(1))[0][:10]
    for a, b in zip(recon, want_recon):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("window", [0, 2, 8, 24, 48, 64])
def test_other_window_sizes_against_oracle(window):
    """WindowSize 8 / 24 / 48 / 64: window/16 = 0, 1, 3, 4 -> 1x1, 3x3, 7x7, 9x9 quarter-pel windows (the batched variants of the
    on-the-fly feature computation, stage-1 key sets up to 1296) and window/2 = 4 .. 32 integer search ranges."""
    from oracle import port
    w, h, qp, maxdiff = 176, 144, 27, 3
    clip = synth.SynthClip(w, h, 31 + window)
    ref, cur = clip.frame(0), clip.frame(1)
    o = port.Oracle(w, h)
    assert not o.phase_r(ref[0])
    want_rec, want_recon = o.encode_p(cur, ref, qp, window, maxdiff)
    with fh.Session(w, h) as s:
        s.upload_recon(0, *ref)
        s.upload_source(0, *cur)
        got = fh.records_to_ints(s.encode_p(qp, window, maxdiff)[0])
        recon = s.download_recon(0)
    assert np.array_equal(got, want_rec), "window %d: records differ at MBs %s" % (window, np.nonzero((got != want_rec).any(1))[0][:10])
    for a, b in zip(recon, want_recon):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("force_miss", ["0", "1"])
@pytest.mark.parametrize("w,h", [(16, 16), (48, 16), (16, 64), (32, 32)])
def test_tiny_pictures_against_oracle(w, h, force_miss, monkeypatch):
    """One macroblock, one macroblock row, one macroblock column: every neighbour-availability corner of the wavefront, search
    windows and the 64x64 index tiles larger than the picture. force_miss: the warp-level phase B searches every partition itself
    (windows that leave the picture, P_Skip trials whose prediction crosses the picture edge)."""
    from oracle import port
    monkeypatch.setenv("FH264_PBW", force_miss)
    monkeypatch.setenv("FH264_PBW_FORCE_MISS", force_miss)
    qp, window, maxdiff = 30, 32, 3
    clip = synth.SynthClip(w, h, 50 + w + h, square=False)
    for t in (1, 2):
        ref, cur = clip.frame(t - 1), clip.frame(t)
        o = port.Oracle(w, h)
        if o.phase_r(ref[0]):
            continue
        want_rec, want_recon = o.encode_p(cur, ref, qp, window, maxdiff)
        with fh.Session(w, h) as s:
            s.upload_recon(0, *ref)
            s.upload_source(0, *cur)
            got = fh.records_to_ints(s.encode_p(qp, window, maxdiff)[0])
            recon = s.download_recon(0)
            bits = s.cavlc_p()[0][1]
        assert np.array_equal(got, want_rec), "%dx%d picture %d: records differ" % (w, h, t)
        for a, b in zip(recon, want_recon):
            assert np.array_equal(a, b)
        assert bits > 0


@pytest.mark.parametrize("spec", ["1", "0", "block", "miss", "miss-adaptive", "miss-basic"])
def test_speculative_fast_path_and_full_search_agree_with_the_oracle(spec, monkeypatch):
    """Phase S (spec.cuh) guesses the integer predictor and leaves finalists; phase B falls back to the full search on a wrong
    guess. All paths against the oracle over chained pictures (the second picture also has the temporal guess): the fast path
    switched off (FH264_SPEC=0: every partition takes the block-level full search) and on (most partitions must hit; "block":
    with the block-level kernel k_phase_b instead of the warp-level k_phase_b_warp), and the warp-level kernel with every lookup
    treated as a miss (FH264_PBW_FORCE_MISS=1: every partition reruns the warp-level search for its true predictor and every
    P_Skip trial is measured in the wavefront; also with adaptive MAXDIFF and with BasicInterEncoding)."""
    monkeypatch.setenv("FH264_SPEC", "0" if spec == "0" else "1")
    monkeypatch.setenv("FH264_PBW", "0" if spec == "block" else "1")
    if spec.startswith("miss"):
        monkeypatch.setenv("FH264_PBW_FORCE_MISS", "1")
    w, h, qp, window, maxdiff = 320, 208, 27, 32, 3
    basic = 0
    if spec == "miss-adaptive":
        maxdiff = -1
    if spec == "miss-basic":
        basic = 1
    clip = synth.SynthClip(w, h, 57)
    o = port.Oracle(w, h)
    ref = clip.frame(0)
    with fh.Session(w, h) as s:
        s.upload_recon(0, *ref)
        for t in range(1, 4):
            cur = clip.frame(t)
            assert not o.phase_r(ref[0])
            erec, erecon = o.encode_p(cur, ref, qp, window, maxdiff, basic)
            s.upload_source(0, *cur)
            got = fh.records_to_ints(s.encode_p(qp, window, maxdiff, basic)[0])
            assert np.array_equal(got, erec), np.argwhere(got != erec)[:6]
            assert all(np.array_equal(a, b) for a, b in zip(s.download_recon(0), erecon))
            st = s.debug_status(0)
            hits, misses = int(st[13]), int(st[14])
            coded = int((erec[:, 0] != 31).sum())
            assert hits + misses == 4 * coded, (hits, misses, coded)
            if spec == "0" or spec.startswith("miss"):
                assert hits == 0
            else:
                assert hits > 2 * misses, (hits, misses)
            ref = erecon
