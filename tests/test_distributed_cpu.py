"""CPU suite: the N>1 host logic with world_size 2 over gloo (no GPU): sequence sharding covers every sequence exactly once,
band arithmetic, max-over-ranks time reduction, and the reference arm's rank gating of bench.py."""
import os
import subprocess
import sys

import pytest
import torch.multiprocessing as mp

from h264_fer_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.sequences_for_rank(13, rank, world)
    allseq = sharding.gather_counts(mine)
    tmax = sharding.reduce_max(10.0 + rank)
    q.put((rank, mine, allseq, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing():
    world, port = 2, 29571
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6, 8, 10, 12] and res[1][1] == [1, 3, 5, 7, 9, 11]
    for _, _, allseq, tmax in res:
        assert sorted(sum(allseq, [])) == list(range(13))      # every sequence exactly once
        assert tmax == 11.0                                     # max over ranks, identical on every rank


def test_mb_row_bands():
    assert [b - a for a, b in sharding.mb_row_bands(67, 8)] == [9, 9, 9, 8, 8, 8, 8, 8]
    bands = sharding.mb_row_bands(67, 8)
    assert bands[0][0] == 0 and bands[-1][1] == 67 and all(bands[i][1] == bands[i + 1][0] for i in range(7))
    assert sharding.mb_row_bands(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    with pytest.raises(ValueError):
        sharding.sequences_for_rank(4, 2, 2)


def test_reference_arm_non_zero_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == b""


def test_host_mirror_frame_type_rule():
    """SequenceEncoder.select_nal_unit_type == selectNALUnitType (ref_frames.cpp:185-234) on a fake session."""
    from h264_fer_b200.encoder import SequenceEncoder, NAL_IDR, NAL_NON_IDR

    class Fake:
        nmb = 99
        def __init__(self): self.sad = 0
        def scene_sad(self, seq): return self.sad

    f = Fake()
    e = SequenceEncoder(f, 0, intra_every=4)
    assert e.select_nal_unit_type() == NAL_IDR                       # no dpb yet
    e.have_dpb = True; e.curr_frame_count = 1
    f.sad = 99 << 12
    assert e.select_nal_unit_type() == NAL_NON_IDR                   # strictly greater than MBs << 12 is required
    f.sad = (99 << 12) + 1
    assert e.select_nal_unit_type() == NAL_IDR
    f.sad = 0; e.curr_frame_count = 4
    assert e.select_nal_unit_type() == NAL_IDR                       # currFrameCount % IntraEvery == 0


def test_host_mirror_routes_idr_pictures():
    """SequenceEncoder.encode_picture: IDR pictures go to fh264_encode_i on the device, or to the intra_coder callback followed by
    fh264_upload_recon when one is given (the reference's host path); P pictures to fh264_encode_p."""
    from h264_fer_b200.encoder import SequenceEncoder, NAL_IDR, NAL_NON_IDR

    class Fake:
        nmb = 99
        def __init__(self): self.calls = []
        def upload_source(self, seq, y, cb, cr): self.calls.append("src")
        def scene_sad(self, seq): return 0
        def encode_i(self, qp, seq0=0, nseq=None): self.calls.append("encode_i"); return ["I"]
        def encode_p(self, qp, window, maxdiff_set, basic, seq0=0, nseq=None): self.calls.append("encode_p"); return ["P"]
        def upload_recon(self, seq, y, cb, cr): self.calls.append("upload_recon")

    f = Fake()
    e = SequenceEncoder(f, 0, intra_every=3)
    out = [e.encode_picture(None, None, None) for _ in range(4)]
    assert [n for n, _ in out] == [NAL_IDR, NAL_NON_IDR, NAL_NON_IDR, NAL_IDR]
    assert [r for _, r in out] == ["I", "P", "P", "I"]
    assert f.calls == ["src", "encode_i", "src", "encode_p", "src", "encode_p", "src", "encode_i"]

    f = Fake()
    e = SequenceEncoder(f, 0, intra_every=1000, intra_coder=lambda y, cb, cr: (y, cb, cr))
    assert e.encode_picture(None, None, None) == (NAL_IDR, None)
    assert e.encode_picture(None, None, None)[0] == NAL_NON_IDR
    assert f.calls == ["src", "upload_recon", "src", "encode_p"]


def test_band_wait_sets_are_symmetric_and_cover_the_halo():
    """Host mirror of fh264_band_peers: at 1080p on 8 GPUs the first rank waits for four bands, not eight; every pair is mutual; the
    neighbours are always in; with two ranks (or a small picture) everybody waits for everybody."""
    from h264_fer_b200 import sharding
    bands = sharding.mb_row_bands(67, 8)
    ws = sharding.band_wait_sets(bands, 1072)
    assert ws[0] == [0, 1, 2, 3] and ws[7] == [4, 5, 6, 7]
    for a in range(8):
        assert a in ws[a]
        for b in ws[a]:
            assert a in ws[b]
        if a + 1 < 8:
            assert a + 1 in ws[a]
        lo, hi = bands[a][0] * 16 - 304, bands[a][1] * 16 + 304
        for b, (q0, q1) in enumerate(bands):
            if q0 * 16 < hi and q1 * 16 > lo:
                assert b in ws[a], "rank %d reads rows of band %d" % (a, b)
    assert sharding.band_wait_sets(sharding.mb_row_bands(67, 2), 1072) == [[0, 1], [0, 1]]
    assert sharding.band_wait_sets(sharding.mb_row_bands(18, 4), 288) == [[0, 1, 2, 3]] * 4


def test_batch_host_mirror_routes_pictures_by_the_reference_rules():
    """BatchEncoder (encoder.py) on a fake session: first picture and every IntraEvery-th picture are IDR pictures for the whole batch
    decided on the host (ref_frames.cpp:191); in between the batch goes through ONE streaming step, and a sequence the device-side
    scene gate stopped (:210-224) is coded as an IDR picture right after, its source picture still current."""
    import numpy as np
    from h264_fer_b200.encoder import BatchEncoder, NAL_IDR, NAL_NON_IDR

    class Block:
        def __init__(self, n, nbytes): self.array = np.zeros((n, nbytes), np.uint8); self.ptr = 1234

    class Out:
        def __init__(self, n): self.n = n; self.gated = set()
        def coded(self): return [b not in self.gated for b in range(self.n)]
        def slices(self): return [(np.array([b], np.uint8), 8) for b in range(self.n)]

    class Fake:
        w, h, batch, nmb = 32, 16, 3, 2
        def __init__(self): self.calls = []
        def upload_source_batch(self, ptr, stride): self.calls.append(("upload", ptr, stride))
        def encode_p_stream(self, qp, window, maxdiff, basic, scene_gate, out): self.calls.append(("stream", scene_gate))
        def sync(self): self.calls.append(("sync",))
        def picture_status(self, b): pass
        def encode_i(self, qp, seq0=0, nseq=None): self.calls.append(("encode_i", seq0))
        def cavlc_i(self, first_bit=0, seq0=0, nseq=None): return [(np.array([100 + seq0], np.uint8), 16)]

    f, out = Fake(), Out(3)
    be = BatchEncoder(f, intra_every=3, buffers=(Block(3, 32 * 16 * 3 // 2), out))
    pics = [(np.full((16, 32), 10 * b, np.uint8), np.zeros((8, 16), np.uint8), np.zeros((8, 16), np.uint8)) for b in range(3)]
    r0 = be.encode_pictures(pics)
    assert [t for t, _, _ in r0] == [NAL_IDR] * 3 and [c[0] for c in f.calls] == ["upload", "encode_i", "encode_i", "encode_i"]
    assert be.block.array[1, 0] == 10 and be.block.array[2, 32 * 16 - 1] == 20 and f.calls[0][1:] == (1234, 32 * 16 * 3 // 2)
    f.calls.clear()
    out.gated = {1}                                           # picture 1: the gate stops sequence 1
    r1 = be.encode_pictures(pics)
    assert [t for t, _, _ in r1] == [NAL_NON_IDR, NAL_IDR, NAL_NON_IDR]
    assert [c[0] for c in f.calls] == ["upload", "stream", "sync", "encode_i"] and f.calls[1] == ("stream", 1) and f.calls[3] == ("encode_i", 1)
    assert int(r1[1][1][0]) == 101 and r1[1][2] == 16 and int(r1[2][1][0]) == 2
    f.calls.clear(); out.gated = set()
    assert [t for t, _, _ in be.encode_pictures(pics)] == [NAL_NON_IDR] * 3
    f.calls.clear()
    assert [t for t, _, _ in be.encode_pictures(pics)] == [NAL_IDR] * 3          # currFrameCount 3 % IntraEvery 3 == 0
    assert "stream" not in [c[0] for c in f.calls]
