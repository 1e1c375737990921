"""GPU suite, needs >= 2 GPUs (skipped otherwise): band mode — one picture split into macroblock-row bands over the GPUs,
wavefront crossing GPUs through mirrored progress flags, reconstruction exchanged by peer stores inside phase C.
Bit-exact against the oracle, picture after picture (each picture predicts from the exchanged reconstruction)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, w, h, seed, npics, qp, window, maxdiff, q, stream=False):
    import torch
    import torch.distributed as dist
    import h264_fer_b200 as fh
    from h264_fer_b200 import synth
    from h264_fer_b200.bands import BandSession
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    clip = synth.SynthClip(w, h, seed)
    bs = BandSession(w, h, device=rank)
    bs.upload_recon(*clip.frame(0))
    prev_recon_y = clip.frame(0)[0]
    out = []
    for t in range(1, npics):
        if stream:
            # the streaming step (fh264_encode_p_stream) in band mode: one upload call, scene SAD measured on the device, records home
            # on the copy stream; every rank holds the whole picture, so every rank must report the same scene SAD
            from h264_fer_b200 import native
            from oracle import port as oport
            y, cb, cr = clip.frame(t)
            blk = native.PinnedArray((1, w * h * 3 // 2), np.uint8)
            blk.array[0] = np.concatenate([y.ravel(), cb.ravel(), cr.ravel()])
            so = native.StreamOut(1, bs.s.nmb, records=True)
            bs.s.upload_source_batch(blk.ptr, w * h * 3 // 2)
            bs.s.encode_p_stream(qp, window, maxdiff, scene_gate=2, out=so)
            bs.s.sync()
            bs.s.picture_status(0)
            assert so.coded() == [True]
            # with the peers' bands known a rank measures its own band: the ranks' sums add up to the picture's (ref_frames.cpp:210-224)
            parts = [None] * world
            dist.all_gather_object(parts, int(so.scene_sad()[0]))
            assert sum(parts) == oport.scene_sad(y, prev_recon_y), "scene SAD of picture %d" % t
            band = so.records.array[0][bs.mb_slice].copy()
        else:
            bs.upload_source(*clip.frame(t))
            band = bs.encode_p(qp, window, maxdiff)
        rec = bs.gather_records(band)
        recon = bs.download_recon()
        prev_recon_y = recon[0]
        out.append((fh.records_to_ints(rec), recon))
    q.put((rank, out))
    bs.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="band mode needs at least 2 GPUs")
@pytest.mark.parametrize("world,w,h,window,maxdiff,stream", [(2, 352, 288, 32, 3, False), (2, 640, 480, 32, -1, False), (4, 352, 288, 32, 3, False),
                                                             (8, 640, 480, 32, 3, False), (2, 352, 288, 32, 3, True), (8, 640, 480, 32, 3, True)])
def test_band_mode_matches_oracle(world, w, h, window, maxdiff, stream):
    if world > 2 and _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    from h264_fer_b200 import synth
    from oracle import port
    world = min(world, _ngpu())
    seed, npics, qp = 9, 4, 28
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29641, w, h, seed, npics, qp, window, maxdiff, q, stream)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    clip = synth.SynthClip(w, h, seed)
    o = port.Oracle(w, h)
    ref = clip.frame(0)
    for t in range(1, npics):
        assert not o.phase_r(ref[0])
        want, want_recon = o.encode_p(clip.frame(t), ref, qp, window, maxdiff)
        for r in range(world):
            got, recon = res[r][t - 1]
            assert np.array_equal(got, want), "rank %d picture %d: %s" % (r, t, np.argwhere(got != want)[:6].tolist())
            assert all(np.array_equal(a, b) for a, b in zip(recon, want_recon)), "rank %d picture %d recon" % (r, t)
        ref = want_recon


def _worker_missing_peer(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import h264_fer_b200 as fh
    from h264_fer_b200 import synth
    from h264_fer_b200.bands import BandSession
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h = 352, 288
    clip = synth.SynthClip(w, h, 9)
    bs = BandSession(w, h, device=rank)
    bs.upload_recon(*clip.frame(0))
    bs.upload_source(*clip.frame(1))
    bs.encode_p(28, 32, 3)                                    # picture 1: every rank takes part
    outcome = "ok"
    if rank == 0:
        bs.upload_source(*clip.frame(2))
        bs.encode_p(28, 32, 3)                                # picture 2: rank 1 is missing; its band never arrives, the picture barrier times out
        bs.upload_source(*clip.frame(3))
        try:
            bs.encode_p(28, 32, 3)                            # picture 3 would predict from an incomplete reference picture
            outcome = "no error"
        except fh.Fh264Error as e:
            outcome = "error %d" % e.code
    q.put((rank, outcome))
    bs.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="band mode needs at least 2 GPUs")
def test_band_mode_reports_a_missing_peer_instead_of_coding_from_an_incomplete_reference():
    """A rank that skips a picture leaves the others' reference picture without its band: the picture barrier's bounded wait
    must surface as FH264_E_STATE on the next picture (it used to be cleared by phase R before anybody read it)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_missing_peer, args=(r, 2, 29643, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert res[0] == "error -4", res


def _worker_async(rank, world, port, w, h, seed, npics, qp, window, maxdiff, q):
    """All pictures enqueued without any host synchronisation between them (fh264_upload_source_batch + fh264_encode_p_stream): with
    the peers' bands known (fh264_band_peers) a rank whose halo is complete runs ahead of the ranks further down."""
    import torch
    import torch.distributed as dist
    import h264_fer_b200 as fh
    from h264_fer_b200 import native, synth
    from h264_fer_b200.bands import BandSession
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    clip = synth.SynthClip(w, h, seed)
    bs = BandSession(w, h, device=rank)
    bs.upload_recon(*clip.frame(0))
    bs.s.sync()
    dist.barrier()
    pic = w * h * 3 // 2
    blocks, outs = [], []
    for t in range(1, npics):
        blk = native.PinnedArray((1, pic), np.uint8)
        blk.array[0] = np.concatenate([p.ravel() for p in clip.frame(t)])
        so = native.StreamOut(1, bs.s.nmb, records=True)
        blocks.append(blk); outs.append(so)
        bs.s.upload_source_batch(blk.ptr, pic)
        bs.s.encode_p_stream(qp, window, maxdiff, scene_gate=2, out=so)
    bs.s.sync()
    bs.s.picture_status(0)
    dist.barrier()                                            # every rank has finished every picture: the whole reconstruction is everywhere
    recs = [fh.records_to_ints(bs.gather_records(so.records.array[0][bs.mb_slice].copy())) for so in outs]
    sads = []
    for so in outs:
        parts = [None] * world
        dist.all_gather_object(parts, int(so.scene_sad()[0]))
        sads.append(sum(parts))
    q.put((rank, (recs, sads, bs.download_recon())))
    bs.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="band mode needs at least 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_band_mode_pipelined_pictures_match_oracle(world):
    """A tall picture (68 macroblock rows, more than two halos) coded by 2 / 4 / 8 ranks with nothing synchronised between pictures:
    the upper ranks start the next picture while the wavefront of the current one is still in the lower bands. Records of every
    picture, the scene SADs (sum of the ranks' band sums) and the final reconstruction against the oracle."""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    from h264_fer_b200 import synth
    from oracle import port
    w, h, seed, npics, qp, window, maxdiff = 352, 1088, 21, 6, 28, 32, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_async, args=(r, world, 29647, w, h, seed, npics, qp, window, maxdiff, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    clip = synth.SynthClip(w, h, seed)
    o = port.Oracle(w, h)
    ref = clip.frame(0)
    for t in range(1, npics):
        assert not o.phase_r(ref[0])
        want_sad = port.scene_sad(clip.frame(t)[0], ref[0])
        want, ref = o.encode_p(clip.frame(t), ref, qp, window, maxdiff)
        for r in range(world):
            recs, sads, _ = res[r]
            assert np.array_equal(recs[t - 1], want), "rank %d picture %d: %s" % (r, t, np.argwhere(recs[t - 1] != want)[:6].tolist())
            assert sads[t - 1] == want_sad, "picture %d scene SAD" % t
    for r in range(world):
        assert all(np.array_equal(a, b) for a, b in zip(res[r][2], ref)), "rank %d final reconstruction" % r


def _worker_ipi(rank, world, port, name, q):
    """I -> P -> I in band mode on a committed fixture of the compiled reference: the I pictures are coded whole by every rank
    (fh264_encode_i; the second one reads which macroblocks of the P picture were P_Skip — types mirrored by phase C from every band),
    the P picture is split into bands."""
    import torch
    import torch.distributed as dist
    import h264_fer_b200 as fh
    from h264_fer_b200.bands import BandSession
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_intra_host import compare_i_records, golden_pictures
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params, pics = golden_pictures(name)
    qp, window, maxdiff = int(params[4]), int(params[5]), int(params[6])
    h, w = pics[0]["SRCY"].shape
    bs = BandSession(w, h, device=rank)
    outcome = "ok"
    try:
        for n, p in enumerate(pics):
            bs.upload_source(p["SRCY"], p["SRCU"], p["SRCV"])
            if "imbrec" in p:
                rec = bs.s.encode_i(qp)[0]
                compare_i_records(fh.i_records_to_ints(rec), p["imbrec"], "%s picture %d (I) on rank %d" % (name, n, rank))
            else:
                band = bs.encode_p(qp, window, maxdiff)
                mine, ref = fh.records_to_ints(bs.gather_records(band)), p["mbrec"]
                assert np.array_equal(mine[:, :17], ref[:, :17]) and np.array_equal(mine[:, 21:], ref[:, 21:]), "picture %d (P)" % n
            dist.barrier()
            for got, t in zip(bs.download_recon(), ("RECY", "RECU", "RECV")):
                assert np.array_equal(got, p[t]), "picture %d: %s differs on rank %d" % (n, t, rank)
    except AssertionError as e:
        outcome = "FAILED: %s" % e
    q.put((rank, outcome))
    bs.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="band mode needs at least 2 GPUs")
@pytest.mark.parametrize("name", ["intra_static_qp12_ipi", "intra_qcif_qp28_ipi"])
def test_band_mode_codes_i_pictures_whole_on_every_rank(name):
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_ipi, args=(r, world, 29651, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert res == {0: "ok", 1: "ok"}, res


def _worker_cavlc(rank, world, port, golden_path, q):
    """P slices of a committed clip of the compiled reference coded in band mode; rank 0, where phase C of every rank gathers the
    picture's records over NVLink, entropy-codes the slice on the device (fh264_cavlc_p): slice data against the reference's RBSP."""
    import sys
    import torch
    import torch.distributed as dist
    import h264_fer_b200 as fh
    from h264_fer_b200.bands import BandSession
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import Golden
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = Golden(golden_path)
    bs = BandSession(g.w, g.h, device=rank, gather=True)
    outcome, checked = "ok", 0
    try:
        for n, t in enumerate(g.types):
            if t == 5:
                bs.upload_recon(*g.rec(n))
                bs.s.sync()
                dist.barrier()
                continue
            bs.upload_source(*g.src(n))
            bs.encode_p(g.qp, g.window, g.maxdiff, g.basic)
            if rank == 0:
                rbsp, bit0 = g.slice_rbsp(n)
                data, nbits = bs.s.cavlc_p(first_bit=bit0 % 8)[0]
                nd = nbits - bit0 % 8
                ref, mine = np.unpackbits(np.asarray(rbsp, np.uint8)), np.unpackbits(np.asarray(data, np.uint8))
                assert np.array_equal(mine[bit0 % 8:nbits], ref[bit0:bit0 + nd]), "picture %d: slice data differs" % n
                tail = ref[bit0 + nd:]
                assert tail[0] == 1 and not tail[1:].any() and len(tail) <= 8
                checked += 1
            else:
                try:
                    bs.s.cavlc_p()
                    outcome = "FAILED: cavlc_p on rank %d did not refuse" % rank
                except fh.Fh264Error as e:
                    assert e.code == -7
            dist.barrier()
        if rank == 0:
            assert checked == len(g.p_pictures())
    except AssertionError as e:
        outcome = "FAILED: %s" % e
    q.put((rank, outcome))
    bs.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="band mode needs at least 2 GPUs")
def test_band_mode_slice_data_is_entropy_coded_on_rank_0():
    import torch.multiprocessing as mp
    world = 2
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qcif_w16_qp28.npz")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_cavlc, args=(r, world, 29653, path, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert res == {0: "ok", 1: "ok"}, res
