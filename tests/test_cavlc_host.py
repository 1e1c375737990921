"""CPU checks of the device CAVLC core (h264_fer_b200/csrc/cavlc_core.h, compiled for the host by g++): its tables against the
reference's coder tables and the slice data it produces from the golden per-MB records against the reference's RBSP
(tests/golden/*.npz, made from the compiled reference by tests/golden/make_golden.py). SURVEY.md §8(f) rank 1."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def host_lib():
    out = os.path.join(tempfile.mkdtemp(prefix="fh264_cavlc_"), "libcavlc_host.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "cavlc_host.cpp")], check=True)
    return C.CDLL(out)


def bits_of(buf):
    return np.unpackbits(np.asarray(buf, np.uint8))


def records_from_golden(g, n):
    """golden per-MB record (405 ints) -> fh264_mb_result array (the layout the device coder reads)."""
    from h264_fer_b200 import native as fh
    r = g["mbrec_%d" % n].astype(np.int32)
    out = np.zeros(len(r), fh.MB_RESULT_DTYPE)
    out["mb_type"] = r[:, 0]
    nparts = {0: 1, 1: 2, 2: 2, 4: 4, 31: 0}
    out["num_parts"] = [nparts[int(t)] for t in r[:, 0]]
    out["mv"] = r[:, 1:9].reshape(-1, 4, 2)
    out["mvd"] = r[:, 9:17].reshape(-1, 4, 2)
    out["luma"] = r[:, 21:277].reshape(-1, 16, 16)
    out["chroma_dc"] = r[:, 277:285].reshape(-1, 2, 4)
    out["chroma_ac"] = r[:, 285:405].reshape(-1, 2, 4, 15)
    return out


def test_tables_match_the_reference(host_lib):
    ref = np.load(os.path.join(GOLD, "cavlc_tables.npz"))["tables"]
    mine = np.zeros(4096, np.int32)
    n = host_lib.cavlc_host_tables(mine.ctypes.data_as(C.c_void_p))
    assert n == len(ref)
    mine = mine[:n]
    # entries the reference never emits have length 0 there; compare (length, code) wherever the reference defines a code
    pairs = n - 48
    rl, rc = ref[:pairs:2], ref[1:pairs:2]
    ml, mc = mine[:pairs:2], mine[1:pairs:2]
    used = rl > 0
    bad = np.nonzero(used & ((rl != ml) | (rc != mc)))[0]
    assert len(bad) == 0, "table entries differ at pair indices %s: ref %s mine %s" % (bad[:10], list(zip(rl[bad][:10], rc[bad][:10])), list(zip(ml[bad][:10], mc[bad][:10])))
    assert np.array_equal(ref[pairs:], mine[pairs:]), "coded_block_pattern map differs"


@pytest.mark.parametrize("name", ["qcif_w16_qp28", "cif_crop_w32_qp20_adaptive", "small_basic_qp33"])
def test_slice_data_matches_the_reference(host_lib, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    w = int(g["params"][0]) // 16 * 16
    wmb = w // 16
    checked = 0
    for n in range(len(g["types"])):
        if "rbsp_%d" % n not in g.files:
            continue
        rec = records_from_golden(g, n)
        rbsp, bit0 = g["rbsp_%d" % n], int(g["slbit0_%d" % n][0])
        out = np.zeros(len(rbsp) + 64, np.uint8)
        nbits, bad = C.c_int(0), C.c_int(0)
        rc = host_lib.cavlc_host_slice(rec.ctypes.data_as(C.c_void_p), len(rec), wmb, bit0 % 8, out.ctypes.data_as(C.c_void_p), len(out), C.byref(nbits), C.byref(bad))
        assert rc == 0 and bad.value == 0
        nd = nbits.value - bit0 % 8                      # slice_data bits
        ref_bits = bits_of(rbsp)
        mine = bits_of(out)[bit0 % 8:bit0 % 8 + nd]
        assert bit0 + nd + 1 <= len(ref_bits), "slice data longer than the reference's"
        assert np.array_equal(mine, ref_bits[bit0:bit0 + nd]), "picture %d: first differing slice_data bit %d of %d" % (
            n, int(np.nonzero(mine != ref_bits[bit0:bit0 + nd])[0][0]), nd)
        # what follows in the reference is rbsp_trailing_bits: a one, then zeros to the byte boundary, and nothing else
        tail = ref_bits[bit0 + nd:]
        assert tail[0] == 1 and not tail[1:].any() and len(tail) <= 8
        checked += 1
    assert checked >= 2


def test_all_skip_slices_match_a_live_reference_run(host_lib):
    """A static clip at a fine quantiser: P pictures of 98 and 99 P_Skip macroblocks out of 99 — the slice data is (almost) nothing
    but the mb_skip_run that ends the slice (rbsp_encoding.cpp:310)."""
    import tempfile
    from h264_fer_b200 import native as fh, synth
    from oracle import refdump
    if not refdump.have_ref_encoder():
        pytest.skip("oracle/_ref/ref_encoder not built")
    y4m = os.path.join(tempfile.mkdtemp(prefix="fh264_cavlc_static_"), "in.y4m")
    synth.write_y4m(y4m, 176, 144, 31, 4, pan=(0, 0), noise=0.0, square=False)
    _, dump, _ = refdump.run_reference(y4m, 4, qp=12, intra_every=2, dumpmask=refdump.D_MBREC | refdump.D_SLICE)
    skipped = []
    for n, p in enumerate(refdump.parse_dump(dump)):
        if "mbrec" not in p:
            continue
        r = p["mbrec"]
        rec = np.zeros(len(r), fh.MB_RESULT_DTYPE)
        rec["mb_type"] = r[:, 0]
        rec["num_parts"] = [{0: 1, 1: 2, 2: 2, 4: 4, 31: 0}[int(t)] for t in r[:, 0]]
        rec["mv"], rec["mvd"] = r[:, 1:9].reshape(-1, 4, 2), r[:, 9:17].reshape(-1, 4, 2)
        rec["luma"], rec["chroma_dc"], rec["chroma_ac"] = r[:, 21:277].reshape(-1, 16, 16), r[:, 277:285].reshape(-1, 2, 4), r[:, 285:405].reshape(-1, 2, 4, 15)
        bit0 = p["slice_bit0"]
        out = np.zeros(len(p["rbsp"]) + 64, np.uint8)
        nbits, bad = C.c_int(0), C.c_int(0)
        rc = host_lib.cavlc_host_slice(rec.ctypes.data_as(C.c_void_p), len(rec), 11, bit0 % 8, out.ctypes.data_as(C.c_void_p), len(out), C.byref(nbits), C.byref(bad))
        assert rc == 0 and bad.value == 0
        nd = nbits.value - bit0 % 8
        ref_bits = bits_of(p["rbsp"])
        assert np.array_equal(bits_of(out)[bit0 % 8:bit0 % 8 + nd], ref_bits[bit0:bit0 + nd]), "picture %d" % n
        tail = ref_bits[bit0 + nd:]
        assert tail[0] == 1 and not tail[1:].any() and len(tail) <= 8
        skipped.append(int((r[:, 0] == 31).sum()))
    assert skipped and max(skipped) == 99
