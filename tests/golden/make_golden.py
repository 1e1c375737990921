"""Generates the committed golden vectors from the UNMODIFIED reference (oracle/_ref/ref_encoder, built from
/root/reference by `make -C oracle ref`). Run in the build container only:  python tests/golden/make_golden.py [intra]

Each fixture (npz) holds, for a short synthetic clip: the cropped source pictures, the reference's reconstruction
of every picture, picture types, and for every P picture the per-macroblock records
(mb_type, mv[4][2], mvd[4][2], sad[4], 384 levels) plus the .264 md5 — everything the GPU path must reproduce."""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from h264_fer_b200 import synth  # noqa: E402
from oracle import refdump  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (width, height, seed, frames, qp, window, maxdiff, basic, noise, square)
    "qcif_w16_qp28": (176, 144, 1, 4, 28, 16, 3, 0, 1.0, True),
    "cif_crop_w32_qp20_adaptive": (200, 120, 11, 4, 20, 32, -1, 0, 0.0, True),
    "small_basic_qp33": (128, 96, 12, 3, 33, 16, 2, 1, 0.5, False),
    # I-only, low contrast: the reference picks Intra16x16 for most MBs -> pins the luma-DC Hadamard path
    "intra16_lowcontrast_qp30": (176, 144, 13, 1, 30, 16, 3, 0, 0.3, False, 0.08),
    "intra16_lowcontrast_qp40": (176, 144, 14, 1, 40, 16, 3, 0, 0.3, False, 0.08),
}


# I-picture fixtures: clips with periodic IDR pictures; per I picture the reference's per-macroblock
# records (chunk IMBR), per P picture the usual records (they drive the session between the I pictures and decide which
# macroblocks were P_Skip before an I picture — the state the first bit-cost trial reads).
ICASES = {
    # name: (width, height, seed, frames, qp, window, maxdiff, intra_every, noise, square, contrast)
    "intra_qcif_qp28_ipi": (176, 144, 1, 3, 28, 16, 3, 2, 1.0, True, 1.0),
    "intra_lowcontrast_qp30_ipi": (176, 144, 6, 3, 30, 16, 3, 2, 0.3, True, 0.05),
    "intra_small_qp12_iii": (64, 48, 8, 3, 12, 16, 3, 1, 1.0, True, 0.3),
    "intra_crop_qp44_ipi": (200, 120, 9, 3, 44, 16, 3, 2, 1.0, True, 1.0),
    # static scene, fine quantiser: the P picture is all P_Skip, and the I picture after it shows the reference reading the previous
    # picture's mb_type_array in its first bit-cost trial (98 of 99 trial sizes and 2 decisions depend on it)
    "intra_static_qp12_ipi": (176, 144, 31, 3, 12, 16, 3, 2, 0.0, False, 1.0, (0, 0)),
}


def main_intra():
    assert refdump.have_ref_encoder(), "build oracle/_ref/ref_encoder first (make -C oracle ref)"
    for name, case in ICASES.items():
        w, h, seed, frames, qp, window, maxdiff, intra_every, noise, square, contrast = case[:11]
        tmp = tempfile.mkdtemp()
        y4m = os.path.join(tmp, "in.y4m")
        synth.write_y4m(y4m, w, h, seed, frames, noise=noise, square=square, contrast=contrast, pan=case[11] if len(case) > 11 else (2, 1))
        summ, dump, out264 = refdump.run_reference(y4m, frames, qp=qp, window=window, maxdiff=maxdiff, intra_every=intra_every,
                                                   dumpmask=refdump.D_MBREC | refdump.D_RECON | refdump.D_SOURCE | refdump.D_IMBREC | refdump.D_SLICE)
        pics = refdump.parse_dump(dump)
        arrays = dict(params=np.array([w, h, seed, frames, qp, window, maxdiff, intra_every], np.int32),
                      types=np.array([p["nal_type"] for p in pics], np.int32),
                      bitstream_md5=np.frombuffer(hashlib.md5(open(out264, "rb").read()).digest(), np.uint8))
        for n, p in enumerate(pics):
            for t in ("SRCY", "SRCU", "SRCV", "RECY", "RECU", "RECV"):
                arrays["%s_%d" % (t, n)] = p[t]
            if "imbrec" in p:
                arrays["imbrec_%d" % n] = p["imbrec"].astype(np.int16)
                arrays["slbit0_%d" % n] = np.array([p["slice_bit0"]], np.int32)     # first slice_data bit of the I slice's RBSP
                arrays["rbsp_%d" % n] = p["rbsp"]
            if "mbrec" in p:
                arrays["mbrec_%d" % n] = p["mbrec"].astype(np.int16)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(name, summ["types"], os.path.getsize(path), "bytes")


def main():
    assert refdump.have_ref_encoder(), "build oracle/_ref/ref_encoder first (make -C oracle ref)"
    for name, case in CASES.items():
        (w, h, seed, frames, qp, window, maxdiff, basic, noise, square), contrast = case[:10], (case[10] if len(case) > 10 else 1.0)
        tmp = tempfile.mkdtemp()
        y4m = os.path.join(tmp, "in.y4m")
        synth.write_y4m(y4m, w, h, seed, frames, noise=noise, square=square, contrast=contrast)
        summ, dump, out264 = refdump.run_reference(y4m, frames, qp=qp, basic=basic, window=window, maxdiff=maxdiff,
                                                   dumpmask=refdump.D_MBREC | refdump.D_RECON | refdump.D_SOURCE | refdump.D_TQIO | refdump.D_INTRA16 | refdump.D_SLICE | refdump.D_TABLES)
        pics = refdump.parse_dump(dump)
        arrays = dict(params=np.array([w, h, seed, frames, qp, window, maxdiff, basic], np.int32),
                      noise=np.array([noise]), square=np.array([int(square)]), contrast=np.array([contrast]), types=np.array([p["nal_type"] for p in pics], np.int32),
                      bitstream_md5=np.frombuffer(hashlib.md5(open(out264, "rb").read()).digest(), np.uint8),
                      y4m_md5=np.frombuffer(hashlib.md5(open(y4m, "rb").read()).digest(), np.uint8))
        for n, p in enumerate(pics):
            for t in ("SRCY", "SRCU", "SRCV", "RECY", "RECU", "RECV"):
                arrays["%s_%d" % (t, n)] = p[t]
            if "i16" in p:
                arrays["i16_%d" % n] = p["i16"][:96]          # Intra16x16 luma records of this I picture (a13)
            if "mbrec" in p:
                arrays["mbrec_%d" % n] = p["mbrec"].astype(np.int16)
                arrays["tqio_%d" % n] = p["tqio"]
                arrays["counts_%d" % n] = np.array(p["counts"], np.int32)
                arrays["slbit0_%d" % n] = np.array([p["slice_bit0"]], np.int32)     # first slice_data bit in the RBSP (after the slice header)
                arrays["rbsp_%d" % n] = p["rbsp"]                                   # slice RBSP (header + slice data + trailing bits)
        if name == "qcif_w16_qp28":
            np.savez_compressed(os.path.join(HERE, "cavlc_tables.npz"), tables=pics[0]["cavlc_tables"])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(name, summ["types"], os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "intra":      # only the I-picture fixtures
        main_intra()
    else:
        main()
        main_intra()
