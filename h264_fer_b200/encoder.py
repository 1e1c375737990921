"""Host-side mirror of the reference's per-picture driver for the P path (fer_h264.cpp:55-79 NastaviEncode,
ref_frames.cpp:185-234 selectNALUnitType, rbsp_encoding.cpp:139-323 RBSP_encode slice branch).

Intra (IDR) pictures: by default coded on the device as well (``fh264_encode_i``); with an
``intra_coder`` callback the reference's host code codes them, as in the integration build, and hands the
reconstruction to the device (``fh264_upload_recon``)."""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from .native import Session

NAL_NON_IDR, NAL_IDR = 1, 5


class SequenceEncoder:
    """Drives one sequence of a Session with the reference's frame-type rule and parameters
    (Starter::PostaviParametre, fer_h264.cpp:169-178)."""

    def __init__(self, session: Session, seq: int, qp=28, window=16, maxdiff_set=3, basic=0, intra_every=1000,
                 intra_coder: Optional[Callable] = None):
        self.s, self.seq = session, seq
        self.qp, self.window, self.maxdiff_set, self.basic, self.intra_every = qp, window, maxdiff_set, basic, intra_every
        self.intra_coder = intra_coder
        self.curr_frame_count = 0        # currFrameCount (fer_h264.cpp:188,196)
        self.have_dpb = False            # dpb.L != NULL (ref_frames.cpp:191)
        self.nmb = session.nmb

    def select_nal_unit_type(self) -> int:
        """selectNALUnitType (ref_frames.cpp:185-234); the source picture must already be uploaded."""
        if not self.have_dpb or self.curr_frame_count % self.intra_every == 0:
            return NAL_IDR
        sad = self.s.scene_sad(self.seq)
        return NAL_IDR if sad > (self.nmb << 12) else NAL_NON_IDR

    def encode_picture(self, y, cb, cr):
        """One picture. Returns (nal_unit_type, records): fh264_mb_result records for a P picture, fh264_mb_result_i records for
        an IDR picture coded on the device, None for an IDR picture coded by the intra_coder callback (which must return the
        reconstruction (Y, Cb, Cr) of the host-coded picture)."""
        self.s.upload_source(self.seq, y, cb, cr)
        nal = self.select_nal_unit_type()
        rec = None
        if nal == NAL_IDR:
            if self.intra_coder is None:
                rec = self.s.encode_i(self.qp, seq0=self.seq, nseq=1)[0]
            else:
                ry, rcb, rcr = self.intra_coder(y, cb, cr)
                self.s.upload_recon(self.seq, ry, rcb, rcr)
            self.have_dpb = True
        else:
            rec = self.s.encode_p(self.qp, self.window, self.maxdiff_set, self.basic, seq0=self.seq, nseq=1)[0]
        self.curr_frame_count += 1
        return nal, rec


class BatchEncoder:
    """All sequences of a Session in lockstep through the streaming step (fh264_upload_source_batch + fh264_encode_p_stream): what
    one reference process does per picture (selectNALUnitType, then RBSP_encode; fer_h264.cpp:55-79, ref_frames.cpp:185-234,
    rbsp_encoding.cpp:139-323) for a whole batch with one asynchronous step per picture.

    Per picture the host decides only what needs no pixels — first picture / ``currFrameCount % IntraEvery == 0`` (ref_frames.cpp:191)
    — and codes those sequences as IDR pictures (``fh264_encode_i`` + ``fh264_cavlc_i``); every other sequence goes through the
    step, whose scene gate (Σ|frame − dpb| > MBs << 12, :210-224) may still stop it: such a sequence is then coded as an IDR picture
    too, exactly as the reference would have. Returns per sequence (nal_unit_type, slice data bytes, bits)."""

    def __init__(self, session: Session, qp=28, window=16, maxdiff_set=3, basic=0, intra_every=1000, first_bit=0, slice_bytes=500000,
                 buffers=None):
        """buffers: (block, out) to use instead of pinned host memory of the CUDA library (CPU tests of the host logic)."""
        self.s = session
        self.n = session.batch
        self.qp, self.window, self.maxdiff_set, self.basic, self.intra_every, self.first_bit = qp, window, maxdiff_set, basic, intra_every, first_bit
        self.curr_frame_count = 0
        self.have_dpb = False
        self.pic_bytes = session.w * session.h * 3 // 2
        if buffers is None:
            from .native import PinnedArray, StreamOut
            self.block = PinnedArray((self.n, self.pic_bytes), np.uint8)
            self.out = StreamOut(self.n, session.nmb, slice_bytes=slice_bytes, mb_info=True, first_bit=first_bit)
        else:
            self.block, self.out = buffers

    def encode_pictures(self, pictures):
        """pictures: one (Y, Cb, Cr) per sequence. Returns [(nal_unit_type, bytes, nbits)] per sequence."""
        assert len(pictures) == self.n
        wh = self.s.w * self.s.h
        for b, (y, cb, cr) in enumerate(pictures):
            a = self.block.array[b]
            a[:wh] = np.asarray(y, np.uint8).ravel(); a[wh:wh + wh // 4] = np.asarray(cb, np.uint8).ravel(); a[wh + wh // 4:] = np.asarray(cr, np.uint8).ravel()
        self.s.upload_source_batch(self.block.ptr, self.pic_bytes)
        res = [None] * self.n
        host_idr = (not self.have_dpb) or self.curr_frame_count % self.intra_every == 0
        if host_idr:
            idr = list(range(self.n))
        else:
            self.s.encode_p_stream(self.qp, self.window, self.maxdiff_set, self.basic, scene_gate=1, out=self.out)
            self.s.sync()
            for b in range(self.n):
                self.s.picture_status(b)
            coded = self.out.coded()
            slices = self.out.slices()
            idr = [b for b in range(self.n) if not coded[b]]
            for b in range(self.n):
                if coded[b]:
                    res[b] = (NAL_NON_IDR,) + slices[b]
        for b in idr:                                        # (the source picture of a gated sequence is still current)
            self.s.encode_i(self.qp, seq0=b, nseq=1)
            data, nbits = self.s.cavlc_i(first_bit=self.first_bit, seq0=b, nseq=1)[0]
            res[b] = (NAL_IDR, data, nbits)
        self.have_dpb = True
        self.curr_frame_count += 1
        return res
