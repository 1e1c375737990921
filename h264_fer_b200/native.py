"""ctypes binding of include/fh264_b200.h (the drop-in C ABI). No torch types cross this boundary."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
P_L0_16x16, P_L0_L0_16x8, P_L0_L0_8x16, P_8x8ref0, P_SKIP = 0, 1, 2, 4, 31

# struct fh264_mb_result (832 bytes)
MB_RESULT_DTYPE = np.dtype([("mb_type", "<i2"), ("num_parts", "<i2"), ("mv", "<i2", (4, 2)), ("mvd", "<i2", (4, 2)),
                            ("sad", "<u2", (4,)), ("luma", "<i2", (16, 16)), ("chroma_dc", "<i2", (2, 4)),
                            ("chroma_ac", "<i2", (2, 4, 15)), ("reserved", "<i2", (10,))])
assert MB_RESULT_DTYPE.itemsize == 832
# struct fh264_mb_result_i (832 bytes): one macroblock of an I picture
MB_RESULT_I_DTYPE = np.dtype([("mb_type", "<i2"), ("intra16x16_pred_mode", "i1"), ("intra_chroma_pred_mode", "u1"), ("cbp_luma", "u1"),
                              ("cbp_chroma", "u1"), ("bits_intra16x16", "<u2"), ("bits_intra4x4", "<u2"), ("intra4x4_pred_mode", "u1", (16,)),
                              ("prev_intra4x4_pred_mode_flag", "u1", (16,)), ("rem_intra4x4_pred_mode", "u1", (16,)), ("luma", "<i2", (16, 16)),
                              ("chroma_dc", "<i2", (2, 4)), ("chroma_ac", "<i2", (2, 4, 15)), ("reserved", "<i2", (3,))])
assert MB_RESULT_I_DTYPE.itemsize == 832
CAVLC_MB_INFO_DTYPE = np.dtype([("skip", "u1"), ("cbp_luma", "u1"), ("cbp_chroma", "u1"), ("mb_type", "u1"), ("total_coeff_luma", "u1", (16,)),
                                ("total_coeff_chroma", "u1", (2, 4)), ("reserved", "u1", (4,))])
assert CAVLC_MB_INFO_DTYPE.itemsize == 32

ERRORS = {-1: "FH264_E_ARG", -2: "FH264_E_CUDA", -3: "FH264_E_NO_DEVICE", -4: "FH264_E_STATE", -5: "FH264_E_UB_INPUT",
          -6: "FH264_E_CAPACITY", -7: "FH264_E_UNSUPPORTED"}

EXPORTS = ["fh264_open", "fh264_close", "fh264_last_error", "fh264_abi_version", "fh264_set_stream", "fh264_sync",
           "fh264_host_alloc", "fh264_host_free", "fh264_upload_source", "fh264_upload_source_frame", "fh264_upload_source_device", "fh264_upload_recon", "fh264_scene_sad", "fh264_scene_sad_batch",
           "fh264_encode_p", "fh264_encode_p_async", "fh264_picture_status", "fh264_download_recon", "fh264_mode_counts",
           "fh264_tq_macroblocks", "fh264_tq_luma_intra16", "fh264_motion_compensate", "fh264_debug_plane",
           "fh264_debug_feature", "fh264_cavlc_p", "fh264_decode_p", "fh264_encode_i", "fh264_last_intra_ms", "fh264_cavlc_i", "fh264_last_timings", "fh264_last_spec_ms", "fh264_measure_int_peak", "fh264_debug_timeline", "fh264_debug_status", "fh264_band_config", "fh264_ipc_export", "fh264_ipc_import",
           "fh264_encode_p_stream", "fh264_set_pipeline", "fh264_upload_source_batch", "fh264_debug_trace", "fh264_band_peers", "fh264_band_gather"]
IPC_BLOB_BYTES = 11 * 64        # FH264_IPC_HANDLES * FH264_IPC_HANDLE_BYTES
STATUS_WORDS, ST_SAD_LO, ST_SAD_HI, ST_GATE, ST_GATED_TOTAL = 24, 8, 9, 17, 18


class Fh264Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "FH264_E_?"), code, msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [("qp", C.c_int32), ("window", C.c_int32), ("maxdiff_set", C.c_int32), ("basic", C.c_int32)]


class StreamOutStruct(C.Structure):        # struct fh264_stream_out
    _fields_ = [("records", C.c_void_p), ("slice", C.c_void_p), ("slice_stride", C.c_size_t), ("slice_copy_bytes", C.c_size_t),
                ("first_bit", C.c_int), ("slice_stat", C.c_void_p), ("mb_info", C.c_void_p), ("status", C.c_void_p)]


def lib_path() -> str:
    # FH264_B200_LIB: development only — an alternative BUILD of this same library (A/B kernel experiments)
    return os.environ.get("FH264_B200_LIB") or os.path.join(HERE, "libfh264_b200.so")


_lib = None


def load_library():
    """Loads the in-tree CUDA library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.isfile(p):
        raise Fh264Error(-3, "CUDA library %s is missing: run `python -m h264_fer_b200.build` (no CPU fallback exists)" % p)
    L = C.CDLL(p)
    vp, i32, u8p = C.c_void_p, C.c_int, C.c_void_p
    L.fh264_open.argtypes = [i32, i32, i32, i32, C.POINTER(vp)]
    L.fh264_close.argtypes = [vp]
    L.fh264_last_error.restype = C.c_char_p
    L.fh264_set_stream.argtypes = [vp, vp]
    L.fh264_sync.argtypes = [vp]
    L.fh264_host_alloc.restype = vp
    L.fh264_host_alloc.argtypes = [C.c_size_t]
    L.fh264_host_free.argtypes = [vp]
    L.fh264_upload_source.argtypes = [vp, i32, u8p, u8p, u8p]
    L.fh264_upload_source_frame.argtypes = [vp, i32, u8p, i32, i32]
    L.fh264_upload_recon.argtypes = [vp, i32, u8p, u8p, u8p]
    L.fh264_scene_sad.argtypes = [vp, i32, C.POINTER(C.c_uint64)]
    L.fh264_scene_sad_batch.argtypes = [vp, i32, i32, C.POINTER(C.c_uint64)]
    L.fh264_upload_source_device.argtypes = [vp, i32, vp, vp, vp]
    L.fh264_encode_p.argtypes = [vp, i32, i32, C.POINTER(Params), vp]
    L.fh264_encode_p_async.argtypes = [vp, i32, i32, C.POINTER(Params), vp]
    L.fh264_picture_status.argtypes = [vp, i32]
    L.fh264_download_recon.argtypes = [vp, i32, u8p, u8p, u8p]
    L.fh264_mode_counts.argtypes = [vp, i32, C.POINTER(C.c_int32)]
    L.fh264_tq_macroblocks.argtypes = [vp, i32, u8p, u8p, i32, vp, u8p]
    L.fh264_tq_luma_intra16.argtypes = [vp, i32, u8p, u8p, i32, vp, vp, u8p]
    L.fh264_motion_compensate.argtypes = [vp, i32, vp, u8p]
    L.fh264_debug_plane.argtypes = [vp, i32, i32, u8p]
    L.fh264_debug_feature.argtypes = [vp, i32, i32, i32, vp]
    L.fh264_last_timings.argtypes = [vp, C.POINTER(C.c_float)]
    L.fh264_last_spec_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.fh264_measure_int_peak.argtypes = [i32, C.POINTER(C.c_double)]
    L.fh264_debug_timeline.argtypes = [vp, i32, vp]
    L.fh264_cavlc_p.argtypes = [vp, i32, i32, i32, vp, C.c_size_t, vp, vp]
    L.fh264_decode_p.argtypes = [vp, i32, i32, i32, vp]
    L.fh264_encode_i.argtypes = [vp, i32, i32, i32, vp]
    L.fh264_last_intra_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.fh264_cavlc_i.argtypes = [vp, i32, i32, i32, vp, C.c_size_t, vp]
    L.fh264_debug_status.argtypes = [vp, i32, vp]
    L.fh264_band_config.argtypes = [vp, i32, i32, i32, i32]
    L.fh264_ipc_export.argtypes = [vp, i32, vp]
    L.fh264_ipc_import.argtypes = [vp, i32, i32, vp]
    L.fh264_encode_p_stream.argtypes = [vp, i32, i32, C.POINTER(Params), i32, C.POINTER(StreamOutStruct)]
    L.fh264_set_pipeline.argtypes = [vp, i32]
    L.fh264_debug_trace.argtypes = [vp, vp]
    L.fh264_band_peers.argtypes = [vp, i32, C.POINTER(C.c_int)]
    L.fh264_band_gather.argtypes = [vp, i32]
    L.fh264_upload_source_batch.argtypes = [vp, i32, i32, vp, C.c_size_t, i32]
    for name in EXPORTS:
        getattr(L, name)
    _lib = L
    return L


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint8)


def measure_int_peak(device=0):
    """Sustained integer-pipe rate of the GPU in T lane-statements/s: dict of IMAD, VIADDMNMX.S16x2, VIADDMNMX, VABSDIFF4+VIADD."""
    L = load_library()
    t = (C.c_double * 4)()
    rc = L.fh264_measure_int_peak(int(device), t)
    if rc:
        raise Fh264Error(rc, (L.fh264_last_error() or b"").decode())
    return {"imad": t[0], "viaddmnmx_s16x2": t[1], "viaddmnmx": t[2], "vabsdiff4_acc_plus_viadd": t[3]}


class PinnedArray:
    """numpy view over cudaHostAlloc memory (fh264_host_alloc)."""

    def __init__(self, shape, dtype):
        L = load_library()
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(shape)) * self.dtype.itemsize
        self.ptr = L.fh264_host_alloc(max(self.nbytes, 16))
        if not self.ptr:
            raise Fh264Error(-2, "cudaHostAlloc failed")
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(shape)

    def free(self):
        if self.ptr:
            load_library().fh264_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class StreamOut:
    """Pinned host buffers one fh264_encode_p_stream step writes (records / device-CAVLC slice data / side information / status)."""

    def __init__(self, nseq, nmb, records=False, slice_bytes=0, mb_info=False, status=True, first_bit=0):
        self.nseq, self.nmb = nseq, nmb
        self.records = PinnedArray((nseq, nmb), MB_RESULT_DTYPE) if records else None
        self.slice = PinnedArray((nseq, slice_bytes), np.uint8) if slice_bytes else None
        self.slice_stat = PinnedArray((nseq, 2), np.uint32) if slice_bytes else None
        self.mb_info = PinnedArray((nseq, nmb), CAVLC_MB_INFO_DTYPE) if (slice_bytes and mb_info) else None
        self.status = PinnedArray((nseq, STATUS_WORDS), np.uint32) if status else None
        g = lambda a: a.ptr if a is not None else None
        self.struct = StreamOutStruct(g(self.records), g(self.slice), slice_bytes, slice_bytes, first_bit, g(self.slice_stat), g(self.mb_info), g(self.status))

    def coded(self):
        """Per sequence: True = coded as a P picture, False = stopped by the scene gate (code it with encode_i)."""
        return [int(self.status.array[b, ST_GATE]) == 0 for b in range(self.nseq)]

    def scene_sad(self):
        st = self.status.array
        return [int(st[b, ST_SAD_LO]) | (int(st[b, ST_SAD_HI]) << 32) for b in range(self.nseq)]

    def slices(self):
        """[(bytes, nbits)] per sequence (only what slice_copy_bytes brought home)."""
        res = []
        for b in range(self.nseq):
            nb = int(self.slice_stat.array[b, 1])
            res.append((self.slice.array[b, :min((nb + 7) // 8, self.slice.array.shape[1])].copy(), nb))
        return res


class Session:
    """One GPU, ``batch`` sequences in lockstep. Thin 1:1 wrapper over the C ABI."""

    def __init__(self, width, height, batch=1, device=0):
        self.L = load_library()
        self.w, self.h, self.batch, self.device = int(width), int(height), int(batch), int(device)
        self.nmb = (self.w >> 4) * (self.h >> 4)
        h = C.c_void_p()
        self._ck(self.L.fh264_open(self.w, self.h, self.batch, self.device, C.byref(h)))
        self.handle = h

    def _ck(self, rc):
        if rc != 0:
            raise Fh264Error(rc, (self.L.fh264_last_error() or b"").decode())

    def close(self):
        if getattr(self, "handle", None):
            self.L.fh264_close(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- stream / sync
    def set_stream(self, cuda_stream_ptr):
        self._ck(self.L.fh264_set_stream(self.handle, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        self._ck(self.L.fh264_sync(self.handle))

    # -- pictures
    def upload_source(self, seq, y, cb, cr):
        y, cb, cr = _u8(y), _u8(cb), _u8(cr)
        assert y.size == self.w * self.h and cb.size == y.size // 4 and cr.size == y.size // 4
        self._keep = (y, cb, cr)
        self._ck(self.L.fh264_upload_source(self.handle, seq, _ptr(y), _ptr(cb), _ptr(cr)))

    def upload_source_frame(self, seq, frame420, in_w, in_h):
        """One raw Y4M FRAME payload (Y, Cb, Cr of the INPUT size); the reference's centre crop happens in the H2D copy."""
        f = _u8(frame420)
        assert f.size >= in_w * in_h * 3 // 2
        self._keep = (f,)
        self._ck(self.L.fh264_upload_source_frame(self.handle, seq, _ptr(f), in_w, in_h))

    def upload_source_ptrs(self, seq, py, pcb, pcr, device=False):
        """Raw-pointer variant (pinned host memory, or device memory with device=True); asynchronous."""
        fn = self.L.fh264_upload_source_device if device else self.L.fh264_upload_source
        self._ck(fn(self.handle, seq, C.c_void_p(py), C.c_void_p(pcb), C.c_void_p(pcr)))

    def scene_sad_batch(self, seq0=0, nseq=None):
        nseq = self.batch - seq0 if nseq is None else nseq
        v = (C.c_uint64 * nseq)()
        self._ck(self.L.fh264_scene_sad_batch(self.handle, seq0, nseq, v))
        return [int(x) for x in v]

    def upload_recon(self, seq, y, cb, cr):
        y, cb, cr = _u8(y), _u8(cb), _u8(cr)
        assert y.size == self.w * self.h and cb.size == y.size // 4 and cr.size == y.size // 4
        self._ck(self.L.fh264_upload_recon(self.handle, seq, _ptr(y), _ptr(cb), _ptr(cr)))
        self.sync()

    def scene_sad(self, seq) -> int:
        v = C.c_uint64()
        self._ck(self.L.fh264_scene_sad(self.handle, seq, C.byref(v)))
        return int(v.value)

    def encode_p(self, qp, window, maxdiff_set, basic=0, seq0=0, nseq=None, out=None, sync=True, download=True):
        """Returns a structured array [nseq, nmb] of MB_RESULT_DTYPE (a view of `out` if given); with download=False the
        records stay on the device (NULL results pointer) and None is returned."""
        nseq = self.batch - seq0 if nseq is None else nseq
        prm = Params(int(qp), int(window), int(maxdiff_set), int(basic))
        if download and out is None:
            out = np.zeros((nseq, self.nmb), dtype=MB_RESULT_DTYPE)
        fn = self.L.fh264_encode_p if sync else self.L.fh264_encode_p_async
        self._ck(fn(self.handle, seq0, nseq, C.byref(prm), _ptr(out) if download else None))
        return out if download else None

    def set_pipeline(self, on):
        self._ck(self.L.fh264_set_pipeline(self.handle, 1 if on else 0))

    def upload_source_batch(self, ptr, stride, seq0=0, nseq=None, device=False):
        """Source pictures of sequences [seq0, seq0 + nseq) from one block (Y | Cb | Cr per sequence, `stride` bytes apart): raw
        pointer to pinned host memory, or to device memory with device=True; asynchronous."""
        nseq = self.batch - seq0 if nseq is None else nseq
        self._ck(self.L.fh264_upload_source_batch(self.handle, seq0, nseq, C.c_void_p(ptr), stride, 1 if device else 0))

    def encode_p_stream(self, qp, window, maxdiff_set, basic=0, seq0=0, nseq=None, scene_gate=True, out=None):
        """One step without a host round trip (fh264_encode_p_stream). `out`: a StreamOut (pinned buffers) or None; everything it
        receives is valid after sync()."""
        nseq = self.batch - seq0 if nseq is None else nseq
        prm = Params(int(qp), int(window), int(maxdiff_set), int(basic))
        self._ck(self.L.fh264_encode_p_stream(self.handle, seq0, nseq, C.byref(prm), int(scene_gate), C.byref(out.struct) if out is not None else None))

    def picture_status(self, seq):
        self._ck(self.L.fh264_picture_status(self.handle, seq))

    def download_recon(self, seq):
        y = np.zeros((self.h, self.w), np.uint8)
        cb = np.zeros((self.h // 2, self.w // 2), np.uint8)
        cr = np.zeros((self.h // 2, self.w // 2), np.uint8)
        self._ck(self.L.fh264_download_recon(self.handle, seq, _ptr(y), _ptr(cb), _ptr(cr)))
        return y, cb, cr

    def mode_counts(self, seq):
        c = (C.c_int32 * 5)()
        self._ck(self.L.fh264_mode_counts(self.handle, seq, c))
        return list(c)

    def last_timings(self):
        t = (C.c_float * 10)()
        self._ck(self.L.fh264_last_timings(self.handle, t))
        names = ("phase_a_ms", "phase_b_ms", "phase_c_ms", "copy_phase_r_ms", "total_ms", "k_stage3_ms", "k_stage2_ms",
                 "k_interp_ms", "k_features_ms", "k_tile_index_ms")
        d = dict(zip(names, [float(x) for x in t]))
        ms = C.c_float(0)
        self._ck(self.L.fh264_last_spec_ms(self.handle, C.byref(ms)))
        d["k_spec_ms"] = float(ms.value)
        return d

    # -- building blocks
    def tq_macroblocks(self, src384, pred384, qp):
        src384, pred384 = _u8(src384).reshape(-1, 384), _u8(pred384).reshape(-1, 384)
        n = src384.shape[0]
        lv = np.zeros((n, 384), np.int16)
        rc = np.zeros((n, 384), np.uint8)
        self._ck(self.L.fh264_tq_macroblocks(self.handle, n, _ptr(src384), _ptr(pred384), int(qp), _ptr(lv), _ptr(rc)))
        return lv, rc

    def tq_luma_intra16(self, src256, pred256, qp):
        src256, pred256 = _u8(src256).reshape(-1, 256), _u8(pred256).reshape(-1, 256)
        n = src256.shape[0]
        dc = np.zeros((n, 16), np.int16)
        ac = np.zeros((n, 16, 15), np.int16)
        rc = np.zeros((n, 256), np.uint8)
        self._ck(self.L.fh264_tq_luma_intra16(self.handle, n, _ptr(src256), _ptr(pred256), int(qp), _ptr(dc), _ptr(ac), _ptr(rc)))
        return dc, ac, rc

    def motion_compensate(self, seq, qmv):
        qmv = np.ascontiguousarray(qmv, dtype=np.int16).reshape(self.nmb, 4, 2)
        out = np.zeros((self.nmb, 384), np.uint8)
        self._ck(self.L.fh264_motion_compensate(self.handle, seq, _ptr(qmv), _ptr(out)))
        return out

    def decode_p(self, records, qp, seq0=0):
        """Decoder inverse path: reconstruct the P picture(s) described by records [nseq, nmb] (MB_RESULT_DTYPE) from the current
        reference picture; the result becomes the reference picture (read it with download_recon)."""
        r = np.ascontiguousarray(records, dtype=MB_RESULT_DTYPE).reshape(-1, self.nmb)
        self._ck(self.L.fh264_decode_p(self.handle, seq0, r.shape[0], qp, _ptr(r)))

    def encode_i(self, qp, seq0=0, nseq=None):
        """Code the current source picture(s) as I pictures on the device (intraPredictionEncoding + quantizationTransform of every
        macroblock, rbsp_encoding.cpp:196-215); the reconstruction becomes the reference picture. Returns [nseq, nmb] records
        (MB_RESULT_I_DTYPE)."""
        nseq = self.batch - seq0 if nseq is None else nseq
        out = np.zeros((nseq, self.nmb), dtype=MB_RESULT_I_DTYPE)
        self._ck(self.L.fh264_encode_i(self.handle, seq0, nseq, qp, _ptr(out)))
        return out

    def last_intra_ms(self):
        """Device time of the intra wavefront kernel of the last encode_i call (ms)."""
        ms = C.c_float(0)
        self._ck(self.L.fh264_last_intra_ms(self.handle, C.byref(ms)))
        return float(ms.value)

    def cavlc_p(self, first_bit=0, seq0=0, nseq=None, capacity=500064, mb_info=False):
        """Device CAVLC of the P picture(s) last coded by encode_p: list of (bytes, nbits) per sequence; slice data occupies bits
        [first_bit, nbits) of the returned bytes (rbsp_encoding.cpp:175-313)."""
        nseq = self.batch - seq0 if nseq is None else nseq
        out = np.zeros((nseq, capacity), np.uint8)
        nbits = np.zeros(nseq, np.uint32)
        info = np.zeros((nseq, self.nmb), CAVLC_MB_INFO_DTYPE) if mb_info else None
        self._ck(self.L.fh264_cavlc_p(self.handle, seq0, nseq, first_bit, _ptr(out), capacity, _ptr(nbits), _ptr(info) if mb_info else None))
        res = [(out[b, :(int(nbits[b]) + 7) // 8].copy(), int(nbits[b])) for b in range(nseq)]
        return (res, info) if mb_info else res

    def cavlc_i(self, first_bit=0, seq0=0, nseq=None, capacity=500064):
        """Device CAVLC of the I picture(s) last coded by encode_i: list of (bytes, nbits) per sequence, like cavlc_p
        (rbsp_encoding.cpp:221-305)."""
        nseq = self.batch - seq0 if nseq is None else nseq
        out = np.zeros((nseq, capacity), np.uint8)
        nbits = np.zeros(nseq, np.uint32)
        self._ck(self.L.fh264_cavlc_i(self.handle, seq0, nseq, first_bit, _ptr(out), capacity, _ptr(nbits)))
        return [(out[b, :(int(nbits[b]) + 7) // 8].copy(), int(nbits[b])) for b in range(nseq)]

    def debug_timeline(self, seq, read=True):
        if not read:
            self._ck(self.L.fh264_debug_timeline(self.handle, seq, None))
            return None
        out = np.zeros((self.nmb, 24), np.int64)
        self._ck(self.L.fh264_debug_timeline(self.handle, seq, _ptr(out)))
        return out

    # -- band mode
    def band_config(self, rank, world, mb_row0, mb_row1):
        self._ck(self.L.fh264_band_config(self.handle, rank, world, mb_row0, mb_row1))
        self.band = (mb_row0, mb_row1)

    def band_peers(self, bands):
        """bands: [(first_mb_row, end_mb_row)] of every rank (fh264_band_peers): halo-restricted phase R and picture barrier."""
        flat = (C.c_int * (2 * len(bands)))(*[v for b in bands for v in b])
        self._ck(self.L.fh264_band_peers(self.handle, len(bands), flat))

    def band_gather(self, on=True):
        self._ck(self.L.fh264_band_gather(self.handle, 1 if on else 0))

    def ipc_export(self, seq=0) -> bytes:
        buf = np.zeros(IPC_BLOB_BYTES, np.uint8)
        self._ck(self.L.fh264_ipc_export(self.handle, seq, _ptr(buf)))
        return buf.tobytes()

    def ipc_import(self, seq, peer_rank, blob: bytes):
        buf = np.frombuffer(blob, np.uint8).copy()
        assert buf.size == IPC_BLOB_BYTES
        self._ck(self.L.fh264_ipc_import(self.handle, seq, peer_rank, _ptr(buf)))

    def debug_trace(self):
        out = np.zeros((3, 8, 5), np.float32)
        self._ck(self.L.fh264_debug_trace(self.handle, _ptr(out)))
        return out

    def debug_status(self, seq):
        out = np.zeros(16, np.uint32)
        self._ck(self.L.fh264_debug_status(self.handle, seq, _ptr(out)))
        return out

    def debug_plane(self, seq, f):
        out = np.zeros((self.h, self.w), np.uint8)
        self._ck(self.L.fh264_debug_plane(self.handle, seq, f, _ptr(out)))
        return out

    def debug_feature(self, seq, k, f):
        out = np.zeros((self.h, self.w), np.uint16)
        self._ck(self.L.fh264_debug_feature(self.handle, seq, k, f, _ptr(out)))
        return out


def i_records_to_ints(rec: np.ndarray) -> np.ndarray:
    """[nmb] MB_RESULT_I_DTYPE -> [nmb, 439] int32 in the layout of the reference dump's I-picture records (chunk IMBR of
    oracle/ref_harness/driver.cpp): mb_type, Intra16x16PredMode, intra_chroma_pred_mode, bits of the two trials, CBP luma / chroma,
    Intra4x4PredMode[16], prev flags[16], rem modes[16], 256 luma levels, cdc[2][4], cac[2][4][15]."""
    n = rec.shape[0]
    out = np.zeros((n, 439), np.int32)
    out[:, 0] = rec["mb_type"]
    out[:, 1] = rec["intra16x16_pred_mode"]
    out[:, 2] = rec["intra_chroma_pred_mode"]
    out[:, 3] = rec["bits_intra16x16"]
    out[:, 4] = rec["bits_intra4x4"]
    out[:, 5] = rec["cbp_luma"]
    out[:, 6] = rec["cbp_chroma"]
    out[:, 7:23] = rec["intra4x4_pred_mode"]
    out[:, 23:39] = rec["prev_intra4x4_pred_mode_flag"]
    out[:, 39:55] = rec["rem_intra4x4_pred_mode"]
    out[:, 55:311] = rec["luma"].reshape(n, 256)
    out[:, 311:319] = rec["chroma_dc"].reshape(n, 8)
    out[:, 319:439] = rec["chroma_ac"].reshape(n, 120)
    return out


def records_to_ints(rec: np.ndarray) -> np.ndarray:
    """[nmb] MB_RESULT_DTYPE -> [nmb, 405] int32 in the layout of the oracle / reference dump records:
    mb_type, mv[4][2], mvd[4][2], sad[4], luma[16][16], cdc[2][4], cac[2][4][15]."""
    n = rec.shape[0]
    out = np.zeros((n, 405), np.int32)
    out[:, 0] = rec["mb_type"]
    out[:, 1:9] = rec["mv"].reshape(n, 8)
    out[:, 9:17] = rec["mvd"].reshape(n, 8)
    out[:, 17:21] = rec["sad"]
    out[:, 21:277] = rec["luma"].reshape(n, 256)
    out[:, 277:285] = rec["chroma_dc"].reshape(n, 8)
    out[:, 285:405] = rec["chroma_ac"].reshape(n, 120)
    return out
