"""ORACLE / TEST INFRASTRUCTURE: ctypes binding of the plain-C restatement oracle/fh264_oracle.c
(built into oracle/_ref/libfh264_oracle.so by `make -C oracle port`)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libfh264_oracle.so")
REC_INTS = 405
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lib = None


def build():
    subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "fh264_oracle.c")):
            build()
        L = C.CDLL(LIB)
        L.fo_create.restype = C.c_void_p
        L.fo_create.argtypes = [C.c_int, C.c_int]
        L.fo_destroy.argtypes = [C.c_void_p]
        L.fo_phase_r.argtypes = [C.c_void_p, _u8p]
        L.fo_encode_p.argtypes = [C.c_void_p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p]
        L.fo_tq_mb.argtypes = [_u8p, _u8p, C.c_int, _i32p, _u8p]
        L.fo_tq_luma_intra16.argtypes = [_u8p, _u8p, C.c_int, _i32p, _i32p, _u8p]
        L.fo_scene_sad.restype = C.c_uint64
        L.fo_scene_sad.argtypes = [_u8p, _u8p, C.c_size_t]
        L.fo_mc_picture.argtypes = [C.c_void_p, _u8p, _u8p, _u8p, _i32p, _u8p]
        for name, rt in (("fo_plane", C.POINTER(C.c_uint8)), ("fo_sorted", C.POINTER(C.c_int32)), ("fo_bucket_start", C.POINTER(C.c_int32))):
            getattr(L, name).restype = rt
        L.fo_plane.argtypes = [C.c_void_p, C.c_int]
        L.fo_kar.restype = C.POINTER(C.c_uint16)
        L.fo_kar.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.fo_sorted.argtypes = [C.c_void_p, C.c_int]
        L.fo_bucket_start.argtypes = [C.c_void_p]
        L.fo_ub_inputs.argtypes = [C.c_void_p]
        L.fo_stat.restype = C.c_longlong
        L.fo_stat.argtypes = [C.c_void_p, C.c_int]
        L.fo_set_trace.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _lib = L
    return _lib


class Trace(C.Structure):
    _fields_ = [("n", C.c_int * 3), ("cost", (C.c_int * 33) * 3), ("mvx", (C.c_int * 33) * 3), ("mvy", (C.c_int * 33) * 3),
                ("sad", (C.c_int * 33) * 3), ("mvpx", C.c_int), ("mvpy", C.c_int), ("s", C.c_int * 5), ("n_stage2_set", C.c_int)]


class Oracle:
    """Stateful CPU model of one sequence's P-picture path: phase_r(reference luma) then encode_p(...)."""

    def __init__(self, width, height):
        self.w, self.h = width, height
        self.L = lib()
        self.ctx = self.L.fo_create(width, height)

    def close(self):
        if self.ctx:
            self.L.fo_destroy(self.ctx)
            self.ctx = None

    __del__ = close

    def phase_r(self, ref_y):
        self.L.fo_phase_r(self.ctx, np.ascontiguousarray(ref_y, dtype=np.uint8))
        return bool(self.L.fo_ub_inputs(self.ctx))

    def plane(self, f):
        return np.ctypeslib.as_array(self.L.fo_plane(self.ctx, f), shape=(self.h, self.w)).copy()

    def kar(self, k, f):
        return np.ctypeslib.as_array(self.L.fo_kar(self.ctx, k, f), shape=(self.h, self.w)).copy()

    def sorted(self, a):
        return np.ctypeslib.as_array(self.L.fo_sorted(self.ctx, a), shape=(self.h * self.w,)).copy()

    def bucket_start(self):
        return np.ctypeslib.as_array(self.L.fo_bucket_start(self.ctx), shape=(16385,)).copy()

    def encode_p(self, cur, ref, qp, window, maxdiff, basic=0):
        """cur/ref: (Y, Cb, Cr). Returns (records[nmb,405] int32, (reconY, reconCb, reconCr))."""
        y, u, v = (np.array(p, dtype=np.uint8, order="C", copy=True) for p in cur)
        ry, ru, rv = (np.ascontiguousarray(p, dtype=np.uint8) for p in ref)
        rec = np.zeros(((self.w >> 4) * (self.h >> 4), REC_INTS), dtype=np.int32)
        self.L.fo_encode_p(self.ctx, y, u, v, ry, ru, rv, qp, window, maxdiff, basic, rec)
        return rec, (y, u, v)

    def mc_picture(self, ref, qmv):
        ry, ru, rv = (np.ascontiguousarray(p, dtype=np.uint8) for p in ref)
        nmb = (self.w >> 4) * (self.h >> 4)
        out = np.zeros((nmb, 384), dtype=np.uint8)
        self.L.fo_mc_picture(self.ctx, ry, ru, rv, np.ascontiguousarray(qmv, dtype=np.int32).reshape(nmb, 8), out)
        return out

    def stats(self):
        return int(self.L.fo_stat(self.ctx, 0)), int(self.L.fo_stat(self.ctx, 1))


def tq_mb(src384, pred384, qp):
    lv = np.zeros(384, dtype=np.int32)
    rc = np.zeros(384, dtype=np.uint8)
    lib().fo_tq_mb(np.ascontiguousarray(src384, dtype=np.uint8), np.ascontiguousarray(pred384, dtype=np.uint8), qp, lv, rc)
    return lv, rc


def tq_luma_intra16(src256, pred256, qp):
    dc = np.zeros(16, dtype=np.int32)
    ac = np.zeros(240, dtype=np.int32)
    rc = np.zeros(256, dtype=np.uint8)
    lib().fo_tq_luma_intra16(np.ascontiguousarray(src256, dtype=np.uint8), np.ascontiguousarray(pred256, dtype=np.uint8), qp, dc, ac, rc)
    return dc, ac.reshape(16, 15), rc


def scene_sad(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint8).ravel()
    b = np.ascontiguousarray(b, dtype=np.uint8).ravel()
    return int(lib().fo_scene_sad(a, b, a.size))
