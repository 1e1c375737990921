"""h264_fer_b200 — B200 (sm_100a) implementation of the P-picture hot path of zoltanmaric/h264-fer.

Python is only the launcher: this module binds the C ABI of ``include/fh264_b200.h`` with ctypes and mirrors the
reference's per-picture call sequence (``selectNALUnitType`` -> ``RBSP_encode``: ``interEncoding`` +
``quantizationTransform`` per MB -> ``modificationProcess`` -> ``FillInterpolatedRefFrame``;
reference rbsp_encoding.cpp:119-326, fer_h264.cpp:55-79). There is no CPU fallback: if the CUDA library or a B200
is missing, every call raises.
"""
from .native import (Fh264Error, Session, MB_RESULT_DTYPE, CAVLC_MB_INFO_DTYPE, P_SKIP, P_8x8ref0, P_L0_16x16, P_L0_L0_16x8, P_L0_L0_8x16,
                     lib_path, load_library, records_to_ints, MB_RESULT_I_DTYPE, i_records_to_ints)
from .encoder import SequenceEncoder, BatchEncoder, NAL_IDR, NAL_NON_IDR

__all__ = ["Fh264Error", "Session", "SequenceEncoder", "BatchEncoder", "MB_RESULT_DTYPE", "CAVLC_MB_INFO_DTYPE", "lib_path", "load_library", "records_to_ints", "MB_RESULT_I_DTYPE", "i_records_to_ints",
           "P_SKIP", "P_8x8ref0", "P_L0_16x16", "P_L0_L0_16x8", "P_L0_L0_8x16", "NAL_IDR", "NAL_NON_IDR"]
