"""ctypes binding of the C-ABI shared library (``include/xmaps_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` (``x-maps_b200/csrc/build.py``) as
``x-maps_b200/libxmaps_b200.so``.  There is no fallback: if the library is missing or a symbol
the header declares cannot be resolved, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# XMAPS_B200_LIB: an alternative build of the same library (A/B runs of kernel variants); default = the in-tree build
LIB_PATH = os.environ.get("XMAPS_B200_LIB") or os.path.join(_HERE, "libxmaps_b200.so")

# ---- constants of include/xmaps_b200.h --------------------------------------------------------
ABI_VERSION = 1
OK, ERR_INVALID_ARG, ERR_CUDA, ERR_NO_XMAP, ERR_TABLE_RANGE, ERR_UNSUPPORTED = range(6)
VIEW_PROJECTOR, VIEW_CAMERA = 0, 1
OUT_DEPTH, OUT_DISPARITY, OUT_BGR = 0, 1, 2
TBOUNDS_REDUCE, TBOUNDS_SORTED, TBOUNDS_GIVEN = 0, 1, 2
FLAG_POLARITY, FLAG_TIME_F64, FLAG_BILINEAR = 0x1, 0x2, 0x4
STATUS_TBOUNDS_VIOLATED, STATUS_PIXEL_OOB, STATUS_SCATTER_OOB = 0x1, 0x2, 0x4
STATUS_FILTER_POLARITY, STATUS_FILTER_INDEX = 0x8, 0x10
FILTER_FIRST_YT, FILTER_FIRST_XY, FILTER_LAST_XY, FILTER_MEAN_XY = 1, 2, 3, 4


class XmTables(C.Structure):
    _fields_ = [
        ("cam_w", C.c_int32),
        ("cam_h", C.c_int32),
        ("rect_w", C.c_int32),
        ("rect_h", C.c_int32),
        ("proj_w", C.c_int32),
        ("proj_h", C.c_int32),
        ("xmap_w", C.c_int32),
        ("t_px_scale", C.c_int32),
        ("x_offset", C.c_int32),
        ("dilate", C.c_int32),
        ("depth_scale", C.c_double),
        ("lut_x", C.c_void_p),
        ("lut_y", C.c_void_p),
        ("x_map", C.c_void_p),
        ("remap_xy", C.c_void_p),
        ("lut_x_f32", C.c_void_p),
        ("lut_y_f32", C.c_void_p),
    ]


class XmFrameArgs(C.Structure):
    _fields_ = [
        ("d_events", C.c_void_p),
        ("n_events", C.c_int64),
        ("flags", C.c_uint32),
        ("view", C.c_int32),
        ("time_bounds", C.c_int32),
        ("output", C.c_int32),
        ("t_min", C.c_int64),
        ("t_max", C.c_int64),
        ("d_out", C.c_void_p),
        ("z_near", C.c_float),
        ("z_far", C.c_float),
    ]


class XmFrameStatus(C.Structure):
    _fields_ = [
        ("n_events", C.c_int64),
        ("n_valid", C.c_int64),
        ("n_inliers", C.c_int64),
        ("t_min", C.c_int64),
        ("t_max", C.c_int64),
        ("flags", C.c_uint32),
        ("epoch", C.c_uint32),
        ("fixup_ran", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class XmIpcHandle(C.Structure):
    _fields_ = [("bytes", C.c_ubyte * 64)]


_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32

# every symbol the header declares: name -> (restype, argtypes)
SIGNATURES = {
    "xm_abi_version": (C.c_int, []),
    "xm_last_error": (C.c_char_p, []),
    "xm_launch_count": (_I64, []),
    "xm_ctx_create": (C.c_int, [C.POINTER(XmTables), C.c_int, C.POINTER(_P)]),
    "xm_ctx_destroy": (C.c_int, [_P]),
    "xm_ctx_set_xmap": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32]),
    "xm_ctx_set_colormap": (C.c_int, [_P, _P]),
    "xm_ctx_set_option": (C.c_int, [_P, C.c_char_p, _I64]),
    "xm_ctx_get_option": (C.c_int, [_P, C.c_char_p, C.POINTER(_I64)]),
    "xm_frame": (C.c_int, [_P, C.POINTER(XmFrameArgs), _P]),
    "xm_frame_batch": (C.c_int, [_P, C.POINTER(XmFrameArgs), _I32, _P]),
    "xm_frame_status": (C.c_int, [_P, C.POINTER(XmFrameStatus), _P]),
    "xm_frame_host": (C.c_int, [_P, C.POINTER(XmFrameArgs), _P, _P, C.POINTER(XmFrameStatus), _P]),
    "xm_host_alloc": (C.c_int, [C.POINTER(_P), _I64]),
    "xm_host_free": (C.c_int, [_P]),
    "xm_rectify_i16": (C.c_int, [_P, _P, _I64, _P, _P, _P]),
    "xm_rectify_f32": (C.c_int, [_P, _P, _I64, _P, _P, _P]),
    "xm_event_disparity": (C.c_int, [_P, C.POINTER(XmFrameArgs), _P, _P, _P, _P, _P]),
    "xm_compact_i16": (C.c_int, [_P, _P, _P, _I64, _P, _P, _P]),
    "xm_scatter_last_wins": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _I32, _P, _P]),
    "xm_dilate_remap": (C.c_int, [_P, _P, _P, _P]),
    "xm_disp_to_depth": (C.c_int, [_P, _P, _I64, C.c_double, _P, _P]),
    "xm_colorize": (C.c_int, [_P, _P, _I64, C.c_double, C.c_float, C.c_float, _P, _P]),
    "xm_point_cloud": (C.c_int, [_P, _P, _P, _P, _I64, _P, _P, _P]),
    "xm_polarity_filter": (C.c_int, [_P, _P, _I64, _P, _P, _P]),
    "xm_activity_filter": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _P]),
    "xm_activity_reset": (C.c_int, [_P, _P]),
    "xm_filter_events": (C.c_int, [_P, _P, _I64, _I32, _P, _I32, _P, _P, _P]),
    "xm_find_trigger": (C.c_int, [_P, _P, _I64, _I64, C.c_double, _I64, _P, _P]),
    "xm_peer_alloc": (C.c_int, [C.c_int, _I64, C.POINTER(_P), C.POINTER(XmIpcHandle)]),
    "xm_peer_free": (C.c_int, [C.c_int, _P]),
    "xm_peer_open": (C.c_int, [C.c_int, C.c_int, C.POINTER(XmIpcHandle), C.POINTER(_P)]),
    "xm_peer_close": (C.c_int, [C.c_int, _P]),
    "xm_peer_copy": (C.c_int, [_P, C.c_int, _P, C.c_int, _I64, _P]),
    "xm_peer_info": (C.c_int, [C.c_int, C.c_int, C.POINTER(_I32), C.POINTER(_I32)]),
    "xm_build_xmap": (C.c_int, [C.c_int, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "xm_build_inverse_lut": (C.c_int, [C.c_int, _P, _P, C.c_int32, _P, C.c_int32, C.c_int32, _P, _P, _P, _P]),
}


class XmapsError(RuntimeError):
    """A C-ABI call returned a non-zero status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"xmaps_b200 error {code}: {message}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()' or python x-maps_b200/csrc/build.py). "
            "xmaps_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.xm_abi_version() != ABI_VERSION:
        raise ImportError(f"ABI mismatch: library {lib.xm_abi_version()}, binding {ABI_VERSION}")
    return lib


lib = _load()


def check(code: int):
    if code != OK:
        raise XmapsError(code, lib.xm_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(lib.xm_launch_count())
