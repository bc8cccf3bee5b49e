"""ctypes binding of libspano.so (include/spano.h).  There is no CPU fallback: if the CUDA
library has not been built, importing the compute API fails loudly."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libspano.so")

SPHERICAL, CYLINDRICAL, STEREOGRAPHIC = 0, 1, 2
OUT_F32, OUT_U8 = 0, 1
E_INVALID, E_CUDA, E_NOMEM, E_NODEVICE, E_LIMIT = -1, -2, -3, -4, -5
MAX_BANDS = 10

c_u8p = C.POINTER(C.c_uint8)
c_f32p = C.POINTER(C.c_float)
c_intp = C.POINTER(C.c_int)
c_sizep = C.POINTER(C.c_size_t)
c_u8pp = C.POINTER(C.c_void_p)


class ImageDesc(C.Structure):
    """struct spano_image_desc"""
    _fields_ = [
        ("src_bgr", C.c_void_p), ("src_w", C.c_int), ("src_h", C.c_int), ("src_step", C.c_size_t),
        ("K", C.c_float * 9), ("R", C.c_float * 9), ("gain", C.c_double),
        ("mask_cut", C.c_void_p), ("mask_cut_step", C.c_size_t),
        ("tl_x", C.c_int), ("tl_y", C.c_int), ("w", C.c_int), ("h", C.c_int),
        ("valid_mask", C.c_void_p), ("valid_mask_step", C.c_size_t),
        ("mask_cut_w", C.c_int), ("mask_cut_h", C.c_int),
        ("intensity", C.c_void_p), ("intensity_w", C.c_int), ("intensity_h", C.c_int), ("intensity_step", C.c_size_t),
    ]


class Slice(C.Structure):
    """struct spano_slice: rows [row0,row1) (columns [col0,col1); col1 <= col0: all) of one warped tile as stored at its band owner"""
    _fields_ = [("row0", C.c_int), ("row1", C.c_int), ("tile", C.c_void_p), ("tile_step", C.c_size_t),
                ("valid", C.c_void_p), ("valid_step", C.c_size_t), ("col0", C.c_int), ("col1", C.c_int)]


class ShardPlanC(C.Structure):
    """struct spano_shard_plan"""
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("n", C.c_int), ("proj", C.c_int), ("scale", C.c_float), ("bands", C.c_int),
                ("sigma", C.c_double), ("canvas_w", C.c_int), ("min_x", C.c_int), ("min_y", C.c_int), ("row0", C.c_int), ("row1", C.c_int),
                ("images", C.POINTER(ImageDesc)), ("owner", C.POINTER(C.c_int)), ("order", C.POINTER(C.c_int)),
                ("slices", C.POINTER(Slice)), ("flags", C.POINTER(C.c_void_p)), ("canvas", C.c_void_p), ("canvas_step", C.c_size_t),
                ("done_lag", C.c_int)]


class CenterFix(C.Structure):
    """struct spano_center_fix"""
    _fields_ = [("ansatz_x", C.c_int), ("ansatz_y", C.c_int), ("radius", C.c_float), ("quadratic", C.c_int)]


class OverlapInfo(C.Structure):
    """struct spano_overlap_info == gain::OverlapInfo"""
    _fields_ = [("i", C.c_int), ("j", C.c_int), ("area", C.c_double), ("I_i", C.c_double), ("I_j", C.c_double)]


# every symbol include/spano.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "spano_version": (C.c_int, []),
    "spano_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "spano_destroy": (None, [C.c_void_p]),
    "spano_last_error": (C.c_char_p, [C.c_void_p]),
    "spano_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "spano_sync": (C.c_int, [C.c_void_p]),
    "spano_launch_count": (C.c_longlong, [C.c_void_p]),
    "spano_warp_roi": (C.c_int, [C.c_void_p, C.c_int, C.c_float, c_f32p, c_f32p, C.c_int, C.c_int, c_intp, c_intp, c_intp, c_intp]),
    "spano_pan_dimension": (C.c_int, [C.c_int, c_intp, c_intp, c_intp, c_intp, c_intp, c_intp, c_intp, c_intp]),
    "spano_warp": (C.c_int, [C.c_void_p, C.c_int, C.c_float, c_f32p, c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_size_t,
                             C.c_double, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "spano_build_maps": (C.c_int, [C.c_void_p, C.c_int, C.c_float, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "spano_remap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "spano_surrounding_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]),
    "spano_resize_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_size_t]),
    "spano_adjust_intensity": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_size_t]),
    "spano_equalize_intensities_size": (C.c_int, [C.c_int, C.c_int, C.c_float, c_intp, c_intp]),
    "spano_equalize_intensities": (C.c_int, [C.c_void_p, C.c_int, c_u8pp, c_sizep, c_u8pp, c_sizep, c_intp, c_intp, c_intp, c_intp, C.c_float,
                                             c_u8pp, c_sizep]),
    "spano_distance_transform": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t]),
    "spano_dist_cut": (C.c_int, [C.c_void_p, C.c_int, c_u8pp, c_sizep, c_intp, c_intp, c_intp, c_intp, c_u8pp, c_sizep]),
    "spano_simple_blend": (C.c_int, [C.c_void_p, C.c_int, c_u8pp, c_sizep, c_u8pp, c_sizep, c_intp, c_intp, c_intp, c_intp, C.c_void_p, C.c_size_t]),
    "spano_no_blend": (C.c_int, [C.c_void_p, C.c_int, c_u8pp, c_sizep, c_u8pp, c_sizep, c_intp, c_intp, c_intp, c_intp, C.c_void_p, C.c_size_t]),
    "spano_overlap_intensity": (C.c_int, [C.c_void_p, C.c_int, c_u8pp, c_sizep, c_intp, c_intp, c_intp, c_intp, C.POINTER(C.c_double),
                                          C.POINTER(OverlapInfo), c_intp]),
    "spano_apply_gain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_double]),
    "spano_disk_reproj_size": (C.c_int, [C.c_void_p, C.c_int, c_intp, c_intp, c_intp, c_intp, C.c_int, C.c_int, C.c_float,
                                         C.c_int, c_intp, c_intp, c_intp, c_intp]),
    "spano_disk_reproj": (C.c_int, [C.c_void_p, C.c_int, c_u8pp, c_sizep, c_intp, c_intp, c_intp, c_intp, C.c_int, C.c_int,
                                    C.c_float, C.c_int, c_u8pp, c_sizep, c_u8pp, c_sizep]),
    "spano_multiblend": (C.c_int, [C.c_void_p, C.c_int, c_u8pp, c_sizep, c_u8pp, c_sizep, c_u8pp, c_sizep, c_intp, c_intp,
                                   c_intp, c_intp, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_size_t]),
    "spano_composite": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_int, C.POINTER(ImageDesc), C.c_int, C.c_double,
                                  C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "spano_dev_composite": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_int, C.POINTER(ImageDesc), C.c_int, C.c_double,
                                      C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "spano_composite_fixed": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_int, C.POINTER(ImageDesc), C.c_int, C.c_double,
                                        C.POINTER(CenterFix), C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "spano_dev_composite_fixed": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_int, C.POINTER(ImageDesc), C.c_int, C.c_double,
                                            C.POINTER(CenterFix), C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "spano_dev_tile_mask": (C.c_int, [C.c_void_p, C.c_int, C.c_float, c_f32p, c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_size_t,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "spano_dev_warp": (C.c_int, [C.c_void_p, C.c_int, C.c_float, c_f32p, c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_size_t,
                                 C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "spano_dev_multiblend": (C.c_int, [C.c_void_p, C.c_int, c_u8pp, c_sizep, c_u8pp, c_sizep, c_u8pp, c_sizep, c_intp, c_intp,
                                       c_intp, c_intp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "spano_peer_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_ubyte)]),
    "spano_peer_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_ubyte), C.POINTER(C.c_void_p)]),
    "spano_peer_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "spano_peer_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "spano_dev_warp_scatter": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.POINTER(ImageDesc), C.c_int, C.POINTER(Slice)]),
    "spano_warp_scatter": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.POINTER(ImageDesc), C.c_int, C.POINTER(Slice)]),
    "spano_dev_blend_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
    "spano_blend_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                    C.POINTER(ImageDesc), C.c_void_p, C.c_size_t]),
    "spano_dev_blend_prepare": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(ImageDesc), C.POINTER(Slice)]),
    "spano_blend_prepare": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(ImageDesc), C.POINTER(Slice)]),
    "spano_dev_blend_add": (C.c_int, [C.c_void_p, C.POINTER(ImageDesc), C.POINTER(Slice)]),
    "spano_dev_blend_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "spano_blend_add": (C.c_int, [C.c_void_p, C.POINTER(ImageDesc), C.POINTER(Slice)]),
    "spano_blend_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "spano_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "spano_shard_step_owner": (C.c_int, [C.c_void_p, C.POINTER(ShardPlanC), C.c_uint, C.c_int]),
    "spano_shard_step_band": (C.c_int, [C.c_void_p, C.POINTER(ShardPlanC), C.c_uint, C.c_int, C.c_void_p, C.c_size_t]),
    "spano_timers_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "spano_timers_reset": (C.c_int, [C.c_void_p]),
    "spano_timers_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_longlong)]),
    "spano_blend_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.c_int]),
    "spano_fp32_peak": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
}

_lib = None


def load() -> C.CDLL:
    """Load libspano.so; raises ImportError when it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first (python -m simplepanorama_b200.build). "
            "simplepanorama_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SpanoError(RuntimeError):
    """What the reference reports as cv::Exception / std::runtime_error on this path."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"spano error {code}: {msg}")
        self.code = code
