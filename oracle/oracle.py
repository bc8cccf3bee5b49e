"""ctypes wrapper of the C oracle (oracle/spano_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (simplepanorama_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libspano_oracle.so")
SPHERICAL, CYLINDRICAL, STEREOGRAPHIC = 0, 1, 2

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "spano_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "libspano_oracle.so"], check=True, capture_output=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f9(a):
    return np.ascontiguousarray(np.asarray(a, np.float32).reshape(9))


def warp_roi(kind, scale, K32, R32, src_w, src_h):
    """-> ((tl_x, tl_y), (tile_w, tile_h)) with the tile = ROI + 1 (RotationWarperBase::warp)."""
    out = np.zeros(4, np.int32)
    k, r = _f9(K32), _f9(R32)
    lib().orc_warp_roi(int(kind), C.c_float(scale), _p(k, C.c_float), _p(r, C.c_float), int(src_w), int(src_h), _p(out, C.c_int))
    return (int(out[0]), int(out[1])), (int(out[2] - out[0] + 1), int(out[3] - out[1] + 1))


def build_maps(kind, scale, K32, R32, tl, size):
    w, h = size
    xm = np.empty((h, w), np.float32)
    ym = np.empty((h, w), np.float32)
    k, r = _f9(K32), _f9(R32)
    lib().orc_build_maps(int(kind), C.c_float(scale), _p(k, C.c_float), _p(r, C.c_float), int(tl[0]), int(tl[1]), w, h,
                         _p(xm, C.c_float), _p(ym, C.c_float))
    return xm, ym


def remap(img, xm, ym):
    img = np.ascontiguousarray(img)
    h, w = xm.shape
    dst = np.empty((h, w, 3), np.uint8)
    lib().orc_remap_linear_u8c3(_p(img, C.c_uint8), img.shape[1], img.shape[0], C.c_size_t(img.strides[0]),
                                _p(xm, C.c_float), _p(ym, C.c_float), w, h, _p(dst, C.c_uint8), C.c_size_t(dst.strides[0]))
    return dst


def warp(kind, scale, K32, R32, img):
    """cv::detail::*Warper::warp(INTER_LINEAR, BORDER_CONSTANT) -> (corner, tile)."""
    tl, size = warp_roi(kind, scale, K32, R32, img.shape[1], img.shape[0])
    xm, ym = build_maps(kind, scale, K32, R32, tl, size)
    return tl, remap(img, xm, ym)


def surrounding_mask(img, erode_iters=3):
    img = np.ascontiguousarray(img)
    h, w = img.shape[:2]
    m = np.empty((h, w), np.uint8)
    lib().orc_surrounding_mask(_p(img, C.c_uint8), w, h, C.c_size_t(img.strides[0]), int(erode_iters), _p(m, C.c_uint8))
    return m


def apply_gain(img, gain):
    out = np.array(img, copy=True, order="C")
    lib().orc_apply_gain_u8(_p(out, C.c_uint8), C.c_size_t(out.size), C.c_double(gain))
    return out


def gaussian_taps(n, sigma):
    t = np.empty(n, np.float32)
    lib().orc_gaussian_taps(int(n), C.c_double(sigma), _p(t, C.c_float))
    return t


def gaussian_blur(src, ksize, sigma):
    src = np.ascontiguousarray(src, np.float32)
    cn = 1 if src.ndim == 2 else src.shape[2]
    dst = np.empty_like(src)
    lib().orc_gaussian_blur_f32(_p(src, C.c_float), src.shape[1], src.shape[0], cn, int(ksize), C.c_double(sigma), _p(dst, C.c_float))
    return dst


def pan_dimension(corners, sizes):
    n = len(corners)
    tlx = np.array([c[0] for c in corners], np.int32); tly = np.array([c[1] for c in corners], np.int32)
    w = np.array([s[0] for s in sizes], np.int32); h = np.array([s[1] for s in sizes], np.int32)
    out = np.zeros(6, np.int32)
    lib().orc_pan_dimension(n, _p(tlx, C.c_int), _p(tly, C.c_int), _p(w, C.c_int), _p(h, C.c_int), _p(out, C.c_int))
    return int(out[0]), int(out[1]), int(out[2]), int(out[3])


def multi_blend(tiles, masks, masks_orig, corners, bands, sigma):
    """blnd::multi_blend -> float32 canvas (H, W, 3)."""
    n = len(tiles)
    tiles = [np.ascontiguousarray(t) for t in tiles]
    masks = [np.ascontiguousarray(m) for m in masks]
    masks_orig = [np.ascontiguousarray(m) for m in masks_orig]
    sizes = [(t.shape[1], t.shape[0]) for t in tiles]
    W, H, _, _ = pan_dimension(corners, sizes)
    out = np.empty((H, W, 3), np.float32)
    arr = lambda xs: (C.c_void_p * n)(*[x.ctypes.data for x in xs])
    tlx = np.array([c[0] for c in corners], np.int32); tly = np.array([c[1] for c in corners], np.int32)
    w = np.array([s[0] for s in sizes], np.int32); h = np.array([s[1] for s in sizes], np.int32)
    lib().orc_multi_blend(n, arr(tiles), arr(masks), arr(masks_orig), _p(tlx, C.c_int), _p(tly, C.c_int), _p(w, C.c_int),
                          _p(h, C.c_int), int(bands), C.c_double(sigma), _p(out, C.c_float))
    return out


def blend_to_u8(blend):
    blend = np.ascontiguousarray(blend, np.float32)
    out = np.empty(blend.shape, np.uint8)
    lib().orc_blend_to_u8(_p(blend, C.c_float), C.c_size_t(blend.size), _p(out, C.c_uint8))
    return out


def resize_linear_u8(mask, size_wh):
    mask = np.ascontiguousarray(mask)
    dw, dh = size_wh
    out = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8c1(_p(mask, C.c_uint8), mask.shape[1], mask.shape[0], _p(out, C.c_uint8), dw, dh)
    return out


def resize_linear_f32(field, size_wh):
    field = np.ascontiguousarray(field, np.float32)
    dw, dh = size_wh
    out = np.empty((dh, dw), np.float32)
    lib().orc_resize_linear_f32c1(_p(field, C.c_float), field.shape[1], field.shape[0], _p(out, C.c_float), dw, dh)
    return out


def adjust_intensity(img, field):
    """test::adjust_intensity for one image -> new CV_8UC3 image."""
    out = np.array(img, copy=True, order="C")
    field = np.ascontiguousarray(field, np.float32)
    lib().orc_adjust_intensity(_p(out, C.c_uint8), out.shape[1], out.shape[0], C.c_size_t(out.strides[0]), _p(field, C.c_float),
                               field.shape[1], field.shape[0])
    return out


class _DiskParams(C.Structure):
    _fields_ = [("cx", C.c_float), ("cy", C.c_float), ("scale", C.c_float), ("radius_n", C.c_float), ("quadratic", C.c_int)]


def disk_reproj(tiles, corners, ansatz, radius, quadratic=True, erode_iters=3):
    """sten_proj::disk_reproj -> (new tiles, new masks, new centre-relative corners)."""
    n = len(tiles)
    tiles = [np.ascontiguousarray(t) for t in tiles]
    tlx = np.array([c[0] for c in corners], np.int32); tly = np.array([c[1] for c in corners], np.int32)
    w = np.array([t.shape[1] for t in tiles], np.int32); h = np.array([t.shape[0] for t in tiles], np.int32)
    ow = np.zeros(n, np.int32); oh = np.zeros(n, np.int32); ox = np.zeros(n, np.int32); oy = np.zeros(n, np.int32)
    P = _DiskParams()
    lib().orc_disk_reproj_plan(n, _p(tlx, C.c_int), _p(tly, C.c_int), _p(w, C.c_int), _p(h, C.c_int), int(ansatz[0]), int(ansatz[1]),
                               C.c_float(radius), int(bool(quadratic)), _p(ow, C.c_int), _p(oh, C.c_int), _p(ox, C.c_int),
                               _p(oy, C.c_int), C.byref(P))
    outs, msks = [], []
    for i in range(n):
        dst = np.empty((int(oh[i]), int(ow[i]), 3), np.uint8)
        lib().orc_disk_reproj_tile(C.byref(P), _p(tiles[i], C.c_uint8), int(w[i]), int(h[i]), C.c_size_t(tiles[i].strides[0]),
                                   int(ox[i]), int(oy[i]), _p(dst, C.c_uint8), int(ow[i]), int(oh[i]), C.c_size_t(dst.strides[0]),
                                   int(tlx[i]), int(tly[i]))
        outs.append(dst)
        msks.append(surrounding_mask(dst, erode_iters))
    return outs, msks, [(int(a), int(b)) for a, b in zip(tlx, tly)]


def adjusted_camera(K, R, w_ref, h_ref):
    """K_adj / R as float32 (src/math/_projection.cpp:36-49)."""
    K = np.asarray(K, np.float64)
    f_i = K[0, 0]
    K_adj = np.array([[f_i, 0, w_ref - K[0, 2]], [0, f_i, h_ref - K[1, 2]], [0, 0, 1]], np.float64)
    return K_adj.astype(np.float32), np.asarray(R, np.float64).astype(np.float32)


def return_full(images, R, K, kind, focal, gains, masks_cut, bands, sigma, want_float=False, intensities=None):
    """stitch_parameters::return_full (MULTI_BLEND): sources -> (u8 canvas, tiles, masks, corners)."""
    tiles, msks, corners = [], [], []
    for img, r, k in zip(images, R, K):
        K32, R32 = adjusted_camera(k, r, img.shape[1], img.shape[0])
        tl, tile = warp(kind, np.float32(focal), K32, R32, img)
        msks.append(surrounding_mask(tile, 3))
        tiles.append(tile)
        corners.append(tl)
    gained = [apply_gain(t, g) for t, g in zip(tiles, gains)] if gains is not None else tiles
    # cv::resize(mask_cut[i], .., tile size) -- default INTER_LINEAR (src/classes/_panorama.cpp:329-335)
    masks_cut = [m if m.shape == t.shape[:2] else resize_linear_u8(m, (t.shape[1], t.shape[0])) for m, t in zip(masks_cut, tiles)]
    if intensities is not None:   # test::adjust_intensity (conf.blend_intensity), src/classes/_panorama.cpp:337-339
        gained = [adjust_intensity(t, f) for t, f in zip(gained, intensities)]
    blend = multi_blend(gained, masks_cut, msks, corners, bands, sigma)
    out = blend_to_u8(blend)
    if want_float:
        return out, blend, gained, msks, corners
    return out, gained, msks, corners


def distance_transform(mask):
    """cv::distanceTransform(mask, DIST_L2, DIST_MASK_5, CV_32F)."""
    m = np.ascontiguousarray(mask, np.uint8)
    out = np.empty(m.shape, np.float32)
    lib().orc_distance_transform(m.ctypes.data_as(C.c_void_p), m.shape[1], m.shape[0], C.c_size_t(m.strides[0]), _p(out, C.c_float))
    return out


def dist_cut(masks, corners):
    """dcut::dist_cut -> list of cut masks."""
    n = len(masks)
    ms = [np.ascontiguousarray(m, np.uint8) for m in masks]
    outs = [np.empty(m.shape, np.uint8) for m in ms]
    arr = lambda xs: (C.c_void_p * n)(*[x.ctypes.data for x in xs])
    tlx = np.array([c[0] for c in corners], np.int32); tly = np.array([c[1] for c in corners], np.int32)
    w = np.array([m.shape[1] for m in ms], np.int32); h = np.array([m.shape[0] for m in ms], np.int32)
    lib().orc_dist_cut(n, arr(ms), _p(tlx, C.c_int), _p(tly, C.c_int), _p(w, C.c_int), _p(h, C.c_int), arr(outs))
    return outs


def _simple_or_no_blend(fn, tiles, masks, corners):
    n = len(tiles)
    ts = [np.ascontiguousarray(t, np.uint8) for t in tiles]
    ms = [np.ascontiguousarray(m, np.uint8) for m in masks]
    sizes = [(t.shape[1], t.shape[0]) for t in ts]
    W, H, _, _ = pan_dimension(corners, sizes)
    out = np.empty((H, W, 3), np.uint8)
    arr = lambda xs: (C.c_void_p * n)(*[x.ctypes.data for x in xs])
    tlx = np.array([c[0] for c in corners], np.int32); tly = np.array([c[1] for c in corners], np.int32)
    w = np.array([s[0] for s in sizes], np.int32); h = np.array([s[1] for s in sizes], np.int32)
    fn(n, arr(ts), arr(ms), _p(tlx, C.c_int), _p(tly, C.c_int), _p(w, C.c_int), _p(h, C.c_int), out.ctypes.data_as(C.c_void_p))
    return out


def simple_blend(tiles, masks, corners):
    """blnd::simple_blend -> CV_8UC3 canvas."""
    return _simple_or_no_blend(lib().orc_simple_blend, tiles, masks, corners)


def no_blend(tiles, masks, corners):
    """blnd::no_blend -> CV_8UC3 canvas."""
    return _simple_or_no_blend(lib().orc_no_blend, tiles, masks, corners)


def overlap_intensity(tiles, corners, adj):
    """gain::get_overlapp_intensity -> list of (i, j, area, I_i, I_j)."""
    n = len(tiles)
    ts = [np.ascontiguousarray(t, np.uint8) for t in tiles]
    arr = (C.c_void_p * n)(*[t.ctypes.data for t in ts])
    tlx = np.array([c[0] for c in corners], np.int32); tly = np.array([c[1] for c in corners], np.int32)
    w = np.array([t.shape[1] for t in ts], np.int32); h = np.array([t.shape[0] for t in ts], np.int32)
    a = np.ascontiguousarray(adj, np.float64)
    out = np.zeros((n * (n + 1) // 2, 5), np.float64)
    lib().orc_overlap_intensity.restype = C.c_int
    cnt = lib().orc_overlap_intensity(n, arr, _p(tlx, C.c_int), _p(tly, C.c_int), _p(w, C.c_int), _p(h, C.c_int), _p(a, C.c_double),
                                      _p(out, C.c_double))
    return [(int(r[0]), int(r[1]), float(r[2]), float(r[3]), float(r[4])) for r in out[:cnt]]
