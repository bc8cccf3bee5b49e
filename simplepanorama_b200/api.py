"""Host-side mirror of the reference's compositing interface, on top of the C ABI (include/spano.h).

Function names, argument meaning and error behaviour follow the reference's C++ callees so the
parity tests read like tests of the reference itself (paths relative to the upstream tree):

    project / get_proj_parameters   proj::projection::project, proj::get_proj_parameters
                                    (src/math/_projection.cpp:27-84,297-324,422-454)
    create_surrounding_mask         blnd::createSurroundingMask  (src/math/_blending.cpp:278-324)
    apply_gain                      `imgs[i] / gain[i]`          (src/classes/_panorama.cpp:321-327)
    get_pan_dimension               util::get_pan_dimension      (src/system/_util.cpp:204-231)
    multi_blend                     blnd::multi_blend            (src/math/_blending.cpp:186-252)
    blend                           stitch_parameters::blend, MULTI_BLEND (src/classes/_panorama.cpp:242-249)
    return_full                     stitch_parameters::return_full (src/classes/_panorama.cpp:259-354)

All compute happens in libspano.so (hand-written CUDA, sm_100a).  Host arrays are numpy.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import (CYLINDRICAL, OUT_F32, OUT_U8, SPHERICAL, STEREOGRAPHIC, ImageDesc, SpanoError)

__all__ = [
    "SPHERICAL", "CYLINDRICAL", "STEREOGRAPHIC", "Context", "ProjData", "SpanoError", "adjusted_camera", "warp_roi",
    "project", "get_proj_parameters", "create_surrounding_mask", "validity_mask", "apply_gain", "get_pan_dimension",
    "multi_blend", "blend", "return_full", "default_context", "distance_transform", "dist_cut", "simple_blend", "no_blend", "get_overlapp_intensity", "stitch_blend", "NO_BLEND", "SIMPLE_BLEND", "MULTI_BLEND",
]


def _f9(a) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(9))
    return a


def _fp(a: np.ndarray):
    return a.ctypes.data_as(_lib.c_f32p)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(_lib.c_intp)


def _u8img(a, channels: int, what: str) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype != np.uint8:
        raise SpanoError(_lib.E_INVALID, f"{what}: expected uint8, got {a.dtype}")
    want = 3 if channels == 3 else 2
    if a.ndim != want or (channels == 3 and a.shape[2] != 3):
        raise SpanoError(_lib.E_INVALID, f"{what}: expected {'HxWx3' if channels == 3 else 'HxW'}, got {a.shape}")
    if a.size == 0:
        raise SpanoError(_lib.E_INVALID, f"{what}: empty image")
    if a.strides[-1] != 1 or (channels == 3 and a.strides[1] != 3):
        a = np.ascontiguousarray(a)
    return a


class Context:
    """One spano_ctx (one per panorama object / thread, like one pan::panorama per viewer window)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.spano_create(C.byref(h), int(device))
        if rc != 0:
            raise SpanoError(rc, {
                _lib.E_NODEVICE: "no usable CUDA (sm_100) device; this library has no CPU path",
                _lib.E_INVALID: f"invalid device ordinal {device}",
            }.get(rc, "spano_create failed"))
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.spano_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != 0:
            raise SpanoError(rc, self.lib.spano_last_error(self.h).decode("utf-8", "replace"))

    def set_stream(self, cuda_stream: int | None):
        self.check(self.lib.spano_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def sync(self):
        self.check(self.lib.spano_sync(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.spano_launch_count(self.h))

    OPT_BLEND_DENSE, OPT_BLEND_KERNEL, OPT_FLAG_WAIT, OPT_WARP_KERNEL = 1, 2, 3, 4   # include/spano.h: SPANO_OPT_*

    def set_option(self, option: int, value: int):
        self.check(self.lib.spano_set_option(self.h, int(option), int(value)))

    def timers_enable(self, on: bool = True):
        self.check(self.lib.spano_timers_enable(self.h, int(on)))

    def timers_reset(self):
        self.check(self.lib.spano_timers_reset(self.h))

    def timers_read(self):
        ms = (C.c_float * 4)()
        n = (C.c_longlong * 4)()
        self.check(self.lib.spano_timers_read(self.h, ms, n))
        names = ("warp", "mask", "blend", "normalise")
        return {k: float(ms[i]) for i, k in enumerate(names)}, {k: int(n[i]) for i, k in enumerate(names)}

    def blend_stats(self, reset: bool = False):
        """(processed tile px, offered tile px) of the blend launches since the last reset (mask_cut sparsity)."""
        a, b = C.c_ulonglong(), C.c_ulonglong()
        self.check(self.lib.spano_blend_stats(self.h, C.byref(a), C.byref(b), int(reset)))
        return int(a.value), int(b.value)

    def fp32_peak(self, variant: int = 0) -> float:
        v = C.c_double()
        self.check(self.lib.spano_fp32_peak(self.h, int(variant), C.byref(v)))
        return float(v.value)


_default_ctx: Context | None = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ---------------------------------------------------------------------------------------------
# geometry
# ---------------------------------------------------------------------------------------------
def adjusted_camera(K, R, w_ref: int, h_ref: int):
    """K_adj / R as float32, the matrices projection::project hands to OpenCV
    (principal point flipped to (w-cx, h-cy); src/math/_projection.cpp:36-49)."""
    K = np.asarray(K, np.float64)
    f_i = K[0, 0]
    K_adj = np.array([[f_i, 0, w_ref - K[0, 2]], [0, f_i, h_ref - K[1, 2]], [0, 0, 1]], np.float64)
    return K_adj.astype(np.float32), np.asarray(R, np.float64).astype(np.float32)


def warp_roi(kind: int, focal: float, K32, R32, src_w: int, src_h: int, ctx: Context | None = None):
    """Corner and size of the warped tile: ((tl_x, tl_y), (w, h)).  K32 is K_adj (float32).
    Host arithmetic only: works without a context (and without a GPU)."""
    k, r = _f9(K32), _f9(R32)
    out = [C.c_int() for _ in range(4)]
    lib = ctx.lib if ctx is not None else _lib.load()
    rc = lib.spano_warp_roi(ctx.h if ctx is not None else None, int(kind), C.c_float(focal), _fp(k), _fp(r),
                            int(src_w), int(src_h), *[C.byref(o) for o in out])
    if rc != 0:
        if ctx is not None:
            ctx.check(rc)
        raise SpanoError(rc, "spano_warp_roi failed (degenerate ROI or bad argument)")
    return (out[0].value, out[1].value), (out[2].value, out[3].value)


def project(kind: int, focal: float, R, K, img, gain: float = 1.0, get_mask: bool = False, ctx: Context | None = None):
    """projection::project (+ optionally the validity mask and the 8-bit gain of the later stages).

    Returns (corner(x, y), warped tile[, validity mask])."""
    ctx = ctx or default_context()
    img = _u8img(img, 3, "image")
    h_ref, w_ref = img.shape[:2]
    K32, R32 = adjusted_camera(K, R, w_ref, h_ref)
    corner, (w, h) = warp_roi(kind, focal, K32, R32, w_ref, h_ref, ctx)
    dst = np.empty((h, w, 3), np.uint8)
    msk = np.empty((h, w), np.uint8) if get_mask else None
    k, r = _f9(K32), _f9(R32)
    ctx.check(ctx.lib.spano_warp(ctx.h, int(kind), C.c_float(focal), _fp(k), _fp(r), img.ctypes.data, w_ref, h_ref,
                                 img.strides[0], float(gain), dst.ctypes.data, dst.strides[0],
                                 msk.ctypes.data if get_mask else None, msk.strides[0] if get_mask else 0))
    return (corner, dst, msk) if get_mask else (corner, dst)


def build_maps(kind: int, focal: float, K32, R32, corner, size, ctx: Context | None = None):
    """cv::detail::RotationWarperBase::buildMaps -> (xmap, ymap) float32 of a (w, h) tile at `corner`."""
    ctx = ctx or default_context()
    w, h = size
    xm = np.empty((h, w), np.float32)
    ym = np.empty((h, w), np.float32)
    k, r = _f9(K32), _f9(R32)
    ctx.check(ctx.lib.spano_build_maps(ctx.h, int(kind), C.c_float(focal), _fp(k), _fp(r), int(corner[0]), int(corner[1]),
                                       int(w), int(h), xm.ctypes.data, ym.ctypes.data))
    return xm, ym


def remap(img, xmap, ymap, ctx: Context | None = None) -> np.ndarray:
    """cv::remap(img, xmap, ymap, INTER_LINEAR, BORDER_CONSTANT) on CV_8UC3."""
    ctx = ctx or default_context()
    img = _u8img(img, 3, "image")
    xm = np.ascontiguousarray(xmap, np.float32)
    ym = np.ascontiguousarray(ymap, np.float32)
    if xm.ndim != 2 or xm.shape != ym.shape or xm.size == 0:
        raise SpanoError(_lib.E_INVALID, "maps must be equal-sized non-empty 2-D float32 arrays")
    h, w = xm.shape
    dst = np.empty((h, w, 3), np.uint8)
    ctx.check(ctx.lib.spano_remap(ctx.h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], xm.ctypes.data,
                                  ym.ctypes.data, w, h, dst.ctypes.data, dst.strides[0]))
    return dst


@dataclass
class ProjData:
    """proj::proj_data (src/math/_projection.h:15-19)."""
    imgs: list = field(default_factory=list)
    msks: list = field(default_factory=list)
    corners: list = field(default_factory=list)


def get_proj_parameters(images, R, K, con, kind: int, focal: float, get_masks: bool = True,
                        ctx: Context | None = None) -> ProjData:
    """proj::get_proj_parameters: warp every connected image, collect corners and validity masks."""
    out = ProjData()
    for i, img in enumerate(images):
        if con[i] > 0:
            res = project(kind, focal, R[i], K[i], img, 1.0, get_masks, ctx)
            out.corners.append(res[0])
            out.imgs.append(res[1])
            if get_masks:
                out.msks.append(res[2])
    return out


def create_surrounding_mask(img, erode_iters: int = 0, ctx: Context | None = None) -> np.ndarray:
    """blnd::createSurroundingMask(img, true, 1), optionally followed by cv::erode(.., iterations)."""
    ctx = ctx or default_context()
    img = _u8img(img, 3, "image")
    h, w = img.shape[:2]
    m = np.empty((h, w), np.uint8)
    ctx.check(ctx.lib.spano_surrounding_mask(ctx.h, img.ctypes.data, w, h, img.strides[0], int(erode_iters),
                                             m.ctypes.data, m.strides[0]))
    return m


def validity_mask(img, ctx: Context | None = None) -> np.ndarray:
    """createSurroundingMask + cv::erode(mask, Mat(), (-1,-1), 3)  (src/math/_projection.cpp:441-443)."""
    return create_surrounding_mask(img, 3, ctx)


def apply_gain(img, g: float, ctx: Context | None = None) -> np.ndarray:
    """`img / g` on CV_8UC3: saturate(rint(float(v) * float(1/g)))."""
    ctx = ctx or default_context()
    out = np.array(_u8img(img, 3, "image"), copy=True, order="C")
    h, w = out.shape[:2]
    ctx.check(ctx.lib.spano_apply_gain(ctx.h, out.ctypes.data, w, h, out.strides[0], float(g)))
    return out


def get_pan_dimension(top_lefts, images):
    """util::get_pan_dimension -> (width, height, min_x, min_y)."""
    return pan_dimension(top_lefts, [(im.shape[1], im.shape[0]) for im in images])


def pan_dimension(top_lefts, sizes):
    """Canvas bounding box of tiles given as corners and (w, h) sizes."""
    n = len(sizes)
    if n == 0 or n != len(top_lefts):
        raise SpanoError(_lib.E_INVALID, "Input consistency!")
    tlx = np.array([c[0] for c in top_lefts], np.int32)
    tly = np.array([c[1] for c in top_lefts], np.int32)
    w = np.array([s[0] for s in sizes], np.int32)
    h = np.array([s[1] for s in sizes], np.int32)
    out = [C.c_int() for _ in range(4)]
    rc = _lib.load().spano_pan_dimension(n, _ip(tlx), _ip(tly), _ip(w), _ip(h), *[C.byref(o) for o in out])
    if rc != 0:
        raise SpanoError(rc, "spano_pan_dimension failed")
    return tuple(o.value for o in out)


def _blend_call(images, masks, masks_orig, top_lefts, bands, sigma, out_kind, ctx):
    ctx = ctx or default_context()
    n = len(images)
    if n == 0 or n != len(masks) or n != len(masks_orig) or n != len(top_lefts):
        raise SpanoError(_lib.E_INVALID, "Input consistency!")
    imgs = [_u8img(a, 3, f"images[{i}]") for i, a in enumerate(images)]
    mc = [_u8img(a, 1, f"masks[{i}]") for i, a in enumerate(masks)]
    mo = [_u8img(a, 1, f"masks_orig[{i}]") for i, a in enumerate(masks_orig)]
    for i in range(n):
        if mc[i].shape != imgs[i].shape[:2] or mo[i].shape != imgs[i].shape[:2]:
            raise SpanoError(_lib.E_INVALID, f"mask {i} does not match its tile size")
    W, H, _, _ = get_pan_dimension(top_lefts, imgs)
    ptr = lambda arrs: (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    steps = lambda arrs: (C.c_size_t * n)(*[a.strides[0] for a in arrs])
    tlx = np.array([c[0] for c in top_lefts], np.int32)
    tly = np.array([c[1] for c in top_lefts], np.int32)
    w = np.array([im.shape[1] for im in imgs], np.int32)
    h = np.array([im.shape[0] for im in imgs], np.int32)
    out = np.empty((H, W, 3), np.float32 if out_kind == OUT_F32 else np.uint8)
    ctx.check(ctx.lib.spano_multiblend(ctx.h, n, ptr(imgs), steps(imgs), ptr(mc), steps(mc), ptr(mo), steps(mo),
                                       _ip(tlx), _ip(tly), _ip(w), _ip(h), int(bands), float(sigma), out_kind,
                                       out.ctypes.data, out.strides[0]))
    return out


def multi_blend(images, masks, masks_orig, top_lefts, bands: int, sigma: float, ctx: Context | None = None):
    """blnd::multi_blend -> CV_32FC3 canvas."""
    return _blend_call(images, masks, masks_orig, top_lefts, bands, sigma, OUT_F32, ctx)


def blend(images, masks, masks_orig, top_lefts, bands: int, sigma: float, ctx: Context | None = None):
    """stitch_parameters::blend with conf.blend == MULTI_BLEND -> CV_8UC3 canvas."""
    return _blend_call(images, masks, masks_orig, top_lefts, bands, sigma, OUT_U8, ctx)


NO_BLEND, SIMPLE_BLEND, MULTI_BLEND = 0, 1, 2   # pan::Blending (src/classes/_panorama.h:34-36)


def stitch_blend(imgs, msks_cut, msks, corners, blend_mode: int = MULTI_BLEND, bands: int = 3, sigma: float = 7.0,
                 cut: bool = True, ctx: Context | None = None):
    """stitch_parameters::blend(blend_data, config) (src/classes/_panorama.cpp:220-256): the dispatch on conf.blend.
    NO_BLEND copies through msks_cut when conf.cut / conf.cut_seams is set, else through msks; SIMPLE_BLEND feathers
    with msks; MULTI_BLEND is multi_blend(imgs, msks_cut, msks) * 255 -> CV_8UC3."""
    if blend_mode == NO_BLEND:
        return no_blend(imgs, msks_cut if cut else msks, corners, ctx)
    if blend_mode == SIMPLE_BLEND:
        return simple_blend(imgs, msks, corners, ctx)
    if blend_mode == MULTI_BLEND:
        return blend(imgs, msks_cut, msks, corners, bands, sigma, ctx)
    raise SpanoError(_lib.E_INVALID, f"unknown blend mode {blend_mode}")   # the reference returns an empty cv::Mat here


def disk_reproj_size(corners, sizes, ansatz, radius: float, quadratic: bool = True, ctx: Context | None = None):
    """Geometry of sten_proj::disk_reproj: the new (canvas-centre-relative) corners and sizes of every
    tile for the circle (ansatz, radius) that sten_proj::estimate_circle returned.  Host arithmetic."""
    n = len(sizes)
    if n == 0 or n != len(corners):
        raise SpanoError(_lib.E_INVALID, "Input consistency!")
    tlx = np.array([c[0] for c in corners], np.int32); tly = np.array([c[1] for c in corners], np.int32)
    w = np.array([s[0] for s in sizes], np.int32); h = np.array([s[1] for s in sizes], np.int32)
    ox = np.zeros(n, np.int32); oy = np.zeros(n, np.int32); ow = np.zeros(n, np.int32); oh = np.zeros(n, np.int32)
    lib = ctx.lib if ctx is not None else _lib.load()
    rc = lib.spano_disk_reproj_size(ctx.h if ctx is not None else None, n, _ip(tlx), _ip(tly), _ip(w), _ip(h),
                                    int(ansatz[0]), int(ansatz[1]), C.c_float(radius), int(bool(quadratic)),
                                    _ip(ox), _ip(oy), _ip(ow), _ip(oh))
    if rc != 0:
        raise SpanoError(rc, "spano_disk_reproj_size failed")
    return [(int(a), int(b)) for a, b in zip(ox, oy)], [(int(a), int(b)) for a, b in zip(ow, oh)]


def disk_reproj(pd: "ProjData", ansatz, radius: float, quadratic: bool = True, ctx: Context | None = None) -> "ProjData":
    """sten_proj::disk_reproj(proj_data, quadratic) with the circle forced to (ansatz, radius):
    returns the re-projected tiles, their new corners and recomputed validity masks."""
    ctx = ctx or default_context()
    n = len(pd.imgs)
    tiles = [_u8img(t, 3, f"imgs[{i}]") for i, t in enumerate(pd.imgs)]
    sizes = [(t.shape[1], t.shape[0]) for t in tiles]
    new_corners, new_sizes = disk_reproj_size(pd.corners, sizes, ansatz, radius, quadratic, ctx)
    outs = [np.empty((h, w, 3), np.uint8) for (w, h) in new_sizes]
    msks = [np.empty((h, w), np.uint8) for (w, h) in new_sizes]
    tlx = np.array([c[0] for c in pd.corners], np.int32); tly = np.array([c[1] for c in pd.corners], np.int32)
    w = np.array([s[0] for s in sizes], np.int32); h = np.array([s[1] for s in sizes], np.int32)
    ptr = lambda arrs: (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    steps = lambda arrs: (C.c_size_t * n)(*[a.strides[0] for a in arrs])
    ctx.check(ctx.lib.spano_disk_reproj(ctx.h, n, ptr(tiles), steps(tiles), _ip(tlx), _ip(tly), _ip(w), _ip(h),
                                        int(ansatz[0]), int(ansatz[1]), C.c_float(radius), int(bool(quadratic)),
                                        ptr(outs), steps(outs), ptr(msks), steps(msks)))
    return ProjData(imgs=outs, msks=msks, corners=new_corners)


def plan_tiles(images, R, K, kind: int, focal: float, ctx: Context | None = None):
    """Geometry of every warped tile without warping: list of (K32, R32, (tl_x, tl_y), (w, h))."""
    plan = []
    for img, r, k in zip(images, R, K):
        h_ref, w_ref = img.shape[:2]
        K32, R32 = adjusted_camera(k, r, w_ref, h_ref)
        corner, size = warp_roi(kind, focal, K32, R32, w_ref, h_ref, ctx)
        plan.append((K32, R32, corner, size))
    return plan


def make_descs(images, plan, gains, masks_cut, ptr_of=lambda a: a.ctypes.data, step_of=lambda a: a.strides[0]):
    """Array of spano_image_desc for spano_composite / spano_dev_composite."""
    n = len(images)
    descs = (ImageDesc * n)()
    for j in range(n):
        K32, R32, (tlx, tly), (w, h) = plan[j]
        d = descs[j]
        d.src_bgr = ptr_of(images[j])
        d.src_h, d.src_w = int(images[j].shape[0]), int(images[j].shape[1])
        d.src_step = step_of(images[j])
        d.K[:] = [float(v) for v in np.asarray(K32, np.float32).reshape(9)]
        d.R[:] = [float(v) for v in np.asarray(R32, np.float32).reshape(9)]
        d.gain = float(gains[j]) if gains is not None else 1.0
        d.mask_cut = ptr_of(masks_cut[j])
        d.mask_cut_step = step_of(masks_cut[j])
        d.tl_x, d.tl_y, d.w, d.h = int(tlx), int(tly), int(w), int(h)
        mh, mw = int(masks_cut[j].shape[0]), int(masks_cut[j].shape[1])
        if (mw, mh) != (int(w), int(h)):     # preview-scale mask: the fused path resizes it like return_full does
            d.mask_cut_w, d.mask_cut_h = mw, mh
    return descs


def equalize_intensities(images, masks, top_lefts, ratio: float = 0.5, ctx: Context | None = None):
    """test::equalizeIntensities(images, masks, top_lefts, ratio) (src/test/_test.cpp:9-106): the CV_32FC1 intensity-
    correction field of every preview-size warp (validity masks as `masks`), ready for adjust_intensity."""
    ctx = ctx or default_context()
    n = len(images)
    if n == 0 or n != len(masks) or n != len(top_lefts):
        raise SpanoError(_lib.E_INVALID, "Input consistency!")
    imgs = [_u8img(a, 3, f"images[{i}]") for i, a in enumerate(images)]
    msks = [_u8img(a, 1, f"masks[{i}]") for i, a in enumerate(masks)]
    w = np.array([a.shape[1] for a in imgs], np.int32); h = np.array([a.shape[0] for a in imgs], np.int32)
    tlx = np.array([c[0] for c in top_lefts], np.int32); tly = np.array([c[1] for c in top_lefts], np.int32)
    outs = []
    for i in range(n):
        fw, fh = C.c_int(), C.c_int()
        ctx.check(ctx.lib.spano_equalize_intensities_size(int(w[i]), int(h[i]), C.c_float(ratio), C.byref(fw), C.byref(fh)))
        outs.append(np.empty((fh.value, fw.value), np.float32))
    ptr = lambda arrs: (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    steps = lambda arrs: (C.c_size_t * n)(*[a.strides[0] for a in arrs])
    ctx.check(ctx.lib.spano_equalize_intensities(ctx.h, n, ptr(imgs), steps(imgs), ptr(msks), steps(msks), _ip(tlx), _ip(tly), _ip(w), _ip(h),
                                                 C.c_float(ratio), ptr(outs), steps(outs)))
    return outs


def resize_mask(mask, size_wh, ctx: Context | None = None) -> np.ndarray:
    """cv::resize(mask, dst, size) on CV_8UC1 with the default INTER_LINEAR -- return_full's mask_cut
    up-scaling (src/classes/_panorama.cpp:329-335)."""
    ctx = ctx or default_context()
    m = _u8img(mask, 1, "mask")
    dw, dh = int(size_wh[0]), int(size_wh[1])
    if dw <= 0 or dh <= 0:
        raise SpanoError(_lib.E_INVALID, "empty destination size")
    out = np.empty((dh, dw), np.uint8)
    ctx.check(ctx.lib.spano_resize_mask(ctx.h, m.ctypes.data, m.shape[1], m.shape[0], m.strides[0], out.ctypes.data, dw, dh,
                                        out.strides[0]))
    return out


def _simple_or_no_blend(name, images, masks, top_lefts, ctx):
    ctx = ctx or default_context()
    n = len(images)
    if n == 0 or n != len(masks) or n != len(top_lefts):
        raise SpanoError(_lib.E_INVALID, "Input consistency!")
    imgs = [_u8img(a, 3, f"images[{i}]") for i, a in enumerate(images)]
    ms = [_u8img(a, 1, f"masks[{i}]") for i, a in enumerate(masks)]
    for i in range(n):
        if ms[i].shape != imgs[i].shape[:2]:
            raise SpanoError(_lib.E_INVALID, f"mask {i} does not match its tile size")
    W, H, _, _ = get_pan_dimension(top_lefts, imgs)
    ptr = lambda arrs: (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    steps = lambda arrs: (C.c_size_t * n)(*[a.strides[0] for a in arrs])
    tlx = np.array([c[0] for c in top_lefts], np.int32); tly = np.array([c[1] for c in top_lefts], np.int32)
    w = np.array([im.shape[1] for im in imgs], np.int32); h = np.array([im.shape[0] for im in imgs], np.int32)
    out = np.empty((H, W, 3), np.uint8)
    fn = getattr(ctx.lib, name)
    ctx.check(fn(ctx.h, n, ptr(imgs), steps(imgs), ptr(ms), steps(ms), _ip(tlx), _ip(tly), _ip(w), _ip(h), out.ctypes.data, out.strides[0]))
    return out


def simple_blend(images, masks, top_lefts, ctx: Context | None = None) -> np.ndarray:
    """blnd::simple_blend (stitch_parameters::blend, SIMPLE_BLEND) -> CV_8UC3 canvas."""
    return _simple_or_no_blend("spano_simple_blend", images, masks, top_lefts, ctx)


def no_blend(images, masks, top_lefts, ctx: Context | None = None) -> np.ndarray:
    """blnd::no_blend (stitch_parameters::blend, NO_BLEND) -> CV_8UC3 canvas."""
    return _simple_or_no_blend("spano_no_blend", images, masks, top_lefts, ctx)


def get_overlapp_intensity(warped_images, corners, adj, ctx: Context | None = None):
    """gain::get_overlapp_intensity (src/math/_gain_compensation.cpp:7-75): list of (i, j, area, I_i, I_j) for every
    pair i <= j that is adjacent in `adj` (the identity is added, as the reference does)."""
    ctx = ctx or default_context()
    n = len(warped_images)
    a = np.ascontiguousarray(adj, np.float64)
    if n == 0 or n != len(corners) or a.shape != (n, n):
        raise SpanoError(_lib.E_INVALID, "Input consistency!")
    imgs = [_u8img(t, 3, f"warped_images[{i}]") for i, t in enumerate(warped_images)]
    ptr = (C.c_void_p * n)(*[t.ctypes.data for t in imgs])
    steps = (C.c_size_t * n)(*[t.strides[0] for t in imgs])
    tlx = np.array([c[0] for c in corners], np.int32); tly = np.array([c[1] for c in corners], np.int32)
    w = np.array([t.shape[1] for t in imgs], np.int32); h = np.array([t.shape[0] for t in imgs], np.int32)
    out = (_lib.OverlapInfo * (n * (n + 1) // 2))()
    cnt = C.c_int()
    ctx.check(ctx.lib.spano_overlap_intensity(ctx.h, n, ptr, steps, _ip(tlx), _ip(tly), _ip(w), _ip(h),
                                              a.ctypes.data_as(C.POINTER(C.c_double)), out, C.byref(cnt)))
    return [(o.i, o.j, o.area, o.I_i, o.I_j) for o in out[: cnt.value]]


def distance_transform(mask, ctx: Context | None = None) -> np.ndarray:
    """cv::distanceTransform(mask, dist, DIST_L2, DIST_MASK_5, CV_32F) on CV_8UC1 (5x5 chamfer, float32)."""
    ctx = ctx or default_context()
    m = _u8img(mask, 1, "mask")
    out = np.empty(m.shape, np.float32)
    ctx.check(ctx.lib.spano_distance_transform(ctx.h, m.ctypes.data, m.shape[1], m.shape[0], m.strides[0], out.ctypes.data,
                                               out.strides[0]))
    return out


def dist_cut(masks, top_lefts, ctx: Context | None = None):
    """dcut::dist_cut (src/math/_distance_cut.cpp:7-51): the seam masks that hand every overlap pixel to the image
    whose validity mask is deepest there (ties stay with both)."""
    ctx = ctx or default_context()
    n = len(masks)
    if n == 0 or n != len(top_lefts):
        raise SpanoError(_lib.E_INVALID, "Input consistency!")
    ms = [_u8img(a, 1, f"masks[{i}]") for i, a in enumerate(masks)]
    outs = [np.empty(m.shape, np.uint8) for m in ms]
    ptr = lambda arrs: (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    steps = lambda arrs: (C.c_size_t * n)(*[a.strides[0] for a in arrs])
    tlx = np.array([c[0] for c in top_lefts], np.int32); tly = np.array([c[1] for c in top_lefts], np.int32)
    w = np.array([m.shape[1] for m in ms], np.int32); h = np.array([m.shape[0] for m in ms], np.int32)
    ctx.check(ctx.lib.spano_dist_cut(ctx.h, n, ptr(ms), steps(ms), _ip(tlx), _ip(tly), _ip(w), _ip(h), ptr(outs), steps(outs)))
    return outs


def adjust_intensity(img, field, ctx: Context | None = None) -> np.ndarray:
    """test::adjust_intensity for one image (src/test/_test.cpp:110-122): resize the CV_32FC1 correction
    field to the image (float INTER_LINEAR) and divide, through 1/255 and 255 round trips, on 8 bits."""
    ctx = ctx or default_context()
    out = np.array(_u8img(img, 3, "image"), copy=True, order="C")
    f = np.ascontiguousarray(field, np.float32)
    if f.ndim != 2 or f.size == 0:
        raise SpanoError(_lib.E_INVALID, "intensity field must be a non-empty 2-D float32 array")
    ctx.check(ctx.lib.spano_adjust_intensity(ctx.h, out.ctypes.data, out.shape[1], out.shape[0], out.strides[0], f.ctypes.data,
                                             f.shape[1], f.shape[0], f.strides[0]))
    return out


def return_full(images, R, K, kind: int, focal: float, gains, masks_cut, bands: int, sigma: float,
                rows: tuple[int, int] | None = None, ctx: Context | None = None, intensities=None, center_fix=None):
    """stitch_parameters::return_full (MULTI_BLEND, gain optional): decoded sources + K/R + gains +
    mask_cut[] (tile-sized or preview-scale) [+ intensity-correction fields when conf.blend_intensity]
    -> final CV_8UC3 canvas, through the fused device path.
    `rows=(row0,row1)` restricts the result to a band of canvas rows (row-band sharding).
    `center_fix=((ansatz_x, ansatz_y), radius, quadratic)`: the little-planet centre fix (conf.fix_center with the
    stereographic projection, src/classes/_panorama.cpp:292-311) for the circle sten_proj::estimate_circle found; the
    blended tiles then have the corners / sizes of disk_reproj_size, and tile-sized masks must have those sizes."""
    ctx = ctx or default_context()
    n = len(images)
    if n == 0 or n != len(R) or n != len(K) or n != len(masks_cut):
        raise SpanoError(_lib.E_INVALID, "Input consistency!")
    imgs = [_u8img(a, 3, f"images[{i}]") for i, a in enumerate(images)]
    plan = plan_tiles(imgs, R, K, kind, focal, ctx)
    cuts = [_u8img(a, 1, f"mask_cut[{i}]") for i, a in enumerate(masks_cut)]
    corners, sizes = [p[2] for p in plan], [p[3] for p in plan]
    fix = None
    if center_fix is not None:
        (ax, ay), radius, quadratic = center_fix
        corners, sizes = disk_reproj_size(corners, sizes, (ax, ay), radius, quadratic, ctx)
        fix = _lib.CenterFix(int(ax), int(ay), float(radius), int(bool(quadratic)))
    for j in range(n):
        if cuts[j].size == 0:
            raise SpanoError(_lib.E_INVALID, f"mask_cut[{j}] is empty")
        # any size other than the (blended) tile's is taken as the preview-scale mask and resized on the device
    W, H, _, _ = pan_dimension(corners, sizes)
    row0, row1 = rows if rows is not None else (0, H)
    descs = make_descs(imgs, plan, gains, cuts)
    for j in range(n):   # (make_descs compared the mask with the WARP tile; with the centre fix the blended tile decides)
        mh, mw = int(cuts[j].shape[0]), int(cuts[j].shape[1])
        descs[j].mask_cut_w, descs[j].mask_cut_h = (0, 0) if (mw, mh) == tuple(sizes[j]) else (mw, mh)
    fields = None
    if intensities is not None:
        if len(intensities) != n:
            raise SpanoError(_lib.E_INVALID, "Input consistency!")
        fields = [np.ascontiguousarray(f, np.float32) for f in intensities]   # kept alive for the call
        for j, f in enumerate(fields):
            descs[j].intensity = f.ctypes.data
            descs[j].intensity_h, descs[j].intensity_w = int(f.shape[0]), int(f.shape[1])
            descs[j].intensity_step = f.strides[0]
    canvas = np.empty((row1 - row0, W, 3), np.uint8)
    if fix is not None:
        ctx.check(ctx.lib.spano_composite_fixed(ctx.h, int(kind), C.c_float(focal), n, descs, int(bands), float(sigma), C.byref(fix),
                                                int(row0), int(row1), canvas.ctypes.data, canvas.strides[0]))
    else:
        ctx.check(ctx.lib.spano_composite(ctx.h, int(kind), C.c_float(focal), n, descs, int(bands), float(sigma),
                                          int(row0), int(row1), canvas.ctypes.data, canvas.strides[0]))
    return canvas
