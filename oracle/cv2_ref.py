"""cv2_ref.py -- the reference's compositing path restated through the SAME OpenCV entry points.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/spano_oracle.c header for the rule):
used by oracle/gen_golden.py to pin the C oracle, by tests as a second checker when cv2
is importable, and by bench.py's cpu_baseline / --impl reference legs as the reference's
CPU implementation (same OpenCV kernels, same loop structure, redundant blurs included).

The reference (C++/OpenCV/Eigen/GTK) cannot be compiled in the build image (no OpenCV C++
headers, no Eigen, no GTK); python cv2 4.13.0 exposes the very OpenCV functions it calls.
Every function cites the reference file:line it follows (paths relative to the upstream tree).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import cv2
import numpy as np

SPHERICAL, CYLINDRICAL, STEREOGRAPHIC = 0, 1, 2
_KIND = {SPHERICAL: "spherical", CYLINDRICAL: "cylindrical", STEREOGRAPHIC: "stereographic"}


def project(kind: int, focal: float, R: np.ndarray, K: np.ndarray, img: np.ndarray):
    """proj::{spherical,cylindrical,sten}_proj::project  (src/math/_projection.cpp:27-84,297-324).

    K_adj flips the principal point (w-cx, h-cy); K and R are cast double->float32 and handed to
    cv::detail::*Warper::warp(INTER_LINEAR, BORDER_CONSTANT).  Returns (corner(x,y), warped tile).
    """
    h_ref, w_ref = img.shape[:2]
    f_i = K[0, 0]
    K_adj = np.array([[f_i, 0, w_ref - K[0, 2]], [0, f_i, h_ref - K[1, 2]], [0, 0, 1]], np.float64)
    cvK = K_adj.astype(np.float32)
    cvR = np.asarray(R, np.float64).astype(np.float32)
    warper = cv2.PyRotationWarper(_KIND[kind], float(np.float32(focal)))
    corner, warped = warper.warp(img, cvK, cvR, cv2.INTER_LINEAR, cv2.BORDER_CONSTANT)
    return corner, warped


def adjusted_camera(K: np.ndarray, R: np.ndarray, w_ref: int, h_ref: int):
    """The float32 K_adj / R handed to OpenCV (src/math/_projection.cpp:36-49)."""
    f_i = K[0, 0]
    K_adj = np.array([[f_i, 0, w_ref - K[0, 2]], [0, f_i, h_ref - K[1, 2]], [0, 0, 1]], np.float64)
    return K_adj.astype(np.float32), np.asarray(R, np.float64).astype(np.float32)


def create_surrounding_mask(img: np.ndarray) -> np.ndarray:
    """blnd::createSurroundingMask(img, invert=true, thresh=1)  (src/math/_blending.cpp:278-324)."""
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    _, thresh = cv2.threshold(gray, 1, 255, cv2.THRESH_BINARY_INV)
    flood = thresh.copy()
    h, w = flood.shape
    flags = 4 | cv2.FLOODFILL_FIXED_RANGE
    for x in range(w):
        if flood[0, x] == 255:
            cv2.floodFill(flood, None, (x, 0), 0, 0, 0, flags)
        if flood[h - 1, x] == 255:
            cv2.floodFill(flood, None, (x, h - 1), 0, 0, 0, flags)
    for y in range(h):
        if flood[y, 0] == 255:
            cv2.floodFill(flood, None, (0, y), 0, 0, 0, flags)
        if flood[y, w - 1] == 255:
            cv2.floodFill(flood, None, (w - 1, y), 0, 0, 0, flags)
    mask = cv2.subtract(thresh, flood)
    # (the diamond erode of `gray` at _blending.cpp:315-317 has no effect on the result)
    return cv2.bitwise_not(mask)


def validity_mask(warped: np.ndarray) -> np.ndarray:
    """createSurroundingMask + cv::erode(mask, Mat(), (-1,-1), 3)  (src/math/_projection.cpp:441-443)."""
    m = create_surrounding_mask(warped)
    return cv2.erode(m, None, iterations=3)


@dataclass
class ProjData:
    """proj::proj_data (src/math/_projection.h:15-19)."""
    imgs: list = field(default_factory=list)
    msks: list = field(default_factory=list)
    corners: list = field(default_factory=list)


def get_proj_parameters(images, R, K, con, kind: int, focal: float, get_masks: bool = True) -> ProjData:
    """proj::get_proj_parameters (src/math/_projection.cpp:422-454)."""
    out = ProjData()
    for i, img in enumerate(images):
        if con[i] > 0:
            corner, warped = project(kind, focal, R[i], K[i], img)
            out.imgs.append(warped)
            out.corners.append(tuple(int(c) for c in corner))
            if get_masks:
                out.msks.append(validity_mask(warped))
    return out


def get_pan_dimension(corners, imgs):
    """util::get_pan_dimension (src/system/_util.cpp:204-231) -> (width, height, min_x, min_y)."""
    min_x = min(c[0] for c in corners)
    min_y = min(c[1] for c in corners)
    max_x = max(c[0] + im.shape[1] for c, im in zip(corners, imgs))
    max_y = max(c[1] + im.shape[0] for c, im in zip(corners, imgs))
    return max_x - min_x, max_y - min_y, min_x, min_y


def apply_gain(img: np.ndarray, g: float) -> np.ndarray:
    """`imgs[i] / gain[i]` on CV_8UC3 (src/classes/_panorama.cpp:321-327) == convertTo(alpha=1/g)."""
    return _convert_scale_u8(img, 1.0 / g)


def _convert_scale_u8(img, alpha):
    # cv::Mat::convertTo(CV_8U, alpha): saturate(cvRound(float(v) * float(alpha)))
    f = img.astype(np.float32) * np.float32(alpha)
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)


def resize_mask(mask: np.ndarray, size_wh) -> np.ndarray:
    """cv::resize(mask_cut[i], dst, size, cv::INTER_CUBIC) -- INTER_CUBIC lands in the `fx` slot, so
    the interpolation is the default INTER_LINEAR (src/classes/_panorama.cpp:329-335)."""
    return cv2.resize(mask, size_wh, interpolation=cv2.INTER_LINEAR)


def adjust_intensity(img: np.ndarray, field: np.ndarray) -> np.ndarray:
    """test::adjust_intensity for one image (src/test/_test.cpp:110-122): resize the float correction
    field to the tile, tile/255 -> divide by the (clamped) field -> *255 -> CV_8UC3."""
    f = cv2.resize(np.asarray(field, np.float32), (img.shape[1], img.shape[0]), interpolation=cv2.INTER_LINEAR)
    x = img.astype(np.float32) * np.float32(1.0 / 255.0)
    x = elementwise_divide(x, f)
    return np.clip(np.rint(x * np.float32(255.0)), 0, 255).astype(np.uint8)


def elementwise_divide(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """imgm::elementwiseOperation(DIVIDE) (src/math/_img_manipulation.cpp:31-84)."""
    d = np.copysign(np.maximum(np.abs(B), np.float32(1e-6)), B).astype(np.float32)
    inv = (np.float32(1.0) / d).astype(np.float32)
    return (A * inv[..., None]).astype(np.float32)


def multi_blend(images, masks, masks_orig, top_lefts, bands: int, sigma: float, timers=None) -> np.ndarray:
    """blnd::multi_blend (src/math/_blending.cpp:186-252) through cv2.GaussianBlur, same loop
    structure (bands outer, images inner, B(s_{i+1}) re-blurred for the middle bands)."""
    W, H, min_x, min_y = get_pan_dimension(top_lefts, images)
    acc_c = np.zeros((H, W, 3), np.float32)
    acc_a = np.zeros((H, W), np.float32)
    k = 2 * math.ceil(3 * sigma) + 1
    inv255 = np.float32(1.0 / 255.0)
    for i in range(bands):
        sigma_band = math.sqrt(2 * (bands - i) + 1) * sigma
        for j in range(len(masks)):
            i_conv = images[j].astype(np.float32)
            w_conv = masks[j].astype(np.float32)
            I_temp = cv2.GaussianBlur(i_conv, (k, k), sigmaX=sigma_band, sigmaY=sigma_band, borderType=cv2.BORDER_REFLECT)
            w_conv = cv2.GaussianBlur(w_conv, (k, k), sigmaX=sigma_band, sigmaY=sigma_band, borderType=cv2.BORDER_REFLECT)
            w_conv = w_conv * inv255
            if i == bands - 1:
                I_temp = i_conv - I_temp
            elif i > 0:
                sigma_prev = math.sqrt(2 * (bands - i - 1) + 1) * sigma
                prev_I = cv2.GaussianBlur(i_conv, (k, k), sigmaX=sigma_prev, sigmaY=sigma_prev, borderType=cv2.BORDER_REFLECT)
                I_temp = I_temp - prev_I
            w_conv[masks_orig[j] != 255] = 0
            color_tmp = I_temp * w_conv[..., None]
            x0 = top_lefts[j][0] - min_x
            y0 = top_lefts[j][1] - min_y
            h, w = images[j].shape[:2]
            acc_c[y0:y0 + h, x0:x0 + w] += color_tmp
            acc_a[y0:y0 + h, x0:x0 + w] += w_conv
    out = elementwise_divide(acc_c, acc_a)
    divisor = np.float32(255 // bands)
    return out * np.float32(1.0 / float(divisor))


def blend_to_u8(blend: np.ndarray) -> np.ndarray:
    """stitch_parameters::blend MULTI_BLEND tail: blend*255; convertTo(CV_8UC3) (_panorama.cpp:242-249)."""
    return np.clip(np.rint(blend * np.float32(255.0)), 0, 255).astype(np.uint8)


def dist_cut(masks, corners):
    """dcut::dist_cut (src/math/_distance_cut.cpp:7-73): chamfer-5 L2 distance compare in overlaps.
    Produces the `mask_cut[]` INPUT of the hot path (not part of the accelerated path)."""
    # `transformed / 255` is a cv::MatExpr: convertTo with alpha = 1/255., i.e. a float multiplication by (float)(1/255.)
    D = [cv2.distanceTransform(m, cv2.DIST_L2, cv2.DIST_MASK_5) * np.float32(1.0 / 255.0) for m in masks]
    out = [m.copy() for m in masks]
    n = len(masks)
    for i in range(n):
        hi, wi = masks[i].shape
        for j in range(n):
            if i == j:
                continue
            hj, wj = masks[j].shape
            x0 = max(corners[i][0], corners[j][0]); y0 = max(corners[i][1], corners[j][1])
            x1 = min(corners[i][0] + wi, corners[j][0] + wj); y1 = min(corners[i][1] + hi, corners[j][1] + hj)
            if x1 <= x0 or y1 <= y0:
                continue
            a = D[i][y0 - corners[i][1]:y1 - corners[i][1], x0 - corners[i][0]:x1 - corners[i][0]]
            b = D[j][y0 - corners[j][1]:y1 - corners[j][1], x0 - corners[j][0]:x1 - corners[j][0]]
            o = out[i][y0 - corners[i][1]:y1 - corners[i][1], x0 - corners[i][0]:x1 - corners[i][0]]
            o[(a - b) < 0] = 0
    return out


def return_full(images, R, K, kind, focal, gains, masks_cut_fullres, bands, sigma, timers=None):
    """stitch_parameters::return_full, MULTI_BLEND (src/classes/_panorama.cpp:259-354) from decoded
    sources to the final 8-bit canvas: warp + validity masks + gain + multi_blend + convert.
    `masks_cut_fullres[j]` is mask_cut[j] already resized to tile j (the resize is a "next" row)."""
    import time
    t0 = time.perf_counter()
    con = [1.0] * len(images)
    pd = get_proj_parameters(images, R, K, con, kind, focal)
    t1 = time.perf_counter()
    imgs = pd.imgs
    if gains is not None:
        imgs = [_convert_scale_u8(im, 1.0 / g) for im, g in zip(imgs, gains)]
    blend = multi_blend(imgs, masks_cut_fullres, pd.msks, pd.corners, bands, sigma)
    out = blend_to_u8(blend)
    t2 = time.perf_counter()
    if timers is not None:
        timers["warp_s"] = t1 - t0
        timers["blend_s"] = t2 - t1
    return out, pd


def scale_rect(r, xs: float, ys: float):
    """util::scaleRect (src/system/_util.cpp:157-166): (x, y, w, h) with std::round (half away from zero)."""
    rnd = lambda v: int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)
    return rnd(r[0] * xs), rnd(r[1] * ys), rnd(r[2] * xs), rnd(r[3] * ys)


def equalize_intensities(images, masks, top_lefts, ratio: float = 0.5):
    """test::equalizeIntensities (src/test/_test.cpp:9-106): the per-image intensity-correction fields (CV_32FC1 at
    `ratio` of the preview size) that test::adjust_intensity later divides the tiles by.  Same OpenCV calls in the same
    order; MatExpr / scalar arithmetic on CV_32F is float arithmetic."""
    n = len(images)
    inv255 = np.float32(1.0 / 255.0)
    eps = np.float32(0.00001)
    D = [cv2.distanceTransform(m, cv2.DIST_L2, cv2.DIST_MASK_5) * inv255 for m in masks]     # dcut::distance_transform
    _, _, min_x, min_y = get_pan_dimension(top_lefts, images)
    msk_s, inten, idist, rois = [], [], [], []
    for i in range(n):
        ms = cv2.resize(masks[i], None, fx=ratio, fy=ratio, interpolation=cv2.INTER_LINEAR)
        im = cv2.resize(images[i], None, fx=ratio, fy=ratio, interpolation=cv2.INTER_LINEAR)
        D[i] = cv2.resize(D[i], None, fx=ratio, fy=ratio, interpolation=cv2.INTER_LINEAR)
        gray = cv2.cvtColor(im, cv2.COLOR_BGR2GRAY).astype(np.float32) * inv255
        gm = np.where(ms != 0, gray, np.float32(0)).astype(np.float32)
        msk_s.append(ms); inten.append(gm); idist.append((gm * D[i]).astype(np.float32))
        roi = (top_lefts[i][0] - min_x, top_lefts[i][1] - min_y, masks[i].shape[1], masks[i].shape[0])
        rois.append(scale_rect(roi, im.shape[1] / images[i].shape[1], im.shape[0] / images[i].shape[0]))
    out = []
    for i in range(n):
        alpha, it = D[i].copy(), idist[i].copy()
        for j in range(n):
            if i == j:
                continue
            x0, y0 = max(rois[i][0], rois[j][0]), max(rois[i][1], rois[j][1])
            x1 = min(rois[i][0] + rois[i][2], rois[j][0] + rois[j][2]); y1 = min(rois[i][1] + rois[i][3], rois[j][1] + rois[j][3])
            if x1 <= x0 or y1 <= y0:
                continue
            si = (slice(y0 - rois[i][1], y1 - rois[i][1]), slice(x0 - rois[i][0], x1 - rois[i][0]))
            sj = (slice(y0 - rois[j][1], y1 - rois[j][1]), slice(x0 - rois[j][0], x1 - rois[j][0]))
            on = msk_s[i][si] != 0
            it[si] = np.where(on, it[si] + idist[j][sj], it[si])
            alpha[si] = np.where(on, alpha[si] + D[j][sj], alpha[si])
        alpha = alpha + eps
        test = cv2.divide(it, alpha)
        test = test + eps
        test = cv2.divide(inten[i], test)
        test = test + (255 - msk_s[i]).astype(np.float32) * inv255
        out.append(cv2.GaussianBlur(test, (13, 13), sigmaX=7, sigmaY=7, borderType=cv2.BORDER_REFLECT))
    return out


def analyze_components_with_circles(image: np.ndarray, min_area: float):
    """util::analyzeComponentsWithCircles for a CV_8UC1 image (src/system/_util.cpp:8-81): connected components of the
    ZERO pixels, and for every component of at least min_area pixels the minimum enclosing circle of its first external
    contour.  Returns [(center(x, y), radius, distance from the image centre)]."""
    mask = (255 - image).astype(np.uint8)          # Mat::ones * 255 - image
    if cv2.countNonZero(mask) == 0:
        raise RuntimeError("No connected components found: mask is entirely zero")
    cx, cy = np.float32(image.shape[1] / 2.0), np.float32(image.shape[0] / 2.0)
    num, labels, stats, _ = cv2.connectedComponentsWithStats(mask)
    if num <= 1:
        raise RuntimeError("No connected components found")
    out = []
    for i in range(1, num):
        if stats[i, cv2.CC_STAT_AREA] >= min_area:
            comp = np.where(labels == i, 255, 0).astype(np.uint8)
            contours, _ = cv2.findContours(comp, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
            if contours:
                (x, y), r = cv2.minEnclosingCircle(contours[0])
                dx, dy = np.float32(x) - cx, np.float32(y) - cy
                out.append(((np.float32(x), np.float32(y)), np.float32(r), np.float32(math.sqrt(float(dx * dx + dy * dy)))))
    if not out:
        raise RuntimeError("No components found with area >= %g" % min_area)
    return out


def estimate_circle(pd: ProjData):
    """sten_proj::estimate_circle (src/math/_projection.cpp:361-419): the hole of the union of the validity masks that
    lies nearest the canvas centre.  Returns ((ansatz_x, ansatz_y), radius) with cv::Point's rounding of the centre and
    the reference's +3 offset, or (None, -1.0) when there is no midsection."""
    W, H, min_x, min_y = get_pan_dimension(pd.corners, pd.imgs)
    test = np.zeros((H, W), np.uint8)
    for c, m in zip(pd.corners, pd.msks):
        x0, y0 = c[0] - min_x, c[1] - min_y
        roi = test[y0:y0 + m.shape[0], x0:x0 + m.shape[1]]
        roi[m != 0] = m[m != 0]                     # copyTo(dst, mask)
    circles = analyze_components_with_circles(test, 100)
    dist = np.float32(math.sqrt(float(W * W + H * H)))
    cutoff, cutoff_dist = (dist / 2) * np.float32(.5), (dist / 2) * np.float32(.2)
    best = None
    for j, (_, _, d) in enumerate(circles):
        if dist > d:
            dist, best = d, j
    if best is None or dist > cutoff_dist or circles[best][1] > cutoff:
        return None, -1.0
    (x, y), r, _ = circles[best]
    return (int(np.rint(x)), int(np.rint(y))), float(np.float32(r) + np.float32(3))   # cv::Point(Point2f) rounds to nearest


def simple_blend(images, masks, top_lefts) -> np.ndarray:
    """blnd::simple_blend (src/math/_blending.cpp:83-153) through the same OpenCV entry points
    (distanceTransform, normalize(NORM_MINMAX), convertTo, mul, subtract) -> CV_8UC3 canvas."""
    W, H, min_x, min_y = get_pan_dimension(top_lefts, images)
    acc_c = np.zeros((H, W, 3), np.float32)
    acc_a = np.zeros((H, W), np.float32)
    for img, mask, tl in zip(images, masks, top_lefts):
        x0, y0 = tl[0] - min_x, tl[1] - min_y
        h, w = img.shape[:2]
        dt = cv2.distanceTransform(np.ascontiguousarray(mask), cv2.DIST_L2, cv2.DIST_MASK_5)
        mask_float = cv2.normalize(dt, None, 0.0, 1.0, cv2.NORM_MINMAX)
        img_float = img.astype(np.float32) * np.float32(1.0 / 255.0)          # convertTo(CV_32F, 1/255.)
        one_minus = cv2.subtract(np.float64(1.0), acc_a[y0:y0 + h, x0:x0 + w]).astype(np.float32)
        new_color = img_float * mask_float[..., None]
        acc_c[y0:y0 + h, x0:x0 + w] += new_color * one_minus[..., None]
        acc_a[y0:y0 + h, x0:x0 + w] += mask_float * one_minus
    res = np.zeros((H, W, 3), np.float32)
    pos = acc_a > 0
    inv = np.zeros_like(acc_a)
    inv[pos] = np.float32(1.0) / acc_a[pos]                                   # Vec3f / float == * (1.f / a)
    res[pos] = acc_c[pos] * inv[pos][..., None]
    return np.clip(np.rint(res * np.float32(255.0)), 0, 255).astype(np.uint8)  # convertTo(CV_8UC3, 255.0)


def no_blend(images, masks, top_lefts) -> np.ndarray:
    """blnd::no_blend (src/math/_blending.cpp:157-182): images[i].copyTo(panorama(roi), masks[i]) in order."""
    W, H, min_x, min_y = get_pan_dimension(top_lefts, images)
    pano = np.zeros((H, W, 3), np.uint8)
    for img, mask, tl in zip(images, masks, top_lefts):
        x0, y0 = tl[0] - min_x, tl[1] - min_y
        h, w = img.shape[:2]
        roi = pano[y0:y0 + h, x0:x0 + w]
        cv2.copyTo(img, mask, roi)
    return pano


def get_overlapp_intensity(warped_images, corners, adj):
    """gain::get_overlapp_intensity (src/math/_gain_compensation.cpp:7-75) through cv2 (cvtColor, bitwise_and,
    countNonZero, sum) -> list of (i, j, area, I_i, I_j)."""
    n = len(warped_images)
    adj = np.asarray(adj, np.float64) + np.eye(n)
    grays = [cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) for img in warped_images]
    masks = [create_surrounding_mask(img) for img in warped_images]
    res = []
    for i in range(n):
        for j in range(i, n):
            if not adj[i, j] > 0:
                continue
            hi, wi = grays[i].shape; hj, wj = grays[j].shape
            x0 = max(corners[i][0], corners[j][0]); y0 = max(corners[i][1], corners[j][1])
            x1 = min(corners[i][0] + wi, corners[j][0] + wj); y1 = min(corners[i][1] + hi, corners[j][1] + hj)
            if x1 <= x0 or y1 <= y0:
                res.append((i, j, 0.0, 0.0, 0.0))
                continue
            si = (slice(y0 - corners[i][1], y1 - corners[i][1]), slice(x0 - corners[i][0], x1 - corners[i][0]))
            sj = (slice(y0 - corners[j][1], y1 - corners[j][1]), slice(x0 - corners[j][0], x1 - corners[j][0]))
            comb = cv2.bitwise_and(np.ascontiguousarray(masks[i][si]), np.ascontiguousarray(masks[j][sj]))
            area = float(cv2.countNonZero(comb))
            gi = np.ascontiguousarray(grays[i][si]); gj = np.ascontiguousarray(grays[j][sj])
            Ii = cv2.sumElems(cv2.bitwise_and(gi, gi, mask=comb))[0]
            Ij = cv2.sumElems(cv2.bitwise_and(gj, gj, mask=comb))[0]
            res.append((i, j, area, float(Ii), float(Ij)))
    return res
