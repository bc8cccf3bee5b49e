"""Generate tests/golden/*.npz from OpenCV itself (python cv2 4.13.0) through oracle/cv2_ref.py.

Run in the BUILD container only (needs cv2):   python oracle/gen_golden.py
The reference has no tests or golden vectors of its own, and it cannot be
compiled here (no OpenCV C++ headers / Eigen / GTK), so these vectors -- produced by the very
OpenCV functions the reference calls, in the reference's call order -- are the parity pin for
both the C oracle and the CUDA path.  The fixtures are committed; this script documents how
they were made and regenerates them bit-identically (everything is seeded).
"""
from __future__ import annotations

import math
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cv2_ref as ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def rot(yaw, pitch, roll):
    cy, sy = math.cos(yaw), math.sin(yaw)
    cp, sp = math.cos(pitch), math.sin(pitch)
    cr, sr = math.cos(roll), math.sin(roll)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    Rz = np.array([[cr, -sr, 0], [sr, cr, 0], [0, 0, 1]])
    return Ry @ Rx @ Rz


def pattern(rng, h, w, j, noise=0):
    x = np.arange(w, dtype=np.float32); y = np.arange(h, dtype=np.float32)
    img = np.empty((h, w, 3), np.float32)
    for c in range(3):
        img[..., c] = 128 + 100 * np.outer(np.cos(y / (9 + 3 * c)), np.sin(x / (11 + 5 * c) + j))
    if noise:
        img += rng.integers(-noise, noise + 1, img.shape)
    return np.clip(np.rint(img), 16, 240).astype(np.uint8)


def gen_roi_table():
    """KAT (ii): ROI / corner for a table of poses per projection, incl. pole-in-view and seam-straddle."""
    rows = []
    poses = [(0.0, 0.0, 0.0), (0.3, 0.1, 0.01), (-2.9, -0.4, -0.02), (3.1, 0.0, 0.0), (1.0, -0.9, 0.0), (0.0, -1.2, 0.3),
             (0.0, -1.5, 0.0), (0.0, 1.45, 0.1), (2.0, 0.7, -0.3), (-1.3, -0.2, 0.05), (0.5, -0.6, 1.2), (-3.0, 0.3, -0.01)]
    for kind in range(3):
        for (yaw, pitch, roll) in poses:
            if kind == ref.STEREOGRAPHIC and pitch > 0:
                pitch = -pitch  # looking up sends the stereographic plane to infinity
            for (W, H, f) in ((320, 240, 300.0), (200, 300, 180.0)):
                K = np.array([[f * 1.02, 0, W / 2 + 1.5], [0, f * 1.02, H / 2 - 0.75], [0, 0, 1]], np.float32)
                R = rot(yaw, pitch, roll).astype(np.float32)
                w = cv2.PyRotationWarper(ref._KIND[kind], f)
                x, y, ww, hh = w.warpRoi((W, H), K, R)
                rows.append(np.concatenate([[kind, W, H, f], K.ravel(), R.ravel(), [x, y, ww, hh]]))
    return np.array(rows, np.float64)


def gen_warp_cases(rng):
    """KAT (i)/(vi): tiny warps per projection: source, K/R (reference-style doubles), tile, corner, mask."""
    out = {}
    cases = [(ref.SPHERICAL, 0.4, 0.15, 0.02), (ref.SPHERICAL, 0.0, 1.35, 0.0), (ref.CYLINDRICAL, -0.7, 0.1, -0.03),
             (ref.STEREOGRAPHIC, 0.9, -0.7, 0.05), (ref.STEREOGRAPHIC, 0.0, -1.45, 0.0)]
    for idx, (kind, yaw, pitch, roll) in enumerate(cases):
        W, H, f = 96, 64, 80.0
        img = pattern(rng, H, W, idx, noise=(30 if idx % 2 == 0 else 0))
        if idx == 0:
            img[20:30, 40:55] = 0          # a dark island inside (must stay valid)
            img[:6, 10:50] = 1             # dark pixels touching the source border
        K = np.array([[f * 1.1, 0, W / 2 + 0.5], [0, f * 1.1, H / 2 - 0.25], [0, 0, 1]], np.float64)
        R = rot(yaw, pitch, roll)
        corner, tile = ref.project(kind, f, R, K, img)
        mask = ref.validity_mask(tile)
        out[f"warp{idx}_kind"] = np.array(kind)
        out[f"warp{idx}_focal"] = np.array(f)
        out[f"warp{idx}_img"] = img
        out[f"warp{idx}_K"] = K
        out[f"warp{idx}_R"] = R
        out[f"warp{idx}_corner"] = np.array(corner, np.int32)
        out[f"warp{idx}_tile"] = tile
        out[f"warp{idx}_mask"] = mask
    out["warp_count"] = np.array(len(cases))
    return out


def gen_remap_case(rng):
    """KAT (i): remap of seeded noise with seeded maps incl. out-of-range / NaN / huge coordinates (exact)."""
    H, W = 37, 53
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    dh, dw = 40, 61
    xm = rng.uniform(-3, W + 3, (dh, dw)).astype(np.float32)
    ym = rng.uniform(-3, H + 3, (dh, dw)).astype(np.float32)
    xm[0, :8] = [-1, -0.5, W - 1, W - 0.5, 1e9, -1e9, np.nan, 0.015625]
    ym[0, :8] = [-1, 0, H - 1, H - 0.5, 5, 5, 5, 31.984375]
    dst = cv2.remap(img, xm, ym, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    return {"remap_img": img, "remap_x": xm, "remap_y": ym, "remap_dst": dst}


def gen_mask_case(rng):
    H, W = 70, 90
    img = rng.integers(16, 240, (H, W, 3), dtype=np.uint8)
    img[:12, :] = 0; img[:, :9] = 0; img[30:40, 40:60] = 0
    img[55:, 70:] = rng.integers(0, 3, (15, 20, 3))
    img[20:24, :30] = 0
    img[44:47, 50:] = 1
    yy, xx = np.mgrid[:H, :W]
    img[(yy - 50) ** 2 + (xx - 25) ** 2 < 36] = 0     # an island
    return {"mask_img": img, "mask_raw": ref.create_surrounding_mask(img), "mask_eroded": ref.validity_mask(img),
            "gray": cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)}


def gen_taps():
    """KAT (iii): taps for sigma 7, bands 1..10 (+ the divisor table, KAT iv)."""
    taps = np.zeros((10, 10, 43), np.float32)
    for B in range(1, 11):
        for i in range(B):
            taps[B - 1, i] = cv2.getGaussianKernel(43, math.sqrt(2 * (B - i) + 1) * 7.0, cv2.CV_32F).ravel()
    return {"taps_sigma7": taps, "divisors": np.array([255 // b for b in range(1, 11)], np.int32)}


def gen_blend_case(rng):
    """KAT (v)/(vi): tiny multi_blend (3 overlapping tiles, soft seam masks) for bands 1, 2, 3 and 5,
    plus a single-tile all-ones-mask case; float canvas straight from cv2.GaussianBlur."""
    out = {}
    sizes = [(70, 50), (64, 58), (30, 17)]   # the last one is smaller than the blur radius (reflect bounces)
    corners = [(-20, 5), (25, -3), (60, 30)]
    tiles, cuts, valids = [], [], []
    for j, (w, h) in enumerate(sizes):
        tiles.append(pattern(rng, h, w, j, noise=10))
        v = np.full((h, w), 255, np.uint8); v[:3, :] = 0; v[:, :2] = 0
        if j == 1:
            v[20:30, 10:20] = 0
        valids.append(v)
        c = np.zeros((h, w), np.uint8)
        c[:, : (2 * w) // 3] = 255
        c = cv2.GaussianBlur(c, (9, 9), 0)
        c[v == 0] = 0
        cuts.append(c)
    for j in range(3):
        out[f"blend_tile{j}"] = tiles[j]; out[f"blend_cut{j}"] = cuts[j]; out[f"blend_valid{j}"] = valids[j]
    out["blend_corners"] = np.array(corners, np.int32)
    for B in (1, 2, 3, 5):
        f = ref.multi_blend(tiles, cuts, valids, corners, B, 7.0)
        out[f"blend_f32_B{B}"] = f
        out[f"blend_u8_B{B}"] = ref.blend_to_u8(f)
    # other sigma (generic-radius kernel): sigma 3 -> 19 taps
    f = ref.multi_blend(tiles, cuts, valids, corners, 4, 3.0)
    out["blend_f32_B4_s3"] = f
    out["blend_u8_B4_s3"] = ref.blend_to_u8(f)
    return out


def gen_misc(rng):
    a = np.arange(256, dtype=np.uint8)
    gains = np.array([0.8, 0.937, 1.0, 1.25, 3.0, 0.31])
    table = np.stack([cv2.convertScaleAbs(a.reshape(1, -1), alpha=1.0 / g).ravel() for g in gains])
    src = (rng.random((23, 31)) > 0.5).astype(np.uint8) * 255
    return {"gain_values": gains, "gain_table": table, "resize_src": src,
            "resize_dst": cv2.resize(src, (200, 117), interpolation=cv2.INTER_LINEAR)}


def gen_cubic_gather():
    """KAT for a5: cv::remap(INTER_CUBIC, BORDER_CONSTANT) on INTEGER-valued maps (what
    sten_proj::disk_reproj feeds it, src/math/_projection.cpp:259-278) is an exact pixel gather."""
    rng = np.random.default_rng(7)
    H, W = 41, 57
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    xm = rng.integers(-4, W + 4, (50, 66)).astype(np.float32)
    ym = rng.integers(-4, H + 4, (50, 66)).astype(np.float32)
    dst = cv2.remap(img, xm, ym, cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    return {"cubic_img": img, "cubic_x": xm, "cubic_y": ym, "cubic_dst": dst}


def gen_intensity():
    """KAT for test::adjust_intensity (src/test/_test.cpp:110-122): float field resized by cv2.resize,
    tile/255 -> divide -> *255 -> u8."""
    rng = np.random.default_rng(11)
    field = cv2.GaussianBlur((0.6 + 0.8 * rng.random((23, 31))).astype(np.float32), (13, 13), 7, borderType=cv2.BORDER_REFLECT)
    field[3, 4] = 0.0          # exercises the 1e-6 clamp of elementwiseOperation(DIVIDE)
    img = rng.integers(0, 256, (97, 141, 3), dtype=np.uint8)
    return {"int_field": field, "int_img": img, "int_field_resized": cv2.resize(field, (141, 97), interpolation=cv2.INTER_LINEAR),
            "int_out": ref.adjust_intensity(img, field)}


def main():
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "cubic_gather.npz"), **gen_cubic_gather())
    np.savez_compressed(os.path.join(OUT, "intensity.npz"), **gen_intensity())
    rng = np.random.default_rng(20261018)
    np.savez_compressed(os.path.join(OUT, "roi_table.npz"), table=gen_roi_table(), cv2_version=np.array(cv2.__version__))
    np.savez_compressed(os.path.join(OUT, "warp_cases.npz"), **gen_warp_cases(rng))
    d = {}
    d.update(gen_remap_case(rng)); d.update(gen_mask_case(rng)); d.update(gen_taps()); d.update(gen_misc(rng))
    np.savez_compressed(os.path.join(OUT, "kernels.npz"), **d)
    np.savez_compressed(os.path.join(OUT, "blend_cases.npz"), **gen_blend_case(rng))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
