"""Golden vectors for the distance-based seam search, produced by the SAME OpenCV
entry point the reference calls: cv2.distanceTransform(mask, DIST_L2, DIST_MASK_5) (cv2 4.13.0 + IPP, the pinned
build), and dcut::dist_cut restated over it (oracle/cv2_ref.py).  Writes tests/golden/dist.npz.

    python oracle/gen_golden_dist.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cv2_ref  # noqa: E402


def validity_like(rng, w, h):
    """a warped tile's validity mask: 255 inside a curved region, 0 outside, a few interior holes"""
    yy, xx = np.mgrid[0:h, 0:w]
    top = 6 + 10 * np.sin(xx / w * np.pi) + rng.integers(0, 2)
    bot = h - 5 - 8 * np.sin(xx / w * np.pi)
    m = ((yy > top) & (yy < bot) & (xx > 3) & (xx < w - 4)).astype(np.uint8) * 255
    for _ in range(3):
        cx, cy, r = rng.integers(0, w), rng.integers(0, h), rng.integers(1, 6)
        m[(xx - cx) ** 2 + (yy - cy) ** 2 < r * r] = 0
    return m


def main():
    rng = np.random.default_rng(2024)
    out = {}
    masks = {
        "sparse_zeros": (rng.random((61, 83)) > 0.03).astype(np.uint8) * 255,
        "dense_zeros": (rng.random((40, 57)) > 0.6).astype(np.uint8) * 255,
        "nonbinary": rng.integers(0, 3, (33, 47)).astype(np.uint8),           # any non-zero value counts as foreground
        "all_set": np.full((9, 13), 255, np.uint8),                           # no zero pixel at all -> FLT_MAX
        "all_zero": np.zeros((7, 5), np.uint8),
        "single_row": np.where(np.arange(31)[None, :] == 11, 0, 255).astype(np.uint8),
        "single_col": np.where(np.arange(29)[:, None] == 3, 0, 255).astype(np.uint8),
        "one_pixel": np.full((1, 1), 255, np.uint8),
        "far_corner": np.pad(np.zeros((1, 1), np.uint8), ((0, 150), (0, 201)), constant_values=255),
        "validity": validity_like(rng, 211, 140),
    }
    names = sorted(masks)
    out["dt_names"] = np.array(names)
    for k in names:
        out[f"dt_mask_{k}"] = masks[k]
        out[f"dt_ref_{k}"] = cv2.distanceTransform(masks[k], cv2.DIST_L2, cv2.DIST_MASK_5)
    # dist_cut: five overlapping validity masks (one pair disjoint, negative corners)
    sizes = [(120, 90), (100, 110), (90, 60), (70, 150), (40, 40)]
    corners = [(0, 0), (80, 40), (150, -10), (30, 60), (400, 300)]
    ms = [validity_like(rng, w, h) for (w, h) in sizes]
    cuts = cv2_ref.dist_cut(ms, corners)
    out["cut_corners"] = np.array(corners, np.int32)
    for i, (m, c) in enumerate(zip(ms, cuts)):
        out[f"cut_mask_{i}"] = m
        out[f"cut_ref_{i}"] = c
    # simple_blend / no_blend (the other two branches of stitch_parameters::blend) on the same geometry
    tiles = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (w, h) in sizes]
    bmasks = list(ms)
    bmasks[2] = np.full_like(ms[2], 255)                # a mask without any zero pixel: distance FLT_MAX, alpha 0
    for i, (t, m) in enumerate(zip(tiles, bmasks)):
        out[f"blend_tile_{i}"] = t
        out[f"blend_mask_{i}"] = m
    out["simple_ref"] = cv2_ref.simple_blend(tiles, bmasks, corners)
    out["noblend_ref"] = cv2_ref.no_blend(tiles, bmasks, corners)
    # gain::get_overlapp_intensity on warped-like tiles (dark outside a curved region, a few dark blobs inside)
    wt = []
    for t, m in zip(tiles, ms):
        t2 = np.maximum(t, 8)
        t2[m == 0] = 0
        wt.append(np.ascontiguousarray(t2))
    adj = np.zeros((len(wt), len(wt)))
    adj[0, 1] = adj[1, 0] = adj[0, 3] = adj[3, 0] = adj[1, 2] = adj[2, 1] = adj[1, 3] = adj[3, 1] = 1
    adj[2, 4] = adj[4, 2] = 1                             # adjacent in the graph, but the rectangles do not overlap
    for i, t in enumerate(wt):
        out[f"ov_tile_{i}"] = t
    out["ov_adj"] = adj
    out["ov_ref"] = np.array(cv2_ref.get_overlapp_intensity(wt, corners, adj), np.float64)
    path = os.path.join(ROOT, "tests", "golden", "dist.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes, cv2", cv2.__version__)


if __name__ == "__main__":
    main()
