"""Golden vectors of test::equalizeIntensities (reference src/test/_test.cpp:9-106) from OpenCV 4.13 (oracle/cv2_ref.py):

    python oracle/gen_golden_equalize.py   ->  tests/golden/equalize.npz

Two layouts: preview tiles with even sizes in both directions (cv::resize by 0.5 takes the 2x2 area-average path) and
with odd sizes (the fixed-point / float linear path).  Stored: the inputs (tiles, validity masks, corners) and the fields.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cv2_ref  # noqa: E402


def case(seed, sizes, corners):
    rng = np.random.default_rng(seed)
    tiles, masks = [], []
    for k, (w, h) in enumerate(sizes):
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        base = 120 + 60 * np.sin(xx / 23.0 + k) * np.cos(yy / 17.0) + rng.normal(0, 4, (h, w))
        t = np.clip(np.stack([base * (0.8 + 0.1 * c) * (0.85 + 0.1 * k) for c in range(3)], -1), 3, 255).astype(np.uint8)
        m = np.full((h, w), 255, np.uint8)
        m[: 4 + 3 * k] = 0; m[:, : 5 + k] = 0; m[h - 6:] = 0     # un-eroded borders like warped tiles have
        m[h // 3: h // 3 + 9, w // 2: w // 2 + 14] = 0          # a hole
        t[m == 0] = 0
        tiles.append(t); masks.append(m)
    fields = cv2_ref.equalize_intensities(tiles, masks, corners, 0.5)
    return tiles, masks, fields


def main():
    out = {}
    layouts = {"even": ([(96, 64), (80, 64), (100, 70)], [(0, 0), (60, 6), (120, -4)]),
               "odd": ([(97, 65), (81, 63), (101, 71)], [(-3, 2), (58, 7), (117, -5)])}
    for name, (sizes, corners) in layouts.items():
        tiles, masks, fields = case(7 if name == "even" else 8, sizes, corners)
        out[f"{name}_n"] = np.int32(len(sizes))
        out[f"{name}_corners"] = np.array(corners, np.int32)
        for k in range(len(sizes)):
            out[f"{name}_tile{k}"] = tiles[k]; out[f"{name}_mask{k}"] = masks[k]; out[f"{name}_field{k}"] = fields[k]
    path = os.path.join(ROOT, "tests", "golden", "equalize.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes", [out[f"even_field{k}"].shape for k in range(3)], [out[f"odd_field{k}"].shape for k in range(3)])


if __name__ == "__main__":
    main()
