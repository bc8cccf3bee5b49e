"""Seeded synthetic image sets for the BASELINE.json configs.

Inputs only: band-limited colour patterns (so a one-bin slip of the 1/32-px sampler stays below
1 LSB), per-image exposure factors with matching gains, camera matrices laid out like the named
panorama, and soft-edged seam masks standing in for the reference's preview-scale
`mask_cut` after its resize to tile size.  Nothing here is on the timed path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# projection kinds (include/spano.h SPANO_SPHERICAL / _CYLINDRICAL / _STEREOGRAPHIC).  This module imports nothing
# from the package: bench.py's reference arm loads it by path so that the CPU arm never touches the CUDA library.
SPHERICAL, CYLINDRICAL, STEREOGRAPHIC = 0, 1, 2


@dataclass
class Config:
    name: str
    kind: int
    n: int
    width: int
    height: int
    focal: float
    bands: int
    sigma: float
    yaw_deg: list
    pitch_deg: list
    seed: int
    description: str


def _ring(n, start, step):
    return [start + step * k for k in range(n)]


def config(name: str, scale: float = 1.0) -> Config:
    """Named workloads.  `scale` < 1 shrinks image size and focal length together (same field of
    view, same layout) for parity tests."""
    s = scale
    if name == "cfg1":  # 6 x 1920x1080, spherical, 5 bands
        hfov = 2 * math.degrees(math.atan(1920 / 2 / 1500.0))
        yaw = [(k - 2.5) * 0.6 * hfov for k in range(6)]
        c = Config(name, SPHERICAL, 6, 1920, 1080, 1500.0, 5, 7.0, yaw, [0.0] * 6, 1,
                   "6 synthetic 1920x1080 RGB images, spherical, 5-band multiband")
    elif name == "cfg2":  # 24 x 24MP, ~353 deg cylindrical ring, 6 bands (no tile crosses +-pi)
        c = Config(name, CYLINDRICAL, 24, 6000, 4000, 6000.0, 6, 7.0, _ring(24, -150.0, 13.04), [0.0] * 24, 2,
                   "24 synthetic 24MP images, 360deg cylindrical panorama, 6-band multiband")
    elif name == "cfg2a":  # edge-case parity: true 360 deg ring, the tiles that straddle the +-pi seam come out full-width
        c = Config(name, CYLINDRICAL, 24, 6000, 4000, 5000.0, 6, 7.0, _ring(24, -172.5, 15.0), [0.0] * 24, 21,
                   "24 synthetic 24MP images, true 360deg cylindrical ring (f=5000): seam-straddling tiles are full-width")
    elif name == "cfg3":  # 36 x 12MP stereographic little planet, 7 bands (looking down; the lowest ring stops 5 degrees
        # short of the nadir, so the planet has the hole in its middle that sten_proj::estimate_circle looks for and the
        # centre fix closes)
        yaw = [30.0 * (k % 12) for k in range(36)]
        pitch = [(-10.0, -35.0, -48.0)[k // 12] for k in range(36)]
        c = Config(name, STEREOGRAPHIC, 36, 4000, 3000, 2000.0, 7, 7.0, yaw, pitch, 3,
                   "stereographic little-planet render of 36 synthetic 12MP images, 7 bands")
    elif name == "cfg4":  # ~1 Gpx canvas, 200 x 24MP, spherical, 8 bands
        yaw, pitch = [], []
        for row, (p, span) in enumerate(((0.0, 158.0), (22.0, 152.0), (-22.0, 152.0), (44.0, 138.0), (-44.0, 138.0))):
            for k in range(40):
                yaw.append(-span + 2 * span * k / 39.0)
                pitch.append(p)
        c = Config(name, SPHERICAL, 200, 6000, 4000, 8800.0, 8, 7.0, yaw, pitch, 4,
                   "~1 gigapixel canvas from 200 synthetic 24MP images, spherical, 8 bands")
    else:
        raise KeyError(name)
    if s != 1.0:
        c.width = max(16, int(round(c.width * s)))
        c.height = max(16, int(round(c.height * s)))
        c.focal = c.focal * s
    return c


def rotation(yaw_deg: float, pitch_deg: float, roll_deg: float) -> np.ndarray:
    y, p, r = (math.radians(v) for v in (yaw_deg, pitch_deg, roll_deg))
    Ry = np.array([[math.cos(y), 0, math.sin(y)], [0, 1, 0], [-math.sin(y), 0, math.cos(y)]])
    Rx = np.array([[1, 0, 0], [0, math.cos(p), -math.sin(p)], [0, math.sin(p), math.cos(p)]])
    Rz = np.array([[math.cos(r), -math.sin(r), 0], [math.sin(r), math.cos(r), 0], [0, 0, 1]])
    return Ry @ Rx @ Rz


def cameras(cfg: Config):
    """K[], R[] (float64, like the reference's Eigen::MatrixXd) and exposure factors e_j (= gain[j])."""
    rng = np.random.default_rng(cfg.seed)
    K = [np.array([[cfg.focal, 0, cfg.width / 2.0], [0, cfg.focal, cfg.height / 2.0], [0, 0, 1]], np.float64)
         for _ in range(cfg.n)]
    roll = rng.uniform(-1.0, 1.0, cfg.n)
    R = [rotation(cfg.yaw_deg[j], cfg.pitch_deg[j], roll[j]) for j in range(cfg.n)]
    gains = rng.uniform(0.8, 1.25, cfg.n)
    return K, R, [float(g) for g in gains]


def make_image(cfg: Config, j: int, exposure: float, noise: int = 0) -> np.ndarray:
    """Image j: 128 + 100 sin(x/(37+5c)+j) cos(y/(29+3c)) per channel, clamped to [16,240], times
    the exposure factor (then clamped again so that gray > 1 everywhere inside the image)."""
    x = np.arange(cfg.width, dtype=np.float32)
    y = np.arange(cfg.height, dtype=np.float32)
    wl = max(cfg.width / 1920.0, 0.05)  # keep the pattern band-limited relative to the image size
    img = np.empty((cfg.height, cfg.width, 3), np.uint8)
    rng = np.random.default_rng(1000 * cfg.seed + j) if noise else None
    for c in range(3):
        sx = np.sin(x / np.float32((37 + 5 * c) * max(wl, 0.25)) + np.float32(j))
        cy = np.cos(y / np.float32((29 + 3 * c) * max(wl, 0.25)))
        v = np.float32(128.0) + np.float32(100.0) * np.outer(cy, sx).astype(np.float32)
        if noise:
            v += rng.integers(-noise, noise + 1, v.shape).astype(np.float32)
        v = np.clip(v, 16, 240) * np.float32(exposure)
        img[..., c] = np.clip(np.rint(v), 4, 255).astype(np.uint8)
    return img


def make_images(cfg: Config, gains, noise: int = 0):
    return [make_image(cfg, j, gains[j], noise) for j in range(cfg.n)]


def seam_masks(corners, sizes, soft: int = 8, only: int | None = None, coarse: bool = False):
    """Soft-edged 0..255 seam masks (uint8, one per tile): pixel p of tile j is 255 when j's
    centre is the nearest among the tiles whose rectangle contains p (a Voronoi seam), computed on
    a coarse grid and bilinearly up-sampled to tile size like the reference's preview->full resize."""
    n = len(corners)
    cx = np.array([corners[j][0] + sizes[j][0] / 2.0 for j in range(n)])
    cy = np.array([corners[j][1] + sizes[j][1] / 2.0 for j in range(n)])
    x0 = np.array([c[0] for c in corners]); y0 = np.array([c[1] for c in corners])
    x1 = x0 + np.array([s[0] for s in sizes]); y1 = y0 + np.array([s[1] for s in sizes])
    out = []
    for j in (range(n) if only is None else [only]):
        w, h = sizes[j]
        gw, gh = max(2, w // soft + 1), max(2, h // soft + 1)
        gx = corners[j][0] + (np.arange(gw) + 0.5) * (w / gw)
        gy = corners[j][1] + (np.arange(gh) + 0.5) * (h / gh)
        GX, GY = np.meshgrid(gx, gy)
        dj = (GX - cx[j]) ** 2 + (GY - cy[j]) ** 2
        keep = np.ones((gh, gw), bool)
        for i in range(n):
            if i == j or x1[i] <= x0[j] or x0[i] >= x1[j] or y1[i] <= y0[j] or y0[i] >= y1[j]:
                continue
            # only the grid cells inside tile i's rectangle can change (same comparisons, evaluated on that window)
            kx = np.nonzero((gx >= x0[i]) & (gx < x1[i]))[0]
            ky = np.nonzero((gy >= y0[i]) & (gy < y1[i]))[0]
            if kx.size == 0 or ky.size == 0:
                continue
            sl = (slice(ky[0], ky[-1] + 1), slice(kx[0], kx[-1] + 1))
            di = (GX[sl] - cx[i]) ** 2 + (GY[sl] - cy[i]) ** 2
            keep[sl] &= ~((di < dj[sl]) | ((di == dj[sl]) & (i < j)))
        if coarse:   # the preview-scale binary mask itself (what dist_cut / graph_cut hand to return_full)
            out.append(np.ascontiguousarray(keep.astype(np.uint8) * 255))
        else:
            out.append(_upsample_u8(keep.astype(np.float32) * 255.0, w, h))
    return out if only is None else out[0]


def _upsample_u8(coarse: np.ndarray, w: int, h: int) -> np.ndarray:
    """Bilinear up-sampling (half-pixel centres) of a coarse float grid to (h, w) uint8."""
    gh, gw = coarse.shape
    fx = (np.arange(w) + 0.5) * (gw / w) - 0.5
    fy = (np.arange(h) + 0.5) * (gh / h) - 0.5
    ix = np.clip(np.floor(fx).astype(np.int64), 0, gw - 2); ax = np.clip(fx - ix, 0, 1).astype(np.float32)
    iy = np.clip(np.floor(fy).astype(np.int64), 0, gh - 2); ay = np.clip(fy - iy, 0, 1).astype(np.float32)
    rows0 = coarse[iy][:, ix] * (1 - ax) + coarse[iy][:, ix + 1] * ax
    rows1 = coarse[iy + 1][:, ix] * (1 - ax) + coarse[iy + 1][:, ix + 1] * ax
    v = rows0 * (1 - ay)[:, None] + rows1 * ay[:, None]
    return np.ascontiguousarray(np.clip(np.rint(v), 0, 255).astype(np.uint8))
