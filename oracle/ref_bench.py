"""ref_bench.py -- the reference's CPU compositing path, timed for bench.py.

TEST / BASELINE INFRASTRUCTURE ONLY: used by bench.py's `--impl reference` arm and by the `cpu_baseline` /
`parity_at_bench_scale` legs of the CUDA arm.  Nothing here imports the `simplepanorama_b200` package or loads
libspano.so: geometry comes from OpenCV's own `cv2.PyRotationWarper.warpRoi` (what cv::detail::*Warper::warp uses,
reference src/math/_projection.cpp:51,81,321), the arithmetic from oracle/cv2_ref.py (the reference's loop structure
through the same OpenCV kernels), and the synthetic workload definition (simplepanorama_b200/synth.py, pure numpy) is
loaded BY PATH so that the package is never imported in the reference process.

Two measurements:
  * sample(): a bounded sample of the named workload that keeps what makes the job what it is -- THREE ADJACENT images
    with their real overlaps and their real preview-scale seam masks (`mask_cut`), cropped to a central strip of rows
    sized for a given CPU budget; the whole path from decoded sources to the 8-bit canvas is timed (warp, validity
    masks, gain, mask_cut up-scaling, 6-band multi_blend, convert).  The whole-job figure is the sample's tile-pixel rate
    scaled by the job's tile-pixel / canvas-pixel ratio (stated in the description; the full cfg2 job is ~200 s on CPU).
  * full_cfg1(): BASELINE.json configs[0] (6 x 1920x1080, spherical, 5 bands) run COMPLETELY: an un-extrapolated number.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import time

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_KIND = {0: "spherical", 1: "cylindrical", 2: "stereographic"}


def load_synth():
    """simplepanorama_b200/synth.py as a stand-alone module (no package import, no native library)."""
    name = "_spano_synth_by_path"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(_ROOT, "simplepanorama_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def warp_geometry(kind, focal, K, R, w, h):
    """(corner, (tile_w, tile_h)) of cv::detail::*Warper::warp: roi = warpRoi, tile = roi + 1 (RotationWarperBase::warp)."""
    import cv2
    from oracle import cv2_ref
    K32, R32 = cv2_ref.adjusted_camera(K, R, w, h)
    x, y, rw, rh = cv2.PyRotationWarper(_KIND[kind], float(np.float32(focal))).warpRoi((w, h), K32, R32)
    return (int(x), int(y)), (int(rw), int(rh))


def job_geometry(cfg, K, R):
    """corners, sizes, canvas (W, H) and tile pixels T of the whole workload, from OpenCV."""
    corners, sizes = [], []
    for j in range(cfg.n):
        c, s = warp_geometry(cfg.kind, cfg.focal, K[j], R[j], cfg.width, cfg.height)
        corners.append(c)
        sizes.append(s)
    W = max(c[0] + s[0] for c, s in zip(corners, sizes)) - min(c[0] for c in corners)
    H = max(c[1] + s[1] for c, s in zip(corners, sizes)) - min(c[1] for c in corners)
    return corners, sizes, W, H, sum(w * h for w, h in sizes)


def pick_adjacent(cfg, k=3):
    """k consecutive image indices lying in one row of the panorama (equal pitch), near the middle of that row."""
    k = min(k, cfg.n)
    runs, a = [], 0
    for j in range(1, cfg.n + 1):
        if j == cfg.n or cfg.pitch_deg[j] != cfg.pitch_deg[a]:
            runs.append((a, j))
            a = j
    a, b = max(runs, key=lambda r: (min(r[1] - r[0], k), -r[0]))
    j0 = max(a, min(b - k, (a + b) // 2 - k // 2))
    return list(range(j0, min(cfg.n, j0 + k)))


def make_sample(cfg, synth, K, R, gains, images, rows, idx=None):
    """Inputs of the sample: the images `idx` (default: three adjacent ones) cropped to their central `rows` rows (the
    principal point moves with the crop), their cameras, gains and their real preview-scale seam masks."""
    idx = pick_adjacent(cfg) if idx is None else list(idx)
    rows = int(max(8, min(cfg.height, rows)))
    y0 = (cfg.height - rows) // 2
    imgs, Ks, Rs, gs = [], [], [], []
    for j in idx:
        im = images[j] if images is not None and images[j] is not None else synth.make_image(cfg, j, gains[j])
        imgs.append(np.ascontiguousarray(im[y0:y0 + rows]))
        Kc = np.array(K[j], np.float64).copy()
        Kc[1, 2] = Kc[1, 2] - y0
        Ks.append(Kc)
        Rs.append(R[j])
        gs.append(gains[j])
    geo = [warp_geometry(cfg.kind, cfg.focal, Ks[i], Rs[i], cfg.width, rows) for i in range(len(idx))]
    corners, sizes = [g[0] for g in geo], [g[1] for g in geo]
    cuts = synth.seam_masks(corners, sizes, coarse=True)      # preview scale (1/8), as return_full receives mask_cut
    return dict(idx=idx, rows=rows, images=imgs, K=Ks, R=Rs, gains=gs, cuts=cuts, corners=corners, sizes=sizes)


def run_path(cfg, s):
    """stitch_parameters::return_full (MULTI_BLEND) on the sample through cv2: returns (seconds, canvas, tile px)."""
    from oracle import cv2_ref
    t0 = time.perf_counter()
    pd = cv2_ref.get_proj_parameters(s["images"], s["R"], s["K"], [1.0] * len(s["images"]), cfg.kind, cfg.focal)
    tiles = [cv2_ref.apply_gain(t, g) for t, g in zip(pd.imgs, s["gains"])]
    cuts = [cv2_ref.resize_mask(c, (t.shape[1], t.shape[0])) for c, t in zip(s["cuts"], tiles)]
    out = cv2_ref.blend_to_u8(cv2_ref.multi_blend(tiles, cuts, pd.msks, pd.corners, cfg.bands, cfg.sigma))
    dt = time.perf_counter() - t0
    return dt, out, sum(t.shape[0] * t.shape[1] for t in tiles)


def sample(cfg, synth, K, R, gains, images, seconds, steps=1, warmup=0, job=None):
    """Calibrates the strip height for about `seconds` of CPU work per pass, then times `steps` passes.
    Returns dict(value = whole-job canvas Mpx/s, ms, cores, desc, canvas, inputs)."""
    import cv2
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    corners, sizes, W, H, T = job if job is not None else job_geometry(cfg, K, R)
    pilot_rows = int(max(64, min(cfg.height, 192)))
    s = make_sample(cfg, synth, K, R, gains, images, pilot_rows)
    t, _, _ = run_path(cfg, s)
    rows = int(max(pilot_rows, min(cfg.height, pilot_rows * seconds / max(t, 1e-3))))
    if rows != pilot_rows:
        s = make_sample(cfg, synth, K, R, gains, images, rows)
    times, out, px = [], None, 0
    for i in range(warmup + steps):
        t, out, px = run_path(cfg, s)
        if i >= warmup:
            times.append(t)
    t = float(np.mean(times))
    tile_rate = px / t / 1e6
    value = tile_rate * (W * H) / T
    desc = (f"images {s['idx']} of {cfg.n} (adjacent: real overlaps, real preview-scale seam masks), central {s['rows']}-row strip "
            f"({px / 1e6:.2f} tile-Mpx, canvas {out.shape[1]}x{out.shape[0]}): warp + validity masks + gain + mask_cut up-scaling + "
            f"{cfg.bands}-band multi_blend + convert via cv2 {cv2.__version__}, {cores} threads, {t:.2f} s per pass; whole-job figure = "
            f"sample tile-Mpx/s ({tile_rate:.3f}) x C/T (C={W * H / 1e6:.1f} canvas-Mpx, T={T / 1e6:.1f} tile-Mpx)")
    return dict(value=value, ms=t * 1e3, cores=cores, desc=desc, canvas=out, inputs=s, tile_mpx_s=tile_rate)


def full_cfg1(synth=None):
    """BASELINE.json configs[0] completely on the CPU: dict(value = canvas Mpx/s, ms, canvas, inputs)."""
    import cv2
    synth = synth or load_synth()
    cfg = synth.config("cfg1")
    cv2.setNumThreads(os.cpu_count() or 1)
    K, R, gains = synth.cameras(cfg)
    images = synth.make_images(cfg, gains)
    s = make_sample(cfg, synth, K, R, gains, images, cfg.height, idx=range(cfg.n))
    t, out, px = run_path(cfg, s)
    return dict(value=out.shape[0] * out.shape[1] / t / 1e6, ms=t * 1e3, canvas=out, inputs=s, cfg=cfg, tile_mpx=px / 1e6,
                unit="Mpx/s", note="cfg1 (6 x 1920x1080, spherical, 5 bands) run completely, no extrapolation")
