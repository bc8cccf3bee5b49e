"""graph_cut.py -- gcut::graph_cut restated (reference src/math/_graph_cut.cpp:8-118, computeCut :285-341,
define_graph_full :344-405, scharr_gradient src/math/_graph_cut.h:57-97 + _graph_cut.cpp:248-282, graph_object
src/math/_graph_cut_helper.h:24-110 / .cpp:53-104,169-186) on top of the reference's vendored max-flow.

TEST / INPUT-GENERATION INFRASTRUCTURE ONLY.  The graph-cut seam search is OUT of the accelerated path (north_star: "the
graph-cut seam search stays in the reference"); its output `mask_cut[]` is an INPUT of the hot path, and BASELINE.json
configs[4] asks for the band sweep to run on "graph-cut seam masks precomputed by the reference".  This module
precomputes them: `python oracle/graph_cut.py` writes tests/golden/cfg5_masks.npz (preview scale 1/8, bit-packed), which
bench.py --workload cfg5 loads.  The max-flow itself is the reference's own code, compiled from /root/reference/src/max_flow
into oracle/_ref/libmaxflow.so (never copied into the repo); OpenCV calls go through cv2 4.13.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/src/max_flow"
LIB = os.path.join(HERE, "_ref", "libmaxflow.so")
MASKS = os.path.join(ROOT, "tests", "golden", "cfg5_masks.npz")
PREVIEW = 8      # masks are made at 1/8 linear scale (the reference makes them on the preview-size warps)


def build(force: bool = False) -> str:
    """g++ on the reference's few max_flow sources, output only into oracle/_ref/ (needs /root/reference)."""
    if os.path.exists(LIB) and not force:
        return LIB
    if not os.path.isdir(REF):
        raise RuntimeError("the reference tree (src/max_flow) is not present: oracle/_ref/libmaxflow.so cannot be built here")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-w", "-I", REF, os.path.join(HERE, "maxflow_wrap.cpp"),
                    os.path.join(REF, "graph.cpp"), os.path.join(REF, "maxflow.cpp"), "-o", LIB], check=True)
    return LIB


_lib = None


def _mf():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.mf_solve.restype = C.c_float
    return _lib


def _find_contour(mask, thickness=1):
    """graph_object::find_contour: the mask minus its 3x3 erosion (`thickness` iterations), zero-padded."""
    import cv2
    p = thickness + 1
    padded = cv2.copyMakeBorder(mask, p, p, p, p, cv2.BORDER_CONSTANT, value=0)
    er = cv2.erode(padded, None, iterations=thickness)
    cont = cv2.subtract(padded, er)
    return cont[p:p + mask.shape[0], p:p + mask.shape[1]]


def compute_cut(img1, img2, mask1, mask2):
    """gcut::computeCut(gray(panorama roi), gray(image), scene roi, element mask) -> new mask of the element."""
    import cv2
    f1, f2 = img1.astype(np.float32), img2.astype(np.float32)
    g1x, g1y = cv2.Scharr(f1, cv2.CV_32F, 1, 0), cv2.Scharr(f1, cv2.CV_32F, 0, 1)
    g2x, g2y = cv2.Scharr(f2, cv2.CV_32F, 1, 0), cv2.Scharr(f2, cv2.CV_32F, 0, 1)
    adif = cv2.absdiff(f1, f2)
    obj = np.where(mask2 != 0, mask1, 0).astype(np.uint8)          # extract_object(scene, element)
    rows, cols = obj.shape
    idx = np.flatnonzero(obj.reshape(-1))                           # raster order == the reference's index[]
    n = int(idx.size)
    if n == 0:
        return mask2.copy()
    inv = np.full(rows * cols, -1, np.int32)
    inv[idx] = np.arange(n, dtype=np.int32)
    r, c = idx // cols, idx % cols
    cont_obj = _find_contour(obj)
    sink = (np.where(_find_contour(mask1) != 0, cont_obj, 0).reshape(-1)[idx] > 0).astype(np.uint8)     # graph channel 1
    source = (np.where(_find_contour(mask2) != 0, cont_obj, 0).reshape(-1)[idx] > 0).astype(np.uint8)   # graph channel 2
    flat = lambda a: a.reshape(-1)
    eps = np.float32(1e-6)
    # horizontal edge to (row, col + 1), vertical edge to (row + 1, col), when that pixel belongs to the object
    right = np.where(c < cols - 1, idx + 1, 0)
    h_ok = (c < cols - 1) & (flat(obj)[right] > 0)
    h_conn = np.where(h_ok, inv[right], -1).astype(np.int32)
    down = np.where(r < rows - 1, idx + cols, 0)
    v_ok = (r < rows - 1) & (flat(obj)[down] > 0)
    v_conn = np.where(v_ok, inv[down], -1).astype(np.int32)
    A = flat(adif)
    ay1, ay2, ax1, ax2 = np.abs(flat(g1y)), np.abs(flat(g2y)), np.abs(flat(g1x)), np.abs(flat(g2x))

    def weight(j, a1, a2):   # scharr_gradient::read(i, j): float arithmetic, left to right
        return ((A[idx] + A[j]) / (((a1[idx] + a1[j]) + a2[idx]) + a2[j] + eps)).astype(np.float32)
    h_cap = np.where(h_ok, weight(right, ay1, ay2), 0).astype(np.float32)
    v_cap = np.where(v_ok, weight(down, ax1, ax2), 0).astype(np.float32)
    label = np.zeros(n, np.int32)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    _mf().mf_solve(n, p(h_conn, C.c_int), p(h_cap, C.c_float), p(v_conn, C.c_int), p(v_cap, C.c_float), p(sink, C.c_ubyte),
                   p(source, C.c_ubyte), int(h_ok.sum() + v_ok.sum()), p(label, C.c_int))
    cut = mask2.copy()                                               # graph_object::write_cut
    cut.reshape(-1)[idx] = (255 * label).astype(np.uint8)
    return cut


def graph_cut(images, masks, top_lefts, seq):
    """gcut::graph_cut(images, masks, top_lefts, seq): the images are pasted in order `seq`; every later image is cut
    against what has been pasted so far; finally pixels claimed by a later image are removed from the earlier ones."""
    import cv2
    n = len(images)
    min_x = min(c[0] for c in top_lefts); min_y = min(c[1] for c in top_lefts)
    W = max(c[0] + im.shape[1] for c, im in zip(top_lefts, images)) - min_x
    H = max(c[1] + im.shape[0] for c, im in zip(top_lefts, images)) - min_y
    pano = np.zeros((H, W, 3), np.uint8)
    scene = np.zeros((H, W), np.uint8)
    roi = [(top_lefts[i][0] - min_x, top_lefts[i][1] - min_y, images[i].shape[1], images[i].shape[0]) for i in range(n)]
    view = lambda a, i: a[roi[i][1]:roi[i][1] + roi[i][3], roi[i][0]:roi[i][0] + roi[i][2]]
    out = [m.copy() for m in masks]
    s0 = seq[0]
    view(pano, s0)[out[s0] != 0] = images[s0][out[s0] != 0]
    view(scene, s0)[masks[s0] != 0] = masks[s0][masks[s0] != 0]
    for s in seq[1:]:
        g1 = cv2.cvtColor(np.ascontiguousarray(view(pano, s)), cv2.COLOR_BGR2GRAY)
        g2 = cv2.cvtColor(images[s], cv2.COLOR_BGR2GRAY)
        cmat = compute_cut(g1, g2, np.ascontiguousarray(view(scene, s)), masks[s])
        view(scene, s)[cmat != 0] = cmat[cmat != 0]
        view(pano, s)[cmat != 0] = images[s][cmat != 0]
        out[s] = cmat
    added = []
    for s in seq:
        for a in added:
            x0 = max(roi[a][0], roi[s][0]); y0 = max(roi[a][1], roi[s][1])
            x1 = min(roi[a][0] + roi[a][2], roi[s][0] + roi[s][2]); y1 = min(roi[a][1] + roi[a][3], roi[s][1] + roi[s][3])
            if x1 <= x0 or y1 <= y0:
                continue
            ia = out[a][y0 - roi[a][1]:y1 - roi[a][1], x0 - roi[a][0]:x1 - roi[a][0]]
            rm = out[s][y0 - roi[s][1]:y1 - roi[s][1], x0 - roi[s][0]:x1 - roi[s][0]]
            ia[rm != 0] = 0
        added.append(s)
    return out


def seam_masks_for(cfg, K, R, gains, corners, sizes):
    """Preview-scale graph-cut masks for the full-size tiles (corners, sizes) of `cfg`: loaded from the committed
    fixture when it matches the layout.  Returns (masks, description)."""
    if not os.path.exists(MASKS):
        raise RuntimeError("tests/golden/cfg5_masks.npz is missing (python oracle/graph_cut.py)")
    z = np.load(MASKS)
    if z["name"].item() != cfg.name or int(z["n"]) != cfg.n or [tuple(s) for s in z["full_sizes"]] != [tuple(s) for s in sizes]:
        raise RuntimeError("the precomputed graph-cut masks belong to another layout")
    out = []
    for j in range(cfg.n):
        h, w = (int(v) for v in z["shapes"][j])
        out.append(np.ascontiguousarray(np.unpackbits(z[f"m{j}"])[: h * w].reshape(h, w) * np.uint8(255)))
    return out, ("graph-cut seams (gcut::graph_cut restated in oracle/graph_cut.py on the reference's vendored max-flow), precomputed at "
                 "1/%d scale on the preview-size warps: tests/golden/cfg5_masks.npz" % PREVIEW)


def generate(name="cfg2"):
    """The set_config part of the reference at preview scale: warp the 1/8-size sources, validity masks, graph_cut in
    array order; the masks are stored bit-packed next to the FULL-size layout they belong to."""
    sys.path.insert(0, ROOT)
    from oracle import cv2_ref, ref_bench
    synth = ref_bench.load_synth()
    full = synth.config(name)
    Kf, Rf, gf = synth.cameras(full)
    _, full_sizes, _, _, _ = ref_bench.job_geometry(full, Kf, Rf)
    cfg = synth.config(name, 1.0 / PREVIEW)
    K, R, gains = synth.cameras(cfg)
    images = synth.make_images(cfg, gains, noise=2)
    pd = cv2_ref.get_proj_parameters(images, R, K, [1.0] * cfg.n, cfg.kind, cfg.focal)
    cuts = graph_cut(pd.imgs, pd.msks, pd.corners, list(range(cfg.n)))
    out = {"name": np.array(full.name), "n": np.int32(cfg.n), "full_sizes": np.array(full_sizes, np.int32),
           "shapes": np.array([c.shape for c in cuts], np.int32)}
    for j, c in enumerate(cuts):
        out[f"m{j}"] = np.packbits((c != 0).astype(np.uint8).reshape(-1))
    np.savez_compressed(MASKS, **out)
    kept = [float((c != 0).mean()) for c in cuts]
    print(MASKS, os.path.getsize(MASKS), "bytes; kept fraction per tile: min %.2f max %.2f" % (min(kept), max(kept)))
    return cuts, pd


if __name__ == "__main__":
    build(force=True)
    generate()
