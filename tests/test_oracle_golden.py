"""CPU: the C oracle (oracle/spano_oracle.c) against the golden vectors produced by OpenCV 4.13
(oracle/gen_golden.py).  This is what pins the oracle; the CUDA path is then held to the oracle."""
import math

import numpy as np
import pytest


def test_roi_table(golden, oracle):
    t = golden("roi_table.npz")["table"]
    assert len(t) >= 36
    for row in t:
        kind, W, H, f = int(row[0]), int(row[1]), int(row[2]), float(row[3])
        K = row[4:13].astype(np.float32); R = row[13:22].astype(np.float32)
        tl, size = oracle.warp_roi(kind, np.float32(f), K, R, W, H)
        assert (tl[0], tl[1], size[0], size[1]) == tuple(int(v) for v in row[22:26]), row[:4]


def test_warp_cases_bit_exact(golden, oracle):
    g = golden("warp_cases.npz")
    for i in range(int(g["warp_count"])):
        kind, f = int(g[f"warp{i}_kind"]), float(g[f"warp{i}_focal"])
        img = g[f"warp{i}_img"]
        K32, R32 = oracle.adjusted_camera(g[f"warp{i}_K"], g[f"warp{i}_R"], img.shape[1], img.shape[0])
        tl, tile = oracle.warp(kind, np.float32(f), K32, R32, img)
        assert tuple(tl) == tuple(g[f"warp{i}_corner"])
        assert np.array_equal(tile, g[f"warp{i}_tile"])
        assert np.array_equal(oracle.surrounding_mask(tile, 3), g[f"warp{i}_mask"])


def test_remap_kat(golden, oracle):
    g = golden("kernels.npz")
    assert np.array_equal(oracle.remap(g["remap_img"], g["remap_x"], g["remap_y"]), g["remap_dst"])


def test_mask_kat(golden, oracle):
    g = golden("kernels.npz")
    assert np.array_equal(oracle.surrounding_mask(g["mask_img"], 0), g["mask_raw"])
    assert np.array_equal(oracle.surrounding_mask(g["mask_img"], 3), g["mask_eroded"])


def test_taps_and_divisors(golden, oracle):
    g = golden("kernels.npz")
    for B in range(1, 11):
        for i in range(B):
            t = oracle.gaussian_taps(43, math.sqrt(2 * (B - i) + 1) * 7.0)
            assert np.array_equal(t, g["taps_sigma7"][B - 1, i]), (B, i)
    assert list(g["divisors"]) == [255, 127, 85, 63, 51, 42, 36, 31, 28, 25]


def test_gain_table(golden, oracle):
    g = golden("kernels.npz")
    a = np.arange(256, dtype=np.uint8).reshape(1, -1, 1).repeat(3, axis=2)
    for k, gain in enumerate(g["gain_values"]):
        assert np.array_equal(oracle.apply_gain(a, float(gain))[0, :, 0], g["gain_table"][k])


def test_resize_kat(golden, oracle):
    g = golden("kernels.npz")
    assert np.array_equal(oracle.resize_linear_u8(g["resize_src"], (200, 117)), g["resize_dst"])


@pytest.mark.parametrize("key,bands,sigma", [("B1", 1, 7.0), ("B2", 2, 7.0), ("B3", 3, 7.0), ("B5", 5, 7.0), ("B4_s3", 4, 3.0)])
def test_multi_blend_golden(golden, oracle, key, bands, sigma):
    g = golden("blend_cases.npz")
    tiles = [g[f"blend_tile{j}"] for j in range(3)]
    cuts = [g[f"blend_cut{j}"] for j in range(3)]
    valids = [g[f"blend_valid{j}"] for j in range(3)]
    corners = [tuple(int(v) for v in c) for c in g["blend_corners"]]
    f = oracle.multi_blend(tiles, cuts, valids, corners, bands, sigma)
    ref = g[f"blend_f32_{key}"]
    assert f.shape == ref.shape
    # float intermediates within 1e-5 relative (scale: the canvas' own magnitude)
    tol = 1e-5 * max(1.0, float(np.abs(ref).max()))
    assert float(np.abs(f - ref).max()) <= tol
    u8 = oracle.blend_to_u8(f)
    assert int(np.abs(u8.astype(int) - g[f"blend_u8_{key}"].astype(int)).max()) <= 1


def test_single_tile_closed_form(oracle):
    """KAT (v): one tile, all-ones masks.  Sum of the reference's bands (Q2, they do not telescope):
    I + G0 + G1 - 2 G_{B-1} for B >= 3;  out = that / B / (255/B)."""
    rng = np.random.default_rng(5)
    tile = rng.integers(16, 240, (40, 48, 3), dtype=np.uint8)
    ones = np.full((40, 48), 255, np.uint8)
    B, sigma = 4, 7.0
    out = oracle.multi_blend([tile], [ones], [ones], [(0, 0)], B, sigma)
    I = tile.astype(np.float32)
    G = [oracle.gaussian_blur(I, 43, math.sqrt(2 * (B - i) + 1) * sigma) for i in range(B)]
    expect = (I + G[0] + G[1] - 2 * G[B - 1]) / B / float(255 // B)
    assert np.allclose(out, expect, rtol=2e-5, atol=1e-6)


def test_cubic_remap_on_integer_maps_is_a_gather(golden):
    """a5 pin: sten_proj::disk_reproj hands cv::remap(INTER_CUBIC) integer-valued maps
    (denormalizePoint returns cv::Point); OpenCV's result is then an exact pixel gather with
    BORDER_CONSTANT(0) outside -- which is how the oracle and the CUDA kernel implement it."""
    g = golden("cubic_gather.npz")
    img, xm, ym, dst = g["cubic_img"], g["cubic_x"], g["cubic_y"], g["cubic_dst"]
    H, W = img.shape[:2]
    xi, yi = xm.astype(int), ym.astype(int)
    ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
    ref = np.zeros_like(dst)
    ref[ok] = img[yi[ok], xi[ok]]
    assert np.array_equal(ref, dst)


def _stereo_tiles(oracle, scale=0.05, n_max=None):
    from simplepanorama_b200 import synth
    cfg = synth.config("cfg3", scale)
    K, R, gains = synth.cameras(cfg)
    tiles, corners = [], []
    for j in range(cfg.n if n_max is None else n_max):
        img = synth.make_image(cfg, j, 1.0)
        K32, R32 = oracle.adjusted_camera(K[j], R[j], cfg.width, cfg.height)
        tl, tile = oracle.warp(cfg.kind, np.float32(cfg.focal), K32, R32, img)
        tiles.append(tile); corners.append(tl)
    return tiles, corners


def test_disk_reproj_oracle_properties(oracle):
    """The centre fix pulls content towards the circle centre: a pixel well outside the circle keeps
    its direction (same polar angle about the centre) and the tiles stay 8-bit gathers of the input."""
    tiles, corners = _stereo_tiles(oracle, 0.04, 6)
    sizes = [(t.shape[1], t.shape[0]) for t in tiles]
    W, H, mx, my = oracle.pan_dimension(corners, sizes)
    ansatz, radius = (W // 2 + 3, H // 2 - 2), 9.0
    for quad in (True, False):
        outs, msks, new_corners = oracle.disk_reproj(tiles, corners, ansatz, radius, quad)
        assert len(outs) == len(tiles)
        for t, o, m in zip(tiles, outs, msks):
            assert o.dtype == np.uint8 and o.shape[:2] == m.shape
            vals = set(map(tuple, o.reshape(-1, 3)[:: 37]))
            src = set(map(tuple, t.reshape(-1, 3))) | {(0, 0, 0)}
            assert vals <= src          # every output pixel is a copy of an input pixel (or border 0)


def test_adjust_intensity_golden(golden, oracle):
    """test::adjust_intensity: the interpolated field within float rounding of cv2.resize, the 8-bit result
    within 1 LSB of the cv2-made fixture."""
    g = golden("intensity.npz")
    f = oracle.resize_linear_f32(g["int_field"], (g["int_img"].shape[1], g["int_img"].shape[0]))
    assert np.abs(f - g["int_field_resized"]).max() <= 1e-6 * np.abs(g["int_field_resized"]).max()
    out = oracle.adjust_intensity(g["int_img"], g["int_field"])
    d = np.abs(out.astype(int) - g["int_out"].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
