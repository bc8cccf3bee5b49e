"""cfg2a: the seam-straddling edge case (true 360-degree cylindrical ring; the tiles at the +-pi seam are as wide as the
panorama).  tests/golden/cfg2a.npz holds what OpenCV 4.13 produces (oracle/gen_golden_cfg2a.py); the C oracle (CPU test)
and the CUDA path through the C ABI (GPU test) are compared with it: corners / sizes and warped tiles + validity masks
exactly, the 8-bit canvas of return_full within 1 LSB."""
import numpy as np
import pytest


def _inputs(g):
    from oracle import ref_bench
    synth = ref_bench.load_synth()
    cfg = synth.config("cfg2a", float(g["scale"]))
    K, R, gains = synth.cameras(cfg)
    images = synth.make_images(cfg, gains)
    corners = [tuple(int(v) for v in c) for c in g["corners"]]
    sizes = [tuple(int(v) for v in s) for s in g["sizes"]]
    cuts = synth.seam_masks(corners, sizes, coarse=True)
    return cfg, K, R, gains, images, corners, sizes, cuts


def test_oracle_matches_cv2_on_seam_straddling_tiles(oracle, golden):
    g = golden("cfg2a.npz")
    cfg, K, R, gains, images, corners, sizes, cuts = _inputs(g)
    assert len(g["wide"]) >= 2 and all(sizes[j][0] >= max(s[0] for s in sizes) - 1 for j in g["wide"])
    for j in list(g["wide"]) + [0]:
        K32, R32 = oracle.adjusted_camera(K[j], R[j], cfg.width, cfg.height)
        tl, size = oracle.warp_roi(cfg.kind, np.float32(cfg.focal), K32, R32, cfg.width, cfg.height)
        assert tuple(tl) == corners[j] and tuple(size) == sizes[j]
        _, tile = oracle.warp(cfg.kind, np.float32(cfg.focal), K32, R32, images[j])
        assert np.array_equal(tile, g[f"tile{j}"])
        assert np.array_equal(oracle.surrounding_mask(tile, 3), g[f"mask{j}"])
    out, _, _, _ = oracle.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma)
    assert out.shape == g["canvas"].shape
    assert np.abs(out.astype(int) - g["canvas"].astype(int)).max() <= 1


@pytest.mark.gpu
def test_gpu_matches_cv2_on_seam_straddling_tiles(ctx, golden):
    from simplepanorama_b200 import api
    g = golden("cfg2a.npz")
    cfg, K, R, gains, images, corners, sizes, cuts = _inputs(g)
    plan = api.plan_tiles(images, R, K, cfg.kind, cfg.focal)
    assert [tuple(p[2]) for p in plan] == corners and [tuple(p[3]) for p in plan] == sizes
    for j in list(g["wide"]) + [0]:
        corner, tile, mask = api.project(cfg.kind, cfg.focal, R[j], K[j], images[j], 1.0, True, ctx)
        assert tuple(corner) == corners[j]
        assert np.array_equal(tile, g[f"tile{j}"]) and np.array_equal(mask, g[f"mask{j}"])
    out = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx)
    assert out.shape == g["canvas"].shape
    assert np.abs(out.astype(int) - g["canvas"].astype(int)).max() <= 1
