"""Parity at BENCHMARK scale: one full-size tile per projection through the C ABI against OpenCV itself (cv2 on the GPU
box: the same entry points the reference calls), and the blend of a full-size tile with its real seam mask at B = 6
(32-column strips) and B = 8 (16-column strips).  Catches what the scaled-down cases cannot: 32-bit overflow in byte
offsets, tiles of > 170 strips, plans with thousands of pieces, the row-buffer ring over thousands of steps.
Bars: warped tile and validity mask BIT-identical to cv2's; blended float canvas within 1e-5 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

cv2 = pytest.importorskip("cv2")


def _tile(name, j):
    from oracle import cv2_ref, ref_bench
    synth = ref_bench.load_synth()
    cfg = synth.config(name)
    K, R, gains = synth.cameras(cfg)
    img = synth.make_image(cfg, j, gains[j], noise=3)
    return cfg, K, R, gains, img, cv2_ref


@pytest.mark.timeout(600)
@pytest.mark.parametrize("name,j", [("cfg2", 11), ("cfg4", 139), ("cfg3", 30), ("cfg2a", 0)])
def test_full_size_warp_is_bit_identical_to_cv2(ctx, name, j):
    from simplepanorama_b200 import api
    cfg, K, R, gains, img, cv2_ref = _tile(name, j)
    corner_ref, tile_ref = cv2_ref.project(cfg.kind, cfg.focal, R[j], K[j], img)
    mask_ref = cv2_ref.validity_mask(tile_ref)
    corner, tile, mask = api.project(cfg.kind, cfg.focal, R[j], K[j], img, 1.0, True, ctx)
    assert tuple(corner) == tuple(int(v) for v in corner_ref) and tile.shape == tile_ref.shape
    assert tile.shape[0] * tile.shape[1] > (10e6 if name != "cfg3" else 2e6)   # (stereographic tiles looking down are small)
    assert np.array_equal(tile, tile_ref), int((tile != tile_ref).sum())
    assert np.array_equal(mask, mask_ref), int((mask != mask_ref).sum())
    _, gained, _ = api.project(cfg.kind, cfg.focal, R[j], K[j], img, gains[j], True, ctx)
    assert np.array_equal(gained, cv2_ref.apply_gain(tile_ref, gains[j]))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("bands", [6, 8])
def test_full_size_blend_vs_cv2(ctx, bands):
    """one cfg2 tile (5.5k x 4k, 174 strips of 32 / 348 of 16) with its real preview-scale seam mask up-scaled by cv2"""
    from simplepanorama_b200 import api
    from oracle import ref_bench
    cfg, K, R, gains, img, cv2_ref = _tile("cfg2", 11)
    synth = ref_bench.load_synth()
    corners, sizes, W, H, T = ref_bench.job_geometry(cfg, K, R)
    corner, tile = cv2_ref.project(cfg.kind, cfg.focal, R[11], K[11], img)
    valid = cv2_ref.validity_mask(tile)
    tile = cv2_ref.apply_gain(tile, gains[11])
    small = synth.seam_masks(corners, sizes, only=11, coarse=True)
    cut = cv2_ref.resize_mask(small, (tile.shape[1], tile.shape[0]))
    assert np.array_equal(api.resize_mask(small, (tile.shape[1], tile.shape[0]), ctx), cut)
    ref = cv2_ref.multi_blend([tile], [cut], [valid], [tuple(corner)], bands, cfg.sigma)
    ctx.blend_stats(reset=True)
    got = api.multi_blend([tile], [cut], [valid], [tuple(corner)], bands, cfg.sigma, ctx)
    done, offered = ctx.blend_stats()
    assert offered == tile.shape[0] * tile.shape[1] and 0 < done < offered      # the sparse path ran
    assert got.shape == ref.shape
    scale = float(np.abs(ref).max())
    assert float(np.abs(got - ref).max()) <= 1e-5 * scale, (float(np.abs(got - ref).max()), scale)
    u8 = api.blend_to_u8(got) if hasattr(api, "blend_to_u8") else None
    if u8 is not None:
        assert np.abs(u8.astype(int) - cv2_ref.blend_to_u8(ref).astype(int)).max() <= 1
