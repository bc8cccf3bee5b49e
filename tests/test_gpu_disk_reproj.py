"""GPU parity of a5, sten_proj::disk_reproj (the stereographic centre fix), through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _stereo(oracle, scale, n):
    from simplepanorama_b200 import synth
    cfg = synth.config("cfg3", scale)
    K, R, gains = synth.cameras(cfg)
    tiles, corners = [], []
    for j in range(0, cfg.n, max(1, cfg.n // n)):
        img = synth.make_image(cfg, j, 1.0, noise=6)
        K32, R32 = oracle.adjusted_camera(K[j], R[j], cfg.width, cfg.height)
        tl, tile = oracle.warp(cfg.kind, np.float32(cfg.focal), K32, R32, img)
        tiles.append(tile); corners.append(tl)
    return tiles, corners


@pytest.mark.parametrize("quadratic", [True, False])
def test_disk_reproj_bit_exact(ctx, oracle, quadratic):
    """Integer source coordinates, gather and recomputed validity masks: exact."""
    from simplepanorama_b200 import api
    tiles, corners = _stereo(oracle, 0.08, 9)
    sizes = [(t.shape[1], t.shape[0]) for t in tiles]
    W, H, mx, my = oracle.pan_dimension(corners, sizes)
    ansatz, radius = (W // 2 + 5, H // 2 - 3), 17.0
    outs_o, msks_o, corners_o = oracle.disk_reproj(tiles, corners, ansatz, radius, quadratic)
    pd = api.disk_reproj(api.ProjData(imgs=tiles, msks=[], corners=corners), ansatz, radius, quadratic, ctx)
    assert pd.corners == corners_o
    for a, b in zip(pd.imgs, outs_o):
        assert a.shape == b.shape
        assert np.array_equal(a, b), float((a != b).mean())
    for a, b in zip(pd.msks, msks_o):
        assert np.array_equal(a, b)


def test_disk_reproj_then_blend(ctx, oracle):
    """config-3 flow with the fix: warp -> disk_reproj -> gain -> multi_blend, <= 1 LSB on the canvas."""
    from simplepanorama_b200 import api, synth
    tiles, corners = _stereo(oracle, 0.06, 6)
    sizes = [(t.shape[1], t.shape[0]) for t in tiles]
    W, H, mx, my = oracle.pan_dimension(corners, sizes)
    ansatz, radius = (W // 2, H // 2), 12.0
    outs_o, msks_o, corners_o = oracle.disk_reproj(tiles, corners, ansatz, radius, True)
    pd = api.disk_reproj(api.ProjData(imgs=tiles, msks=[], corners=corners), ansatz, radius, True, ctx)
    cuts = synth.seam_masks(pd.corners, [(t.shape[1], t.shape[0]) for t in pd.imgs])
    ref = oracle.blend_to_u8(oracle.multi_blend(outs_o, cuts, msks_o, corners_o, 3, 7.0))
    out = api.blend(pd.imgs, cuts, pd.msks, pd.corners, 3, 7.0, ctx)
    assert np.abs(out.astype(int) - ref.astype(int)).max() <= 1


def test_disk_reproj_errors(ctx):
    from simplepanorama_b200 import api
    with pytest.raises(api.SpanoError):
        api.disk_reproj(api.ProjData(imgs=[], msks=[], corners=[]), (0, 0), 5.0, True, ctx)


@pytest.mark.parametrize("quadratic", [True, False])
def test_return_full_with_center_fix(ctx, oracle, quadratic):
    """The fused path with the little-planet centre fix (spano_composite_fixed: warp -> disk_reproj -> validity masks ->
    gain -> multi_blend, all on the device) against the oracle's composition of the same stages, for the circle that the
    restated sten_proj::estimate_circle finds in the union of the oracle's validity masks: <= 1 LSB on the 8-bit canvas."""
    cv2 = pytest.importorskip("cv2")
    from simplepanorama_b200 import api, synth
    from oracle import cv2_ref
    cfg = synth.config("cfg3", 0.08)
    K, R, gains = synth.cameras(cfg)
    images = synth.make_images(cfg, gains)
    tiles, msks, corners = [], [], []
    for j in range(cfg.n):
        K32, R32 = oracle.adjusted_camera(K[j], R[j], cfg.width, cfg.height)
        tl, t = oracle.warp(cfg.kind, np.float32(cfg.focal), K32, R32, images[j])
        tiles.append(t); corners.append(tuple(tl)); msks.append(oracle.surrounding_mask(t, 3))
    ansatz, radius = cv2_ref.estimate_circle(cv2_ref.ProjData(imgs=tiles, msks=msks, corners=corners))
    assert ansatz is not None and radius > 5, "the synthetic little planet must have a hole in its middle"
    new_tiles, new_msks, new_corners = oracle.disk_reproj(tiles, corners, ansatz, radius, quadratic)
    new_sizes = [(t.shape[1], t.shape[0]) for t in new_tiles]
    c2, s2 = api.disk_reproj_size(corners, [(t.shape[1], t.shape[0]) for t in tiles], ansatz, radius, quadratic, ctx)
    assert c2 == [tuple(c) for c in new_corners] and s2 == new_sizes
    small = synth.seam_masks(new_corners, new_sizes, coarse=True)
    cuts = [oracle.resize_linear_u8(c, s) for c, s in zip(small, new_sizes)]
    gained = [oracle.apply_gain(t, g) for t, g in zip(new_tiles, gains)]
    ref = oracle.blend_to_u8(oracle.multi_blend(gained, cuts, new_msks, new_corners, cfg.bands, cfg.sigma))
    got = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, small, cfg.bands, cfg.sigma, ctx=ctx,
                          center_fix=(ansatz, radius, quadratic))
    assert got.shape == ref.shape
    d = np.abs(got.astype(int) - ref.astype(int))
    assert d.max() <= 1, (int(d.max()), float((d > 1).mean()))
    plain = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, synth.seam_masks(corners, [(t.shape[1], t.shape[0]) for t in tiles], coarse=True),
                            cfg.bands, cfg.sigma, ctx=ctx)
    assert plain.shape != got.shape or not np.array_equal(plain, got)
    # tile-sized masks (already at the NEW tile size) take the same path
    got2 = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx, center_fix=(ansatz, radius, quadratic))
    assert np.array_equal(got2, got)


def test_center_fix_errors(ctx):
    from simplepanorama_b200 import api, synth
    cfg = synth.config("cfg1", 0.1)
    K, R, gains = synth.cameras(cfg)
    images = synth.make_images(cfg, gains)
    plan = api.plan_tiles(images, R, K, cfg.kind, cfg.focal)
    cuts = synth.seam_masks([p[2] for p in plan], [p[3] for p in plan], coarse=True)
    with pytest.raises(api.SpanoError):   # the centre fix belongs to the stereographic projection
        api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx, center_fix=((10, 10), 5.0, True))
