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
