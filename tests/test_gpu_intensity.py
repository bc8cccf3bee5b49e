"""GPU parity of test::adjust_intensity ("next" #1; reference src/test/_test.cpp:110-122,
called from return_full at src/classes/_panorama.cpp:337-339 when conf.blend_intensity is on)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_adjust_intensity_golden(ctx, golden):
    """Against the cv2-made fixture: <= 1 LSB (the interpolated float field differs in the last ulp where
    OpenCV's SIMD path fuses a multiply-add), almost everywhere identical."""
    from simplepanorama_b200 import api
    g = golden("intensity.npz")
    out = api.adjust_intensity(g["int_img"], g["int_field"], ctx)
    d = np.abs(out.astype(int) - g["int_out"].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3


@pytest.mark.parametrize("fw,fh,w,h", [(70, 47, 560, 375), (35, 20, 701, 398), (10, 10, 333, 777), (1, 1, 17, 9)])
def test_adjust_intensity_vs_oracle(ctx, oracle, fw, fh, w, h):
    from simplepanorama_b200 import api
    rng = np.random.default_rng(fw + h)
    field = (0.5 + rng.random((fh, fw))).astype(np.float32)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(api.adjust_intensity(img, field, ctx), oracle.adjust_intensity(img, field))


def test_return_full_with_intensity_fields(ctx, oracle):
    """Default configuration of the reference (blend_intensity = true): fused path with fields, <= 1 LSB."""
    from simplepanorama_b200 import api, synth
    cfg = synth.config("cfg1", 0.2)
    K, R, gains = synth.cameras(cfg)
    images = synth.make_images(cfg, gains)
    plan = api.plan_tiles(images, R, K, cfg.kind, cfg.focal)
    sizes = [p[3] for p in plan]
    cuts = synth.seam_masks([p[2] for p in plan], sizes, coarse=True)
    rng = np.random.default_rng(5)
    fields = [(0.8 + 0.4 * rng.random((max(2, h // 16), max(2, w // 16)))).astype(np.float32) for (w, h) in sizes]
    out = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx, intensities=fields)
    ref, _, _, _ = oracle.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, intensities=fields)
    d = np.abs(out.astype(int) - ref.astype(int))
    assert d.max() <= 1, (d.max(), (d > 1).mean())   # the warp is bit-identical now (glibc-exact trig): <= 1 LSB everywhere
    plain = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx)
    assert not np.array_equal(plain, out)


def test_adjust_intensity_errors(ctx):
    from simplepanorama_b200 import api
    with pytest.raises(api.SpanoError):
        api.adjust_intensity(np.zeros((4, 4, 3), np.uint8), np.zeros((0, 0), np.float32), ctx)
