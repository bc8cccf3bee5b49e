"""GPU parity of the mask_cut up-scaling ("next" #2): cv::resize(CV_8UC1, INTER_LINEAR)
as stitch_parameters::return_full applies it to every seam mask (src/classes/_panorama.cpp:329-335)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_resize_golden(ctx, golden):
    """Against cv2.resize itself (fixture made by oracle/gen_golden.py): bit-exact."""
    from simplepanorama_b200 import api
    g = golden("kernels.npz")
    assert np.array_equal(api.resize_mask(g["resize_src"], (200, 117), ctx), g["resize_dst"])


@pytest.mark.parametrize("sw,sh,dw,dh", [(100, 80, 400, 333), (173, 91, 701, 350), (50, 60, 50, 60), (30, 20, 1000, 777),
                                         (700, 466, 5591, 4004), (64, 48, 31, 17), (1, 1, 9, 5), (5, 1, 40, 7)])
def test_resize_vs_oracle(ctx, oracle, sw, sh, dw, dh):
    from simplepanorama_b200 import api
    rng = np.random.default_rng(sw * 1000 + dh)
    src = np.where(rng.random((sh, sw)) > 0.5, 255, 0).astype(np.uint8)
    src[: sh // 3] = rng.integers(0, 256, (sh // 3, sw))
    assert np.array_equal(api.resize_mask(src, (dw, dh), ctx), oracle.resize_linear_u8(src, (dw, dh)))


def test_return_full_with_preview_scale_masks(ctx, oracle):
    """The fused path takes mask_cut at preview scale (as return_full does) and resizes it on the device;
    result == resizing first and compositing with full-resolution masks, and <= 1 LSB from the oracle."""
    from simplepanorama_b200 import api, synth
    cfg = synth.config("cfg1", 0.2)
    K, R, gains = synth.cameras(cfg)
    images = synth.make_images(cfg, gains)
    plan = api.plan_tiles(images, R, K, cfg.kind, cfg.focal)
    sizes = [p[3] for p in plan]
    full_cuts = synth.seam_masks([p[2] for p in plan], sizes)
    small = [np.ascontiguousarray(c[::4, ::4]) for c in full_cuts]           # stand-in for the preview-scale masks
    fused = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, small, cfg.bands, cfg.sigma, ctx=ctx)
    resized = [api.resize_mask(m, s, ctx) for m, s in zip(small, sizes)]
    staged = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, resized, cfg.bands, cfg.sigma, ctx=ctx)
    assert np.array_equal(fused, staged)
    ref, _, _, _ = oracle.return_full(images, R, K, cfg.kind, cfg.focal, gains, small, cfg.bands, cfg.sigma)
    assert np.abs(fused.astype(int) - ref.astype(int)).max() <= 1


def test_resize_errors(ctx):
    from simplepanorama_b200 import api
    with pytest.raises(api.SpanoError):
        api.resize_mask(np.zeros((0, 0), np.uint8), (4, 4), ctx)
    with pytest.raises(api.SpanoError):
        api.resize_mask(np.zeros((4, 4), np.uint8), (0, 4), ctx)
