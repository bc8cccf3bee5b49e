"""GPU parity of test::equalizeIntensities (src/test/_test.cpp:9-106) through the C ABI (spano_equalize_intensities):
against the fields OpenCV 4.13 produced (tests/golden/equalize.npz: even sizes = the 2x2 area path of cv::resize, odd
sizes = the linear path) and against cv2 on a synthetic panorama's preview warps.  Bar: 1e-5 relative (float field)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _close(a, b):
    assert a.shape == b.shape
    scale = float(np.abs(b).max())
    assert float(np.abs(a - b).max()) <= 1e-5 * scale, (float(np.abs(a - b).max()), scale)


@pytest.mark.parametrize("name", ["even", "odd"])
def test_equalize_golden(ctx, golden, name):
    from simplepanorama_b200 import api
    g = golden("equalize.npz")
    n = int(g[f"{name}_n"])
    tiles = [g[f"{name}_tile{k}"] for k in range(n)]
    masks = [g[f"{name}_mask{k}"] for k in range(n)]
    corners = [tuple(int(v) for v in c) for c in g[f"{name}_corners"]]
    fields = api.equalize_intensities(tiles, masks, corners, 0.5, ctx)
    for k in range(n):
        _close(fields[k], g[f"{name}_field{k}"])


def test_equalize_on_preview_warps_and_feeds_adjust_intensity(ctx):
    cv2 = pytest.importorskip("cv2")
    from simplepanorama_b200 import api, synth
    from oracle import cv2_ref
    cfg = synth.config("cfg1", 0.3)
    K, R, gains = synth.cameras(cfg)
    images = synth.make_images(cfg, gains)
    tiles, masks, corners = [], [], []
    for j in range(cfg.n):
        c, t, m = api.project(cfg.kind, cfg.focal, R[j], K[j], images[j], 1.0, True, ctx)
        tiles.append(t); masks.append(m); corners.append(tuple(c))
    ref = cv2_ref.equalize_intensities(tiles, masks, corners, 0.5)
    got = api.equalize_intensities(tiles, masks, corners, 0.5, ctx)
    for a, b in zip(got, ref):
        _close(a, b)
    # the fields are what test::adjust_intensity divides by
    adj = api.adjust_intensity(tiles[0], got[0], ctx)
    assert np.abs(adj.astype(int) - cv2_ref.adjust_intensity(tiles[0], ref[0]).astype(int)).max() <= 1


def test_equalize_errors(ctx):
    from simplepanorama_b200 import api
    t = np.zeros((8, 8, 3), np.uint8); m = np.zeros((8, 8), np.uint8)
    with pytest.raises(api.SpanoError):
        api.equalize_intensities([t], [m, m], [(0, 0)], 0.5, ctx)
    with pytest.raises(api.SpanoError):
        api.equalize_intensities([t], [m], [(0, 0)], 2.0, ctx)
