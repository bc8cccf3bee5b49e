"""GPU parity of the seam search by distance through the C ABI: the wavefront chamfer
transform against the oracle's sequential two-pass restatement and OpenCV's golden vectors (bit-exact floats), and
dcut::dist_cut (bit-exact masks)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_distance_transform_golden(ctx, golden):
    from simplepanorama_b200 import api
    g = golden("dist.npz")
    for name in g["dt_names"]:
        got = api.distance_transform(g[f"dt_mask_{name}"], ctx)
        ref = g[f"dt_ref_{name}"]
        if name == "far_corner":   # float ties beyond 32 px: <= 1 ulp against this OpenCV build, see tests/test_oracle_dist.py
            ulp = np.abs(got.view(np.int32).astype(np.int64) - ref.view(np.int32))
            assert ulp.max() <= 1 and (ulp != 0).mean() < 0.01
            continue
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), name


@pytest.mark.parametrize("w,h,p", [(701, 467, 0.001), (97, 1300, 0.01), (2050, 33, 0.02), (5, 3, 0.3), (1500, 1100, 0.00002)])
def test_distance_transform_vs_oracle(ctx, oracle, w, h, p):
    """sizes beyond one wavefront of 1024 threads, long thin images, distances of many hundred pixels"""
    from simplepanorama_b200 import api
    rng = np.random.default_rng(w * 7 + h)
    m = (rng.random((h, w)) > p).astype(np.uint8) * 255
    m[h // 2, w // 3] = 0
    got = api.distance_transform(m, ctx)
    ref = oracle.distance_transform(m)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_dist_cut_golden_and_oracle(ctx, oracle, golden):
    from simplepanorama_b200 import api
    g = golden("dist.npz")
    corners = [tuple(int(v) for v in c) for c in g["cut_corners"]]
    masks = [g[f"cut_mask_{i}"] for i in range(len(corners))]
    cuts = api.dist_cut(masks, corners, ctx)
    for i, c in enumerate(cuts):
        assert np.array_equal(c, g[f"cut_ref_{i}"]), i
    # preview-scale set of a real configuration: the validity masks of cfg1 at 1/4 scale
    from simplepanorama_b200 import synth
    cfg = synth.config("cfg1", 0.25)
    K, R, gains = synth.cameras(cfg)
    images = synth.make_images(cfg, gains, 0)
    pd = api.get_proj_parameters(images, R, K, [1.0] * cfg.n, cfg.kind, cfg.focal, True, ctx)
    got = api.dist_cut(pd.msks, pd.corners, ctx)
    ref = oracle.dist_cut(pd.msks, pd.corners)
    assert all(np.array_equal(a, b) for a, b in zip(got, ref))
    # every canvas pixel covered by some valid tile keeps at least one owner
    W, H, mx, my = api.pan_dimension(pd.corners, [(m.shape[1], m.shape[0]) for m in pd.msks])
    cover = np.zeros((H, W), bool); owned = np.zeros((H, W), bool)
    for m, c, (x, y) in zip(pd.msks, got, pd.corners):
        cover[y - my:y - my + m.shape[0], x - mx:x - mx + m.shape[1]] |= m > 0
        owned[y - my:y - my + m.shape[0], x - mx:x - mx + m.shape[1]] |= c > 0
    assert np.array_equal(cover, owned)


def test_dist_errors(ctx):
    from simplepanorama_b200 import api
    with pytest.raises(api.SpanoError):
        api.dist_cut([], [], ctx)
    with pytest.raises(api.SpanoError):
        api.dist_cut([np.zeros((4, 4), np.uint8)], [(0, 0), (1, 1)], ctx)
    with pytest.raises(api.SpanoError):
        api.distance_transform(np.zeros((0, 0), np.uint8), ctx)


def test_simple_and_no_blend(ctx, oracle, golden):
    """The SIMPLE_BLEND and NO_BLEND branches of stitch_parameters::blend: <= 1 LSB / bit-exact."""
    from simplepanorama_b200 import api
    g = golden("dist.npz")
    corners = [tuple(int(v) for v in c) for c in g["cut_corners"]]
    tiles = [g[f"blend_tile_{i}"] for i in range(len(corners))]
    masks = [g[f"blend_mask_{i}"] for i in range(len(corners))]
    s = api.simple_blend(tiles, masks, corners, ctx)
    assert np.abs(s.astype(int) - g["simple_ref"].astype(int)).max() <= 1
    assert np.abs(s.astype(int) - oracle.simple_blend(tiles, masks, corners).astype(int)).max() <= 1
    assert np.array_equal(api.no_blend(tiles, masks, corners, ctx), g["noblend_ref"])
    # a real configuration: warped tiles + validity masks of cfg1 at 1/4 scale
    from simplepanorama_b200 import synth
    cfg = synth.config("cfg1", 0.25)
    K, R, gains = synth.cameras(cfg)
    images = synth.make_images(cfg, gains, 0)
    pd = api.get_proj_parameters(images, R, K, [1.0] * cfg.n, cfg.kind, cfg.focal, True, ctx)
    got = api.simple_blend(pd.imgs, pd.msks, pd.corners, ctx)
    ref = oracle.simple_blend(pd.imgs, pd.msks, pd.corners)
    assert got.shape == ref.shape and np.abs(got.astype(int) - ref.astype(int)).max() <= 1
    cuts = api.dist_cut(pd.msks, pd.corners, ctx)
    assert np.array_equal(api.no_blend(pd.imgs, cuts, pd.corners, ctx), oracle.no_blend(pd.imgs, cuts, pd.corners))
    with pytest.raises(api.SpanoError):
        api.simple_blend(tiles, masks[:-1], corners, ctx)       # "Input consistency!"


def test_overlap_intensity(ctx, oracle, golden):
    """gain::get_overlapp_intensity (the reduction feeding the gain solve): exact."""
    from simplepanorama_b200 import api
    g = golden("dist.npz")
    corners = [tuple(int(v) for v in c) for c in g["cut_corners"]]
    tiles = [g[f"ov_tile_{i}"] for i in range(len(corners))]
    got = np.array(api.get_overlapp_intensity(tiles, corners, g["ov_adj"], ctx), np.float64)
    assert np.array_equal(got, g["ov_ref"])
    # real warped tiles (cfg1 at 1/4 scale), chain adjacency
    from simplepanorama_b200 import synth
    cfg = synth.config("cfg1", 0.25)
    K, R, gains = synth.cameras(cfg)
    images = synth.make_images(cfg, gains, 0)
    pd = api.get_proj_parameters(images, R, K, [1.0] * cfg.n, cfg.kind, cfg.focal, False, ctx)
    adj = np.zeros((cfg.n, cfg.n)); idx = np.arange(cfg.n - 1); adj[idx, idx + 1] = adj[idx + 1, idx] = 1
    got = api.get_overlapp_intensity(pd.imgs, pd.corners, adj, ctx)
    ref = oracle.overlap_intensity(pd.imgs, pd.corners, adj)
    assert got == ref and len(got) == 2 * cfg.n - 1 and all(r[2] > 0 for r in got)
    with pytest.raises(api.SpanoError):
        api.get_overlapp_intensity(tiles, corners[:-1], g["ov_adj"], ctx)
