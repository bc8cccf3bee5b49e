"""GPU: the mask_cut sparsity of the blend is EXACT.  A tile pixel whose whole 43x43 window of mask_cut is zero
has weight 0 in every band (blnd::multi_blend, src/math/_blending.cpp:205-222: the weight is GaussianBlur(mask)),
so the marching kernel skips it; the canvas must be bit-identical to the dense evaluation of the same kernel
(SPANO_OPT_BLEND_DENSE) and within 1e-5 of the oracle, for adversarial sparsity patterns."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _patterns(rng, w, h):
    yy, xx = np.mgrid[0:h, 0:w]
    pats = {}
    pats["empty"] = np.zeros((h, w), np.uint8)
    one = np.zeros((h, w), np.uint8); one[h // 2, w // 3] = 1
    pats["one_pixel_value_1"] = one
    corner = np.zeros((h, w), np.uint8); corner[0, 0] = 255; corner[h - 1, w - 1] = 7
    pats["two_corners"] = corner
    blobs = np.zeros((h, w), np.uint8)
    for _ in range(4):
        cx, cy, r = rng.integers(0, w), rng.integers(0, h), rng.integers(3, 40)
        blobs[(xx - cx) ** 2 + (yy - cy) ** 2 < r * r] = rng.integers(1, 256)
    pats["blobs"] = blobs
    line = np.zeros((h, w), np.uint8); line[:, w // 2] = 200; line[h // 4, :] = 90
    pats["cross_lines"] = line
    ring = ((np.hypot(xx - w / 2, yy - h / 2) > min(w, h) / 3) * 255).astype(np.uint8)
    pats["hole"] = ring
    band = np.zeros((h, w), np.uint8); band[:, : max(1, w // 5)] = 255
    pats["left_band"] = band
    pats["dense_noise"] = rng.integers(0, 256, (h, w), dtype=np.uint8)
    return pats


@pytest.mark.parametrize("bands", [1, 3, 6, 7, 10])
def test_sparse_blend_is_bit_identical_to_dense(ctx, oracle, bands):
    from simplepanorama_b200 import api
    rng = np.random.default_rng(100 + bands)
    sizes = [(333, 217), (150, 260), (97, 64), (40, 300)]
    corners = [(0, 0), (250, -30), (120, 150), (310, 10)]
    tiles = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (w, h) in sizes]
    valids = []
    for (w, h) in sizes:
        v = np.full((h, w), 255, np.uint8); v[: h // 7] = 0
        valids.append(v)
    names = list(_patterns(rng, 8, 8).keys())
    for trial in range(len(names)):
        cuts = []
        for t, (w, h) in enumerate(sizes):
            pats = _patterns(rng, w, h)
            cuts.append(pats[names[(trial + t) % len(names)]])
        ctx.set_option(ctx.OPT_BLEND_DENSE, 1)
        dense = api.multi_blend(tiles, cuts, valids, corners, bands, 7.0, ctx)
        ctx.set_option(ctx.OPT_BLEND_DENSE, 0)
        ctx.blend_stats(reset=True)
        sparse = api.multi_blend(tiles, cuts, valids, corners, bands, 7.0, ctx)
        done, offered = ctx.blend_stats()
        assert offered == sum(w * h for (w, h) in sizes)
        assert np.array_equal(dense.view(np.uint32), sparse.view(np.uint32)), f"trial {trial}"
        if trial == 0:
            ref = oracle.multi_blend(tiles, cuts, valids, corners, bands, 7.0)
            scale = max(1e-6, float(np.abs(ref).max()))
            assert float(np.abs(sparse - ref).max()) <= 1e-5 * scale


def test_sparse_blend_skips_work(ctx):
    """An all-zero mask_cut costs nothing, a narrow band costs about its dilated width."""
    from simplepanorama_b200 import api
    rng = np.random.default_rng(7)
    w, h = 1600, 600
    tile = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    valid = np.full((h, w), 255, np.uint8)
    cut = np.zeros((h, w), np.uint8)
    ctx.blend_stats(reset=True)
    out = api.multi_blend([tile], [cut], [valid], [(0, 0)], 6, 7.0, ctx)
    done, offered = ctx.blend_stats(reset=True)
    assert offered == w * h and done == 0
    assert not out.any()
    cut[:, 640:960] = 255          # strips 20..29 -> with the neighbour strips 19..30: 12 strips of 32 columns
    api.multi_blend([tile], [cut], [valid], [(0, 0)], 6, 7.0, ctx)
    done, offered = ctx.blend_stats(reset=True)
    assert done == 12 * 32 * h


def test_two_contexts_with_different_bands_concurrently(ctx):
    """One spano_ctx per thread (one pan::panorama per viewer window): the Gaussian tap tables are per device, so two
    contexts blending with different (bands, sigma) at the same time must not see each other's tables."""
    import threading
    from simplepanorama_b200 import api
    rng = np.random.default_rng(11)
    w, h = 700, 300
    tile = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    cut = rng.integers(0, 256, (h, w), dtype=np.uint8)
    valid = np.full((h, w), 255, np.uint8)
    args = ([tile], [cut], [valid], [(0, 0)])
    want = {b: api.multi_blend(*args, b, 7.0, ctx) for b in (3, 6)}
    other = api.Context(0)
    errors = []

    def worker(c, bands):
        try:
            for _ in range(12):
                got = api.multi_blend(*args, bands, 7.0, c)
                if not np.array_equal(got.view(np.uint32), want[bands].view(np.uint32)):
                    errors.append(bands)
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=worker, args=(ctx, 3)), threading.Thread(target=worker, args=(other, 6))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    other.close()
    assert not errors, errors
