"""GPU parity: the CUDA path, called through the C ABI (ctypes -> libspano.so), against the C oracle
on the same seeded inputs and against the golden vectors OpenCV 4.13 produced.

Bars (BASELINE.json north_star): bit-exact for integer/byte/index work (ROI, fixed-point
sampler on given maps, gray/dark flag, flood fill, erosion, gain, u8 conversion of a given
float); float intermediates within 1e-5 relative; <= 1 LSB per channel on the 8-bit canvas.
Coordinate generation evaluates sinf/cosf/atan2f/atanf with the host libm's own arithmetic
(csrc/glibc_trig.cuh, checked exhaustively against glibc in tests/test_trig_port.py), so the float maps and
therefore the warped tiles are BIT-IDENTICAL to OpenCV's for all three projections: no bin-slip allowance.
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FLOAT_RTOL = 1e-5   # north_star: float intermediates within 1e-5 relative
U8_TOL = 1          # north_star: <= 1 LSB per channel on the 8-bit canvas


def _rot(yaw, pitch, roll):
    from simplepanorama_b200 import synth
    return synth.rotation(math.degrees(yaw), math.degrees(pitch), math.degrees(roll))


# ------------------------------------------------------------------ a3: maps / sampler / warp
def test_remap_kat_bit_exact(ctx, golden):
    """Fixed-point sampler on given maps incl. out-of-range / NaN / huge coordinates: exact."""
    from simplepanorama_b200 import api
    g = golden("kernels.npz")
    out = api.remap(g["remap_img"], g["remap_x"], g["remap_y"], ctx)
    assert np.array_equal(out, g["remap_dst"])


@pytest.mark.parametrize("aligned", [True, False])
def test_remap_random_bit_exact(ctx, oracle, aligned):
    from simplepanorama_b200 import api
    rng = np.random.default_rng(11)
    W = 328 if aligned else 331   # 3*W multiple of 8 or not: both sampling paths
    H = 203
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    xm = rng.uniform(-4, W + 4, (240, 517)).astype(np.float32)
    ym = rng.uniform(-4, H + 4, (240, 517)).astype(np.float32)
    assert np.array_equal(api.remap(img, xm, ym, ctx), oracle.remap(img, xm, ym))


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_build_maps_bit_identical(ctx, oracle, kind):
    """Float maps bit-identical to the oracle's (which are bit-identical to cv2's buildMaps)."""
    from simplepanorama_b200 import api
    W, H, f = 640, 480, 520.0
    K = np.array([[f * 1.03, 0, W / 2 + 3.5], [0, f * 1.03, H / 2 - 2.25], [0, 0, 1]])
    R = _rot(0.3, -0.35, 0.02)
    K32, R32 = api.adjusted_camera(K, R, W, H)
    tl, size = api.warp_roi(kind, f, K32, R32, W, H, ctx)
    assert (tl, size) == oracle.warp_roi(kind, np.float32(f), K32, R32, W, H)
    xm, ym = api.build_maps(kind, f, K32, R32, tl, size, ctx)
    xo, yo = oracle.build_maps(kind, np.float32(f), K32, R32, tl, size)
    assert np.array_equal(xm.view(np.uint32), xo.view(np.uint32))
    assert np.array_equal(ym.view(np.uint32), yo.view(np.uint32))


def test_warp_golden_cases(ctx, golden):
    """End-to-end warp of the OpenCV-made fixtures (noise and band-limited images, all projections): corner, tile
    and validity mask bit-identical to cv2's."""
    from simplepanorama_b200 import api
    g = golden("warp_cases.npz")
    for i in range(int(g["warp_count"])):
        kind, f = int(g[f"warp{i}_kind"]), float(g[f"warp{i}_focal"])
        corner, tile, mask = api.project(kind, f, g[f"warp{i}_R"], g[f"warp{i}_K"], g[f"warp{i}_img"], 1.0, True, ctx)
        ref = g[f"warp{i}_tile"]
        assert tuple(corner) == tuple(g[f"warp{i}_corner"]) and tile.shape == ref.shape
        assert np.array_equal(tile, ref), (i, kind, int((tile != ref).sum()))
        assert np.array_equal(mask, g[f"warp{i}_mask"]), (i, kind)


@pytest.mark.parametrize("name,scale", [("cfg1", 0.25), ("cfg2", 0.06), ("cfg3", 0.08)])
def test_warp_vs_oracle_band_limited(ctx, oracle, name, scale):
    from simplepanorama_b200 import api, synth
    cfg = synth.config(name, scale)
    K, R, gains = synth.cameras(cfg)
    for j in (0, cfg.n // 2, cfg.n - 1):
        img = synth.make_image(cfg, j, gains[j])
        K32, R32 = api.adjusted_camera(K[j], R[j], cfg.width, cfg.height)
        tl_o, tile_o = oracle.warp(cfg.kind, np.float32(cfg.focal), K32, R32, img)
        corner, tile, mask = api.project(cfg.kind, cfg.focal, R[j], K[j], img, 1.0, True, ctx)
        assert corner == tl_o
        assert np.array_equal(tile, tile_o), int((tile != tile_o).sum())
        assert np.array_equal(mask, oracle.surrounding_mask(tile_o, 3))
        # the fused gain is exactly the 8-bit gain applied to the un-gained warp
        _, gained, mask2 = api.project(cfg.kind, cfg.focal, R[j], K[j], img, gains[j], True, ctx)
        assert np.array_equal(gained, oracle.apply_gain(tile_o, gains[j]))
        assert np.array_equal(mask2, mask)


# ------------------------------------------------------------------ a4 / a6: masks and gain
def test_mask_golden(ctx, golden):
    from simplepanorama_b200 import api
    g = golden("kernels.npz")
    assert np.array_equal(api.create_surrounding_mask(g["mask_img"], 0, ctx), g["mask_raw"])
    assert np.array_equal(api.validity_mask(g["mask_img"], ctx), g["mask_eroded"])


@pytest.mark.parametrize("seed,w,h", [(0, 257, 131), (1, 64, 64), (2, 1031, 517), (3, 33, 700), (4, 5, 3), (5, 1, 1)])
def test_mask_random_structures(ctx, oracle, seed, w, h):
    """Flood fill on adversarial dark structures (mazes, spirals, islands, dark noise): exact."""
    from simplepanorama_b200 import api
    rng = np.random.default_rng(seed)
    img = rng.integers(16, 240, (h, w, 3), dtype=np.uint8)
    dark = rng.random((h, w)) < 0.45                       # percolating noise
    if w > 40 and h > 40:
        dark[h // 3: h // 3 + 3, : w - 7] = True           # long corridors
        dark[h // 3: 2 * h // 3, w - 10: w - 7] = True
        dark[2 * h // 3 - 3: 2 * h // 3, 5: w - 7] = True
        dark[h // 2 - 4: h // 2 + 4, w // 2 - 4: w // 2 + 4] = False
    img[dark] = rng.integers(0, 2, (int(dark.sum()), 3))
    for it in (0, 3):
        assert np.array_equal(api.create_surrounding_mask(img, it, ctx), oracle.surrounding_mask(img, it))


def test_mask_spiral(ctx, oracle):
    """A one-pixel-wide spiral corridor reaching the border: the longest possible label chain."""
    from simplepanorama_b200 import api
    n = 201
    img = np.full((n, n, 3), 200, np.uint8)
    x = y = 0
    dx, dy = 1, 0
    lo, hi = 0, n - 1
    seen = np.zeros((n, n), bool)
    for _ in range(n * n):
        seen[y, x] = True
        nx, ny = x + dx, y + dy
        if not (0 <= nx < n and 0 <= ny < n) or seen[ny, nx] or (0 <= nx + dx < n and 0 <= ny + dy < n and seen[ny + dy, nx + dx]):
            dx, dy = -dy, dx
            nx, ny = x + dx, y + dy
            if not (0 <= nx < n and 0 <= ny < n) or seen[ny, nx] or (0 <= nx + dx < n and 0 <= ny + dy < n and seen[ny + dy, nx + dx]):
                break
        x, y = nx, ny
    img[seen] = 0
    assert np.array_equal(api.create_surrounding_mask(img, 0, ctx), oracle.surrounding_mask(img, 0))


def test_gain_exact(ctx, golden, oracle):
    from simplepanorama_b200 import api
    g = golden("kernels.npz")
    a = np.arange(256, dtype=np.uint8).reshape(1, -1, 1).repeat(3, axis=2).copy()
    for k, gain in enumerate(g["gain_values"]):
        assert np.array_equal(api.apply_gain(a, float(gain), ctx)[0, :, 0], g["gain_table"][k])
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (77, 129, 3), dtype=np.uint8)
    assert np.array_equal(api.apply_gain(img, 1.1732, ctx), oracle.apply_gain(img, 1.1732))


# ------------------------------------------------------------------ a7 / a8 / a10: the blend
def _blend_inputs(golden):
    g = golden("blend_cases.npz")
    tiles = [g[f"blend_tile{j}"] for j in range(3)]
    cuts = [g[f"blend_cut{j}"] for j in range(3)]
    valids = [g[f"blend_valid{j}"] for j in range(3)]
    corners = [tuple(int(v) for v in c) for c in g["blend_corners"]]
    return g, tiles, cuts, valids, corners


def _assert_float_close(a, ref):
    tol = FLOAT_RTOL * max(1.0, float(np.abs(ref).max()))
    err = float(np.abs(a - ref).max())
    assert err <= tol, (err, tol)


@pytest.mark.parametrize("key,bands,sigma", [("B1", 1, 7.0), ("B2", 2, 7.0), ("B3", 3, 7.0), ("B5", 5, 7.0), ("B4_s3", 4, 3.0)])
def test_multi_blend_golden(ctx, golden, key, bands, sigma):
    """Against cv2.GaussianBlur-made fixtures (3 overlapping tiles, one smaller than the blur radius)."""
    from simplepanorama_b200 import api
    g, tiles, cuts, valids, corners = _blend_inputs(golden)
    f = api.multi_blend(tiles, cuts, valids, corners, bands, sigma, ctx)
    _assert_float_close(f, g[f"blend_f32_{key}"])
    u8 = api.blend(tiles, cuts, valids, corners, bands, sigma, ctx)
    assert np.abs(u8.astype(int) - g[f"blend_u8_{key}"].astype(int)).max() <= U8_TOL


@pytest.mark.parametrize("bands", list(range(1, 11)))
def test_multi_blend_all_band_counts(ctx, oracle, bands):
    from simplepanorama_b200 import api
    rng = np.random.default_rng(100 + bands)
    sizes = [(150, 97), (131, 140), (90, 60)]
    corners = [(0, 10), (100, -20), (60, 70)]
    tiles = [rng.integers(16, 240, (h, w, 3), dtype=np.uint8) for (w, h) in sizes]
    valids = []
    cuts = []
    for (w, h) in sizes:
        v = np.full((h, w), 255, np.uint8); v[:5] = 0; v[:, -4:] = 0
        valids.append(v)
        c = (rng.random((h, w)) * 255).astype(np.uint8); c[:, : w // 2] = 255
        cuts.append(c)
    ref = oracle.multi_blend(tiles, cuts, valids, corners, bands, 7.0)
    f = api.multi_blend(tiles, cuts, valids, corners, bands, 7.0, ctx)
    _assert_float_close(f, ref)
    u8 = api.blend(tiles, cuts, valids, corners, bands, 7.0, ctx)
    assert np.abs(u8.astype(int) - oracle.blend_to_u8(ref).astype(int)).max() <= U8_TOL


@pytest.mark.parametrize("sigma", [1.0, 2.5, 5.0, 10.0])
def test_multi_blend_other_sigmas(ctx, oracle, sigma):
    """Radius != 21 takes the generic-radius kernel."""
    from simplepanorama_b200 import api
    rng = np.random.default_rng(7)
    tiles = [rng.integers(16, 240, (80, 120, 3), dtype=np.uint8), rng.integers(16, 240, (70, 90, 3), dtype=np.uint8)]
    ones = [np.full(t.shape[:2], 255, np.uint8) for t in tiles]
    cuts = [(rng.random(t.shape[:2]) * 255).astype(np.uint8) for t in tiles]
    corners = [(0, 0), (70, 30)]
    ref = oracle.multi_blend(tiles, cuts, ones, corners, 3, sigma)
    _assert_float_close(api.multi_blend(tiles, cuts, ones, corners, 3, sigma, ctx), ref)


@pytest.mark.parametrize("bands", [1, 2, 5, 6, 7, 10])
def test_blend_kernels_agree(ctx, oracle, bands):
    """Warp-specialised marching kernel (default), the 8-warp marching kernel (same arithmetic in the same order:
    BIT-identical) and the generic-radius kernel (different summation order: within 1e-5), all against the oracle."""
    from simplepanorama_b200 import api
    rng = np.random.default_rng(40 + bands)
    sizes = [(233, 310), (75, 40), (19, 300), (640, 97)]     # taller than a segment, smaller than the radius, narrow, wide
    corners = [(0, 0), (200, 100), (120, 5), (10, 150)]
    tiles = [rng.integers(16, 240, (h, w, 3), dtype=np.uint8) for (w, h) in sizes]
    cuts = [(rng.random((h, w)) * 255).astype(np.uint8) for (w, h) in sizes]
    cuts[3][:, 200:520] = 0                                   # a sparse tile: several pieces per CTA
    valids = [np.where(rng.random((h, w)) < 0.9, 255, 0).astype(np.uint8) for (w, h) in sizes]
    outs = []
    for mode in (0, 2, 3, 4, 1):
        ctx.set_option(ctx.OPT_BLEND_KERNEL, mode)
        try:
            outs.append(api.multi_blend(tiles, cuts, valids, corners, bands, 7.0, ctx))
        finally:
            ctx.set_option(ctx.OPT_BLEND_KERNEL, 0)
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    assert np.array_equal(outs[0].view(np.uint32), outs[2].view(np.uint32))
    assert np.array_equal(outs[0].view(np.uint32), outs[3].view(np.uint32))
    _assert_float_close(outs[0], outs[4])
    _assert_float_close(outs[0], oracle.multi_blend(tiles, cuts, valids, corners, bands, 7.0))


def test_single_tile_closed_form(ctx, oracle):
    """All-ones masks: sum of the reference's bands = I + G0 + G1 - 2 G_{B-1} (they do not telescope)."""
    from simplepanorama_b200 import api
    rng = np.random.default_rng(5)
    tile = rng.integers(16, 240, (140, 200, 3), dtype=np.uint8)
    ones = np.full(tile.shape[:2], 255, np.uint8)
    B, sigma = 6, 7.0
    out = api.multi_blend([tile], [ones], [ones], [(0, 0)], B, sigma, ctx)
    I = tile.astype(np.float32)
    G = [oracle.gaussian_blur(I, 43, math.sqrt(2 * (B - i) + 1) * sigma) for i in range(B)]
    expect = (I + G[0] + G[1] - 2 * G[B - 1]) / B / float(255 // B)
    assert np.allclose(out, expect, rtol=2e-5, atol=2e-6)


def test_uncovered_canvas_is_black(ctx):
    """alpha == 0 -> clamp to 1e-6 -> 0/1e-6 = 0 (elementwiseOperation DIVIDE)."""
    from simplepanorama_b200 import api
    t = np.full((30, 30, 3), 100, np.uint8)
    ones = np.full((30, 30), 255, np.uint8)
    out = api.blend([t, t], [ones, ones], [ones, ones], [(0, 0), (60, 50)], 2, 7.0, ctx)
    assert out.shape == (80, 90, 3)
    assert out[40:50, 35:60].max() == 0 and out[5:25, 5:25].min() > 0


# ------------------------------------------------------------------ a1: the fused path
def _fused_case(name, scale, noise=0):
    from simplepanorama_b200 import api, synth
    cfg = synth.config(name, scale)
    K, R, gains = synth.cameras(cfg)
    images = synth.make_images(cfg, gains, noise)
    plan = api.plan_tiles(images, R, K, cfg.kind, cfg.focal)
    cuts = synth.seam_masks([p[2] for p in plan], [p[3] for p in plan])
    return cfg, K, R, gains, images, cuts


@pytest.mark.parametrize("name,scale", [("cfg1", 0.2), ("cfg2", 0.04), ("cfg3", 0.06)])
def test_return_full_vs_oracle(ctx, oracle, name, scale):
    """stitch_parameters::return_full: sources -> 8-bit canvas, <= 1 LSB per channel."""
    from simplepanorama_b200 import api
    cfg, K, R, gains, images, cuts = _fused_case(name, scale)
    ref, tiles, msks, corners = oracle.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma)
    out = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx)
    assert out.shape == ref.shape
    assert np.abs(out.astype(int) - ref.astype(int)).max() <= U8_TOL
    # the staged API gives the same canvas as the fused one
    pd = api.get_proj_parameters(images, R, K, [1.0] * cfg.n, cfg.kind, cfg.focal, True, ctx)
    gained = [api.apply_gain(t, g, ctx) for t, g in zip(pd.imgs, gains)]
    staged = api.blend(gained, cuts, pd.msks, pd.corners, cfg.bands, cfg.sigma, ctx)
    assert np.array_equal(staged, out)


def test_row_bands_equal_full_canvas(ctx):
    """Row-band sharding: any partition of the canvas rows reproduces the full canvas bit for bit."""
    from simplepanorama_b200 import api, dist
    cfg, K, R, gains, images, cuts = _fused_case("cfg1", 0.2)
    full = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx)
    plan = api.plan_tiles(images, R, K, cfg.kind, cfg.focal)
    W, H, mx, my = api.pan_dimension([p[2] for p in plan], [p[3] for p in plan])
    for world in (2, 3, 8):
        bands = dist.plan_row_bands([(p[2], p[3]) for p in plan], world, my, H)
        parts = [api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, rows=b, ctx=ctx)
                 for b in bands if b[1] > b[0]]
        assert np.array_equal(np.concatenate(parts, axis=0), full)


# ------------------------------------------------------------------ error behaviour
def test_errors(ctx):
    from simplepanorama_b200 import api
    t = np.zeros((10, 10, 3), np.uint8); m = np.zeros((10, 10), np.uint8)
    with pytest.raises(api.SpanoError):                      # "Input consistency!" (simple_blend's check)
        api.multi_blend([t], [m, m], [m], [(0, 0)], 2, 7.0, ctx)
    with pytest.raises(api.SpanoError):
        api.multi_blend([], [], [], [], 2, 7.0, ctx)
    with pytest.raises(api.SpanoError):
        api.multi_blend([t], [m], [m], [(0, 0)], 0, 7.0, ctx)      # bands < 1 (255/0 in the reference)
    with pytest.raises(api.SpanoError):
        api.multi_blend([t], [m], [m], [(0, 0)], 11, 7.0, ctx)
    with pytest.raises(api.SpanoError):
        api.multi_blend([t], [m], [m], [(0, 0)], 2, 0.0, ctx)
    with pytest.raises(api.SpanoError) as e:
        api.multi_blend([t], [m], [m], [(0, 0)], 2, 20.0, ctx)     # radius 60 > supported 32
    assert e.value.code == -5
    with pytest.raises(api.SpanoError):
        api.apply_gain(t, 0.0, ctx)
    with pytest.raises(api.SpanoError):
        api.create_surrounding_mask(np.zeros((0, 0, 3), np.uint8), 0, ctx)
    with pytest.raises(api.SpanoError):
        api.remap(np.zeros((4, 40000, 3), np.uint8), np.zeros((2, 2), np.float32), np.zeros((2, 2), np.float32), ctx)
    # the context survives errors
    assert api.apply_gain(t + 10, 2.0, ctx)[0, 0, 0] == 5


def test_band_composite_with_supplied_masks(ctx):
    """Multi-GPU data path on one GPU: validity masks computed per tile with spano_dev_tile_mask (as the
    owning rank would), handed to spano_dev_composite, which then warps only the rows each band reads.
    Every band must equal the corresponding rows of the ordinary full-canvas result, bit for bit."""
    import ctypes as C
    import torch
    from simplepanorama_b200 import api, dist
    cfg, K, R, gains, images, cuts = _fused_case("cfg1", 0.25)
    full = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx)
    plan = api.plan_tiles(images, R, K, cfg.kind, cfg.focal)
    W, H, mx, my = api.pan_dimension([p[2] for p in plan], [p[3] for p in plan])
    dev = torch.device("cuda", 0)
    d_img = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in images]
    d_cut = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in cuts]
    d_val = [torch.zeros((p[3][1], p[3][0]), dtype=torch.uint8, device=dev) for p in plan]
    torch.cuda.synchronize()
    descs = api.make_descs(d_img, plan, gains, d_cut, lambda t: t.data_ptr(), lambda t: t.stride(0))
    lib = ctx.lib
    for j in range(cfg.n):
        d = descs[j]
        ctx.check(lib.spano_dev_tile_mask(ctx.h, cfg.kind, C.c_float(cfg.focal), d.K, d.R, d.src_bgr, d.src_w, d.src_h, d.src_step,
                                          d.tl_x, d.tl_y, d.w, d.h, d_val[j].data_ptr(), d_val[j].stride(0)))
        d.valid_mask = d_val[j].data_ptr()
        d.valid_mask_step = d_val[j].stride(0)
    for world in (1, 3, 7):
        bands = dist.plan_row_bands([(p[2], p[3]) for p in plan], world, my, H)
        parts = []
        for (a, b) in bands:
            if b <= a:
                continue
            out = torch.empty((b - a, W, 3), dtype=torch.uint8, device=dev)
            ctx.check(lib.spano_dev_composite(ctx.h, cfg.kind, C.c_float(cfg.focal), cfg.n, descs, cfg.bands, cfg.sigma, a, b,
                                              out.data_ptr(), out.stride(0)))
            ctx.sync()
            parts.append(out.cpu().numpy())
        assert np.array_equal(np.concatenate(parts, axis=0), full), world


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("width", [640, 643])
def test_tma_staged_warp_equals_plain_warp(ctx, oracle, kind, width):
    """The TMA-staged warp kernel (SPANO_OPT_WARP_KERNEL = 1, taken when the source pitch is a multiple of 16 bytes) and the
    un-staged kernel (also what a source with another pitch gets) produce the same bytes, and both equal the oracle:
    noise image, strong rotation (the footprint of a block is then much larger than the block: blocks fall back), a pose
    that leaves part of the tile outside the source (border taps)."""
    from simplepanorama_b200 import api
    rng = np.random.default_rng(70 + kind)
    W, H, f = width, 480, 520.0
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    for (yaw, pitch, roll) in ((0.3, -0.35, 0.02), (0.1, -0.9 if kind == 2 else 0.2, 0.8)):
        K = np.array([[f * 1.03, 0, W / 2 + 3.5], [0, f * 1.03, H / 2 - 2.25], [0, 0, 1]])
        R = _rot(yaw, pitch, roll)
        K32, R32 = api.adjusted_camera(K, R, W, H)
        _, ref = oracle.warp(kind, np.float32(f), K32, R32, img)
        outs = []
        for mode in (0, 1):
            ctx.set_option(ctx.OPT_WARP_KERNEL, mode)
            try:
                outs.append(api.project(kind, f, R, K, img, 1.3, True, ctx))
            finally:
                ctx.set_option(ctx.OPT_WARP_KERNEL, 0)
        assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
        assert np.array_equal(outs[0][1], oracle.apply_gain(ref, 1.3))
