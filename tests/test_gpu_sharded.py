"""GPU parity of the tile-sharded multi-GPU path (include/spano.h: spano_*_warp_scatter, spano_*_blend_*),
run on ONE device: the G band arenas are plain device buffers, so the owner-side scatter (warp + validity
mask kernels storing rows into every band's arena) and the band-side incremental blend are exercised exactly
as on G GPUs, minus the NVLink hop (the IPC mapping itself is covered by bench.py --gpus 2).
Bar: the bands concatenate to the single-GPU canvas of stitch_parameters::return_full BIT FOR BIT."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(name, scale, coarse):
    from simplepanorama_b200 import api, synth
    cfg = synth.config(name, scale)
    K, R, gains = synth.cameras(cfg)
    images = synth.make_images(cfg, gains, 0)
    plan = api.plan_tiles(images, R, K, cfg.kind, cfg.focal)
    corners, sizes = [p[2] for p in plan], [p[3] for p in plan]
    cuts = [synth.seam_masks(corners, sizes, only=j, coarse=True) for j in range(cfg.n)] if coarse else synth.seam_masks(corners, sizes)
    return cfg, K, R, gains, images, plan, cuts


def _run_sharded(ctx, cfg, gains, images, plan, cuts, world, host):
    import torch
    from simplepanorama_b200 import api, dist
    dev = torch.device("cuda", 0)
    sp = dist.plan_tile_shards([p[2] for p in plan], [p[3] for p in plan], world, cfg.sigma)
    arenas = [torch.zeros(sp.arena_bytes[k], dtype=torch.uint8, device=dev) for k in range(world)]
    ptrs = [a.data_ptr() for a in arenas]
    if host:
        imgs = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in images]
        cts = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in cuts]
    else:
        imgs = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in images]
        cts = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in cuts]
    torch.cuda.synchronize()
    descs = api.make_descs(imgs, plan, gains, cts, lambda t: t.data_ptr(), lambda t: t.stride(0))
    for j in range(cfg.n):                       # owner side (every "rank" in turn)
        dist.scatter_tile(ctx, sp, j, descs[j], ptrs, cfg.kind, cfg.focal, host=host)
    parts = []
    for k in range(world):                       # band side
        r0, r1 = sp.bands[k]
        if r1 <= r0:
            continue
        # host variant: announce the images (and the canvas, for the early column download) for even bands only, so
        # both the staged / early-flush and the on-demand paths run
        out = torch.zeros((r1 - r0, sp.canvas_w, 3), dtype=torch.uint8).pin_memory() if host else None
        announce = host and k % 2 == 0
        dist.blend_begin(ctx, sp, k, cfg.bands, cfg.sigma, host_descs=descs if announce else None,
                         host_canvas=(out.data_ptr(), out.stride(0)) if announce else (0, 0))
        if k % 3 != 1:     # with and without the ahead-of-time preparation (mask up-scaling + plans on the aux stream)
            dist.blend_prepare(ctx, sp, k, descs, ptrs[k], host=host)
        for j in range(cfg.n):
            dist.blend_add(ctx, sp, k, j, descs, ptrs[k], host=host)
        if host:
            dist.blend_finish(ctx, out.data_ptr(), out.stride(0), host=True)
            parts.append(out.numpy().copy())
        else:
            out = torch.empty((r1 - r0, sp.canvas_w, 3), dtype=torch.uint8, device=dev)
            dist.blend_finish(ctx, out.data_ptr(), out.stride(0))
            ctx.sync()
            parts.append(out.cpu().numpy())
    ctx.sync()
    return np.concatenate(parts, axis=0), sp


@pytest.mark.parametrize("name,scale,coarse", [("cfg1", 0.2, False), ("cfg2", 0.04, True), ("cfg3", 0.06, True)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_sharded_equals_single_gpu(ctx, name, scale, coarse, world):
    from simplepanorama_b200 import api
    cfg, K, R, gains, images, plan, cuts = _case(name, scale, coarse)
    full = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx)
    got, sp = _run_sharded(ctx, cfg, gains, images, plan, cuts, world, host=False)
    assert got.shape == full.shape
    assert np.array_equal(got, full)


def test_sharded_host_buffers(ctx):
    """The host-buffer entry points (spano_warp_scatter / spano_blend_add / spano_blend_finish): same canvas."""
    from simplepanorama_b200 import api
    cfg, K, R, gains, images, plan, cuts = _case("cfg2", 0.04, True)
    full = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx)
    got, _ = _run_sharded(ctx, cfg, gains, images, plan, cuts, 4, host=True)
    assert np.array_equal(got, full)
    cfg, K, R, gains, images, plan, cuts = _case("cfg1", 0.2, False)      # tile-sized host masks
    full = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx)
    got, _ = _run_sharded(ctx, cfg, gains, images, plan, cuts, 2, host=True)
    assert np.array_equal(got, full)


def test_sharded_errors(ctx):
    import torch
    from simplepanorama_b200 import api, dist
    from simplepanorama_b200._lib import Slice
    cfg, K, R, gains, images, plan, cuts = _case("cfg1", 0.2, False)
    dev = torch.device("cuda", 0)
    sp = dist.plan_tile_shards([p[2] for p in plan], [p[3] for p in plan], 2, cfg.sigma)
    imgs = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in images]
    cts = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in cuts]
    descs = api.make_descs(imgs, plan, gains, cts, lambda t: t.data_ptr(), lambda t: t.stride(0))
    arena = torch.zeros(sp.arena_bytes[0], dtype=torch.uint8, device=dev)
    lib = ctx.lib
    # add without begin
    s = dist.band_slice(sp, 0, 0, arena.data_ptr())
    assert lib.spano_dev_blend_add(ctx.h, C.byref(descs[0]), C.byref(s)) == -1
    # slice that does not cover the rows the band reads
    dist.blend_begin(ctx, sp, 0, cfg.bands, cfg.sigma)
    bad = Slice(s.row0 + 1, s.row1, s.tile, s.tile_step, s.valid, s.valid_step) if s.row0 + 1 < s.row1 else s
    h = plan[0][3][1]
    if h >= 4 * sp.radius or s.row0 == 0:
        bad = Slice(s.row0, s.row1 - 1, s.tile, s.tile_step, s.valid, s.valid_step)
    assert lib.spano_dev_blend_add(ctx.h, C.byref(descs[0]), C.byref(bad)) == -1
    out = torch.empty((sp.bands[0][1] - sp.bands[0][0], sp.canvas_w, 3), dtype=torch.uint8, device=dev)
    dist.blend_finish(ctx, out.data_ptr(), out.stride(0))
    # scatter: slice rows outside the tile
    oob = (Slice * 1)(Slice(0, h + 1, arena.data_ptr(), sp.tile_step[0], arena.data_ptr(), sp.valid_step[0]))
    assert lib.spano_dev_warp_scatter(ctx.h, cfg.kind, C.c_float(cfg.focal), C.byref(descs[0]), 1, oob) == -1
    ctx.sync()
