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


def _band_shape(sp, k):
    """(rows, columns) of rank k's band"""
    cw, _, r0, r1 = sp.band_geometry(k)
    return max(0, r1 - r0), cw


def _join(sp, parts):
    return np.concatenate(parts, axis=1 if sp.orient == "cols" else 0)


def _run_sharded(ctx, cfg, gains, images, plan, cuts, world, host, orient="rows"):
    import torch
    from simplepanorama_b200 import api, dist
    dev = torch.device("cuda", 0)
    sp = dist.plan_tile_shards([p[2] for p in plan], [p[3] for p in plan], world, cfg.sigma, orient=orient)
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
        rows, bw = _band_shape(sp, k)
        if rows <= 0 or bw <= 0:
            continue
        # host variant: announce the images (and the canvas, for the early column download) for even bands only, so
        # both the staged / early-flush and the on-demand paths run
        out = torch.zeros((rows, bw, 3), dtype=torch.uint8).pin_memory() if host else None
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
            out = torch.empty((rows, bw, 3), dtype=torch.uint8, device=dev)
            dist.blend_finish(ctx, out.data_ptr(), out.stride(0))
            ctx.sync()
            parts.append(out.cpu().numpy())
    ctx.sync()
    return _join(sp, parts), sp


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


# ---------------------------------------------------------------------------------------------------------------
# spano_shard_step_owner / spano_shard_step_band: the library-driven step, ordered by readiness flags
# ---------------------------------------------------------------------------------------------------------------
def _run_shard_steps(cfg, gains, images, plan, cuts, world, host, steps=3, poll_kernel=False, two_arenas=False, orient="rows"):
    """`world` ranks emulated on one device: per rank one band context + one owner context (their own streams), arenas and
    flag blocks are plain device buffers.  From the second step on (scratch buffers have their final size: nothing calls
    cudaFree, which would wait for the blocked streams) the band phases are enqueued BEFORE the owner phases, so the band
    streams really sit in their flag waits until the owner streams get there (stream memory operations occupy no SM);
    with the polling-kernel fallback the owners always go first (a kernel must never wait for a later launch).
    Every stream that may block needs its own hardware queue, or it would hold up an unrelated stream queued behind it:
    tests/conftest.py raises CUDA_DEVICE_MAX_CONNECTIONS to 32 (world 4 = 12 streams here)."""
    import torch
    from simplepanorama_b200 import api, dist
    dev = torch.device("cuda", 0)
    sp = dist.plan_tile_shards([p[2] for p in plan], [p[3] for p in plan], world, cfg.sigma, orient=orient)
    n = cfg.n
    arenas = [torch.full((sp.arena_bytes[k],), 0xAB, dtype=torch.uint8, device=dev) for k in range(world)]
    arenas_b = [torch.full((sp.arena_bytes[k],), 0xCD, dtype=torch.uint8, device=dev) for k in range(world)] if two_arenas else None
    flags = [torch.zeros(n + world, dtype=torch.int32, device=dev) for _ in range(world)]
    canvas = torch.zeros((sp.canvas_h, sp.canvas_w, 3), dtype=torch.uint8, device=dev)
    if host:
        imgs = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in images]
        cts = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in cuts]
        hcv = [torch.zeros((max(1, _band_shape(sp, k)[0]), max(1, _band_shape(sp, k)[1]), 3), dtype=torch.uint8).pin_memory() for k in range(world)]
    else:
        imgs = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in images]
        cts = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in cuts]
    torch.cuda.synchronize()
    descs = api.make_descs(imgs, plan, gains, cts, lambda t: t.data_ptr(), lambda t: t.stride(0))
    band_ctx = [api.Context(0) for _ in range(world)]
    own_ctx = [api.Context(0) for _ in range(world)]
    sessions = []
    for k in range(world):
        sessions.append(dist.ShardSession(sp, k, cfg.kind, cfg.focal, cfg.bands, cfg.sigma, [a.data_ptr() for a in arenas],
                                          [f.data_ptr() for f in flags], canvas.data_ptr() + sp.band_origin(k, canvas.stride(0)), canvas.stride(0),
                                          arena_ptrs2=[a.data_ptr() for a in arenas_b] if two_arenas else None))
        if poll_kernel:
            band_ctx[k].set_option(3, 1)   # SPANO_OPT_FLAG_WAIT
            own_ctx[k].set_option(3, 1)
    results = []
    for s in range(steps):
        if s == 1:   # poison the arenas between steps: every row a band reads must be rewritten by its owner in this step
            for b in band_ctx:
                b.sync()
            for a in arenas + (arenas_b or []):
                a.fill_(0x5C)
            canvas.zero_()
            torch.cuda.synchronize()
        for sess in sessions:
            sess.next_step()
        def bands():
            for k in range(world):
                if host:
                    sessions[k].step_band(band_ctx[k], descs, host=True, host_canvas=(hcv[k].data_ptr(), hcv[k].stride(0)))
                else:
                    sessions[k].step_band(band_ctx[k], descs, host=False)
        def owners():
            for k in reversed(range(world)):
                sessions[k].step_owner(own_ctx[k], descs, host=host)
        if poll_kernel or host or s == 0:   # (the host variant of the band phase blocks until its canvas is down: owners first)
            owners(); bands()
        else:
            bands(); owners()
        for c in band_ctx + own_ctx:
            c.sync()
        if host:
            results.append(_join(sp, [hcv[k].numpy()[: _band_shape(sp, k)[0], : _band_shape(sp, k)[1]] for k, (b0, b1) in enumerate(sp.bands) if b1 > b0]).copy())
        else:
            results.append(canvas.cpu().numpy())
    for c in band_ctx + own_ctx:
        c.close()
    return results, sp


@pytest.mark.timeout(120)
@pytest.mark.parametrize("name,scale,coarse", [("cfg1", 0.2, False), ("cfg2", 0.04, True), ("cfg3", 0.06, True), ("cfg4", 0.03, True)])
@pytest.mark.parametrize("world", [1, 2, 4])
def test_shard_step_equals_single_gpu(ctx, name, scale, coarse, world):
    from simplepanorama_b200 import api
    cfg, K, R, gains, images, plan, cuts = _case(name, scale, coarse)
    full = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx)
    res, sp = _run_shard_steps(cfg, gains, images, plan, cuts, world, host=False)
    assert sorted(sp.order) == list(range(cfg.n)) and max(sp.owner.count(k) for k in range(world)) <= -(-cfg.n // world)
    for got in res:
        assert got.shape == full.shape and np.array_equal(got, full)


@pytest.mark.timeout(120)
def test_shard_step_host_and_poll_fallback(ctx):
    from simplepanorama_b200 import api
    cfg, K, R, gains, images, plan, cuts = _case("cfg2", 0.04, True)
    full = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx)
    res, _ = _run_shard_steps(cfg, gains, images, plan, cuts, 3, host=True)
    for got in res:
        assert np.array_equal(got, full)
    res, _ = _run_shard_steps(cfg, gains, images, plan, cuts, 2, host=False, poll_kernel=True)
    for got in res:
        assert np.array_equal(got, full)


@pytest.mark.timeout(120)
def test_shard_step_two_arena_sets(ctx):
    """done_lag = 2: even and odd steps use different arenas, the owners of a step do not wait for the previous step's bands"""
    from simplepanorama_b200 import api
    cfg, K, R, gains, images, plan, cuts = _case("cfg2", 0.04, True)
    full = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx)
    res, _ = _run_shard_steps(cfg, gains, images, plan, cuts, 2, host=False, steps=5, two_arenas=True)
    for got in res:
        assert np.array_equal(got, full)


def test_shard_step_errors(ctx):
    from simplepanorama_b200._lib import ShardPlanC
    p = ShardPlanC()
    assert ctx.lib.spano_shard_step_owner(ctx.h, C.byref(p), 1, 0) == -1
    assert ctx.lib.spano_shard_step_band(ctx.h, C.byref(p), 1, 0, None, 0) == -1
    assert ctx.lib.spano_shard_step_band(ctx.h, None, 1, 0, None, 0) == -1


def _ipc_worker(rank, world, port, q):
    """one process per GPU: the real thing (cudaIpc arenas / flags / canvas, NVLink peer stores)"""
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as tdist
    from simplepanorama_b200 import api, dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, K, R, gains, images, plan, cuts = _case("cfg2", 0.05, True)
    sp = dist.plan_tile_shards([p[2] for p in plan], [p[3] for p in plan], world, cfg.sigma)
    bctx, octx = api.Context(rank), api.Context(rank)
    arenas, flags, pc = dist.PeerArenas(bctx, sp, rank), dist.PeerFlags(bctx, sp, rank), dist.PeerCanvas(bctx, sp, rank)
    imgs = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) if sp.owner[j] == rank else None for j, a in enumerate(images)]
    cts = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in cuts]

    class _A:
        shape = (1, 1, 3)
    descs = api.make_descs([t if t is not None else _A() for t in imgs], plan, gains, cts,
                           lambda t: t.data_ptr() if not isinstance(t, _A) else 0, lambda t: t.stride(0) if not isinstance(t, _A) else 0)
    for j in range(cfg.n):
        descs[j].src_h, descs[j].src_w = cfg.height, cfg.width
    sess = dist.ShardSession(sp, rank, cfg.kind, cfg.focal, cfg.bands, cfg.sigma, arenas.ptrs, flags.ptrs, pc.band_ptr(sp.bands[rank][0]), pc.step)
    for s in range(3):
        sess.next_step()
        if s == 0:   # first step: scratch buffers are still being sized (see _run_shard_steps)
            sess.step_owner(octx, descs)
            sess.step_band(bctx, descs)
        else:
            sess.step_band(bctx, descs)
            sess.step_owner(octx, descs)
    bctx.sync(); octx.sync()
    tdist.barrier()
    if rank == 0:
        cv = torch.empty((sp.canvas_h, pc.step), dtype=torch.uint8, device=dev)
        C.CDLL("libcudart.so.12").cudaMemcpy(C.c_void_p(cv.data_ptr()), C.c_void_p(pc.ptr), C.c_size_t(pc.bytes), 3)
        full = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=bctx)
        q.put(bool(np.array_equal(cv[:, : 3 * sp.canvas_w].reshape(sp.canvas_h, sp.canvas_w, 3).cpu().numpy(), full)))
    tdist.barrier()
    arenas.close(); flags.close(); pc.close()
    tdist.destroy_process_group()


@pytest.mark.timeout(400)
def test_shard_step_two_processes_ipc():
    """Two processes on two GPUs, cudaIpc-mapped arenas, flag blocks and canvas: needs >= 2 devices (gpurun --gpus 2)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_ipc_worker, args=(r, 2, 29541, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) is True


# ---------------------------------------------------------------------------------------------------------------
# column ranges: the accumulator of a blend session may be a COLUMN range of the canvas (spano_dev_blend_begin with
# canvas_w = the range's width and min_x shifted): tiles that stick out of it are clipped, the result is those
# columns of the full canvas bit for bit
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.timeout(120)
@pytest.mark.parametrize("name,scale,coarse", [("cfg2", 0.04, True), ("cfg1", 0.2, False), ("cfg4", 0.03, True)])
@pytest.mark.parametrize("kernel", [0, 2, 1])
def test_column_range_of_the_canvas(ctx, name, scale, coarse, kernel):
    import torch
    from simplepanorama_b200 import api
    from simplepanorama_b200._lib import Slice, ImageDesc
    cfg, K, R, gains, images, plan, cuts = _case(name, scale, coarse)
    full = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx)
    dev = torch.device("cuda", 0)
    corners, sizes = [p[2] for p in plan], [p[3] for p in plan]
    W, H, min_x, min_y = api.pan_dimension(corners, sizes)
    imgs = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in images]
    cts = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in cuts]
    descs = api.make_descs(imgs, plan, gains, cts, lambda t: t.data_ptr(), lambda t: t.stride(0))
    al16 = lambda v: (v + 15) // 16 * 16
    tiles, valids = [], []
    for j in range(cfg.n):   # whole tiles, warped once
        w, h = sizes[j]
        t = torch.zeros((h, al16(3 * w)), dtype=torch.uint8, device=dev)
        m = torch.zeros((h, al16(w)), dtype=torch.uint8, device=dev)
        d = descs[j]
        ctx.check(ctx.lib.spano_dev_warp(ctx.h, cfg.kind, C.c_float(cfg.focal), d.K, d.R, d.src_bgr, d.src_w, d.src_h, d.src_step, C.c_double(d.gain),
                                         d.tl_x, d.tl_y, d.w, d.h, t.data_ptr(), t.stride(0), m.data_ptr(), m.stride(0)))
        tiles.append(t); valids.append(m)
    ctx.set_option(ctx.OPT_BLEND_KERNEL, kernel)
    try:
        rng = np.random.default_rng(7)
        cuts_at = sorted(set([0, W] + [int(v) for v in rng.integers(1, W, 3)] + [W // 2, W // 2 + 1]))
        for c0, c1 in zip(cuts_at, cuts_at[1:]):
            ctx.check(ctx.lib.spano_dev_blend_begin(ctx.h, c1 - c0, min_x + c0, min_y, 0, H, cfg.bands, C.c_double(cfg.sigma)))
            for j in range(cfg.n):
                s = Slice(0, sizes[j][1], tiles[j].data_ptr(), tiles[j].stride(0), valids[j].data_ptr(), valids[j].stride(0))
                pj = C.cast(C.addressof(descs) + j * C.sizeof(ImageDesc), C.POINTER(ImageDesc))
                ctx.check(ctx.lib.spano_dev_blend_add(ctx.h, pj, C.byref(s)))
            out = torch.zeros((H, 3 * (c1 - c0)), dtype=torch.uint8, device=dev)
            ctx.check(ctx.lib.spano_dev_blend_finish(ctx.h, C.c_void_p(out.data_ptr()), out.stride(0)))
            ctx.sync()
            got = out.cpu().numpy().reshape(H, c1 - c0, 3)
            assert np.array_equal(got, full[:, c0:c1]), f"columns [{c0},{c1}) of {W}"
    finally:
        ctx.set_option(ctx.OPT_BLEND_KERNEL, 0)


# ---------------------------------------------------------------------------------------------------------------
# column bands (dist.plan_tile_shards(orient="cols")): slices are column ranges of the tiles, the owners scatter by
# column, every band blends a column range of the canvas -- same bar: the bands join to the single-GPU canvas bit for bit
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.timeout(120)
@pytest.mark.parametrize("name,scale,coarse", [("cfg1", 0.2, False), ("cfg2", 0.04, True), ("cfg3", 0.06, True), ("cfg4", 0.03, True)])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_column_bands_equal_single_gpu(ctx, name, scale, coarse, world):
    from simplepanorama_b200 import api
    cfg, K, R, gains, images, plan, cuts = _case(name, scale, coarse)
    full = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx)
    got, sp = _run_sharded(ctx, cfg, gains, images, plan, cuts, world, host=False, orient="cols")
    assert sp.orient == "cols" and got.shape == full.shape and np.array_equal(got, full)


@pytest.mark.timeout(120)
def test_column_bands_host_buffers(ctx):
    from simplepanorama_b200 import api
    cfg, K, R, gains, images, plan, cuts = _case("cfg2", 0.04, True)
    full = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx)
    got, _ = _run_sharded(ctx, cfg, gains, images, plan, cuts, 3, host=True, orient="cols")
    assert np.array_equal(got, full)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("name,scale,coarse,world", [("cfg2", 0.04, True, 4), ("cfg4", 0.03, True, 3), ("cfg1", 0.2, False, 2)])
def test_shard_step_column_bands(ctx, name, scale, coarse, world):
    from simplepanorama_b200 import api
    cfg, K, R, gains, images, plan, cuts = _case(name, scale, coarse)
    full = api.return_full(images, R, K, cfg.kind, cfg.focal, gains, cuts, cfg.bands, cfg.sigma, ctx=ctx)
    res, sp = _run_shard_steps(cfg, gains, images, plan, cuts, world, host=False, steps=4, two_arenas=True, orient="cols")
    assert sp.orient == "cols"
    for got in res:
        assert got.shape == full.shape and np.array_equal(got, full)
    res, _ = _run_shard_steps(cfg, gains, images, plan, cuts, world, host=True, orient="cols")
    for got in res:
        assert np.array_equal(got, full)
