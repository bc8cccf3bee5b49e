"""CPU, world_size 2 (gloo): the multi-GPU plumbing of the row-band path -- band planning by tile-pixel
work, per-rank band production and the gather of the finished 8-bit bands -- without a GPU.  The per-band
compositor is injected (the oracle here; libspano on the GPU box, see
test_gpu_parity.py::test_row_bands_equal_full_canvas for the CUDA side of the same property)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene():
    rng = np.random.default_rng(3)
    sizes = [(120, 90), (100, 110), (90, 60), (70, 150)]
    corners = [(0, 0), (80, 40), (150, -10), (30, 60)]
    tiles = [rng.integers(16, 240, (h, w, 3), dtype=np.uint8) for (w, h) in sizes]
    cuts = [(rng.random((h, w)) * 255).astype(np.uint8) for (w, h) in sizes]
    valids = [np.full((h, w), 255, np.uint8) for (w, h) in sizes]
    return tiles, cuts, valids, corners, sizes


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from simplepanorama_b200 import dist as sdist
    tiles, cuts, valids, corners, sizes = _scene()
    W, H, mx, my = orc.pan_dimension(corners, sizes)
    bands = sdist.plan_row_bands(list(zip(corners, sizes)), world, my, H)
    row0, row1 = bands[rank]
    # the rank's band: here the oracle's canvas rows (every rank could compute only its rows; the oracle
    # has no band mode, so it computes the canvas and keeps its slice)
    full = orc.blend_to_u8(orc.multi_blend(tiles, cuts, valids, corners, 2, 7.0))
    band = torch.from_numpy(np.ascontiguousarray(full[row0:row1]))
    canvas = sdist.gather_bands(band, bands, W, rank, world)
    if rank == 0:
        assert canvas is not None and tuple(canvas.shape) == (H, W, 3)
        np.save(out_path, canvas.numpy())
        np.save(out_path + ".ref.npy", full)
    else:
        assert canvas is None
    tdist.barrier()
    tdist.destroy_process_group()


@pytest.mark.timeout(300)
def test_row_band_gather_world2(tmp_path):
    world = 2
    out = str(tmp_path / "canvas.npy")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got, ref = np.load(out), np.load(out + ".ref.npy")
    assert np.array_equal(got, ref)


def test_bands_are_work_balanced():
    sys.path.insert(0, ROOT)
    from simplepanorama_b200 import dist as sdist, synth, api
    cfg = synth.config("cfg4", 0.02)
    K, R, _ = synth.cameras(cfg)
    tiles = []
    for j in range(cfg.n):
        K32, R32 = api.adjusted_camera(K[j], R[j], cfg.width, cfg.height)
        tiles.append(api.warp_roi(cfg.kind, cfg.focal, K32, R32, cfg.width, cfg.height, ctx=None))
    W, H, mx, my = api.pan_dimension([t[0] for t in tiles], [t[1] for t in tiles])
    for world in (2, 4, 8):
        bands = sdist.plan_row_bands(tiles, world, my, H)
        work = []
        for (a, b) in bands:
            tot = 0
            for (tl, (w, h)) in tiles:
                y0, y1 = max(a, tl[1] - my), min(b, tl[1] - my + h)
                tot += max(0, y1 - y0) * w
            work.append(tot)
        assert max(work) / (sum(work) / world) < 1.05, (world, work)


# ---------------------------------------------------------------------------------------------
# tile-sharded path: owners warp + mask their images once, every band receives only the tile rows its
# plan says it reads (here over gloo send/recv instead of NVLink peer stores), blends its band and the
# bands are gathered.  The compositor is the oracle; rows a band did NOT receive are poisoned, so a wrong
# slice range in dist.plan_tile_shards shows up as a wrong canvas.
# ---------------------------------------------------------------------------------------------
def _sharded_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from simplepanorama_b200 import dist as sdist, synth
    cfg = synth.config("cfg1", 0.12)
    K, R, gains = synth.cameras(cfg)
    bands_n, sigma = 2, 7.0
    # geometry is host arithmetic every rank can do; pixels only by the owner
    geo = []
    for j in range(cfg.n):
        K32, R32 = orc.adjusted_camera(K[j], R[j], cfg.width, cfg.height)
        geo.append((K32, R32) + tuple(orc.warp_roi(cfg.kind, cfg.focal, K32, R32, cfg.width, cfg.height)))
    corners = [(g[2][0], g[2][1]) for g in geo]
    sizes = [(g[3][0], g[3][1]) for g in geo]
    sp = sdist.plan_tile_shards(corners, sizes, world, sigma)
    cuts = synth.seam_masks(corners, sizes)
    # owner side
    mine = {}
    for j in range(cfg.n):
        if sp.owner[j] != rank:
            continue
        img = synth.make_image(cfg, j, gains[j])
        _, tile = orc.warp(cfg.kind, cfg.focal, geo[j][0], geo[j][1], img)
        mask = orc.surrounding_mask(tile, 3)
        mine[j] = (orc.apply_gain(tile, gains[j]), mask)
    # exchange: rows [r0, r1) of tile j go to every band k whose plan lists them
    rng = np.random.default_rng(100 + rank)
    tiles, valids = [], []
    for j in range(cfg.n):
        w, h = sizes[j]
        for k in range(world):
            sl = sp.slices[k][j]
            if sl is None:
                continue
            r0, r1 = sl
            if sp.owner[j] == rank and k != rank:
                tdist.send(torch.from_numpy(np.ascontiguousarray(mine[j][0][r0:r1])), dst=k)
                tdist.send(torch.from_numpy(np.ascontiguousarray(mine[j][1][r0:r1])), dst=k)
        sl = sp.slices[rank][j]
        t = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)      # poison: rows this band was never sent
        v = rng.integers(0, 2, (h, w), dtype=np.uint8) * 255
        if sl is not None:
            r0, r1 = sl
            if sp.owner[j] == rank:
                t[r0:r1], v[r0:r1] = mine[j][0][r0:r1], mine[j][1][r0:r1]
            else:
                bt = torch.empty((r1 - r0, w, 3), dtype=torch.uint8); bv = torch.empty((r1 - r0, w), dtype=torch.uint8)
                tdist.recv(bt, src=sp.owner[j]); tdist.recv(bv, src=sp.owner[j])
                t[r0:r1], v[r0:r1] = bt.numpy(), bv.numpy()
        tiles.append(t); valids.append(v)
    # band side (the oracle has no band mode: it blends everything and the band keeps its rows)
    row0, row1 = sp.bands[rank]
    full = orc.blend_to_u8(orc.multi_blend(tiles, cuts, valids, corners, bands_n, sigma))
    band = torch.from_numpy(np.ascontiguousarray(full[row0:row1]))
    canvas = sdist.gather_bands(band, sp.bands, sp.canvas_w, rank, world)
    if rank == 0:
        np.save(out_path, canvas.numpy())
    # the single-process answer, from the true tiles
    if rank == 0:
        true_t, true_v = [], []
        for j in range(cfg.n):
            img = synth.make_image(cfg, j, gains[j])
            _, tile = orc.warp(cfg.kind, cfg.focal, geo[j][0], geo[j][1], img)
            true_v.append(orc.surrounding_mask(tile, 3)); true_t.append(orc.apply_gain(tile, gains[j]))
        np.save(out_path + ".ref.npy", orc.blend_to_u8(orc.multi_blend(true_t, cuts, true_v, corners, bands_n, sigma)))
    tdist.barrier()
    tdist.destroy_process_group()


@pytest.mark.timeout(600)
def test_tile_sharded_exchange_world2(tmp_path):
    world = 2
    out = str(tmp_path / "sharded.npy")
    mp.spawn(_sharded_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got, ref = np.load(out), np.load(out + ".ref.npy")
    assert got.shape == ref.shape and np.array_equal(got, ref)


# ---------------------------------------------------------------------------------------------
# the same with COLUMN bands (dist.plan_tile_shards(orient="cols")): a band receives a column range of every tile
# it touches (all rows), everything else is poisoned, and keeps its canvas columns
# ---------------------------------------------------------------------------------------------
def _sharded_cols_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from simplepanorama_b200 import dist as sdist, synth
    cfg = synth.config("cfg1", 0.12)
    K, R, gains = synth.cameras(cfg)
    bands_n, sigma = 2, 7.0
    geo = []
    for j in range(cfg.n):
        K32, R32 = orc.adjusted_camera(K[j], R[j], cfg.width, cfg.height)
        geo.append((K32, R32) + tuple(orc.warp_roi(cfg.kind, cfg.focal, K32, R32, cfg.width, cfg.height)))
    corners = [(g[2][0], g[2][1]) for g in geo]
    sizes = [(g[3][0], g[3][1]) for g in geo]
    sp = sdist.plan_tile_shards(corners, sizes, world, sigma, orient="cols")
    assert sp.orient == "cols"
    cuts = synth.seam_masks(corners, sizes)
    true = {}
    for j in range(cfg.n):
        if sp.owner[j] != rank and rank != 0:
            continue
        img = synth.make_image(cfg, j, gains[j])
        _, tile = orc.warp(cfg.kind, cfg.focal, geo[j][0], geo[j][1], img)
        true[j] = (orc.apply_gain(tile, gains[j]), orc.surrounding_mask(tile, 3))
    rng = np.random.default_rng(200 + rank)
    tiles, valids = [], []
    for j in range(cfg.n):
        w, h = sizes[j]
        for k in range(world):
            if sp.slices[k][j] is None or not (sp.owner[j] == rank and k != rank):
                continue
            c0, c1 = sp.cols[k][j]
            tdist.send(torch.from_numpy(np.ascontiguousarray(true[j][0][:, c0:c1])), dst=k)
            tdist.send(torch.from_numpy(np.ascontiguousarray(true[j][1][:, c0:c1])), dst=k)
        t = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)      # poison: columns this band was never sent
        v = rng.integers(0, 2, (h, w), dtype=np.uint8) * 255
        if sp.slices[rank][j] is not None:
            c0, c1 = sp.cols[rank][j]
            if sp.owner[j] == rank:
                t[:, c0:c1], v[:, c0:c1] = true[j][0][:, c0:c1], true[j][1][:, c0:c1]
            else:
                bt = torch.empty((h, c1 - c0, 3), dtype=torch.uint8); bv = torch.empty((h, c1 - c0), dtype=torch.uint8)
                tdist.recv(bt, src=sp.owner[j]); tdist.recv(bv, src=sp.owner[j])
                t[:, c0:c1], v[:, c0:c1] = bt.numpy(), bv.numpy()
        tiles.append(t); valids.append(v)
    b0, b1 = sp.bands[rank]
    full = orc.blend_to_u8(orc.multi_blend(tiles, cuts, valids, corners, bands_n, sigma))
    band = torch.from_numpy(np.ascontiguousarray(full[:, b0:b1]))
    if rank == 0:
        canvas = np.zeros_like(full)
        canvas[:, b0:b1] = band.numpy()
        for k in range(1, world):
            k0, k1 = sp.bands[k]
            buf = torch.empty((sp.canvas_h, k1 - k0, 3), dtype=torch.uint8)
            tdist.recv(buf, src=k)
            canvas[:, k0:k1] = buf.numpy()
        np.save(out_path, canvas)
        tt = [true[j][0] for j in range(cfg.n)]; tv = [true[j][1] for j in range(cfg.n)]
        np.save(out_path + ".ref.npy", orc.blend_to_u8(orc.multi_blend(tt, cuts, tv, corners, bands_n, sigma)))
    else:
        tdist.send(band, dst=0)
    tdist.barrier()
    tdist.destroy_process_group()


@pytest.mark.timeout(600)
def test_tile_sharded_exchange_column_bands_world2(tmp_path):
    world = 2
    out = str(tmp_path / "sharded_cols.npy")
    mp.spawn(_sharded_cols_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got, ref = np.load(out), np.load(out + ".ref.npy")
    assert got.shape == ref.shape and np.array_equal(got, ref)
