"""Row-band sharding of the canvas across the GPUs of one box.

One process per GPU.  The canvas is cut into `world` horizontal bands; rank r composites the
band [row0, row1) from every tile that touches it (the blur never reaches across a tile border
-- BORDER_REFLECT is resolved inside the tile -- so a rank needs the whole extent of exactly
those tiles and nothing from its neighbours: no halo exchange, no data-path collective).
The only communication is the gather of the finished 8-bit bands (3 B/px).
Band edges are chosen by cumulative tile-pixel work, not by equal height.
"""
from __future__ import annotations

import numpy as np


def plan_row_bands(tiles, world: int, min_y: int, canvas_h: int):
    """tiles: list of ((tl_x, tl_y), (w, h)).  Returns `world` half-open canvas-row ranges
    [(row0, row1), ...] that partition [0, canvas_h) with (approximately) equal tile-pixel work."""
    work = np.zeros(canvas_h + 1, np.float64)
    for (tlx, tly), (w, h) in tiles:
        y0 = max(0, tly - min_y)
        y1 = min(canvas_h, tly - min_y + h)
        if y1 > y0:
            work[y0] += w
            work[y1] -= w
    per_row = np.cumsum(work[:-1])          # tile pixels blended on each canvas row
    cum = np.concatenate([[0.0], np.cumsum(per_row)])
    total = cum[-1]
    edges = [0]
    for r in range(1, world):
        target = total * r / world
        e = int(np.searchsorted(cum, target, side="left"))
        e = min(max(e, edges[-1]), canvas_h)
        edges.append(e)
    edges.append(canvas_h)
    return [(edges[i], edges[i + 1]) for i in range(world)]


def gather_bands(band, bands, canvas_w: int, rank: int, world: int, device=None):
    """Gather the finished uint8 bands to rank 0 with torch.distributed (NCCL on GPU, gloo on CPU).
    `band` is this rank's (rows, canvas_w, 3) uint8 torch tensor.  Returns the full canvas on rank 0,
    None elsewhere.  Bands are padded to the tallest band so a single gather suffices."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return band
    max_rows = max(b[1] - b[0] for b in bands)
    send = torch.zeros((max_rows, canvas_w, 3), dtype=torch.uint8, device=band.device)
    send[: band.shape[0]] = band
    recv = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
    dist.gather(send, recv, dst=0)
    if rank != 0:
        return None
    return torch.cat([recv[r][: bands[r][1] - bands[r][0]] for r in range(world)], dim=0)


# ---------------------------------------------------------------------------------------------
# Tile-sharded path (include/spano.h, "tile-sharded multi-GPU path"): images shard by owner for the
# warp + validity mask, the canvas shards by row band for the blend, and the owners store every tile
# row straight into the memory of the band(s) that read it (NVLink peer stores from the warp / mask
# kernels).  The plan below is pure host arithmetic, identical on every rank.
# ---------------------------------------------------------------------------------------------
from dataclasses import dataclass, field


def _al(v: int, a: int) -> int:
    return (v + a - 1) // a * a


@dataclass
class ShardPlan:
    world: int
    radius: int
    canvas_w: int
    canvas_h: int
    min_x: int
    min_y: int
    corners: list
    sizes: list
    bands: list                      # [(row0, row1)] per rank
    owner: list                      # owner rank of tile j
    order: list                      # the order in which owners process tiles (a permutation of range(n))
    rounds: list                     # [[tile ids]]: round t holds at most one tile per owner (array order; legacy callers)
    slices: list                     # slices[k][j] = (r0, r1) tile rows of j stored at rank k, or None
    offsets: list                    # offsets[k][j] = (tile_off, valid_off) into rank k's arena, or None
    arena_bytes: list                # per rank
    tile_step: list = field(default_factory=list)
    valid_step: list = field(default_factory=list)
    # orient == "cols": bands[k] = (col0, col1) canvas COLUMNS of rank k; a slice then holds all rows of a column range of
    # the tile: cols[k][j] = (c0, c1) tile columns (multiples of 32 / the tile width), steps[k][j] = (tile_step, valid_step)
    orient: str = "rows"
    cols: list = field(default_factory=list)
    steps: list = field(default_factory=list)

    def band_geometry(self, k: int):
        """(canvas_w, min_x, row0, row1) of rank k's blend session and the byte offset of its first pixel in a canvas of
        pitch `step`: see include/spano.h, spano_slice (a column band is a session whose canvas is the column range)."""
        b0, b1 = self.bands[k]
        if self.orient == "cols":
            return (b1 - b0, self.min_x + b0, 0, self.canvas_h)
        return (self.canvas_w, self.min_x, b0, b1)

    def band_origin(self, k: int, step: int) -> int:
        b0, _ = self.bands[k]
        return 3 * b0 if self.orient == "cols" else b0 * step


def plan_area_bands(world: int, canvas_h: int):
    """`world` row bands of (almost) equal height.  In the tile-sharded path the band side only blends, and the blend
    skips every tile pixel whose seam-mask window is zero: its work follows the canvas AREA of the band (every canvas
    pixel belongs to one image, plus the seam margins), not the number of tiles stacked over it -- balancing by tile
    pixels (plan_row_bands) would make the bands under many overlapping tiles thin and leave the others with most of
    the work (measured on the gigapixel configuration: 14 ms of blend on one rank, 70 ms on another)."""
    edges = [round(canvas_h * r / world) for r in range(world + 1)]
    return [(edges[i], edges[i + 1]) for i in range(world)]


def assign_owners(corners, sizes, bands, min_y: int):
    """Owner rank of every tile.  A tile's rows are stored into the arenas of the bands that read them, so the owner
    with the most rows of the tile inside its own band turns most of the scatter into LOCAL stores (less NVLink
    traffic); every owner takes at most ceil(n / world) tiles so that the warp + mask work stays balanced.  Tiles are
    placed in order of how strongly they prefer one band; what is left falls to the least loaded rank.  On a one-row
    panorama (every tile spans every band equally) this degenerates to the round robin j % world."""
    n, world = len(sizes), len(bands)
    cap = -(-n // world)
    rows = []   # rows[j][k] = rows of tile j inside band k
    for j in range(n):
        cy, h = corners[j][1] - min_y, sizes[j][1]
        rows.append([max(0, min(cy + h, b1) - max(cy, b0)) for (b0, b1) in bands])
    load = [0] * world
    owner = [-1] * n
    pref = sorted(range(n), key=lambda j: (-(max(rows[j]) / max(1, sizes[j][1])), j))
    for j in pref:
        best = max(rows[j])
        if best * world <= sum(rows[j]) + 1e-9 * best:   # no preference: spread evenly
            continue
        for k in sorted(range(world), key=lambda k: (-rows[j][k], load[k], k)):
            if load[k] < cap and rows[j][k] > 0:
                owner[j] = k
                load[k] += 1
                break
    for j in range(n):
        if owner[j] < 0:
            k = min(range(world), key=lambda k: (load[k], (k - j) % world))
            owner[j] = k
            load[k] += 1
    return owner


def owner_order(n: int, slices):
    """Order in which the owners process tiles.  Every band blends its tiles in array order (the accumulation order of
    the single-GPU path) and can only start a tile once its rows have arrived: visiting the bands round robin and
    taking each band's next outstanding tile keeps every band supplied (a set laid out row by row would otherwise
    feed one band after the other)."""
    world = len(slices)
    need = [[j for j in range(n) if slices[k][j] is not None] for k in range(world)]
    pos = [0] * world
    done = [False] * n
    order = []
    while len(order) < n:
        progressed = False
        for k in range(world):
            while pos[k] < len(need[k]) and done[need[k][pos[k]]]:
                pos[k] += 1
            if pos[k] < len(need[k]):
                j = need[k][pos[k]]
                done[j] = True
                order.append(j)
                progressed = True
        if not progressed:
            order.extend(j for j in range(n) if not done[j])   # tiles no band reads
            break
    return order


def band_orientation(rows_plan: "ShardPlan", cols_plan: "ShardPlan") -> str:
    """Row bands or column bands, from the two candidate plans.  What a band costs beyond its share of the pixels is the
    number of tiles it touches: every (band, tile) pair is a slice, a flag, a mask up-scaling, a sparsity plan and a blend
    launch whose pipeline fill and 42-row halo are amortised over fewer rows the thinner the slice is.  Measured at 8
    ranks: the 36999 x 4004 one-row panorama (24 tiles under every row band, at most 8 under a column band) 3.26 ms with
    row bands, 2.58 ms with column bands; the 54360 x 17451 sphere (80 against 63) 23.9 against 25.1 ms.  Column bands are
    taken when they cut the busiest band's tile count to 60 % or less."""
    def busiest(p):
        return max(sum(1 for s in p.slices[k] if s is not None) for k in range(p.world))
    return "cols" if busiest(cols_plan) <= 0.6 * busiest(rows_plan) else "rows"


def plan_tile_shards(corners, sizes, world: int, sigma: float = 7.0, balance: str = "area", owners: str = "locality",
                     orient: str = "rows") -> ShardPlan:
    """corners[j] = (tl_x, tl_y), sizes[j] = (w, h) of every warped tile (spano_warp_roi).
    balance: "area" (equal band heights, see plan_area_bands) or "tile_pixels" (plan_row_bands).
    owners: "locality" (assign_owners) or "round_robin" (j % world).
    orient: "rows" (canvas row bands), "cols" (canvas column bands) or "auto" (both are planned, band_orientation picks)."""
    import math
    n = len(sizes)
    radius = int(math.ceil(3 * sigma))
    xs0 = min(c[0] for c in corners); ys0 = min(c[1] for c in corners)
    xs1 = max(c[0] + s[0] for c, s in zip(corners, sizes)); ys1 = max(c[1] + s[1] for c, s in zip(corners, sizes))
    W, H = xs1 - xs0, ys1 - ys0           # == util::get_pan_dimension
    if orient == "auto":
        r = plan_tile_shards(corners, sizes, world, sigma, balance, owners, "rows")
        if world == 1:
            return r
        c = _plan_column_shards(corners, sizes, world, radius, W, H, xs0, ys0, owners)
        return c if band_orientation(r, c) == "cols" else r
    if orient == "cols":
        return _plan_column_shards(corners, sizes, world, radius, W, H, xs0, ys0, owners)
    bands = plan_area_bands(world, H) if balance == "area" else plan_row_bands(list(zip(corners, sizes)), world, ys0, H)
    owner = assign_owners(corners, sizes, bands, ys0) if owners == "locality" else [j % world for j in range(n)]
    rounds = [list(range(t, min(n, t + world))) for t in range(0, n, world)]
    tile_step = [_al(3 * w, 16) for (w, h) in sizes]
    valid_step = [_al(w, 16) for (w, h) in sizes]
    slices, offsets, arena = [], [], []
    for k in range(world):
        b0, b1 = bands[k]
        sl, of, fill = [], [], 0
        for j in range(n):
            (w, h), cy = sizes[j], corners[j][1] - ys0
            first, last = max(0, b0 - cy), min(h, b1 - cy)
            if last <= first:
                sl.append(None); of.append(None)
                continue
            r0, r1 = (0, h) if h < 4 * radius else (max(0, first - radius), min(h, last + radius))
            sl.append((r0, r1))
            t_off = fill
            fill = _al(fill + tile_step[j] * (r1 - r0), 256)
            v_off = fill
            fill = _al(fill + valid_step[j] * (r1 - r0), 256)
            of.append((t_off, v_off))
        slices.append(sl); offsets.append(of); arena.append(max(fill, 256))
    return ShardPlan(world, radius, W, H, xs0, ys0, list(corners), list(sizes), bands, owner, owner_order(n, slices), rounds, slices,
                     offsets, arena, tile_step, valid_step)


def _plan_column_shards(corners, sizes, world, radius, W, H, xs0, ys0, owners) -> ShardPlan:
    """Column bands of equal width.  Band k blends the canvas columns [b0, b1): of tile j it reads the 32-column strips
    that intersect its columns plus the blur radius (BORDER_REFLECT stays inside the tile), all rows; the stored range is
    widened to multiples of 32 (one word of the mask kernel's scatter)."""
    n = len(sizes)
    edges = [round(W * r / world) for r in range(world + 1)]
    bands = [(edges[i], edges[i + 1]) for i in range(world)]
    flipped_c, flipped_s = [(c[1], c[0]) for c in corners], [(s[1], s[0]) for s in sizes]
    owner = assign_owners(flipped_c, flipped_s, bands, xs0) if owners == "locality" else [j % world for j in range(n)]
    rounds = [list(range(t, min(n, t + world))) for t in range(0, n, world)]
    slices, cols, steps, offsets, arena = [], [], [], [], []
    for k in range(world):
        b0, b1 = bands[k]
        sl, cl, st, of, fill = [], [], [], [], 0
        for j in range(n):
            (w, h), cx = sizes[j], corners[j][0] - xs0
            wx0, wx1 = max(0, b0 - cx), min(w, b1 - cx)
            if wx1 <= wx0 or b1 <= b0:
                sl.append(None); cl.append(None); st.append(None); of.append(None)
                continue
            if w < 4 * radius:
                c0, c1 = 0, w
            else:
                c0 = max(0, (wx0 & ~31) - radius) & ~31
                c1 = min(w, _al(min(w, _al(wx1, 32) + radius), 32))
            ts, vs = _al(3 * (c1 - c0), 16), _al(c1 - c0, 16)
            sl.append((0, h)); cl.append((c0, c1)); st.append((ts, vs))
            t_off = fill
            fill = _al(fill + ts * h, 256)
            v_off = fill
            fill = _al(fill + vs * h, 256)
            of.append((t_off, v_off))
        slices.append(sl); cols.append(cl); steps.append(st); offsets.append(of); arena.append(max(fill, 256))
    return ShardPlan(world, radius, W, H, xs0, ys0, list(corners), list(sizes), bands, owner, owner_order(n, slices), rounds, slices,
                     offsets, arena, [_al(3 * w, 16) for (w, h) in sizes], [_al(w, 16) for (w, h) in sizes], "cols", cols, steps)


def _slice_of(plan: ShardPlan, k: int, j: int, arena_ptr: int):
    from ._lib import Slice
    r0, r1 = plan.slices[k][j]
    t_off, v_off = plan.offsets[k][j]
    if plan.orient == "cols":
        (c0, c1), (ts, vs) = plan.cols[k][j], plan.steps[k][j]
        return Slice(r0, r1, arena_ptr + t_off, ts, arena_ptr + v_off, vs, c0, c1)
    return Slice(r0, r1, arena_ptr + t_off, plan.tile_step[j], arena_ptr + v_off, plan.valid_step[j], 0, 0)


def scatter_slices(plan: ShardPlan, j: int, arena_ptrs):
    """ctypes array of spano_slice: where the rows of tile j go (arena_ptrs[k] = base of rank k's arena as
    seen from THIS process: its own allocation or a peer mapping)."""
    from ._lib import Slice
    out = []
    for k in range(plan.world):
        if plan.slices[k][j] is None:
            continue
        out.append(_slice_of(plan, k, j, arena_ptrs[k]))
    return (Slice * max(1, len(out)))(*out), len(out)


def band_slice(plan: ShardPlan, k: int, j: int, arena_ptr: int):
    """spano_slice of tile j inside rank k's own arena (None when the tile does not touch band k)."""
    from ._lib import Slice
    if plan.slices[k][j] is None:
        return None
    return _slice_of(plan, k, j, arena_ptr)


class PeerArenas:
    """One slice arena per rank, every arena mapped into every process (cudaIpc through the C ABI).
    ptrs[k] = rank k's arena as addressable from this process."""

    def __init__(self, ctx, plan: ShardPlan, rank: int, group=None):
        import ctypes as C
        import torch.distributed as dist
        self.ctx, self.rank, self.world = ctx, rank, plan.world
        own = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        ctx.check(ctx.lib.spano_peer_alloc(ctx.h, plan.arena_bytes[rank], C.byref(own), handle))
        self.own = own.value
        handles = [None] * plan.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.ptrs = []
        for k in range(plan.world):
            if k == rank:
                self.ptrs.append(self.own)
                continue
            p = C.c_void_p()
            hb = (C.c_ubyte * 64).from_buffer_copy(handles[k])
            ctx.check(ctx.lib.spano_peer_open(ctx.h, hb, C.byref(p)))
            self.ptrs.append(p.value)

    def close(self):
        import ctypes as C
        for k, p in enumerate(self.ptrs):
            if k != self.rank and p:
                self.ctx.lib.spano_peer_close(self.ctx.h, C.c_void_p(p))
        if self.own:
            self.ctx.lib.spano_peer_free(self.ctx.h, C.c_void_p(self.own))
        self.ptrs, self.own = [], None


class PeerCanvas:
    """The final 8-bit canvas lives on rank 0 and is mapped into every process: each rank's normalise kernel
    (spano_dev_blend_finish) stores its finished band straight into rank 0's memory over NVLink, so there is no
    gather collective -- only the barrier that ends the step.  ptr = canvas base as addressable from this process."""

    def __init__(self, ctx, plan: ShardPlan, rank: int, group=None):
        import ctypes as C
        import torch.distributed as dist
        self.ctx, self.rank = ctx, rank
        self.step = _al(3 * plan.canvas_w, 16)
        self.bytes = self.step * plan.canvas_h
        self.own = None
        handle = None
        if rank == 0:
            own = C.c_void_p()
            hb = (C.c_ubyte * 64)()
            ctx.check(ctx.lib.spano_peer_alloc(ctx.h, self.bytes, C.byref(own), hb))
            self.own = own.value
            handle = bytes(hb)
        box = [handle]
        dist.broadcast_object_list(box, src=0, group=group)
        if rank == 0:
            self.ptr = self.own
        else:
            p = C.c_void_p()
            hb = (C.c_ubyte * 64).from_buffer_copy(box[0])
            ctx.check(ctx.lib.spano_peer_open(ctx.h, hb, C.byref(p)))
            self.ptr = p.value

    def band_ptr(self, row0: int) -> int:
        return self.ptr + row0 * self.step

    def origin(self, plan: ShardPlan, k: int) -> int:
        """address of the first canvas pixel of rank k's band (row band or column band)"""
        return self.ptr + plan.band_origin(k, self.step)

    def close(self):
        import ctypes as C
        if self.rank == 0 and self.own:
            self.ctx.lib.spano_peer_free(self.ctx.h, C.c_void_p(self.own))
        elif self.ptr:
            self.ctx.lib.spano_peer_close(self.ctx.h, C.c_void_p(self.ptr))
        self.ptr = self.own = None


class PeerFlags:
    """Readiness-flag block of every rank (include/spano.h, spano_shard_plan): n + world 32-bit counters per rank, zeroed
    once, every block mapped into every process.  ptrs[k] = rank k's block as addressable from this process."""

    def __init__(self, ctx, plan: ShardPlan, rank: int, group=None):
        import ctypes as C
        import torch.distributed as dist
        self.ctx, self.rank, self.world = ctx, rank, plan.world
        self.bytes = 4 * (len(plan.sizes) + plan.world)
        own = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        ctx.check(ctx.lib.spano_peer_alloc(ctx.h, max(256, self.bytes), C.byref(own), handle))
        self.own = own.value
        C.CDLL("libcudart.so.12").cudaMemset(C.c_void_p(self.own), 0, C.c_size_t(max(256, self.bytes)))
        handles = [None] * plan.world
        dist.all_gather_object(handles, bytes(handle), group=group)   # (also orders the memset before any peer's first store)
        self.ptrs = []
        for k in range(plan.world):
            if k == rank:
                self.ptrs.append(self.own)
                continue
            p = C.c_void_p()
            hb = (C.c_ubyte * 64).from_buffer_copy(handles[k])
            ctx.check(ctx.lib.spano_peer_open(ctx.h, hb, C.byref(p)))
            self.ptrs.append(p.value)

    def close(self):
        import ctypes as C
        for k, p in enumerate(self.ptrs):
            if k != self.rank and p:
                self.ctx.lib.spano_peer_close(self.ctx.h, C.c_void_p(p))
        if self.own:
            self.ctx.lib.spano_peer_free(self.ctx.h, C.c_void_p(self.own))
        self.ptrs, self.own = [], None


class ShardSession:
    """The spano_shard_plan of one rank plus the step counter: step_owner / step_band enqueue one pass over the image
    set (include/spano.h: spano_shard_step_owner / spano_shard_step_band).  `arena_ptrs[k]` / `flag_ptrs[k]` = rank
    k's slice arena / flag block as addressable from this process; `canvas_ptr` = row 0 of THIS band in the destination
    canvas (device pointer, local or peer), `descs` the spano_image_desc array (device or host pointers).
    `arena_ptrs2`: a second set of arenas; even and odd steps then use different memory, and the owners of a step need
    not wait for the bands of the previous one (done_lag = 2)."""

    def __init__(self, plan: ShardPlan, rank: int, kind: int, focal: float, bands: int, sigma: float, arena_ptrs, flag_ptrs,
                 canvas_ptr: int = 0, canvas_step: int = 0, arena_ptrs2=None):
        import ctypes as C
        from ._lib import ShardPlanC, Slice
        n, world = len(plan.sizes), plan.world
        self.plan, self.rank, self.step = plan, rank, 0
        self._owner = (C.c_int * n)(*plan.owner)
        self._order = (C.c_int * n)(*plan.order)
        self._slices = []
        for ptrs in ([arena_ptrs] if arena_ptrs2 is None else [arena_ptrs, arena_ptrs2]):
            arr = (Slice * (world * n))()
            for k in range(world):
                for j in range(n):
                    s = band_slice(plan, k, j, ptrs[k])
                    if s is not None:
                        arr[k * n + j] = s
            self._slices.append(arr)
        self._flags = (C.c_void_p * world)(*flag_ptrs)
        self.c = ShardPlanC()
        c = self.c
        c.world, c.rank, c.n, c.proj, c.scale, c.bands, c.sigma = world, rank, n, int(kind), float(focal), int(bands), float(sigma)
        # (a column band is a session whose canvas is the band's column range: the blend clips every tile to it)
        c.canvas_w, c.min_x, c.row0, c.row1 = plan.band_geometry(rank)
        c.min_y = plan.min_y
        c.owner, c.order, c.slices = self._owner, self._order, self._slices[0]
        c.flags = C.cast(self._flags, C.POINTER(C.c_void_p))
        c.canvas, c.canvas_step = canvas_ptr or None, canvas_step
        c.done_lag = len(self._slices)
        self._descs = None

    def _bind(self, descs):
        import ctypes as C
        from ._lib import ImageDesc
        self._descs = descs
        self.c.images = C.cast(descs, C.POINTER(ImageDesc))
        self.c.slices = self._slices[self.step % len(self._slices)]

    def next_step(self):
        self.step += 1
        return self.step

    def step_owner(self, ctx, descs, host: bool = False):
        import ctypes as C
        self._bind(descs)
        ctx.check(ctx.lib.spano_shard_step_owner(ctx.h, C.byref(self.c), self.step, 1 if host else 0))

    def step_band(self, ctx, descs, host: bool = False, host_canvas=(0, 0)):
        import ctypes as C
        self._bind(descs)
        ctx.check(ctx.lib.spano_shard_step_band(ctx.h, C.byref(self.c), self.step, 1 if host else 0,
                                                C.c_void_p(host_canvas[0] or None), host_canvas[1]))


def scatter_tile(ctx, plan: ShardPlan, j: int, desc, arena_ptrs, kind: int, focal: float, host: bool = False):
    """Owner side of one image: warp + validity mask, rows stored into the band arenas."""
    import ctypes as C
    sl, n = scatter_slices(plan, j, arena_ptrs)
    fn = ctx.lib.spano_warp_scatter if host else ctx.lib.spano_dev_warp_scatter
    ctx.check(fn(ctx.h, int(kind), C.c_float(focal), C.byref(desc), n, sl))


def blend_prepare(ctx, plan: ShardPlan, k: int, descs, arena_ptr: int, host: bool = False):
    """Band side, once per step after blend_begin: mask_cut up-scaling and sparsity plan of every image that touches
    band k, on the library's auxiliary stream (overlaps the wait for the owners' tile rows)."""
    from ._lib import Slice
    n = len(plan.sizes)
    arr = (Slice * n)()
    for j in range(n):
        s = band_slice(plan, k, j, arena_ptr)
        if s is not None:
            arr[j] = s
    fn = ctx.lib.spano_blend_prepare if host else ctx.lib.spano_dev_blend_prepare
    ctx.check(fn(ctx.h, n, descs, arr))


def blend_add(ctx, plan: ShardPlan, k: int, j: int, descs, arena_ptr: int, host: bool = False):
    """Band side of image j of the spano_image_desc array `descs` (no-op when tile j does not touch band k)."""
    import ctypes as C
    from ._lib import ImageDesc
    s = band_slice(plan, k, j, arena_ptr)
    if s is None:
        return
    fn = ctx.lib.spano_blend_add if host else ctx.lib.spano_dev_blend_add
    # element j by address, so that the library can recognise it as one of the images announced at begin
    pj = C.cast(C.addressof(descs) + j * C.sizeof(ImageDesc), C.POINTER(ImageDesc))
    ctx.check(fn(ctx.h, pj, C.byref(s)))


def blend_begin(ctx, plan: ShardPlan, k: int, bands: int, sigma: float, host_descs=None, host_canvas=(0, 0)):
    """host_descs: the spano_image_desc array with HOST mask_cut pointers (host-buffer variant: the preview-scale
    masks are uploaded right away, ahead of the owners' source uploads).  host_canvas = (pointer, step) announces
    the destination of blend_finish so that finished canvas columns are downloaded while blending continues."""
    import ctypes as C
    cw, mx, r0, r1 = plan.band_geometry(k)
    if host_descs is not None:
        ctx.check(ctx.lib.spano_blend_begin(ctx.h, cw, mx, plan.min_y, r0, r1, int(bands), float(sigma),
                                            len(host_descs), host_descs, C.c_void_p(host_canvas[0] or None), host_canvas[1]))
    else:
        ctx.check(ctx.lib.spano_dev_blend_begin(ctx.h, cw, mx, plan.min_y, r0, r1, int(bands), float(sigma)))


def blend_finish(ctx, canvas_ptr: int, canvas_step: int, host: bool = False):
    import ctypes as C
    fn = ctx.lib.spano_blend_finish if host else ctx.lib.spano_dev_blend_finish
    ctx.check(fn(ctx.h, C.c_void_p(canvas_ptr), canvas_step))


def gather_bands_into(canvas, band, bands, rank: int, world: int):
    """Gather the finished bands into rank 0's preallocated (H, W, 3) uint8 `canvas` (NCCL send/recv straight
    into the canvas rows; rank 0's own band is expected to be a view of `canvas` already).  `band` is this
    rank's (rows, W, 3) tensor."""
    import torch.distributed as dist
    if world == 1:
        return canvas
    ops = []
    if rank == 0:
        for k in range(1, world):
            r0, r1 = bands[k]
            if r1 > r0:
                ops.append(dist.P2POp(dist.irecv, canvas[r0:r1], k))
    else:
        r0, r1 = bands[rank]
        if r1 > r0:
            ops.append(dist.P2POp(dist.isend, band, 0))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return canvas if rank == 0 else None
