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
