"""CPU: the C-ABI library loads, exports every symbol include/spano.h declares, refuses to run
without a GPU (no CPU fallback), and its host-side geometry matches the oracle / golden vectors."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from simplepanorama_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "spano.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spano_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(spano_lib):
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(spano_lib, n), f"libspano.so does not export {n}"
        assert n in L.SYMBOLS, f"ctypes binding misses {n}"
    assert spano_lib.spano_version() == 200


def test_no_cpu_fallback(spano_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert spano_lib.spano_create(C.byref(h), 0) == L.E_NODEVICE
    from simplepanorama_b200 import api
    with pytest.raises(api.SpanoError):
        api.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "simplepanorama_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("SURVEY", ""), f"{f} mentions the oracle"


def test_roi_matches_golden_table(spano_lib, golden):
    from simplepanorama_b200 import api
    t = golden("roi_table.npz")["table"]
    for row in t:
        kind, W, H, f = int(row[0]), int(row[1]), int(row[2]), float(row[3])
        tl, size = api.warp_roi(kind, f, row[4:13], row[13:22], W, H, ctx=None)
        assert (tl[0], tl[1], size[0], size[1]) == tuple(int(v) for v in row[22:26])


def test_roi_matches_oracle_on_config_layouts(spano_lib, oracle):
    from simplepanorama_b200 import api, synth
    for name, scale in (("cfg1", 0.25), ("cfg2", 0.05), ("cfg3", 0.05), ("cfg4", 0.03)):
        cfg = synth.config(name, scale)
        K, R, _ = synth.cameras(cfg)
        for j in range(0, cfg.n, max(1, cfg.n // 8)):
            K32, R32 = api.adjusted_camera(K[j], R[j], cfg.width, cfg.height)
            assert api.warp_roi(cfg.kind, cfg.focal, K32, R32, cfg.width, cfg.height, ctx=None) == \
                oracle.warp_roi(cfg.kind, np.float32(cfg.focal), K32, R32, cfg.width, cfg.height)


def test_roi_rejects_bad_arguments(spano_lib):
    from simplepanorama_b200 import api
    K = np.eye(3, dtype=np.float32); R = np.eye(3, dtype=np.float32)
    with pytest.raises(api.SpanoError):
        api.warp_roi(7, 100.0, K, R, 10, 10)
    with pytest.raises(api.SpanoError):
        api.warp_roi(0, 100.0, K, R, 0, 10)
    with pytest.raises(api.SpanoError):
        api.warp_roi(0, -1.0, K, R, 10, 10)


def test_pan_dimension(spano_lib, oracle):
    from simplepanorama_b200 import api
    corners = [(-20, 5), (25, -3), (60, 30)]
    sizes = [(70, 50), (64, 58), (30, 17)]
    assert api.pan_dimension(corners, sizes) == oracle.pan_dimension(corners, sizes) == (110, 58, -20, -3)
    with pytest.raises(api.SpanoError):
        api.pan_dimension([], [])


def test_band_planner():
    from simplepanorama_b200 import dist
    tiles = [((0, 0), (100, 40)), ((50, 30), (100, 60)), ((0, 80), (120, 20))]
    for world in (1, 2, 3, 4, 8):
        bands = dist.plan_row_bands(tiles, world, min_y=0, canvas_h=100)
        assert len(bands) == world
        assert bands[0][0] == 0 and bands[-1][1] == 100
        for a, b in zip(bands, bands[1:]):
            assert a[1] == b[0] and a[0] <= a[1]
    # work-balanced: a tall, wide tile at the bottom pulls the cut downwards
    eq = dist.plan_row_bands([((0, 0), (10, 50)), ((0, 50), (1000, 50))], 2, 0, 100)
    assert eq[0][1] > 50


def test_disk_reproj_geometry_matches_oracle(spano_lib, oracle):
    """a5 host geometry (normaliser, normalised radius, stretched bounding boxes) == oracle restatement."""
    from simplepanorama_b200 import api
    rng = np.random.default_rng(9)
    for trial in range(6):
        n = int(rng.integers(1, 9))
        sizes = [(int(rng.integers(40, 400)), int(rng.integers(40, 400))) for _ in range(n)]
        corners = [(int(rng.integers(-300, 300)), int(rng.integers(-300, 300))) for _ in range(n)]
        tiles = [np.zeros((h, w, 3), np.uint8) for (w, h) in sizes]
        W, H, mx, my = oracle.pan_dimension(corners, sizes)
        ansatz = (W // 2 + int(rng.integers(-20, 20)), H // 2 + int(rng.integers(-20, 20)))
        radius = float(rng.uniform(5, 60))
        for quad in (True, False):
            new_c, new_s = api.disk_reproj_size(corners, sizes, ansatz, radius, quad, ctx=None)
            outs, _, ocorners = oracle.disk_reproj(tiles, corners, ansatz, radius, quad, erode_iters=0)
            assert new_c == ocorners
            assert new_s == [(o.shape[1], o.shape[0]) for o in outs]


def test_tile_shard_plan_covers_every_band():
    """plan_tile_shards (host arithmetic of the tile-sharded multi-GPU path): bands partition the canvas, every
    tile row a band reads (its rows +- the blur radius, inside the tile) is in that band's slice, arena slots do
    not overlap, owners are balanced and the processing order is a permutation."""
    from simplepanorama_b200 import dist
    rng = np.random.default_rng(5)
    for world in (1, 2, 3, 4, 8):
        n = int(rng.integers(1, 30))
        sizes = [(int(rng.integers(1, 700)), int(rng.integers(1, 500))) for _ in range(n)]
        corners = [(int(rng.integers(-300, 3000)), int(rng.integers(-200, 400))) for _ in range(n)]
        sp = dist.plan_tile_shards(corners, sizes, world, 7.0)
        assert sp.radius == 21
        assert sp.bands[0][0] == 0 and sp.bands[-1][1] == sp.canvas_h
        assert all(sp.bands[k][1] == sp.bands[k + 1][0] for k in range(world - 1))
        # owners: balanced (at most ceil(n / world) tiles each); the owners' processing order is a permutation that serves
        # every band's FIRST tile within the first `world` positions (no band starts late)
        assert max(sp.owner.count(k) for k in range(world)) <= -(-n // world)
        assert all(0 <= o < world for o in sp.owner) and sorted(sp.order) == list(range(n))
        for k in range(world):
            need = [j for j in range(n) if sp.slices[k][j] is not None]
            assert not need or sp.order.index(need[0]) < world
        for k in range(world):
            b0, b1 = sp.bands[k]
            spans = []
            for j in range(n):
                (w, h), cy = sizes[j], corners[j][1] - sp.min_y
                first, last = max(0, b0 - cy), min(h, b1 - cy)
                if last <= first:
                    assert sp.slices[k][j] is None
                    continue
                r0, r1 = sp.slices[k][j]
                assert 0 <= r0 <= max(0, first - 21) and min(h, last + 21) <= r1 <= h
                if h < 84:
                    assert (r0, r1) == (0, h)
                t_off, v_off = sp.offsets[k][j]
                spans.append((t_off, t_off + sp.tile_step[j] * (r1 - r0)))
                spans.append((v_off, v_off + sp.valid_step[j] * (r1 - r0)))
                assert t_off % 256 == 0 and v_off % 256 == 0 and sp.tile_step[j] >= 3 * w and sp.valid_step[j] >= w
            spans.sort()
            assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
            assert not spans or spans[-1][1] <= sp.arena_bytes[k]
        # every tile row is stored somewhere (the bands cover the canvas)
        for j in range(n):
            covered = np.zeros(sizes[j][1], bool)
            for k in range(world):
                if sp.slices[k][j] is not None:
                    covered[sp.slices[k][j][0]:sp.slices[k][j][1]] = True
            assert covered.all()


def test_stitch_blend_dispatch_rejects_unknown_mode(spano_lib):
    """stitch_parameters::blend's switch (src/classes/_panorama.cpp:220-256): an unknown mode is an error here (the
    reference hands back an empty cv::Mat); the argument check needs no device."""
    from simplepanorama_b200 import api
    assert (api.NO_BLEND, api.SIMPLE_BLEND, api.MULTI_BLEND) == (0, 1, 2)
    with pytest.raises(api.SpanoError):
        api.stitch_blend([], [], [], [], blend_mode=7)


def test_ctypes_structs_match_the_header(tmp_path):
    """The ctypes mirrors of spano_image_desc / spano_slice / spano_overlap_info have the C layout of include/spano.h
    (a C compiler is the judge)."""
    import subprocess
    fields = {"spano_image_desc": (L.ImageDesc, ["src_bgr", "src_w", "src_h", "src_step", "K", "R", "gain", "mask_cut", "mask_cut_step",
                                                 "tl_x", "tl_y", "w", "h", "valid_mask", "valid_mask_step", "mask_cut_w", "mask_cut_h",
                                                 "intensity", "intensity_w", "intensity_h", "intensity_step"]),
              "spano_slice": (L.Slice, ["row0", "row1", "tile", "tile_step", "valid", "valid_step", "col0", "col1"]),
              "spano_overlap_info": (L.OverlapInfo, ["i", "j", "area", "I_i", "I_j"])}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "spano.h"', 'int main(void) {']
    for name, (_, fs) in fields.items():
        src.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for f in fs:
            src.append(f'  printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    src += ['  return 0;', '}']
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(c)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, (cls, fs) in fields.items():
        assert int(out[name]) == C.sizeof(cls), name
        for f in fs:
            assert int(out[f"{name}.{f}"]) == getattr(cls, f).offset, f"{name}.{f}"


def test_column_shard_plan_covers_every_band():
    """plan_tile_shards(orient="cols"): the column bands partition the canvas, every tile column a band reads (the 32-column
    strips that intersect its columns, +- the blur radius, inside the tile) is in that band's slice, slices start on
    multiples of 32, arena slots do not overlap; orient="auto" takes column bands when they cut the busiest band's tile count."""
    from simplepanorama_b200 import dist
    # a one-row strip of 24 wide tiles: every row band sees all 24, a column band a handful -> columns; stacked the other way: rows
    strip_c, strip_s = [(1540 * j, 0) for j in range(24)], [(5591, 4004)] * 24
    assert dist.plan_tile_shards(strip_c, strip_s, 8, 7.0, orient="auto").orient == "cols"
    assert dist.plan_tile_shards([(y, x) for (x, y) in strip_c], [(h, w) for (w, h) in strip_s], 8, 7.0, orient="auto").orient == "rows"
    assert dist.plan_tile_shards(strip_c, strip_s, 1, 7.0, orient="auto").orient == "rows"
    rng = np.random.default_rng(11)
    for world in (1, 2, 3, 4, 8):
        n = int(rng.integers(1, 30))
        sizes = [(int(rng.integers(1, 900)), int(rng.integers(1, 500))) for _ in range(n)]
        corners = [(int(rng.integers(-300, 4000)), int(rng.integers(-200, 400))) for _ in range(n)]
        sp = dist.plan_tile_shards(corners, sizes, world, 7.0, orient="cols")
        assert sp.orient == "cols" and sp.radius == 21
        assert sp.bands[0][0] == 0 and sp.bands[-1][1] == sp.canvas_w
        assert all(sp.bands[k][1] == sp.bands[k + 1][0] for k in range(world - 1))
        assert max(sp.owner.count(k) for k in range(world)) <= -(-n // world)
        assert all(0 <= o < world for o in sp.owner) and sorted(sp.order) == list(range(n))
        for k in range(world):
            b0, b1 = sp.bands[k]
            cw, mx, r0, r1 = sp.band_geometry(k)
            assert (cw, mx, r0, r1) == (b1 - b0, sp.min_x + b0, 0, sp.canvas_h) and sp.band_origin(k, 12345) == 3 * b0
            spans = []
            for j in range(n):
                (w, h), cx = sizes[j], corners[j][0] - sp.min_x
                wx0, wx1 = max(0, b0 - cx), min(w, b1 - cx)
                if wx1 <= wx0:
                    assert sp.slices[k][j] is None
                    continue
                assert sp.slices[k][j] == (0, h)
                c0, c1 = sp.cols[k][j]
                ts, vs = sp.steps[k][j]
                assert c0 % 32 == 0 and (c1 % 32 == 0 or c1 == w) and 0 <= c0 < c1 <= w
                if w < 84:
                    assert (c0, c1) == (0, w)
                else:   # what spano_dev_blend_add asks of a slice
                    assert c0 <= max(0, (wx0 & ~31) - 21) and c1 >= min(w, ((wx1 + 31) & ~31) + 21)
                assert ts >= 3 * (c1 - c0) and vs >= c1 - c0 and ts % 16 == 0 and vs % 16 == 0
                t_off, v_off = sp.offsets[k][j]
                assert t_off % 256 == 0 and v_off % 256 == 0
                spans += [(t_off, t_off + ts * h), (v_off, v_off + vs * h)]
            spans.sort()
            assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
            assert not spans or spans[-1][1] <= sp.arena_bytes[k]
        for j in range(n):   # every tile column is stored somewhere
            covered = np.zeros(sizes[j][0], bool)
            for k in range(world):
                if sp.slices[k][j] is not None:
                    covered[sp.cols[k][j][0]:sp.cols[k][j][1]] = True
            assert covered.all()
