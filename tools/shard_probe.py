"""What does ONE rank of an N-GPU run cost when nothing else interferes?  python tools/shard_probe.py [world] [rank] [cfg] [rows|cols|auto]
All `world` ranks of the tile-sharded path are emulated on one device (plain device buffers as arenas / flag blocks, as in
tests/test_gpu_sharded.py); only rank r's work is timed: its band phase of step s next to its owner phase of step s+1 (what
overlaps in a real run with two arena sets), the band phase alone, and the owner phase alone.  The other ranks' owners run
untimed in between.  Comparing with the per-step time of a real N-GPU run separates the rank's own work from what the
interplay of the ranks (flag latencies, NVLink stores, stragglers) adds."""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from simplepanorama_b200 import api, synth, dist

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
r = int(sys.argv[2]) if len(sys.argv) > 2 else world // 2
name = sys.argv[3] if len(sys.argv) > 3 else "cfg2"
orient = sys.argv[4] if len(sys.argv) > 4 else "rows"
dev = torch.device("cuda", 0)
cfg = synth.config(name, 1.0)
K, R, gains = synth.cameras(cfg)
plan = api.plan_tiles([bench._Shape(cfg.height, cfg.width)] * cfg.n, R, K, cfg.kind, cfg.focal)
corners, sizes = [p[2] for p in plan], [p[3] for p in plan]
sp = dist.plan_tile_shards(corners, sizes, world, cfg.sigma, orient=orient)
n = cfg.n
cuts = [synth.seam_masks(corners, sizes, only=j, coarse=True) for j in range(n)]
imgs = [bench.make_image_torch(torch, cfg, j, gains[j], dev) for j in range(n)]
cts = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in cuts]
descs = api.make_descs(imgs, plan, gains, cts, lambda t: t.data_ptr(), lambda t: t.stride(0))
arenas = [[torch.zeros(sp.arena_bytes[k], dtype=torch.uint8, device=dev) for k in range(world)] for _ in range(2)]
flags = [torch.zeros(n + world, dtype=torch.int32, device=dev) for _ in range(world)]
canvas = torch.zeros((sp.canvas_h, sp.canvas_w, 3), dtype=torch.uint8, device=dev)


def session(k):
    return dist.ShardSession(sp, k, cfg.kind, cfg.focal, cfg.bands, cfg.sigma, [a.data_ptr() for a in arenas[0]], [f.data_ptr() for f in flags],
                             canvas.data_ptr() + sp.band_origin(k, canvas.stride(0)), canvas.stride(0), arena_ptrs2=[a.data_ptr() for a in arenas[1]])


own_ctx = [api.Context(0) for _ in range(world)]
own_sess = [session(k) for k in range(world)]
band_ctx, band_sess = api.Context(0), session(r)
sb, so = torch.cuda.Stream(), torch.cuda.Stream()
band_ctx.set_stream(sb.cuda_stream)
own_ctx[r].set_stream(so.cuda_stream)


def fake_done(step):   # the bands that are not run "have finished" step `step`
    for f in flags:
        f[n:] = step
    torch.cuda.synchronize()


def owners(step, which):
    for k in which:
        own_sess[k].step = step
        own_sess[k].step_owner(own_ctx[k], descs)


others = [k for k in range(world) if k != r]
owners(1, range(world))
for c in own_ctx:
    c.sync()
fake_done(1)
res = {"band || owner(next)": [], "band alone": [], "owner alone": []}
K_STEPS = 8
for s in range(1, K_STEPS + 1):
    owners(s + 1, others)            # untimed: the other ranks' tiles of the next step
    for c in own_ctx:
        c.sync()
    mode = ("band || owner(next)", "band alone", "owner alone")[s % 3] if s > 2 else "band || owner(next)"
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    band_sess.step = s
    torch.cuda.synchronize()
    if mode == "band || owner(next)":
        e0.record(sb); so.wait_event(e0)
        band_sess.step_band(band_ctx, descs)
        owners(s + 1, [r])
        e1.record(sb); e2.record(so)
        torch.cuda.synchronize()
        t = max(e0.elapsed_time(e1), e0.elapsed_time(e2))
    elif mode == "band alone":
        e0.record(sb)
        band_sess.step_band(band_ctx, descs)
        e1.record(sb)
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        owners(s + 1, [r]); own_ctx[r].sync()
    else:
        band_sess.step_band(band_ctx, descs); band_ctx.sync()
        e0.record(so)
        owners(s + 1, [r])
        e1.record(so)
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
    if s > 2:
        res[mode].append(t)
    fake_done(s)
b0, b1 = sp.bands[r]
print(f"{name} world {world} rank {r}: band {'columns' if sp.orient == 'cols' else 'rows'} {b0}..{b1} of {sp.canvas_w if sp.orient == 'cols' else sp.canvas_h}, owns {sp.owner.count(r)} of {n} images, "
      f"{sum(1 for j in range(n) if sp.slices[r][j] is not None)} tiles touch the band")
for k, v in res.items():
    if v:
        print(f"  {k:22s} {np.mean(v):7.3f} ms  (n={len(v)}, min {min(v):.3f})")
