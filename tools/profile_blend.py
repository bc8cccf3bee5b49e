"""One cfg2-sized tile through the blend (for ncu): python tools/profile_blend.py [bands] [reps] [active_cols] [kernel]
active_cols = width of the non-zero mask_cut band in the middle of the tile (default: the whole tile);
kernel = SPANO_OPT_BLEND_KERNEL (0 warp-specialised, 2 the 8-warp marching kernel)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplepanorama_b200 import api
B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w, h = 5591, 4004
cols = int(sys.argv[3]) if len(sys.argv) > 3 and int(sys.argv[3]) > 0 else w
kernel = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ctx = api.Context(0)
ctx.set_option(ctx.OPT_BLEND_KERNEL, kernel)
rng = np.random.default_rng(0)
tile = rng.integers(16, 240, (h, w, 3), dtype=np.uint8)
ones = np.full((h, w), 255, np.uint8)
cut = np.zeros((h, w), np.uint8)
cut[:, (w - cols) // 2:(w - cols) // 2 + cols] = 255
api.blend([tile], [cut], [ones], [(0, 0)], B, 7.0, ctx)   # warm-up (allocations)
ctx.timers_enable(True); ctx.timers_reset(); ctx.blend_stats(reset=True)
for _ in range(reps):
    api.blend([tile], [cut], [ones], [(0, 0)], B, 7.0, ctx)
ms, n = ctx.timers_read()
done, offered = ctx.blend_stats()
t = ms["blend"] / reps
px = done / reps
peak = max(ctx.fp32_peak(0), ctx.fp32_peak(1), ctx.fp32_peak(2))
print(f"B={B} kernel={kernel} active_cols={cols}: blend {t:.3f} ms, filtered {px/1e6:.2f} of {w*h/1e6:.1f} MP -> "
      f"{688*B*px/t/1e9:.1f} TFLOP/s algorithmic on the filtered pixels = {688*B*px/t/1e9/peak:.3f} of the measured FFMA peak {peak:.1f}")
