"""One cfg2-sized tile through the blend (for ncu): python tools/profile_blend.py [bands] [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplepanorama_b200 import api
B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = api.Context(0)
rng = np.random.default_rng(0)
w, h = 5591, 4004
tile = rng.integers(16, 240, (h, w, 3), dtype=np.uint8)
ones = np.full((h, w), 255, np.uint8)
ctx.timers_enable(True); ctx.timers_reset()
for _ in range(reps):
    api.blend([tile], [ones], [ones], [(0, 0)], B, 7.0, ctx)
ms, n = ctx.timers_read()
t = ms["blend"] / reps
print(f"B={B}: blend {t:.3f} ms for {w*h/1e6:.1f} MP -> {688*B*w*h/t/1e9:.1f} TFLOP/s algorithmic")
