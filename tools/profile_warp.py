"""One cfg2 image through warp + validity mask (for ncu): python tools/profile_warp.py [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplepanorama_b200 import api, synth
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = synth.config(sys.argv[2] if len(sys.argv) > 2 else "cfg2")
K, R, gains = synth.cameras(cfg)
img = synth.make_image(cfg, 3, gains[3])
ctx = api.Context(0)
ctx.timers_enable(True); ctx.timers_reset()
for _ in range(reps):
    corner, tile, mask = api.project(cfg.kind, cfg.focal, R[3], K[3], img, gains[3], True, ctx)
ms, n = ctx.timers_read()
T = tile.shape[0] * tile.shape[1]
print(f"tile {tile.shape[1]}x{tile.shape[0]}: warp {ms['warp']/reps:.3f} ms ({7*T/(ms['warp']/reps)/1e6:.0f} GB/s algorithmic), mask {ms['mask']/reps:.3f} ms, valid frac {mask.mean()/255:.3f}")
