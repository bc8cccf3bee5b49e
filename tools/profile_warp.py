"""One tile through the warp kernel (for ncu): python tools/profile_warp.py [cfg] [image] [reps] [scale] [SPANO_OPT_WARP_KERNEL]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from simplepanorama_b200 import api, synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
j = int(sys.argv[2]) if len(sys.argv) > 2 else 11
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
scale = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 0
cfg = synth.config(name, scale)
K, R, gains = synth.cameras(cfg)
img = synth.make_image(cfg, j, gains[j])
ctx = api.Context(0)
ctx.set_option(ctx.OPT_WARP_KERNEL, mode)
dev = torch.device("cuda", 0)
plan = api.plan_tiles([img], [R[j]], [K[j]], cfg.kind, cfg.focal)
K32, R32, (tlx, tly), (w, h) = plan[0]
d_img = torch.from_numpy(img).to(dev)
al16 = lambda v: (v + 15) // 16 * 16
d_tile = torch.empty(al16(3 * w) * h, dtype=torch.uint8, device=dev)
d_mask = torch.empty(al16(w) * h, dtype=torch.uint8, device=dev)
Kc = (C.c_float * 9)(*[float(v) for v in np.asarray(K32).reshape(9)])
Rc = (C.c_float * 9)(*[float(v) for v in np.asarray(R32).reshape(9)])
def run():
    ctx.check(ctx.lib.spano_dev_warp(ctx.h, cfg.kind, C.c_float(cfg.focal), Kc, Rc, d_img.data_ptr(), cfg.width, cfg.height, d_img.stride(0),
                                     C.c_double(gains[j]), tlx, tly, w, h, d_tile.data_ptr(), al16(3 * w), d_mask.data_ptr(), al16(w)))
run(); ctx.sync()
ctx.timers_enable(True); ctx.timers_reset()
for _ in range(reps):
    run()
ctx.sync()
ms, n = ctx.timers_read()
t = ms["warp"] / reps
print(f"{name} x{scale} kernel={mode} image {j}: tile {w}x{h} = {w*h/1e6:.2f} Mpx, warp {t:.4f} ms ({7*w*h/t/1e6:.0f} GB/s algorithmic = {7*w*h/t/1e6/6535.7:.3f} of 6535.7), mask {ms['mask']/reps:.4f} ms")
