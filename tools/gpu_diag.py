"""Diagnostics run on the GPU box (not a test): prints parity statistics and micro-benchmarks."""
import json, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplepanorama_b200 import api, synth, build as sb
from oracle import oracle as orc
sb.build(); orc.build()
ctx = api.Context(0)
out = {}
out["fp32_peak_tflops"] = {v: ctx.fp32_peak(v) for v in (0, 1, 2)}
print("fp32 peak", out["fp32_peak_tflops"], flush=True)
# maps parity
for kind in (0, 1, 2):
    W, H, f = 1280, 960, 1040.0
    K = np.array([[f * 1.03, 0, W / 2 + 3.5], [0, f * 1.03, H / 2 - 2.25], [0, 0, 1]])
    R = synth.rotation(17.0, -20.0, 1.0)
    K32, R32 = api.adjusted_camera(K, R, W, H)
    tl, size = api.warp_roi(kind, f, K32, R32, W, H, ctx)
    xm, ym = api.build_maps(kind, f, K32, R32, tl, size, ctx)
    xo, yo = orc.build_maps(kind, np.float32(f), K32, R32, tl, size)
    inside = (xo > -1) & (xo < W) & (yo > -1) & (yo < H)
    same = (xm.view(np.uint32) == xo.view(np.uint32)) & (ym.view(np.uint32) == yo.view(np.uint32))
    slip = (np.rint(xm * 32) != np.rint(xo * 32)) | (np.rint(ym * 32) != np.rint(yo * 32))
    print(f"kind {kind}: maps identical {same[inside].mean():.4f}, bin slip {slip[inside].mean():.6f}, max abs err {np.abs(xm-xo)[inside].max():.3e} {np.abs(ym-yo)[inside].max():.3e}", flush=True)
# blend timing on a mid-size tile
rng = np.random.default_rng(0)
for B in (2, 6, 8, 10):
    w, h = 4096, 2048
    tile = rng.integers(16, 240, (h, w, 3), dtype=np.uint8)
    ones = np.full((h, w), 255, np.uint8)
    ctx.timers_enable(True); ctx.timers_reset()
    for _ in range(3):
        api.blend([tile], [ones], [ones], [(0, 0)], B, 7.0, ctx)
    ms, n = ctx.timers_read(); ctx.timers_enable(False)
    t = ms["blend"] / 3
    print(f"B={B}: blend {t:.3f} ms for {w*h/1e6:.1f} MP -> {w*h/1e6/t*1e3:.0f} tile-Mpx/s, {688*B*w*h/t/1e9:.1f} TFLOP/s algorithmic", flush=True)
    out[f"blend_ms_B{B}"] = t
json.dump(out, open("gpurun_out/diag.json", "w"))
