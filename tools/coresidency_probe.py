"""Do other kernels share the SMs with a running blend kernel?  python tools/coresidency_probe.py [blend kernel mode]
Stream A: dense blends of a full-size tile (~1.7 ms each, persistent CTAs on every SM).  Stream B, started while A runs:
(1) a tiny torch elementwise kernel, (2) a warp of a small image through the library (its own context / stream).
Reports how long B's work took from issue to completion next to the time left on A."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from simplepanorama_b200 import api, synth
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dev = torch.device("cuda", 0)
ctx = api.Context(0); ctx.set_option(ctx.OPT_BLEND_KERNEL, mode)
sa = torch.cuda.Stream(); ctx.set_stream(sa.cuda_stream)
w, h, B = 5591, 4004, 6
rng = np.random.default_rng(0)
al16 = lambda v: (v + 15) // 16 * 16
tile = torch.from_numpy(rng.integers(16, 240, (h, al16(3 * w)), dtype=np.uint8)).to(dev)
ones = torch.full((h, al16(w)), 255, dtype=torch.uint8, device=dev)
out = torch.empty((h, w, 3), dtype=torch.float32, device=dev)
def blend():
    n = 1
    P = lambda t: (C.c_void_p * n)(t.data_ptr()); S = lambda t: (C.c_size_t * n)(t.stride(0)); I = lambda v: (C.c_int * n)(v)
    ctx.check(ctx.lib.spano_dev_multiblend(ctx.h, n, P(tile), S(tile), P(ones), S(ones), P(ones), S(ones), I(0), I(0), I(w), I(h), B, C.c_double(7.0), 0, h, 0,
                                           C.c_void_p(out.data_ptr()), out.stride(0) * 4))
blend(); torch.cuda.synchronize()
# B: a small warp through a second context
cfg = synth.config("cfg1"); K, R, g = synth.cameras(cfg)
img = torch.from_numpy(synth.make_image(cfg, 2, g[2])).to(dev)
ctx2 = api.Context(0); sb = torch.cuda.Stream(); ctx2.set_stream(sb.cuda_stream)
plan = api.plan_tiles([img.cpu().numpy()], [R[2]], [K[2]], cfg.kind, cfg.focal)
K32, R32, (tlx, tly), (tw, th) = plan[0]
Kc = (C.c_float * 9)(*[float(v) for v in np.asarray(K32).reshape(9)]); Rc = (C.c_float * 9)(*[float(v) for v in np.asarray(R32).reshape(9)])
dt = torch.empty(al16(3 * tw) * th, dtype=torch.uint8, device=dev); dm = torch.empty(al16(tw) * th, dtype=torch.uint8, device=dev)
def small_warp():
    ctx2.check(ctx2.lib.spano_dev_warp(ctx2.h, cfg.kind, C.c_float(cfg.focal), Kc, Rc, img.data_ptr(), cfg.width, cfg.height, img.stride(0), C.c_double(1.1),
                                       tlx, tly, tw, th, dt.data_ptr(), al16(3 * tw), dm.data_ptr(), al16(tw)))
small_warp(); torch.cuda.synchronize()
x = torch.zeros(1024, device=dev)
for what in ("torch elementwise kernel", "library warp + mask of a 1.8 Mpx tile"):
    # isolated time of B
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(sb):
        e0.record(sb)
        if what.startswith("torch"): x.add_(1.0)
        else: small_warp()
        e1.record(sb)
    torch.cuda.synchronize(); iso = e0.elapsed_time(e1)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(sa):
        a0.record(sa)
        for _ in range(4): blend()
        a1.record(sa)
    time.sleep(0.002)    # A is now in the middle of its blends
    with torch.cuda.stream(sb):
        b0.record(sb)
        if what.startswith("torch"): x.add_(1.0)
        else: small_warp()
        b1.record(sb)
    torch.cuda.synchronize()
    print(f"mode {mode}: {what}: isolated {iso:.3f} ms; issued {a0.elapsed_time(b0):.2f} ms into A (A = {a0.elapsed_time(a1):.2f} ms), finished {a0.elapsed_time(b1):.2f} ms into A"
          f" -> B took {b0.elapsed_time(b1):.3f} ms while A was running")
