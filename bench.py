#!/usr/bin/env python
"""bench.py -- composited output Mpx/s of the warp + gain + multiband path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   the reference's CPU path (cv2), host cores

One "step" = one full pass of the hot path over the synthetic image set: every source image is
warped (inverse projection + fixed-point bilinear + gain), its validity mask is built, all tiles are
multiband-blended and the canvas is normalised to 8 bit.  At N > 1 image j is warped and masked once, by its owner
rank, whose kernels store the tile rows straight into the memory of the rank(s) whose canvas row band reads them (NVLink
peer stores); every rank blends its band and its normalise kernel stores the finished 8-bit rows into the canvas on rank 0.
Ranks order their work through readiness flags in device memory (stream memory operations): no collective, no host
rendezvous inside the timed region.

`value`  : canvas Mpx / step time with sources, K/R, gains and mask_cut already resident in HBM.
`e2e`    : the same through the host-buffer C-ABI call a user makes: pinned host sources + masks copied H2D and the
           8-bit canvas copied D2H inside the timed region.
Besides the primary workload (BASELINE.json configs[1], `--workload cfg2`) the line carries
  `cfg1`  (N = 1): configs[0] on the GPU next to the reference's CPU path run COMPLETELY (no extrapolation) + the
           canvas difference between the two,
  `cfg4`  (every N): the gigapixel configuration north_star scales on (sources synthesised on the device),
  `parity_at_bench_scale` (N = 1): the CPU baseline's sample re-run on the GPU and compared.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# every stream that may sit in a flag wait (cuStreamWaitValue32, tile-sharded path) needs its own hardware queue, otherwise it
# holds up unrelated streams that share the queue; must be set before the CUDA context exists
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

METRIC = "composited output Mpx/s (warp+gain+multiband)"
UNIT = "Mpx/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--scale", type=float, default=1.0, help="debug: shrink the workload (not a valid bench number)")
    ap.add_argument("--bands", type=int, default=0, help="override the band count of the workload (cfg5 sweeps 1..10 itself)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cfg1", action="store_true")
    ap.add_argument("--no-cfg4", action="store_true")
    ap.add_argument("--no-center-fix", action="store_true", help="cfg3 without sten_proj::disk_reproj")
    ap.add_argument("--band-orient", default="auto", choices=["auto", "rows", "cols"],
                    help="N > 1: canvas row bands or column bands (auto: whichever leaves the bands closer to square)")
    ap.add_argument("--blend-kernel", type=int, default=0, help="SPANO_OPT_BLEND_KERNEL (0 default; 2 = the 8-warp marching kernel of round 1)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU time budget of the cpu_baseline sample")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        self.index = index
        self.enabled = enabled
        self.samples = []
        self.stop = threading.Event()
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        if self.enabled:
            self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.enabled:
            self.th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# workload description shared by both arms
# ----------------------------------------------------------------------------------------------
def config_json(cfg, W, H, T, args, world, note=None, orient="rows"):
    d = {"workload": f"{cfg.name}: {cfg.description}", "images": cfg.n, "image_size": [cfg.width, cfg.height],
         "projection": ["spherical", "cylindrical", "stereographic"][cfg.kind], "focal": cfg.focal, "bands": cfg.bands,
         "sigma": cfg.sigma, "canvas": [W, H], "tile_mpx": round(T / 1e6, 1),
         "sharding": (f"tile-sharded warp+mask (owner by band locality) -> NVLink peer stores -> {'column' if orient == 'cols' else 'row'}-band blend x{world}, ordered by "
                      f"readiness flags (cuStreamWaitValue32); two alternating sets of slice arenas, so the owners' warps of step s+1 overlap the blends of step s" if world > 1 else "single GPU"),
         "mask_cut": "preview scale (1/8), resized to tile size on the device inside the step",
         "l2": "inputs larger than L2 (sources 1.7 GB vs 126 MB)", "scale": args.scale}
    if note:
        d["note"] = note
    return d


def apply_overrides(cfg, args):
    if args.bands:
        cfg.bands = int(args.bands)
    return cfg


# ----------------------------------------------------------------------------------------------
# the reference arm: the reference's own CPU path (oracle/ref_bench.py -> cv2), no CUDA library in the process
# ----------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_bench
    synth = ref_bench.load_synth()
    name = "cfg2" if args.workload == "cfg5" else args.workload
    cfg = apply_overrides(synth.config(name, args.scale), args)
    K, R, gains = synth.cameras(cfg)
    job = ref_bench.job_geometry(cfg, K, R)
    budget = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    r = ref_bench.sample(cfg, synth, K, R, gains, None, budget, args.steps, args.warmup, job=job)
    full = None
    if not args.no_cfg1 and args.scale == 1.0:
        f = ref_bench.full_cfg1(synth)
        full = {"value": f["value"], "unit": UNIT, "ms_per_step": f["ms"], "canvas": [int(f["canvas"].shape[1]), int(f["canvas"].shape[0])],
                "tile_mpx": f["tile_mpx"], "note": f["note"]}
    assert "simplepanorama_b200" not in sys.modules, "the reference arm must not import the CUDA package"
    _, _, W, H, T = job
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_json(cfg, W, H, T, args, 1),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["desc"], "full_cfg1": full},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(torch, index: int):
    """One process per GPU: run (and first-touch the pinned staging buffers) on the CPU cores next to the GPU, so that N
    ranks uploading at once do not all pull from one socket's memory.  sysfs first, `nvidia-smi topo -m` as a fallback
    (virtualised hosts report numa_node = -1).  Best effort; returns what was applied."""
    def parse_list(txt):
        cpus = set()
        for part in txt.strip().split(","):
            a, _, b = part.partition("-")
            if a.strip().isdigit():
                cpus.update(range(int(a), int(b or a) + 1))
        return cpus
    seen = {}
    try:
        seen["nodes_online"] = open("/sys/devices/system/node/online").read().strip()
    except Exception:
        seen["nodes_online"] = None
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        seen["sysfs_numa_node"] = node
        if node >= 0:
            cpus = parse_list(open(f"/sys/devices/system/node/node{node}/cpulist").read()) & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                return {"numa_node": node, "source": "sysfs", "cpus": len(cpus)}
    except Exception:
        pass
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=10).stdout
        rows = [l for l in out.splitlines() if l.startswith("GPU%d" % index)]
        hdr = [l for l in out.splitlines() if "CPU Affinity" in l]
        if rows and hdr:
            cols = [c for c in hdr[0].split("\t") if c.strip()]
            vals = [c for c in rows[0].split("\t") if c.strip()]
            k = cols.index([c for c in cols if "CPU Affinity" in c][0]) + 1    # the row has its own name in front
            cpus = parse_list(vals[k]) & os.sched_getaffinity(0)
            seen["topo_cpu_affinity"] = vals[k].strip()
            if cpus and len(cpus) < len(os.sched_getaffinity(0)):
                os.sched_setaffinity(0, cpus)
                return {"numa_node": None, "source": "nvidia-smi topo", "cpus": len(cpus)}
    except Exception:
        pass
    # nothing to bind to: say what the host exposes (a single node / no PCI locality = nothing to fix by binding)
    return {"numa_node": None, "source": "not bound", "cpus": len(os.sched_getaffinity(0)), **seen}


class _Shape:
    """stands in for a source image this rank does not hold (geometry only)"""
    def __init__(self, h, w):
        self.shape = (h, w, 3)


def make_image_torch(torch, cfg, j, exposure, device):
    """synth.make_image on the device (same formula in float32; rounding may differ from numpy's by 1 LSB on a few
    pixels, so this is used only where nothing is compared with the CPU path: the gigapixel block)."""
    x = torch.arange(cfg.width, dtype=torch.float32, device=device)
    y = torch.arange(cfg.height, dtype=torch.float32, device=device)
    wl = max(cfg.width / 1920.0, 0.05)
    img = torch.empty((cfg.height, cfg.width, 3), dtype=torch.uint8, device=device)
    for c in range(3):
        sx = torch.sin(x / float((37 + 5 * c) * max(wl, 0.25)) + float(j))
        cy = torch.cos(y / float((29 + 3 * c) * max(wl, 0.25)))
        v = 128.0 + 100.0 * torch.outer(cy, sx)
        v = torch.clamp(v, 16, 240) * float(exposure)
        img[..., c] = torch.clamp(torch.round(v), 4, 255).to(torch.uint8)
    return img


class Runner:
    """Our arm on one workload: inputs resident on the device (and pinned on the host for the e2e variant), one
    spano context (+ one for the owner side at N > 1), step functions and timers."""

    def __init__(self, args, name, torch, tdist, dev, rank, world, local, scale=1.0, device_synth=False, want_host=True, bands=0):
        from concurrent.futures import ThreadPoolExecutor
        from simplepanorama_b200 import api, synth, dist as sdist
        self.torch, self.tdist, self.dev, self.rank, self.world, self.local = torch, tdist, dev, rank, world, local
        self.api, self.sdist = api, sdist
        cfg = synth.config(name, scale)
        if bands:
            cfg.bands = int(bands)
        self.cfg = cfg
        K, R, gains = synth.cameras(cfg)
        self.K, self.R, self.gains = K, R, gains
        shapes = [_Shape(cfg.height, cfg.width)] * cfg.n
        self.plan = api.plan_tiles(shapes, R, K, cfg.kind, cfg.focal)                      # host geometry only
        self.corners = [p[2] for p in self.plan]
        self.sizes = [p[3] for p in self.plan]
        self.warp_sizes = list(self.sizes)
        # cfg3 "with center-fix" (stitch_parameters::return_full with conf.fix_center, _panorama.cpp:292-311): the circle comes
        # from sten_proj::estimate_circle, which stays in the reference (here: restated through cv2 on the validity masks of a
        # 1/8-scale copy of the set, outside the timed region, and scaled up); sten_proj::disk_reproj runs inside the step
        self.fix = None
        if name == "cfg3" and world == 1 and not getattr(args, "no_center_fix", False):
            self.fix, self.fix_note = self._estimate_circle(api, synth, name, scale)
            if self.fix is not None:
                self.corners, self.sizes = api.disk_reproj_size(self.corners, self.sizes, (self.fix.ansatz_x, self.fix.ansatz_y),
                                                                self.fix.radius, bool(self.fix.quadratic))
        self.W, self.H, _, self.min_y = api.pan_dimension(self.corners, self.sizes)
        self.T = sum(w * h for w, h in self.sizes)
        self.sp = sdist.plan_tile_shards(self.corners, self.sizes, world, cfg.sigma, orient=getattr(args, "band_orient", "auto") if world > 1 else "rows")
        self.band_w, _, self.row0, self.row1 = self.sp.band_geometry(rank)   # this rank's band: band_w columns x rows [row0, row1)
        self.mine = set(j for j in range(cfg.n) if self.sp.owner[j] == rank) if world > 1 else set(range(cfg.n))
        workers = max(1, min(16, (os.cpu_count() or 1) // max(1, world)))
        with ThreadPoolExecutor(max_workers=workers) as ex:   # numpy releases the GIL
            # mask_cut stays at preview scale (1/8), as stitch_parameters::return_full receives it; the up-scaling to
            # tile size (cv::resize, 8-bit INTER_LINEAR) is part of the path and happens on the device
            self.cuts = list(ex.map(lambda j: synth.seam_masks(self.corners, self.sizes, only=j, coarse=True), range(cfg.n)))
            if device_synth:
                self.images = None
            else:   # a rank synthesises only the sources it owns
                self.images = list(ex.map(lambda j: synth.make_image(cfg, j, gains[j]) if j in self.mine else None, range(cfg.n)))
        self.ctx = api.Context(local)
        # a dedicated (non-default) stream shared by torch and the library, so that torch.cuda.Event brackets exactly the
        # kernels the library launches (the legacy default stream has handle 0 = "use the context's own stream")
        # (high priority, so that the blend's 148 large CTAs are placed before the small CTAs of the next image's warp when both
        # become runnable at once; measured to make no difference -- what decides whether the two share the SMs is whether every
        # kernel of the warp / mask chain fits next to a blend CTA)
        self.stream = torch.cuda.Stream(device=dev, priority=-1)
        assert self.stream.cuda_stream != 0
        self.ctx.set_stream(self.stream.cuda_stream)
        if getattr(args, "blend_kernel", 0):
            self.ctx.set_option(self.ctx.OPT_BLEND_KERNEL, args.blend_kernel)
        self.h_img = None
        if device_synth:
            self.d_img = [make_image_torch(torch, cfg, j, gains[j], dev) if j in self.mine else None for j in range(cfg.n)]
        else:
            self.h_img = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() if a is not None else None for a in self.images]
            self.d_img = [t.to(dev, non_blocking=True) if t is not None else None for t in self.h_img]
            if not want_host:
                self.h_img = None
        self.h_cut = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in self.cuts]
        self.d_cut = [t.to(dev, non_blocking=True) for t in self.h_cut]
        rows = max(1, self.row1 - self.row0)
        self.d_canvas = torch.empty((rows, self.W, 3), dtype=torch.uint8, device=dev) if world == 1 else None
        self.h_canvas = torch.empty((rows, max(1, self.band_w), 3), dtype=torch.uint8).pin_memory() if (want_host and self.h_img is not None) else None
        torch.cuda.synchronize()

        class _Absent:   # placeholder for a source this rank does not own (never dereferenced)
            shape = (1, 1, 3)
        ptr = lambda t: t.data_ptr() if not isinstance(t, _Absent) else 0
        step_of = lambda t: t.stride(0) if not isinstance(t, _Absent) else 0
        fill = lambda lst: [t if t is not None else _Absent() for t in lst]
        self.descs_dev = api.make_descs(fill(self.d_img), self.plan, gains, self.d_cut, ptr, step_of)
        self.descs_host = api.make_descs(fill(self.h_img), self.plan, gains, self.h_cut, ptr, step_of) if self.h_img is not None else None
        for dd in (self.descs_dev, self.descs_host):
            if dd is not None:
                for j in range(cfg.n):
                    dd[j].src_h, dd[j].src_w = cfg.height, cfg.width
        self.ctx_s = self.aux = self.arenas = self.flags = self.peer_canvas = self.session = None
        if world > 1:
            self.ctx_s = api.Context(local)                     # owner-side work runs on its own stream / context
            self.aux = torch.cuda.Stream(device=dev)
            self.ctx_s.set_stream(self.aux.cuda_stream)
            self.arenas = sdist.PeerArenas(self.ctx, self.sp, rank)
            self.arenas2 = sdist.PeerArenas(self.ctx, self.sp, rank)   # even / odd steps use different arenas (done_lag = 2)
            self.flags = sdist.PeerFlags(self.ctx, self.sp, rank)
            self.peer_canvas = sdist.PeerCanvas(self.ctx, self.sp, rank)   # the canvas lives on rank 0
            self.session = sdist.ShardSession(self.sp, rank, cfg.kind, cfg.focal, cfg.bands, cfg.sigma, self.arenas.ptrs, self.flags.ptrs,
                                              self.peer_canvas.origin(self.sp, rank), self.peer_canvas.step, arena_ptrs2=self.arenas2.ptrs)
        self.enqueue_ms = []

    def _estimate_circle(self, api, synth, name, scale):
        from simplepanorama_b200 import _lib
        try:
            from oracle import cv2_ref
            small = synth.config(name, scale / 8.0)
            Ks, Rs, gs = synth.cameras(small)
            c = api.Context(self.local)
            pd = cv2_ref.ProjData()
            for j in range(small.n):
                corner, tile, mask = api.project(small.kind, small.focal, Rs[j], Ks[j], synth.make_image(small, j, gs[j]), 1.0, True, c)
                pd.imgs.append(tile); pd.msks.append(mask); pd.corners.append(tuple(corner))
            c.close()
            ansatz, radius = cv2_ref.estimate_circle(pd)
            if ansatz is None:
                return None, "estimate_circle found no midsection: no centre fix"
            Ws, Hs, _, _ = cv2_ref.get_pan_dimension(pd.corners, pd.imgs)
            Wf, Hf, _, _ = api.pan_dimension(self.corners, self.sizes)
            # canvas pixel coordinates scale with the canvas; the +3 px safety offset of estimate_circle does not
            ax, ay = int(round(ansatz[0] * Wf / Ws)), int(round(ansatz[1] * Hf / Hs))
            r = (radius - 3.0) * (Wf / Ws) + 3.0
            return _lib.CenterFix(ax, ay, float(r), 1), f"circle ({ax}, {ay}) r = {r:.1f} px (estimate_circle at 1/8 scale, scaled up), QUADRATIC_SCALING"
        except Exception as e:
            return None, "no centre fix (%s)" % repr(e)[:160]

    # ---- one step -------------------------------------------------------------------------------------------
    def step_dev(self):
        import ctypes as C
        cfg = self.cfg
        if self.world > 1:
            t0 = time.perf_counter()
            self.session.next_step()
            self.session.step_owner(self.ctx_s, self.descs_dev, host=False)
            self.session.step_band(self.ctx, self.descs_dev, host=False)
            self.enqueue_ms.append((time.perf_counter() - t0) * 1e3)
            return
        if self.fix is not None:
            self.ctx.check(self.ctx.lib.spano_dev_composite_fixed(self.ctx.h, cfg.kind, C.c_float(cfg.focal), cfg.n, self.descs_dev, cfg.bands,
                                                                  cfg.sigma, C.byref(self.fix), self.row0, self.row1, self.d_canvas.data_ptr(),
                                                                  self.d_canvas.stride(0)))
            return
        self.ctx.check(self.ctx.lib.spano_dev_composite(self.ctx.h, cfg.kind, C.c_float(cfg.focal), cfg.n, self.descs_dev, cfg.bands, cfg.sigma,
                                                        self.row0, self.row1, self.d_canvas.data_ptr(), self.d_canvas.stride(0)))

    def step_host(self):
        import ctypes as C
        cfg = self.cfg
        if self.world > 1:
            self.session.next_step()
            self.session.step_owner(self.ctx_s, self.descs_host, host=True)
            self.session.step_band(self.ctx, self.descs_host, host=True, host_canvas=(self.h_canvas.data_ptr(), self.h_canvas.stride(0)))
            return
        if self.fix is not None:
            self.ctx.check(self.ctx.lib.spano_composite_fixed(self.ctx.h, cfg.kind, C.c_float(cfg.focal), cfg.n, self.descs_host, cfg.bands,
                                                              cfg.sigma, C.byref(self.fix), self.row0, self.row1, self.h_canvas.data_ptr(),
                                                              self.h_canvas.stride(0)))
            return
        self.ctx.check(self.ctx.lib.spano_composite(self.ctx.h, cfg.kind, C.c_float(cfg.focal), cfg.n, self.descs_host, cfg.bands, cfg.sigma,
                                                    self.row0, self.row1, self.h_canvas.data_ptr(), self.h_canvas.stride(0)))

    def barrier(self):
        if self.world > 1:
            self.tdist.barrier()
        self.torch.cuda.synchronize()

    def launch_count(self):
        return self.ctx.launch_count + (self.ctx_s.launch_count if self.ctx_s is not None else 0)

    def timed(self, fn, steps, warmup, with_timers=False):
        torch = self.torch
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                fn()
            self.barrier()
            if with_timers:
                self.ctx.timers_enable(True)
                self.ctx.timers_reset()
                self.ctx.blend_stats(reset=True)
            l0 = self.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            for _ in range(steps):
                fn()
            e1.record(self.stream)
            self.barrier()
            ms = e0.elapsed_time(e1) / steps
            stage = None
            if with_timers:
                stage = self.ctx.timers_read() + (self.ctx.blend_stats(reset=True),)
                self.ctx.timers_enable(False)
            launches = self.launch_count() - l0
            if self.world > 1:
                t = torch.tensor([ms], device=self.dev)
                self.tdist.all_reduce(t, op=self.tdist.ReduceOp.MAX)
                ms = float(t.item())
        return ms, launches, stage

    def canvas_device(self):
        """the finished canvas of the last device step as an (H, 3W-ish) tensor on rank 0 (None elsewhere)"""
        import ctypes as C
        torch = self.torch
        if self.rank != 0:
            return None
        if self.world == 1:
            return self.d_canvas
        cv = torch.empty((self.H, self.peer_canvas.step), dtype=torch.uint8, device=self.dev)
        C.CDLL("libcudart.so.12").cudaMemcpy(C.c_void_p(cv.data_ptr()), C.c_void_p(self.peer_canvas.ptr), C.c_size_t(self.peer_canvas.bytes), 3)
        return cv[:, : 3 * self.W]

    def checksum(self):
        """checksum of the finished canvas (rank 0): identical at every N -- the multi-GPU canvas is bit-identical to the
        single-GPU one"""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            self.step_dev()
            self.barrier()
        cv = self.canvas_device()
        if cv is None:
            return None
        flat = cv.reshape(-1).to(torch.int64)
        return int((flat * (torch.arange(flat.numel(), device=self.dev, dtype=torch.int64) % 65521 + 1)).sum().item() % (1 << 61))

    def close(self):
        torch = self.torch
        torch.cuda.synchronize()
        if self.world > 1:
            self.tdist.barrier()
            self.arenas.close()
            self.arenas2.close()
            self.flags.close()
            self.peer_canvas.close()
            self.ctx_s.close()
        self.ctx.close()
        self.d_img = self.d_cut = self.h_img = self.h_cut = self.d_canvas = self.h_canvas = self.images = None
        torch.cuda.empty_cache()


def run_ours(args):
    import torch
    import torch.distributed as tdist
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: simplepanorama_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    if world > 1:
        tdist.init_process_group("nccl", device_id=dev)
    from simplepanorama_b200 import api

    sweep = args.workload == "cfg5"
    name = "cfg2" if sweep else args.workload
    rn = Runner(args, name, torch, tdist, dev, rank, world, local, scale=args.scale, bands=args.bands)
    cfg, ctx, lib = rn.cfg, rn.ctx, rn.ctx.lib
    canvas_mpx = rn.W * rn.H / 1e6

    with ClockSampler(local, enabled=(rank == 0)) as clk:   # one nvidia-smi poller per job, not per rank
        ms, launches, stage = rn.timed(rn.step_dev, args.steps, args.warmup, with_timers=True)
    clocks = clk.summary()
    value = canvas_mpx / (ms * 1e-3)

    # roofline of the dominant kernel (the blend) and of the warp kernel, from the live stage timers
    stage_ms, stage_n, (px_done, px_offered) = stage
    px_done /= args.steps        # tile pixels the blend kernels filtered per step (mask_cut sparsity)
    px_offered /= args.steps
    warp_T = sum(rn.warp_sizes[j][0] * rn.warp_sizes[j][1] for j in sorted(rn.mine))   # tiles this rank warps (all at N = 1)
    fp32_peak = max(ctx.fp32_peak(0), ctx.fp32_peak(1), ctx.fp32_peak(2))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    blend_s = stage_ms["blend"] * 1e-3 / args.steps
    # Inside a step the warp + mask kernels of image i+1 overlap the blend of image i (auxiliary stream), so
    # their in-step event times measure the overlap, not the kernels.  Their own roofline is taken from an
    # isolated pass over the same tiles right here (same buffers, CUDA events on the launching stream).
    iso_tiles = sorted(rn.mine)
    al16 = lambda v: (v + 15) // 16 * 16
    iso_tile = torch.empty(max(al16(3 * w) * h for (w, h) in rn.warp_sizes), dtype=torch.uint8, device=dev)
    iso_mask = torch.empty(max(al16(w) * h for (w, h) in rn.warp_sizes), dtype=torch.uint8, device=dev)
    iso_reps = 3
    with torch.cuda.stream(rn.stream):
        for rep in range(iso_reps + 1):
            if rep == 1:
                torch.cuda.synchronize()
                ctx.timers_enable(True)
                ctx.timers_reset()
            for j in iso_tiles:
                d = rn.descs_dev[j]
                ctx.check(lib.spano_dev_warp(ctx.h, cfg.kind, C.c_float(cfg.focal), d.K, d.R, d.src_bgr, d.src_w, d.src_h, d.src_step,
                                             C.c_double(d.gain), d.tl_x, d.tl_y, d.w, d.h, iso_tile.data_ptr(), al16(3 * d.w),
                                             iso_mask.data_ptr(), al16(d.w)))
        torch.cuda.synchronize()
    iso_ms, _ = ctx.timers_read()
    ctx.timers_enable(False)
    del iso_tile, iso_mask
    warp_s = iso_ms["warp"] * 1e-3 / iso_reps
    mask_iso_ms = iso_ms["mask"] / iso_reps
    blend_flops = 688.0 * cfg.bands * px_done         # 4 ch x B sigmas x 2 passes x 43 MACs per FILTERED tile pixel
    n_blend = max(1, stage_n["blend"] // args.steps)
    n_warp = max(1, len(iso_tiles))   # one warp_kernel (+ one tiny table kernel) per tile
    roofline = {
        "kernel": f"march::blend_ws_kernel<{cfg.bands},{32 if cfg.bands <= 6 else 16},{256 if cfg.bands <= 8 else 384}>", "bound": "fp32",
        "achieved": blend_flops / blend_s / 1e12 if blend_s > 0 else None, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": (blend_flops / blend_s / 1e12 / fp32_peak) if blend_s > 0 else None,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full capture, scaled to this run
        "traffic": 31.3 * px_done / n_blend, "traffic_unit": "bytes per launch (ncu captures under profiles/, scaled by filtered tile pixels)",
        "peak_source": "FFMA microbenchmark run by this bench (MEASURED_PEAKS.json has no fp32 entry)",
        "launches_per_step": n_blend, "avg_launch_ms": blend_s * 1e3 / n_blend,
        "algorithmic_flop_per_tile_px": 688 * cfg.bands,
        "tile_px_filtered_per_step": px_done, "tile_px_in_band_per_step": px_offered,
        "active_fraction": (px_done / px_offered) if px_offered else None,
        "timing_note": "the blend stage timer brackets the plan kernels and the accumulator clear as well as the blend kernels",
        "sparsity_note": "tile pixels whose whole 43x43 mask_cut window is zero have zero weight in every band and are skipped "
                         "(bit-identical canvas); achieved counts only the pixels actually filtered",
        "hbm": {"achieved": (37.0 * px_done) / blend_s / 1e9 if blend_s > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                "bytes_per_tile_px": 37, "note": "5 B u8 inputs + 32 B float4 accumulator read-modify-write; not the binding roof"},
    }
    roofline_warp = {
        "kernel": "warp_kernel", "bound": "hbm", "achieved": 7.0 * warp_T / warp_s / 1e9 if warp_s > 0 else None,
        "peak": hbm_peak, "unit": "GB/s", "frac": (7.0 * warp_T / warp_s / 1e9 / hbm_peak) if warp_s > 0 else None,
        "traffic": 5.32 * warp_T / n_warp, "traffic_unit": "bytes per launch (ncu capture profiles/r1j_warp_raw.csv: 72.3 + 47.1 MB for 22.46 Mpx; part of the tile is still dirty in L2 at kernel end)",
        "peak_source": hbm_src, "launches_per_step": n_warp, "avg_launch_ms": warp_s * 1e3 / n_warp,
        "algorithmic_bytes_per_tile_px": 7,
    }

    # the same step with the mask_cut sparsity switched off (every tile pixel filtered): what the path costs when the
    # seam masks are dense (e.g. all-soft masks); reported next to the headline, never instead of it
    dense = None
    if world == 1:
        ctx.set_option(ctx.OPT_BLEND_DENSE, 1)
        ms_d, _, _ = rn.timed(rn.step_dev, 2, 1)
        cks_dense = rn.checksum()
        ctx.set_option(ctx.OPT_BLEND_DENSE, 0)
        dense = {"value": canvas_mpx / (ms_d * 1e-3), "unit": UNIT, "ms_per_step": ms_d,
                 "note": "blend sparsity disabled (SPANO_OPT_BLEND_DENSE): all tile pixels filtered"}

    checksum = rn.checksum()
    # size-independent properties at the full bench size: the canvas with the sparsity switched off, and the canvas from the
    # round-1 blend kernel (same arithmetic, different schedule), must be bit-identical to the default one
    invariants = None
    if world == 1:
        ctx.set_option(ctx.OPT_BLEND_KERNEL, 2)
        cks_march = rn.checksum()
        ctx.set_option(ctx.OPT_BLEND_KERNEL, getattr(args, "blend_kernel", 0) or 0)
        invariants = {"dense_canvas_identical": cks_dense == checksum, "march_kernel_canvas_identical": cks_march == checksum}

    e2e = None
    if not args.no_e2e:
        ms_e, _, _ = rn.timed(rn.step_host, max(1, min(args.steps, 3)), 1)
        h2d = sum(a.numel() for a in rn.h_img if a is not None) + sum(a.numel() for a in rn.h_cut)
        d2h = (rn.row1 - rn.row0) * rn.band_w * 3
        # what the host link delivers on its own: the same pinned sources copied H2D back to back (all ranks at once)
        link = None
        try:
            pairs = [(d, h) for d, h in zip(rn.d_img, rn.h_img) if d is not None and h is not None]
            rn.barrier()
            with torch.cuda.stream(rn.stream):
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record(rn.stream)
                for d, h in pairs:
                    d.copy_(h, non_blocking=True)
                c1.record(rn.stream)
            rn.barrier()
            link = sum(h.numel() for _, h in pairs) / (c0.elapsed_time(c1) * 1e-3) / 1e9
        except Exception:
            pass
        e2e = {"value": canvas_mpx / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": ms_e, "api": ("spano_composite (host buffers, pinned)" if world == 1 else
                                            "spano_shard_step_owner + spano_shard_step_band (host buffers, pinned)"),
               "h2d_gbs_in_step": h2d / (ms_e * 1e-3) / 1e9, "h2d_link_gbs_measured_this_rank": link,
               "note": "the step is bound by the host link when h2d_gbs_in_step is close to the measured link rate"}

    # ---- cfg5: band-count sweep 1..10 on the same set with irregular (graph-cut) seam masks ------------------
    band_sweep = None
    if sweep and world == 1:
        band_sweep = run_band_sweep(args, rn, torch, fp32_peak)

    # ---- CPU baseline on a bounded sample + the same sample on the GPU (parity at bench scale) ----------------
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import ref_bench
        synth = ref_bench.load_synth()
        r = ref_bench.sample(cfg, synth, rn.K, rn.R, rn.gains, rn.images, args.cpu_seconds,
                             job=(rn.corners, rn.sizes, rn.W, rn.H, rn.T))
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["desc"]}
        s = r["inputs"]
        pctx = api.Context(local)
        got = api.return_full(s["images"], s["R"], s["K"], cfg.kind, cfg.focal, s["gains"], s["cuts"], cfg.bands, cfg.sigma, ctx=pctx)
        pctx.close()
        d = np.abs(got.astype(np.int16) - r["canvas"].astype(np.int16))
        parity = {"max_abs_diff_lsb": int(d.max()), "differing_bytes": int((d > 0).sum()), "bytes": int(d.size), "shape_equal": got.shape == r["canvas"].shape,
                  "sample": f"images {s['idx']}, {s['rows']}-row strips, canvas {got.shape[1]}x{got.shape[0]}: the cpu_baseline sample's own inputs "
                            f"through spano_composite vs the cv2 result", "bar": "<= 1 LSB per channel"}
    prim = {"W": rn.W, "H": rn.H, "T": rn.T, "orient": rn.sp.orient, "note": ("centre fix: " + rn.fix_note) if name == "cfg3" and world == 1 and hasattr(rn, "fix_note") else None}
    enqueue = float(np.median(rn.enqueue_ms)) if rn.enqueue_ms else None
    rn.close()
    del rn

    # ---- cfg1 completely: GPU (device-resident and end to end) next to the reference's CPU path, canvases compared ----
    cfg1 = None
    if world == 1 and not args.no_cfg1 and args.scale == 1.0 and name != "cfg1":
        r1 = Runner(args, "cfg1", torch, tdist, dev, rank, world, local)
        m1, _, _ = r1.timed(r1.step_dev, 20, 5)
        m1e, _, _ = r1.timed(r1.step_host, 10, 3)
        c1 = r1.W * r1.H / 1e6
        cfg1 = {"workload": f"cfg1: {r1.cfg.description}", "canvas": [r1.W, r1.H], "value": c1 / (m1 * 1e-3), "ms_per_step": m1,
                "e2e_value": c1 / (m1e * 1e-3), "e2e_ms_per_step": m1e, "unit": UNIT}
        if not args.no_cpu_baseline:
            from oracle import ref_bench
            f = ref_bench.full_cfg1()
            gpu_canvas = r1.h_canvas.numpy()
            dd = np.abs(gpu_canvas.astype(np.int16) - f["canvas"].astype(np.int16)) if gpu_canvas.shape == f["canvas"].shape else None
            cfg1["cpu_full"] = {"value": f["value"], "ms_per_step": f["ms"], "unit": UNIT, "note": f["note"]}
            cfg1["gpu_vs_cpu_canvas"] = {"max_abs_diff_lsb": int(dd.max()) if dd is not None else None,
                                         "differing_bytes": int((dd > 0).sum()) if dd is not None else None, "shape_equal": dd is not None}
            cfg1["speedup_e2e_vs_cpu_full"] = cfg1["e2e_value"] / f["value"]
            if cpu is not None:
                cpu["full_cfg1"] = cfg1["cpu_full"]
        r1.close()
        del r1

    # ---- cfg4: the gigapixel configuration, at this N (sources synthesised on the device) ---------------------
    cfg4 = None
    if not args.no_cfg4 and args.scale == 1.0 and name != "cfg4":
        try:
            r4 = Runner(args, "cfg4", torch, tdist, dev, rank, world, local, device_synth=True, want_host=False)
            m4, l4, st4 = r4.timed(r4.step_dev, 3, 2, with_timers=True)
            c4 = r4.W * r4.H / 1e6
            cfg4 = {"workload": f"cfg4: {r4.cfg.description}", "canvas": [r4.W, r4.H], "tile_mpx": round(r4.T / 1e6, 1), "n_gpus": world,
                    "bands": ("column" if r4.sp.orient == "cols" else "row") if world > 1 else None, "value": c4 / (m4 * 1e-3), "ms_per_step": m4, "unit": UNIT, "steps": 3, "warmup": 2, "gpu_launches": int(l4),
                    "canvas_checksum": r4.checksum(), "data": "synthetic, generated on the device",
                    "stage_ms_per_step_rank0": {k: v / 3 for k, v in st4[0].items()},
                    "cpu_enqueue_ms_per_step": (float(np.median(r4.enqueue_ms)) if r4.enqueue_ms else None)}
            r4.close()
            del r4
        except Exception as e:   # the block must never take the primary number down with it
            cfg4 = {"error": repr(e)[:300]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config_json(cfg, prim["W"], prim["H"], prim["T"], args, world, prim.get("note"), prim.get("orient", "rows")), "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "numa": numa,
                "canvas_checksum": checksum, "cpu_enqueue_ms_per_step": enqueue,
                "roofline": roofline, "roofline_warp": roofline_warp, "cpu_baseline": cpu, "parity_at_bench_scale": parity,
                "dense_masks": dense, "invariants_at_bench_scale": invariants, "cfg1": cfg1, "cfg4": cfg4, "band_sweep": band_sweep,
                "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
                "stage_note": "in-step warp/mask times overlap the blend (auxiliary stream); isolated: warp %.3f ms, mask %.3f ms per step" % (warp_s * 1e3, mask_iso_ms),
                "tile_mpx_per_s": prim["T"] / 1e6 / (ms * 1e-3)}
        print(json.dumps(line))
    if world > 1:
        torch.cuda.synchronize()
        tdist.barrier()
        tdist.destroy_process_group()


def run_band_sweep(args, rn, torch, fp32_peak):
    """BASELINE.json configs[4]: bands 1..10 on the 24 x 24 MP set with irregular seam masks (graph-cut seams restated
    from the reference, oracle/graph_cut.py, computed on the CPU at preview scale outside the timed region)."""
    out = {"masks": None, "per_band": []}
    cuts = None
    try:
        from oracle import graph_cut
        cuts, how = graph_cut.seam_masks_for(rn.cfg, rn.K, rn.R, rn.gains, rn.corners, rn.sizes)
        out["masks"] = how
    except Exception as e:
        out["masks"] = "Voronoi seams (graph-cut masks unavailable: %s)" % repr(e)[:120]
    if cuts is not None:
        rn.h_cut = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in cuts]
        rn.d_cut = [t.to(rn.dev, non_blocking=True) for t in rn.h_cut]
        torch.cuda.synchronize()
        for j in range(rn.cfg.n):
            d = rn.descs_dev[j]
            d.mask_cut, d.mask_cut_step = rn.d_cut[j].data_ptr(), rn.d_cut[j].stride(0)
            d.mask_cut_h, d.mask_cut_w = int(cuts[j].shape[0]), int(cuts[j].shape[1])
    keep = rn.cfg.bands
    for B in range(1, 11):
        rn.cfg.bands = B
        ms, _, st = rn.timed(rn.step_dev, 5, 2, with_timers=True)
        stage_ms, stage_n, (done, offered) = st
        done /= 5
        offered /= 5
        bs = stage_ms["blend"] * 1e-3 / 5
        out["per_band"].append({"bands": B, "value": rn.W * rn.H / 1e6 / (ms * 1e-3), "ms_per_step": ms, "blend_ms_per_step": bs * 1e3,
                                "roofline_frac": (688.0 * B * done / bs / 1e12 / fp32_peak) if bs > 0 else None,
                                "kernel": f"blend_ws_kernel<{B},{32 if B <= 6 else 16},{256 if B <= 8 else 384}>", "active_fraction": done / offered if offered else None})
    rn.cfg.bands = keep
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
