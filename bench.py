#!/usr/bin/env python
"""bench.py -- composited output Mpx/s of the warp + gain + multiband path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   the reference's CPU path (cv2), host cores

One "step" = one full pass of the hot path over the synthetic image set: every source image is
warped (inverse projection + fixed-point bilinear + gain), its validity mask is built, all tiles are
multiband-blended and the canvas is normalised to 8 bit.  At N > 1 the canvas is cut into row bands
(one per GPU); every rank's normalise kernel stores its finished 8-bit band straight into the canvas on rank 0
(NVLink peer stores) inside the timed region.

`value`  : canvas Mpx / step time with sources, K/R, gains and mask_cut already resident in HBM.
`e2e`    : the same through the host-buffer C-ABI call a user makes (spano_composite): pinned host
           sources + masks copied H2D and the 8-bit canvas copied D2H inside the timed region.
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
# workload
# ----------------------------------------------------------------------------------------------
class _Shape:
    """stands in for a source image this rank does not hold (geometry only)"""
    def __init__(self, h, w):
        self.shape = (h, w, 3)


def build_workload(args, rank=0, world=1):
    from simplepanorama_b200 import api, synth
    cfg = synth.config(args.workload, args.scale)
    K, R, gains = synth.cameras(cfg)
    from concurrent.futures import ThreadPoolExecutor
    workers = max(1, min(16, (os.cpu_count() or 1) // max(1, world)))
    with ThreadPoolExecutor(max_workers=workers) as ex:   # numpy releases the GIL
        # at N > 1 a rank synthesises only the sources it owns (image j belongs to rank j % N)
        images = list(ex.map(lambda j: synth.make_image(cfg, j, gains[j]) if j % world == rank else None, range(cfg.n)))
        plan = api.plan_tiles([im if im is not None else _Shape(cfg.height, cfg.width) for im in images], R, K, cfg.kind, cfg.focal)   # host geometry only
        corners = [p[2] for p in plan]
        sizes = [p[3] for p in plan]
        # mask_cut stays at preview scale (1/8), as stitch_parameters::return_full receives it; the up-scaling to
        # tile size (cv::resize, 8-bit INTER_LINEAR) is part of the path and happens on the device
        cuts = list(ex.map(lambda j: synth.seam_masks(corners, sizes, only=j, coarse=True), range(cfg.n)))
    W, H, mx, my = api.pan_dimension(corners, sizes)
    T = sum(w * h for w, h in sizes)
    return dict(cfg=cfg, K=K, R=R, gains=gains, images=images, plan=plan, corners=corners, sizes=sizes, cuts=cuts,
                W=W, H=H, min_y=my, T=T)


def config_json(wl, args, world):
    cfg = wl["cfg"]
    return {"workload": f"{cfg.name}: {cfg.description}", "images": cfg.n, "image_size": [cfg.width, cfg.height],
            "projection": ["spherical", "cylindrical", "stereographic"][cfg.kind], "focal": cfg.focal, "bands": cfg.bands,
            "sigma": cfg.sigma, "canvas": [wl["W"], wl["H"]], "tile_mpx": round(wl["T"] / 1e6, 1),
            "sharding": (f"tile-sharded warp+mask (owner j % {world}) -> NVLink peer stores -> row-band blend x{world}" if world > 1 else "single GPU"),
            "mask_cut": "preview scale (1/8), resized to tile size on the device inside the step",
            "l2": "inputs larger than L2 (sources 1.7 GB vs 126 MB)", "scale": args.scale}


# ----------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/cv2_ref.py) on a bounded sample
# ----------------------------------------------------------------------------------------------
def cpu_reference_sample(wl, seconds: float, steps: int = 1, warmup: int = 0):
    """Times the reference's CPU implementation (same OpenCV kernels, all host threads) on a bounded
    sample of the workload: image 0 cropped to a horizontal strip whose height is calibrated so that
    one pass costs about `seconds`.  Returns (canvas-equivalent Mpx/s, description, ms per step)."""
    import cv2
    from oracle import cv2_ref
    cfg = wl["cfg"]
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)

    def run(rows):
        y0 = (cfg.height - rows) // 2
        img = np.ascontiguousarray(wl["images"][0][y0:y0 + rows])
        K = wl["K"][0].copy(); K[1, 2] = rows / 2.0
        t0 = time.perf_counter()
        corner, tile = cv2_ref.project(cfg.kind, cfg.focal, wl["R"][0], K, img)
        msk = cv2_ref.validity_mask(tile)
        tile = cv2_ref.apply_gain(tile, wl["gains"][0])
        small = np.full((max(1, tile.shape[0] // 8), max(1, tile.shape[1] // 8)), 255, np.uint8)
        cut = cv2_ref.resize_mask(small, (tile.shape[1], tile.shape[0]))   # return_full's mask_cut up-scaling
        out = cv2_ref.blend_to_u8(cv2_ref.multi_blend([tile], [cut], [msk], [corner], cfg.bands, cfg.sigma))
        return time.perf_counter() - t0, tile.shape[0] * tile.shape[1], out

    pilot_rows = max(64, min(cfg.height, 256))
    t, px, _ = run(pilot_rows)
    rows = int(max(pilot_rows, min(cfg.height, pilot_rows * seconds / max(t, 1e-3))))
    times = []
    for i in range(warmup + steps):
        t, px, _ = run(rows)
        if i >= warmup:
            times.append(t)
    t = float(np.mean(times))
    tile_mpx_s = px / t / 1e6
    canvas_mpx_s = tile_mpx_s * (wl["W"] * wl["H"]) / wl["T"]   # whole job: T tile-px for C canvas-px
    desc = (f"image 0 of {cfg.n}, central {rows}-row strip ({px / 1e6:.2f} tile-Mpx): warp+mask+gain+{cfg.bands}-band "
            f"multi_blend+convert via cv2 {cv2.__version__}, {cores} threads; scaled to the whole job by tile pixels "
            f"(T={wl['T'] / 1e6:.0f} Mpx for C={wl['W'] * wl['H'] / 1e6:.0f} canvas-Mpx)")
    return canvas_mpx_s, desc, t * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = build_workload(args)
    budget = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    v, desc, ms, cores = cpu_reference_sample(wl, budget, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_json(wl, args, 1),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def bind_to_gpu_numa_node(torch, index: int):
    """One process per GPU: run (and first-touch the pinned staging buffers) on the CPU cores of the NUMA node the
    GPU hangs off, so that N ranks uploading at once do not all pull from one socket's memory.  Best effort."""
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as tdist
    from simplepanorama_b200 import api, dist as sdist

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

    wl = build_workload(args, rank, world)
    cfg = wl["cfg"]
    ctx = api.Context(local)
    # a dedicated (non-default) stream shared by torch and the library, so that torch.cuda.Event
    # brackets exactly the kernels the library launches (the legacy default stream has handle 0,
    # which spano_set_stream reads as "use the context's own stream")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)

    # N = 1: the fused single-GPU path (spano_dev_composite / spano_composite).
    # N > 1: the tile-sharded path (include/spano.h): image j is uploaded, warped and masked once, by its owner
    # rank j % N, whose warp / mask kernels store every tile row straight into the memory of the rank(s) whose
    # canvas row band reads it (NVLink peer stores); rank k blends its band from its own arena.  Per round of N
    # images one 4-byte all-reduce orders "all owners have written" before "blend"; the finished 8-bit bands are
    # received straight into rank 0's canvas.
    sp = sdist.plan_tile_shards(wl["corners"], wl["sizes"], world, cfg.sigma)
    bands = sp.bands
    row0, row1 = bands[rank]
    owned = [j for j in range(cfg.n) if sp.owner[j] == rank]
    mine = set(owned) if world > 1 else set(range(cfg.n))

    # host (pinned) and device copies of the inputs; at N > 1 a rank holds only the sources it owns
    h_img = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() if (j in mine and a is not None) else None for j, a in enumerate(wl["images"])]
    h_cut = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in wl["cuts"]]
    d_img = [t.to(dev, non_blocking=True) if t is not None else None for t in h_img]
    d_cut = [t.to(dev, non_blocking=True) for t in h_cut]
    d_canvas = torch.empty((max(1, row1 - row0), wl["W"], 3), dtype=torch.uint8, device=dev) if world == 1 else None
    h_canvas = torch.empty((max(1, row1 - row0), wl["W"], 3), dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()

    class _Absent:   # placeholder for a source this rank does not own (never dereferenced)
        shape = (1, 1, 3)
    ptr = lambda t: t.data_ptr() if not isinstance(t, _Absent) else 0
    step_of = lambda t: t.stride(0) if not isinstance(t, _Absent) else 0
    fill = lambda lst: [t if t is not None else _Absent() for t in lst]
    descs_dev = api.make_descs(fill(d_img), wl["plan"], wl["gains"], d_cut, ptr, step_of)
    descs_host = api.make_descs(fill(h_img), wl["plan"], wl["gains"], h_cut, ptr, step_of)
    for j in range(cfg.n):
        for dd in (descs_dev, descs_host):
            dd[j].src_h, dd[j].src_w = cfg.height, cfg.width
    lib = ctx.lib
    import ctypes as C

    ctx_s = aux = arenas = tok = peer_canvas = None
    if world > 1:
        ctx_s = api.Context(local)                     # owner-side work runs on its own stream / context
        aux = torch.cuda.Stream(device=dev)
        ctx_s.set_stream(aux.cuda_stream)
        arenas = sdist.PeerArenas(ctx, sp, rank)
        peer_canvas = sdist.PeerCanvas(ctx, sp, rank)   # the canvas lives on rank 0; every rank stores its band into it
        tok = torch.zeros(1, device=dev)
        # one barrier per group of rounds: every round for up to 12 rounds, coarser for many small images (each barrier
        # is an all-reduce every rank has to reach on the CPU as well)
        per = max(1, -(-len(sp.rounds) // 12))
        groups = [sp.rounds[i:i + per] for i in range(0, len(sp.rounds), per)]

    trace = {"enqueue_ms": [], "phases": []}

    def step_sharded(host):
        t_cpu0 = time.perf_counter()
        try:
            return _step_sharded(host)
        finally:
            trace["enqueue_ms"].append((time.perf_counter() - t_cpu0) * 1e3)

    def _step_sharded(host):
        descs = descs_host if host else descs_dev
        have = row1 > row0
        pe = [torch.cuda.Event(enable_timing=True) for _ in range(6)]   # phase marks (main stream; [4], [5] on aux)
        pe[0].record(stream)
        if have:   # (host variant: queues the small mask uploads ahead of the large source uploads below)
            sdist.blend_begin(ctx, sp, rank, cfg.bands, cfg.sigma, host_descs=descs_host if host else None,
                              host_canvas=(h_canvas.data_ptr(), h_canvas.stride(0)))
            # everything that does not depend on the owners' data (mask up-scaling, sparsity plans) starts now
            sdist.blend_prepare(ctx, sp, rank, descs, arenas.own, host=host)
        # All barriers live on the OWNER stream: "every rank has finished blending the previous step" (its arenas may be
        # overwritten), then per group of rounds "every owner has written this group".  The band side only waits for
        # the local event behind each of them, so a rank never waits for another rank's blends -- owners run ahead,
        # bands blend as their rows arrive (on many-row panoramas the images of one round all land in the same bands).
        aux.wait_stream(stream)
        evs = []
        with torch.cuda.stream(aux):
            tdist.all_reduce(tok, op=tdist.ReduceOp.MAX)
            pe[1].record(aux)
            pe[4].record(aux)
            for g in groups:
                for rnd in g:
                    for j in rnd:
                        if sp.owner[j] == rank:
                            sdist.scatter_tile(ctx_s, sp, j, descs[j], arenas.ptrs, cfg.kind, cfg.focal, host=host)
                tdist.all_reduce(tok, op=tdist.ReduceOp.MAX)   # the rounds of this group are in every arena
                ev = torch.cuda.Event()
                ev.record(aux)
                evs.append(ev)
            pe[5].record(aux)
        for t, g in enumerate(groups):
            stream.wait_event(evs[t])
            if have:
                for rnd in g:
                    for j in rnd:
                        sdist.blend_add(ctx, sp, rank, j, descs, arenas.own, host=host)
        if have:
            if host:
                sdist.blend_finish(ctx, h_canvas.data_ptr(), h_canvas.stride(0), host=True)
            else:
                # the normalise kernel writes the finished 8-bit band straight into rank 0's canvas (NVLink peer stores)
                sdist.blend_finish(ctx, peer_canvas.band_ptr(row0), peer_canvas.step)
        pe[2].record(stream)
        pe[3].record(stream)
        trace["phases"].append(pe)

    def step_dev():
        if world > 1:
            return step_sharded(False)
        ctx.check(lib.spano_dev_composite(ctx.h, cfg.kind, C.c_float(cfg.focal), cfg.n, descs_dev, cfg.bands, cfg.sigma,
                                          row0, row1, d_canvas.data_ptr(), d_canvas.stride(0)))
        return d_canvas

    def step_host():
        if world > 1:
            return step_sharded(True)
        ctx.check(lib.spano_composite(ctx.h, cfg.kind, C.c_float(cfg.focal), cfg.n, descs_host, cfg.bands, cfg.sigma,
                                      row0, row1, h_canvas.data_ptr(), h_canvas.stride(0)))

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, with_timers=False):
        for _ in range(warmup):
            fn()
        barrier()
        if with_timers:
            ctx.timers_enable(True)
            ctx.timers_reset()
            ctx.blend_stats(reset=True)
        count = lambda: ctx.launch_count + (ctx_s.launch_count if ctx_s is not None else 0)
        l0 = count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / steps
        stage = None
        if with_timers:
            stage = ctx.timers_read() + (ctx.blend_stats(reset=True),)
            ctx.timers_enable(False)
        launches = count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, stage

    canvas_mpx = wl["W"] * wl["H"] / 1e6
    phase_ms = None
    with ClockSampler(local, enabled=(rank == 0)) as clk:   # one nvidia-smi poller per job, not per rank
        ms, launches, stage = timed(step_dev, args.steps, args.warmup, with_timers=True)
    clocks = clk.summary()
    value = canvas_mpx / (ms * 1e-3)
    if world > 1 and trace["phases"]:
        torch.cuda.synchronize()
        pe = trace["phases"][-1]     # last timed step of the device-resident run
        phase_ms = {"step_start_barrier": pe[0].elapsed_time(pe[1]), "rounds_barriers_blend_normalise": pe[1].elapsed_time(pe[2]),
                    "owner_stream_warp_mask_scatter": pe[4].elapsed_time(pe[5])}
        trace["phases"].clear()

    # roofline of the dominant kernel (the blend) and of the warp kernel, from the live stage timers
    stage_ms, stage_n, (px_done, px_offered) = stage
    px_done /= args.steps        # tile pixels the blend kernels filtered per step (mask_cut sparsity)
    px_offered /= args.steps
    my_T = 0   # tile pixels inside this rank's band
    for (tlx, tly), (w, h) in zip(wl["corners"], wl["sizes"]):
        cy = tly - wl["min_y"]
        a, b = max(row0, cy), min(row1, cy + h)
        if b > a:
            my_T += w * (b - a)
    warp_T = sum(wl["sizes"][j][0] * wl["sizes"][j][1] for j in sorted(mine))   # tiles this rank warps (all at N = 1)
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
    iso_tiles = sorted(mine)
    al16 = lambda v: (v + 15) // 16 * 16
    iso_tile = torch.empty(max(al16(3 * w) * h for (w, h) in wl["sizes"]), dtype=torch.uint8, device=dev)
    iso_mask = torch.empty(max(al16(w) * h for (w, h) in wl["sizes"]), dtype=torch.uint8, device=dev)
    iso_reps = 3
    for rep in range(iso_reps + 1):
        if rep == 1:
            torch.cuda.synchronize()
            ctx.timers_enable(True)
            ctx.timers_reset()
        for j in iso_tiles:
            d = descs_dev[j]
            ctx.check(lib.spano_dev_warp(ctx.h, cfg.kind, C.c_float(cfg.focal), d.K, d.R, d.src_bgr, d.src_w, d.src_h, d.src_step,
                                         C.c_double(d.gain), d.tl_x, d.tl_y, d.w, d.h, iso_tile.data_ptr(), al16(3 * d.w),
                                         iso_mask.data_ptr(), al16(d.w)))
    torch.cuda.synchronize()
    iso_ms, _ = ctx.timers_read()
    ctx.timers_enable(False)
    warp_s = iso_ms["warp"] * 1e-3 / iso_reps
    mask_iso_ms = iso_ms["mask"] / iso_reps
    blend_flops = 688.0 * cfg.bands * px_done         # 4 ch x B sigmas x 2 passes x 43 MACs per FILTERED tile pixel
    n_blend = max(1, stage_n["blend"] // args.steps)
    n_warp = max(1, len(iso_tiles))   # one warp_kernel (+ one tiny table kernel) per tile
    roofline = {
        "kernel": f"march::blend_march_kernel<{cfg.bands},{32 if cfg.bands <= 6 else 16}>", "bound": "fp32",
        "achieved": blend_flops / blend_s / 1e12 if blend_s > 0 else None, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": (blend_flops / blend_s / 1e12 / fp32_peak) if blend_s > 0 else None,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full capture
        # (profiles/r1j_blend_march6_sparse_raw.csv: 149.7 + 54.9 MB for 6.54 Mpx filtered = 31.3 B/px), scaled to this run
        "traffic": 31.3 * px_done / n_blend, "traffic_unit": "bytes per launch (ncu capture profiles/r1j_blend_march6_sparse_raw.csv, scaled by filtered tile pixels)",
        "peak_source": "FFMA microbenchmark run by this bench (MEASURED_PEAKS.json has no fp32 entry)",
        "launches_per_step": n_blend, "avg_launch_ms": blend_s * 1e3 / n_blend,
        "algorithmic_flop_per_tile_px": 688 * cfg.bands,
        "tile_px_filtered_per_step": px_done, "tile_px_in_band_per_step": px_offered,
        "active_fraction": (px_done / px_offered) if px_offered else None,
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
        ms_d, _, _ = timed(step_dev, 2, 1)
        ctx.set_option(ctx.OPT_BLEND_DENSE, 0)
        dense = {"value": canvas_mpx / (ms_d * 1e-3), "unit": UNIT, "ms_per_step": ms_d,
                 "note": "blend sparsity disabled (SPANO_OPT_BLEND_DENSE): all tile pixels filtered"}

    # checksum of the finished canvas (rank 0): identical at every N -- the multi-GPU canvas is bit-identical to the
    # single-GPU one
    checksum = None
    if rank == 0:
        step_dev()
        barrier()
        if world == 1:
            cv = d_canvas
        else:
            cv = torch.empty((wl["H"], peer_canvas.step), dtype=torch.uint8, device=dev)
            C.CDLL("libcudart.so.12").cudaMemcpy(C.c_void_p(cv.data_ptr()), C.c_void_p(peer_canvas.ptr), C.c_size_t(peer_canvas.bytes), 3)
            cv = cv[:, : 3 * wl["W"]]
        flat = cv.reshape(-1).to(torch.int64)
        checksum = int((flat * (torch.arange(flat.numel(), device=dev, dtype=torch.int64) % 65521 + 1)).sum().item() % (1 << 61))
    elif world > 1:
        step_dev()
        barrier()

    e2e = None
    if not args.no_e2e:
        ms_e, _, _ = timed(step_host, max(1, min(args.steps, 3)), 1)
        h2d = sum(a.numel() for a in h_img if a is not None) + sum(a.numel() for a in h_cut)
        d2h = (row1 - row0) * wl["W"] * 3
        # what the host link delivers on its own: the same pinned sources copied H2D back to back (all ranks at once)
        link = None
        try:
            pairs = [(d, h) for d, h in zip(d_img, h_img) if d is not None and h is not None]
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            for d, h in pairs:
                d.copy_(h, non_blocking=True)
            c1.record(stream)
            barrier()
            link = sum(h.numel() for _, h in pairs) / (c0.elapsed_time(c1) * 1e-3) / 1e9
        except Exception:
            pass
        e2e = {"value": canvas_mpx / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": ms_e, "api": ("spano_composite (host buffers, pinned)" if world == 1 else
                                            "spano_warp_scatter + spano_blend_begin/prepare/add/finish (host buffers, pinned)"),
               "h2d_gbs_in_step": h2d / (ms_e * 1e-3) / 1e9, "h2d_link_gbs_measured_this_rank": link,
               "note": "the step is bound by the host link when h2d_gbs_in_step is close to the measured link rate"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, desc, cms, cores = cpu_reference_sample(wl, args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config_json(wl, args, world), "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "numa_node": numa,
                "canvas_checksum": checksum, "cpu_enqueue_ms_per_step": (float(np.median(trace["enqueue_ms"])) if trace["enqueue_ms"] else None),
                "phase_ms_rank0": phase_ms, "roofline": roofline, "roofline_warp": roofline_warp, "cpu_baseline": cpu, "dense_masks": dense,
                "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
                "stage_note": "in-step warp/mask times overlap the blend (auxiliary stream); isolated: warp %.3f ms, mask %.3f ms per step" % (warp_s * 1e3, mask_iso_ms),
                "tile_mpx_per_s": wl["T"] / 1e6 / (ms * 1e-3)}
        print(json.dumps(line))
    if world > 1:
        torch.cuda.synchronize()
        tdist.barrier()
        arenas.close()
        peer_canvas.close()
        tdist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
