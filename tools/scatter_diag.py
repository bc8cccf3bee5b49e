"""torchrun --nproc-per-node 2 tools/scatter_diag.py : owner-side scatter into the local vs the peer arena."""
import ctypes as C, os, sys, numpy as np, torch, torch.distributed as tdist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplepanorama_b200 import api, dist as sdist, synth
from simplepanorama_b200._lib import Slice
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
tdist.init_process_group("nccl", device_id=dev)
cfg = synth.config("cfg2", 1.0)
K, R, gains = synth.cameras(cfg)
img = synth.make_image(cfg, 3, gains[3])
plan = api.plan_tiles([img], [R[3]], [K[3]], cfg.kind, cfg.focal)
(K32, R32, (tlx, tly), (w, h)) = plan[0]
ctx = api.Context(rank)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
ts, vs = (3 * w + 15) // 16 * 16, (w + 15) // 16 * 16
nbytes = (ts + vs) * h + 4096
own = C.c_void_p(); handle = (C.c_ubyte * 64)()
ctx.check(ctx.lib.spano_peer_alloc(ctx.h, nbytes, C.byref(own), handle))
handles = [None] * world
tdist.all_gather_object(handles, bytes(handle))
peer = C.c_void_p()
ctx.check(ctx.lib.spano_peer_open(ctx.h, (C.c_ubyte * 64).from_buffer_copy(handles[1 - rank]), C.byref(peer)))
d_img = torch.from_numpy(img).to(dev)
cut = torch.zeros((8, 8), dtype=torch.uint8, device=dev)
descs = api.make_descs([d_img], plan, [gains[3]], [cut], lambda t: t.data_ptr(), lambda t: t.stride(0))
def run(base, label):
    sl = (Slice * 1)(Slice(0, h, base, ts, base + ts * h, vs))
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tdist.barrier(); torch.cuda.synchronize()
        e0.record(stream)
        ctx.check(ctx.lib.spano_dev_warp_scatter(ctx.h, cfg.kind, C.c_float(cfg.focal), C.byref(descs[0]), 1, sl))
        e1.record(stream); torch.cuda.synchronize()
    if rank == 0:
        print(f"{label}: {e0.elapsed_time(e1):.3f} ms for {w}x{h} ({4 * w * h / 1e6:.0f} MB stored)")
run(own.value, "scatter -> own arena ")
run(peer.value, "scatter -> peer arena")
# plain copy through the mapped pointer
src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
cudart = torch.cuda.cudart()
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tdist.barrier(); torch.cuda.synchronize()
    e0.record(stream)
    C.CDLL("libcudart.so.12").cudaMemcpyAsync(C.c_void_p(peer.value), C.c_void_p(src.data_ptr()), C.c_size_t(nbytes), 4, C.c_void_p(stream.cuda_stream))
    e1.record(stream); torch.cuda.synchronize()
if rank == 0:
    print(f"cudaMemcpyAsync -> peer arena: {nbytes / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
tdist.barrier(); torch.cuda.synchronize()
tdist.destroy_process_group()
