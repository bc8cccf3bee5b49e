"""P2P diagnostics on a multi-GPU box: topology, copy-engine peer bandwidth, SM-store peer bandwidth."""
import subprocess, time, torch
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
print(subprocess.run(["nvidia-smi", "nvlink", "-s", "-i", "0"], capture_output=True, text=True).stdout[:1200])
n = torch.cuda.device_count()
print("devices", n, "can_access_peer(0,1)", torch.cuda.can_device_access_peer(0, 1) if n > 1 else None)
if n > 1:
    a = torch.empty(1 << 28, dtype=torch.uint8, device="cuda:0")
    b = torch.empty(1 << 28, dtype=torch.uint8, device="cuda:1")
    for name, fn in (("copy engine cuda:0 -> cuda:1", lambda: b.copy_(a, non_blocking=True)),):
        fn(); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        dt = (time.perf_counter() - t0) / 10
        print(f"{name}: {a.numel() / dt / 1e9:.1f} GB/s")
