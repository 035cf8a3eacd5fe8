#!/usr/bin/env python
"""Aggregate pinned H2D + D2H bandwidth with every visible GPU busy at once (context for bench.py's e2e number at N > 1):
one process per GPU, each moving 338 MB host->device and 232 MB device->host per iteration on two streams, like one GOF."""
import os, sys, time, subprocess
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    import torch
    dev = int(sys.argv[2]); torch.cuda.set_device(dev)
    try:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench; bench.bind_to_gpu_numa_node(dev)
    except Exception:
        pass
    hi = torch.empty(338 << 20, dtype=torch.uint8).pin_memory(); di = torch.empty_like(hi, device="cuda")
    ho = torch.empty(232 << 20, dtype=torch.uint8).pin_memory(); do = torch.empty_like(ho, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def it():
        with torch.cuda.stream(s1): di.copy_(hi, non_blocking=True)
        with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
    for _ in range(3): it()
    torch.cuda.synchronize()
    go = float(sys.argv[3])
    while time.time() < go: pass
    n = 40; t0 = time.perf_counter()
    for _ in range(n): it()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"gpu {dev}: {dt / n * 1e3:.2f} ms per 570 MB iteration, {n * (570 << 20) / dt / 1e9:.1f} GB/s both directions", flush=True)
else:
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    go = time.time() + 25
    ps = [subprocess.Popen([sys.executable, __file__, "worker", str(d), str(go)]) for d in range(n)]
    for p in ps: p.wait()
