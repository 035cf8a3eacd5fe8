#!/usr/bin/env python
"""Aggregate pinned H2D + D2H bandwidth of this box with 1, 2, 4, ... GPUs busy at once (the host-side ceiling of bench.py's
e2e number): one process per GPU, each moving 338 MB host->device and 232 MB device->host per iteration on two streams, like
one BASELINE-config-2 GOF.  All processes are started once; round k runs with the first N_k GPUs, the others idle.

    python tools/pcie_probe_multi.py [max_gpus]        # prints one line per (round, gpu) and a summary per round
"""
import os, sys, time, subprocess
ROUND_S = 7.0
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    import torch
    dev, n_max, t_first = int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
    torch.cuda.set_device(dev)
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
    rounds = [n for n in (1, 2, 4, 8, 16) if n <= n_max]
    for k, n in enumerate(rounds):
        go = t_first + k * ROUND_S
        if dev >= n:
            continue
        while time.time() < go: pass
        m = 25; t0 = time.perf_counter()
        for _ in range(m): it()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"round {n} gpus | gpu {dev}: {dt / m * 1e3:.2f} ms per 570 MB iteration, {m * (570 << 20) / dt / 1e9:.1f} GB/s both directions", flush=True)
else:
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    t_first = time.time() + 30
    ps = [subprocess.Popen([sys.executable, __file__, "worker", str(d), str(n), str(t_first)], stdout=subprocess.PIPE, text=True) for d in range(n)]
    lines = []
    for p in ps:
        out, _ = p.communicate()
        lines += [l for l in out.splitlines() if l.startswith("round")]
    for l in sorted(lines, key=lambda l: (int(l.split()[1]), int(l.split()[5].rstrip(":")))):
        print(l)
    for r in sorted({int(l.split()[1]) for l in lines}):
        tot = sum(float(l.split("iteration, ")[1].split()[0]) for l in lines if int(l.split()[1]) == r)
        print(f"== {r} GPU(s) busy: {tot:.1f} GB/s aggregate host<->device (H2D + D2H)")
