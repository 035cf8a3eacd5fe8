#!/usr/bin/env python
"""Raw pinned H2D / D2H bandwidth of this box (context for bench.py's e2e number)."""
import torch, time
n = 512 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4): fn()
    torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (4 * n / (time.perf_counter() - t0) / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
print("bidirectional %.1f GB/s total" % (8 * n / (time.perf_counter() - t0) / 1e9))
import os; print("cpus", os.cpu_count()); os.system("nvidia-smi topo -m 2>/dev/null | head -12; nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv")
