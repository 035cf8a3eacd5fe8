#!/usr/bin/env python
"""e2e (pinned host planes in, pinned host frames out) for different numbers of GOFs in flight."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tmc2rs_b200  # noqa
from tmc2rs_b200 import abi, codec, synth
import torch

cfg = synth.config("c2", frames=8)
gof = synth.replicate_gof(synth.make_gof(cfg), 32)
pinned = codec.pinned_copy_of(gof)
view = abi.GofView(pinned)
for depth in (2, 3, 4):
    ctx = codec.Context(devices=(0,), gofs_in_flight=depth)
    def drain(n):
        got = 0
        for _ in range(n):
            got += ctx.next_frame_raw()[0]
        return got
    for _ in range(depth - 1):
        ctx.submit_gof(view)
    for _ in range(depth + 2):
        ctx.submit_gof(view); drain(32)
    for _ in range(depth - 1):
        drain(32)
    torch.cuda.synchronize()
    steps = 12
    t0 = time.perf_counter()
    pts = 0
    for _ in range(depth - 1):
        ctx.submit_gof(view)
    for s in range(steps - (depth - 1)):
        ctx.submit_gof(view); pts += drain(32)
    for _ in range(depth - 1):
        pts += drain(32)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"gofs_in_flight {depth}: {dt / steps * 1e3:.2f} ms/step  {pts / dt / 1e9:.2f} Gpts/s  ({(337.8 + 232.0) / (dt / steps * 1e3):.1f} GB/s over PCIe both ways)")
    ctx.close()
