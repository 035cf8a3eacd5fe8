#!/usr/bin/env python
"""Time tmc2gpu_frame_to_ply on a BASELINE config-1 frame (0.8 M points): device formatting + D2H of the finished file against the
host writer mirror (tmc2rs_b200/ply.py, numpy / Python; the reference's own writer is a `write!` per point, src/writer.rs:62-75)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tmc2rs_b200 import abi, codec, ply, synth  # noqa: E402


def main():
    g = synth.make_gof(synth.config("c2", frames=4))
    view = abi.GofView(g)
    ctx = codec.Context()
    frames = ctx.decode_gof(view)
    out = {"points_per_frame": int(np.mean([len(f) for f in frames]))}
    for name, fmt in (("ascii", abi.PLY_ASCII), ("binary_le", abi.PLY_BINARY_LE)):
        for rep in range(3):                                            # the last repetition is reported (scratch buffers warm)
            ctx.submit_gof(view)
            fo = abi.CFrameOut()
            size = C.c_uint64(0)
            ms, total = [], 0
            buf = np.empty(40 << 20, np.uint8)
            for _ in frames:
                ctx.check(ctx.lib.tmc2gpu_next_frame(ctx.h, C.byref(fo)), "next_frame")
                t0 = time.perf_counter()
                ctx.check(ctx.lib.tmc2gpu_frame_to_ply(ctx.h, C.byref(fo), fmt, buf.ctypes.data, buf.nbytes, C.byref(size)), "ply")
                ms.append((time.perf_counter() - t0) * 1e3)
                total += size.value
                ctx.lib.tmc2gpu_release_frame(ctx.h, C.byref(fo))
        out[name] = {"device_ms_per_frame": round(float(np.mean(ms)), 3), "file_bytes": total // len(frames)}
    f0 = frames[0]
    t0 = time.perf_counter(); ply.binary_ply(f0.positions, f0.colors); out["binary_le"]["host_numpy_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
    n8 = len(f0) // 8
    t0 = time.perf_counter(); ply.ascii_ply(f0.positions[:n8], f0.colors[:n8])
    out["ascii"]["host_python_ms_scaled_from_one_eighth"] = round((time.perf_counter() - t0) * 8e3, 1)
    ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
