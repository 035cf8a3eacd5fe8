#!/usr/bin/env python
"""bench.py -- throughput of the V-PCC rec0 reconstruction hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU (oracle port)
    torchrun ... bench.py --gpus N ...                       # one rank per GPU, frames sharded GOF-wise, no collective

A "step" is one pass of the hot path over one synthetic GOF (BASELINE config 2: 32 frames, 1024x1024 atlas, 10-bit
geometry, 2 maps, occupancy precision 4, grid geometry smoothing + colour smoothing ON).  With N ranks every rank
reconstructs its own GOF per step (weak scaling; frames / GOFs are independent).

  value  points/s with the planes already resident in HBM (kernels only, CUDA events on the launching streams): step i
         reconstructs resident GOF i % 3 on stream i % 3, like the three GOFs in flight of the streaming path, so the
         small latency-bound passes of one GOF run under the emit of another (--resident-gofs 1: one stream, GOF latency)
  e2e    points/s through the public C ABI (tmc2gpu_submit_gof / tmc2gpu_next_frame) from PINNED HOST planes:
         H2D of every plane and D2H of every reconstructed frame inside the timed region
  roofline  algorithmic bytes of the dominant kernel (fused unpack) / its device time, against the measured HBM peak
  cpu_baseline  the reference's single-threaded algorithm (C restatement, oracle/) on this box's host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import tmc2rs_b200  # noqa: E402,F401
from tmc2rs_b200 import abi, shard, synth  # noqa: E402

METRIC = "reconstructed_points_per_sec"
UNIT = "points/s"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def build_workload(name: str, frames: int, distinct: int):
    """GOF of `frames` frames cycling `distinct` distinct synthetic frames (generation cost only)."""
    cfg = synth.config(name, frames=min(distinct, frames))
    base = synth.make_gof(cfg)
    return synth.replicate_gof(base, frames), cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(gpu_index: int):
    """Pin this rank to the CPU cores next to its GPU (NVML affinity) before any pinned memory is allocated, so that the
    staging buffers land on the GPU's own NUMA node.  Best effort: silently keeps the inherited affinity on failure."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if target:
            os.sched_setaffinity(0, target)
        return sorted(target) if target else None
    except Exception:
        return None


def cpu_reference_run(gof, frames_per_step: int, steps: int, warmup: int, threads: int):
    """The reference algorithm on host cores: oracle port (the Rust reference cannot be built here), single-threaded
    like the reference (README.md:7, src/lib.rs:113) unless --ref-threads asks for frame-parallel workers."""
    from oracle import oracle
    view = abi.GofView(gof)
    F = gof.frame_count

    def one_step(step):
        lo = (step * frames_per_step) % F
        idx = [(lo + k) % F for k in range(frames_per_step)]
        if threads <= 1:
            return sum(oracle.time_frames(view, f, 1) for f in idx)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:     # ctypes releases the GIL
            return sum(ex.map(lambda f: oracle.time_frames(view, f, 1), idx))
    for w in range(warmup):
        one_step(w)
    t0 = time.perf_counter()
    pts = sum(one_step(s) for s in range(steps))
    dt = time.perf_counter() - t0
    return pts, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", help="synthetic workload: c2 (default, BASELINE config 2), c1, c3, c4")
    ap.add_argument("--frames", type=int, default=0, help="frames per GOF (default: the config's)")
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic frames cycled inside the GOF")
    ap.add_argument("--ref-threads", type=int, default=1)
    ap.add_argument("--cpu-sample-frames", type=int, default=0, help="frames in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-smoothing", action="store_true", help="reference-exact rec0 path only (no post-processing)")
    ap.add_argument("--resident-gofs", type=int, default=3,
                    help="resident GOFs the kernel-only loop alternates between, each on its own stream (consecutive GOFs "
                         "of a stream overlap: the small latency-bound passes of one run under the emit of the other)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    frames_default = {"c1": 1, "c2": 32, "c3": 32, "c4": 8}.get(args.config, 32)
    frames = args.frames or frames_default
    smoothing = args.config != "c1" and not args.no_smoothing
    cfg_desc = {"workload": f"{args.config}: {frames}-frame GOF, synthetic decoded planes + patch metadata", "frames_per_gof": frames}

    # ------------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        gof, cfg = build_workload(args.config, frames, min(args.distinct, 4))
        gof.params.geometry_smoothing = smoothing
        gof.params.color_smoothing = smoothing
        cfg_desc.update({"atlas": f"{cfg.width}x{cfg.height}", "smoothing": smoothing, "maps": 2,
                         "occupancy_precision": cfg.occupancy_precision})
        sample = max(1, min(frames, 2))                   # frames per step: bounded so K steps end within minutes
        pts, dt = cpu_reference_run(gof, sample, args.steps, args.warmup, args.ref_threads)
        v = pts / dt
        line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u16", "data": "synthetic", "config": cfg_desc,
                "frames_per_sec": sample * args.steps / dt,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": args.ref_threads, "kind": "port",
                                 "sample": f"{sample} frames/step x {args.steps} steps of the same GOF; C restatement of "
                                           "tmc2-rs src/codec.rs (the Rust reference cannot be built: no cargo/rustc)",
                                 "host_cores": os.cpu_count()},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        # context only: the same port with one worker per host core, frames in parallel (the reference itself reconstructs on
        # one thread, src/lib.rs:113 -- the headline `value` above is that)
        cores = len(os.sched_getaffinity(0))
        if cores > 1 and args.ref_threads <= 1:
            par = min(cores, max(frames, 1))
            pts2, dt2 = cpu_reference_run(gof, par, 2, 1, par)
            line["cpu_baseline"]["frame_parallel"] = {"value": pts2 / dt2, "unit": UNIT, "cores": par,
                                                      "sample": f"{par} frames at once x 2 steps"}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from tmc2rs_b200 import codec

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the reconstruction path is CUDA-only (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    cpus = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    gof, cfg = build_workload(args.config, frames, args.distinct)
    gof.params.geometry_smoothing = smoothing
    gof.params.color_smoothing = smoothing
    cfg_desc.update({"atlas": f"{cfg.width}x{cfg.height}", "smoothing": smoothing, "maps": 2,
                     "occupancy_precision": cfg.occupancy_precision, "bitdepth_3d": cfg.bitdepth_3d,
                     "sharding": f"{world} rank(s), one GOF per rank per step, no collective",
                     "e2e_gofs_in_flight": 3,
                     "host_cpus_of_rank0": (f"{cpus[0]}-{cpus[-1]} ({len(cpus)})" if cpus else "inherited"),
                     "l2": f"inputs {gof.input_bytes() / 1e6:.0f} MB/step > 126 MB L2 (no flush needed)"})
    depth = 3                                             # GOFs in flight on the streaming path (keeps both DMA engines busy)
    ctx = codec.Context(devices=(local_rank,), gofs_in_flight=depth)
    pinned = codec.pinned_copy_of(gof)
    view = abi.GofView(pinned)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: planes resident in HBM, kernels only -----------------------------------------------------------------
    n_res = max(1, args.resident_gofs)
    cfg_desc["resident_gofs_in_flight"] = n_res
    residents = [ctx.upload_gof(view) for _ in range(n_res)]
    res = residents[0]
    # dedicated torch streams: torch.cuda.Event only sees the stream it is recorded on, and torch's default stream
    # has handle 0, which the C ABI reads as "use the library's own stream"
    tstreams = [torch.cuda.Stream(device=dev) for _ in range(n_res)]
    tstream = tstreams[0]
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert all(t.cuda_stream != 0 for t in tstreams)
    for _ in range(args.warmup):
        for r, t in zip(residents, tstreams):
            r.reconstruct(t.cuda_stream)
    counts = res.counts()
    points_per_step = int(counts.sum())
    launches_per_step, alg_bytes, _ = ctx.last_launch_info()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    unpack_ms, stage_acc = [], {}
    # step i reconstructs resident GOF i % n_res on stream i % n_res; the timed region starts when every stream has passed
    # e0 and ends when every stream has finished its last step
    e0.record(tstream)
    for t in tstreams[1:]:
        t.wait_event(e0)
    for i in range(args.steps):
        residents[i % n_res].reconstruct(tstreams[i % n_res].cuda_stream)
    for t in tstreams[1:]:
        done = torch.cuda.Event()
        done.record(t)
        tstream.wait_event(done)
    e1.record(tstream)
    barrier()
    kernel_ms = e0.elapsed_time(e1)
    # per-stage device times (library-recorded CUDA events on the same stream), measured on separate launches so that
    # reading them does not serialise the timed loop above
    for _ in range(min(args.steps, 5)):
        res.reconstruct(stream)
        st = ctx.last_stage_ms()
        unpack_ms.append(st["unpack"])
        for k, v in st.items():
            stage_acc.setdefault(k, []).append(v)
    kernel_ms_max, pts_all, frames_all = shard.reduce_metrics(kernel_ms, points_per_step * args.steps, frames * args.steps, dev)
    value = pts_all / (kernel_ms_max * 1e-3)
    for r in residents:
        r.free()

    # ---- e2e: public C ABI, pinned host planes in, pinned host frames out, two GOFs in flight ---------------------------
    def drain(n):
        got = 0
        for _ in range(n):
            cnt, pos_addr, _col_addr = ctx.next_frame_raw()      # frame is in pinned host memory when this returns
            got += cnt
        return got
    # warm-up with the same GOFs-in-flight pattern as the timed loop, so that every GOF slot of the context has its device
    # buffers and pinned result slab allocated before timing starts
    for _ in range(depth - 1):
        ctx.submit_gof(view)
    for _ in range(max(depth + 1, args.warmup)):
        ctx.submit_gof(view)
        drain(frames)
    for _ in range(depth - 1):
        drain(frames)
    barrier()
    trace = os.environ.get("TMC2_TRACE") is not None
    t0 = time.perf_counter()
    got = 0
    ahead = min(depth - 1, args.steps)
    for _ in range(ahead):
        ctx.submit_gof(view)
    for s in range(args.steps - ahead):
        ctx.submit_gof(view)
        got += drain(frames)
        if trace:
            sys.stderr.write(f"[bench] e2e step {s}: {(time.perf_counter() - t0) * 1e3:.2f} ms since start\n")
    for _ in range(ahead):
        got += drain(frames)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()            # sampled over both timed regions (kernel-only steps and the end-to-end steps)
    barrier()
    e2e_ms_max, e2e_pts, _ = shard.reduce_metrics(e2e_ms, got, frames * args.steps, dev)
    e2e_value = e2e_pts / (e2e_ms_max * 1e-3)
    h2d = gof.input_bytes()
    d2h = points_per_step * 9

    # ---- roofline of the dominant kernel -------------------------------------------------------------------------------
    peak, peak_src = measured_peak()
    t_unpack = statistics.mean(unpack_ms) if unpack_ms else float("nan")
    # the emit kernel is launched once per smoothing frame group (8 frames) -- once for the whole GOF without smoothing;
    # algorithmic bytes and device time are per launch (SURVEY.md 8d figure x the frames one launch processes)
    group = int(os.environ.get("TMC2_SMOOTH_GROUP", "32")) if smoothing else frames
    n_emit = max(1, -(-frames // max(group, 1)))
    achieved = alg_bytes / (t_unpack * 1e-3) / 1e9 if t_unpack and t_unpack > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "emit_kernel (fused occupancy upsample / unpack / attribute fetch / YUV->RGB; + boundary + cell statistics when smoothing)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "frac_of_nominal_8TBps": achieved / 8000.0,
                "launches_per_step": n_emit, "frames_per_launch": min(group, frames),
                "algorithmic_bytes_per_launch": alg_bytes / n_emit, "ms_per_launch": t_unpack / n_emit, "traffic": None,
                "stage_ms": {k: statistics.mean(v) for k, v in stage_acc.items()}}
    # the other HBM-streaming kernel of the path: the count pass reads the occupancy video and both geometry planes once
    # (its stage time also holds the tiny per-frame scan launch, so the fraction is a lower bound)
    t_count = statistics.mean(stage_acc.get("count_scan", [0.0])) if stage_acc else 0.0
    count_bytes = frames * (cfg.width // cfg.occupancy_precision) * (cfg.height // cfg.occupancy_precision) + frames * 2 * cfg.width * cfg.height * 2
    if t_count > 0:
        roofline["other_kernels"] = {"count_kernel+slot_scan_kernel": {
            "algorithmic_bytes_per_launch": count_bytes, "ms_per_launch": t_count,
            "achieved": count_bytes / (t_count * 1e-3) / 1e9, "frac": count_bytes / (t_count * 1e-3) / 1e9 / peak}}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            roofline["traffic"] = json.load(f).get(args.config + ("" if smoothing else "_nosmooth"))
    except Exception:
        pass

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16", "data": "synthetic", "config": cfg_desc,
            "frames_per_sec": frames_all / (kernel_ms_max * 1e-3), "points_per_step_per_gpu": points_per_step,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "frames_per_sec": frames * args.steps * world / (e2e_ms_max * 1e-3), "ms_per_step": e2e_ms_max / args.steps},
            "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "roofline": roofline, "clocks": clocks}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # ~10-20 s of single-threaded CPU work: 64 frames (two passes over the GOF) with smoothing, 128 without
        sample = args.cpu_sample_frames or ((64 if smoothing else 128) if frames > 1 else 16)
        if frames > 1:
            per, reps = min(sample, frames), max(1, sample // min(sample, frames))
        else:
            per, reps = 1, sample
        sample = per * reps
        pts, dt = cpu_reference_run(gof, per, reps, 0, 1)
        line["cpu_baseline"] = {"value": pts / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"{sample} frame(s) of the same workload, {dt:.1f} s; single-threaded C restatement "
                                          "of tmc2-rs src/codec.rs (+ this repo's smoothing spec) -- the Rust reference "
                                          "cannot be built in this image",
                                "host_cores": os.cpu_count(), "ms_per_frame": dt / sample * 1e3}
    ctx.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
