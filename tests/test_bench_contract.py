"""bench.py contract on the CPU: the reference arm prints one JSON line with the driver's keys; the B200 arm refuses to run
without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = run("--impl", "reference", "--config", "c1", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "reconstructed_points_per_sec" and d["unit"] == "points/s"
    assert d["steps"] == 1 and d["warmup"] == 0 and d["higher_is_better"] is True and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("c1")


def test_reference_arm_under_torchrun_only_rank0_prints():
    r = run("--impl", "reference", "--config", "c1", "--steps", "1", "--warmup", "0", env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_strong_scaling_plan_covers_the_sequence():
    """bench.py's N > 1 workload (BASELINE config 3, strong scaling): the ranks' GOF lists tile the 300-frame sequence exactly,
    cut at the sequence's GOF boundaries; both arms derive the same `config` dict from the same arguments."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for world in (1, 2, 3, 4, 8):
        sizes = [bench.rank_gofs("strong", 30, r, world) for r in range(world)]
        assert sum(sum(s) for s in sizes) == bench.SEQ_FRAMES
        assert all(0 < g <= bench.SEQ_GOF for s in sizes for g in s)
        f = 0
        for s in sizes:                                   # contiguous slices; no GOF of a rank straddles a sequence GOF boundary
            for g in s:
                assert f // bench.SEQ_GOF == (f + g - 1) // bench.SEQ_GOF
                f += g
    args = argparse.Namespace(config="auto", frames=0, no_smoothing=False, gpus=8)
    mode, sname, frames, smoothing, desc = bench.workload(args, 8)
    assert (mode, sname, frames, smoothing) == ("strong", "c3", 30, True) and desc["frames_per_step"] == 300
    assert bench.workload(args, 1) == bench.workload(args, 8)                 # --gpus 8 without torchrun (the reference arm): same dict
    args1 = argparse.Namespace(config="auto", frames=0, no_smoothing=False, gpus=1)
    mode1, sname1, frames1, _, desc1 = bench.workload(args1, 1)
    assert (mode1, sname1, frames1) == ("weak", "c2", 32) and desc1["atlas"] == "1024x1024"
    assert bench.rank_gofs("weak", 32, 0, 1) == [32] * bench.GOFS_PER_STEP
