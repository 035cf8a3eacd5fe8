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
