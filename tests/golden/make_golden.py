"""Writes tests/golden/small_f0.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference (tmc2-rs) cannot be executed in this image (no rustc/cargo, no ffmpeg) and ships no fixtures, so this golden
file pins the ORACLE against regressions; the oracle itself is pinned by the hand-derived vectors in test_oracle_kat.py."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import tmc2rs_b200  # noqa: E402,F401
from oracle import oracle  # noqa: E402
from tmc2rs_b200 import abi, synth  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


g = synth.make_gof(synth.config("small"))
r = oracle.reconstruct_frame(abi.GofView(g), 0)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "small_f0.npz"), point_count=r["point_count"],
                    sha_positions=digest(r["positions"]), sha_colors=digest(r["colors"]),
                    block_to_patch=r["block_to_patch"], positions_head=r["positions"][:256],
                    colors_head=r["colors"][:256])
print("points", r["point_count"])
