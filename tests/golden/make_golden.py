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


if __name__ == "__main__":
    g = synth.make_gof(synth.config("small"))
    r = oracle.reconstruct_frame(abi.GofView(g), 0)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "small_f0.npz"), point_count=r["point_count"],
                        sha_positions=digest(r["positions"]), sha_colors=digest(r["colors"]),
                        block_to_patch=r["block_to_patch"], positions_head=r["positions"][:256],
                        colors_head=r["colors"][:256])
    print("points", r["point_count"])


# The smoothing stages are this repository's own integer specification (the reference has only stubs):
# a second fixture freezes it, so that the oracle and the kernels cannot drift together unnoticed.
def smoothing_case():
    """The 'small' synthetic GOF with its patches crowded into one 3D region (same recipe as
    tests/test_gpu_parity.py::test_smoothing_dense_overlap_case), both smoothing stages on."""
    gs = synth.make_gof(synth.config("small"))
    for p in gs.patches:
        p["u1"] = 100 + (np.arange(len(p)) % 3)
        p["v1"] = 100 + (np.arange(len(p)) % 2)
        p["d1"] = np.where(p["projection_mode"] == 0, 64, 1024 - 64 - 300)
        p["normal_axis"], p["tangent_axis"], p["bitangent_axis"] = 0, 2, 1
    gs.params.geometry_smoothing = True
    gs.params.color_smoothing = True
    return gs


if __name__ == "__main__":
    rs = oracle.reconstruct_frame(abi.GofView(smoothing_case()), 1)
    moved = np.flatnonzero((rs["positions"] != rs["positions_presmooth"]).any(axis=1))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "small_smooth_f1.npz"), point_count=rs["point_count"],
                        sha_positions=digest(rs["positions"]), sha_colors=digest(rs["colors"]),
                        sha_boundary_type=digest(rs["boundary_type"]),
                        smoothed_positions=rs["smoothed_positions"], smoothed_colors=rs["smoothed_colors"],
                        moved_index_head=moved[:64], moved_positions_head=rs["positions"][moved[:64]])
    print("smoothing fixture: points", rs["point_count"], "moved", rs["smoothed_positions"], "recoloured", rs["smoothed_colors"])
