#!/usr/bin/env python
"""compare_dump.py -- compare a point cloud written by an external decoder (MPEG TMC2 v18, tmc2-rs) with one of ours.

    python tools/compare_dump.py REFERENCE OURS [--tolerance 1] [--unordered] [--json]

Prints the report `north_star` defines for the smoothing stages: the number of points whose position / colour differs at
all, and the number that differ by MORE than the tolerance (default +-1 quantisation step in any coordinate or colour
channel); for the reference-exact stages (no smoothing) the expectation is 0 differing points.

Inputs: PLY files as tmc2-rs writes them (ASCII, src/writer.rs:24-75: `uint x y z`, `uchar red green blue`), the
binary_little_endian form of the same header, or raw dumps `*.u16` (positions, n x 3 little-endian u16) with an optional
`*.u8` colour file next to them (n x 3).  By default points are compared IN ORDER (tmc2-rs and this repository emit them in
the same patch / block / raster order); `--unordered` sorts both clouds first (for decoders with another emission order:
positions, then colours, lexicographically) and then also reports points that exist on one side only.

Nothing upstream is available offline in the build image, so parity of the smoothing stages stays "unpinned" until somebody
runs this against a TMC2 dump; the tool itself is tested on synthetic pairs (tests/test_host.py).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tmc2rs_b200  # noqa: E402,F401
from tmc2rs_b200 import ply  # noqa: E402


def load(path: str):
    """(positions [n,3] int64, colours [n,3] int64 or None)."""
    if path.endswith(".u16"):
        pos = np.fromfile(path, dtype="<u2").reshape(-1, 3)
        cpath = path[:-4] + ".u8"
        col = np.fromfile(cpath, dtype=np.uint8).reshape(-1, 3) if os.path.exists(cpath) else None
    else:
        with open(path, "rb") as f:
            pos, col = ply.read_ply(f.read())
    return pos.astype(np.int64), (None if col is None else col.astype(np.int64))


def compare(ref, ours, tolerance: int = 1, unordered: bool = False) -> dict:
    (rp, rc), (op, oc) = ref, ours
    rep = {"points_reference": int(len(rp)), "points_ours": int(len(op)), "tolerance": tolerance, "ordered": not unordered}
    with_col = rc is not None and oc is not None
    if unordered:
        def order(p, c):
            keys = ([c[:, 2], c[:, 1], c[:, 0]] if c is not None else []) + [p[:, 2], p[:, 1], p[:, 0]]
            i = np.lexsort(keys)
            return p[i], (None if c is None else c[i])
        rp, rc = order(rp, rc)
        op, oc = order(op, oc)
    n = min(len(rp), len(op))
    rep["points_compared"] = int(n)
    rep["count_mismatch"] = int(abs(len(rp) - len(op)))
    dp = np.abs(rp[:n] - op[:n])
    rep["positions_differing"] = int((dp.max(axis=1) > 0).sum()) if n else 0
    rep["positions_beyond_tolerance"] = int((dp.max(axis=1) > tolerance).sum()) if n else 0
    rep["positions_max_abs_diff"] = int(dp.max()) if n else 0
    if with_col:
        dc = np.abs(rc[:n] - oc[:n])
        rep["colors_differing"] = int((dc.max(axis=1) > 0).sum()) if n else 0
        rep["colors_beyond_tolerance"] = int((dc.max(axis=1) > tolerance).sum()) if n else 0
        rep["colors_max_abs_diff"] = int(dc.max()) if n else 0
    rep["bit_exact"] = rep["count_mismatch"] == 0 and rep["positions_differing"] == 0 and rep.get("colors_differing", 0) == 0
    rep["within_tolerance"] = rep["count_mismatch"] == 0 and rep["positions_beyond_tolerance"] == 0 and \
        rep.get("colors_beyond_tolerance", 0) == 0
    return rep


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("reference")
    ap.add_argument("ours")
    ap.add_argument("--tolerance", type=int, default=1)
    ap.add_argument("--unordered", action="store_true")
    ap.add_argument("--json", action="store_true")
    a = ap.parse_args(argv)
    rep = compare(load(a.reference), load(a.ours), a.tolerance, a.unordered)
    if a.json:
        print(json.dumps(rep))
    else:
        print(f"points: reference {rep['points_reference']}, ours {rep['points_ours']} (compared {rep['points_compared']}, "
              f"{'in order' if rep['ordered'] else 'sorted'})")
        print(f"positions: {rep['positions_differing']} differ, {rep['positions_beyond_tolerance']} by more than "
              f"+-{a.tolerance} (max |diff| {rep['positions_max_abs_diff']})")
        if "colors_differing" in rep:
            print(f"colours:   {rep['colors_differing']} differ, {rep['colors_beyond_tolerance']} by more than "
                  f"+-{a.tolerance} (max |diff| {rep['colors_max_abs_diff']})")
        print("verdict:", "bit-exact" if rep["bit_exact"] else "within tolerance" if rep["within_tolerance"] else "DIFFERENT")
    return 0 if rep["within_tolerance"] else 1


if __name__ == "__main__":
    sys.exit(main())
