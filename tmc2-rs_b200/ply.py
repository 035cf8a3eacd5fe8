"""PLY output of a reconstructed frame (the reference's consumer of ``PointSet3``, ``src/writer.rs:15-102``).

``ascii_ply`` reproduces the reference writer byte for byte (header lines of ``write_header`` :31-60, one ``x y z r g b`` line
per point from ``write_body`` :62-75).  ``binary_ply`` is the ``binary_little_endian`` variant the reference lists but leaves
commented out (:10-11, :41-46): the same header properties (``uint`` x y z, ``uchar`` red green blue), 15 bytes per point.
Host-side formatting of frames that are already on the host; nothing here touches the GPU path.
"""
from __future__ import annotations

import numpy as np


def _header(fmt: str, n: int, with_colors: bool) -> bytes:
    lines = ["ply", f"format {fmt} 1.0", f"element vertex {n}", "property uint x", "property uint y", "property uint z"]
    if with_colors:
        lines += ["property uchar red", "property uchar green", "property uchar blue"]
    lines += ["element face 0", "property list uint8 int32 vertex_index", "end_header"]
    return ("\n".join(lines) + "\n").encode("ascii")


def ascii_ply(positions: np.ndarray, colors: np.ndarray | None) -> bytes:
    """src/writer.rs:24-75 (Format::Ascii)."""
    n = len(positions)
    with_colors = colors is not None
    cols = [positions.astype(np.int64)]
    if with_colors:
        cols.append(colors.astype(np.int64))
    table = np.concatenate(cols, axis=1) if n else np.zeros((0, 6 if with_colors else 3), np.int64)
    body = "".join(" ".join(map(str, row)) + "\n" for row in table.tolist())
    return _header("ascii", n, with_colors) + body.encode("ascii")


def binary_ply(positions: np.ndarray, colors: np.ndarray | None) -> bytes:
    """The binary_little_endian form of the same file (src/writer.rs:10, :41-43, commented out upstream)."""
    n = len(positions)
    with_colors = colors is not None
    dt = [("x", "<u4"), ("y", "<u4"), ("z", "<u4")] + ([("r", "u1"), ("g", "u1"), ("b", "u1")] if with_colors else [])
    rec = np.zeros(n, dtype=np.dtype(dt))
    rec["x"], rec["y"], rec["z"] = positions[:, 0], positions[:, 1], positions[:, 2]
    if with_colors:
        rec["r"], rec["g"], rec["b"] = colors[:, 0], colors[:, 1], colors[:, 2]
    return _header("binary_little_endian", n, with_colors) + rec.tobytes()


def read_ply(data: bytes):
    """Positions [n,3] u16 and colours [n,3] u8 (or None) of a file written by either function (used by tools/compare_dump.py)."""
    end = data.index(b"end_header\n") + len(b"end_header\n")
    head = data[:end].decode("ascii").split("\n")
    fmt = next(l.split()[1] for l in head if l.startswith("format"))
    n = next(int(l.split()[2]) for l in head if l.startswith("element vertex"))
    props = [l.split()[2] for l in head if l.startswith("property") and "list" not in l]
    with_colors = "red" in props
    if fmt == "ascii":
        vals = np.array(data[end:].split(), dtype=np.int64).reshape(n, len(props)) if n else np.zeros((0, len(props)), np.int64)
        return vals[:, :3].astype(np.uint16), (vals[:, 3:6].astype(np.uint8) if with_colors else None)
    dt = [("x", "<u4"), ("y", "<u4"), ("z", "<u4")] + ([("r", "u1"), ("g", "u1"), ("b", "u1")] if with_colors else [])
    rec = np.frombuffer(data[end:], dtype=np.dtype(dt), count=n)
    pos = np.stack([rec["x"], rec["y"], rec["z"]], axis=1).astype(np.uint16)
    return pos, (np.stack([rec["r"], rec["g"], rec["b"]], axis=1) if with_colors else None)
