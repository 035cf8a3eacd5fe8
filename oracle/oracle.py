"""ctypes wrapper of the CPU oracle ``oracle/liboracle.so``.  TEST INFRASTRUCTURE ONLY.

May be imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs,
never by the product path (``tmc2-rs_b200/``).  Input structures are the ones of ``include/tmc2gpu.h`` (shared ctypes
mirror in ``tmc2-rs_b200/abi.py``), so the oracle and the CUDA path consume the very same buffers.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, n) for n in ("tmc2_oracle.c", "tmc2_oracle.h")] + \
          [os.path.join(_HERE, "..", "include", "tmc2gpu.h")]
    if force or not os.path.exists(_LIB_PATH) or \
            any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class OrcFrame(C.Structure):
    _fields_ = [("point_count", C.c_uint64), ("positions", C.c_void_p), ("colors", C.c_void_p),
                ("colors16bit", C.c_void_p), ("point_patch_indexes", C.c_void_p), ("partition", C.c_void_p),
                ("point_to_pixel", C.c_void_p), ("occupancy_map", C.c_void_p), ("block_to_patch", C.c_void_p),
                ("block_count", C.c_uint64), ("with_colors", C.c_uint8), ("boundary_type", C.c_void_p),
                ("positions_presmooth", C.c_void_p), ("colors16bit_presmooth", C.c_void_p),
                ("smoothed_positions", C.c_uint64), ("smoothed_colors", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_generate_block_to_patch_from_occupancy_map_video.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_generate_block_to_patch_from_occupancy_map_video.restype = C.c_int
        L.orc_reconstruct_frame.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.POINTER(OrcFrame))]
        L.orc_reconstruct_frame.restype = C.c_int
        L.orc_frame_free.argtypes = [C.POINTER(OrcFrame)]
        L.orc_frame_free.restype = None
        L.orc_convert_yuv10_to_rgb8.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_convert_yuv10_to_rgb8.restype = None
        L.orc_patch_to_canvas.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                          C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_patch_to_canvas.restype = C.c_int
        L.orc_patch_generate_point.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint16, C.c_void_p]
        L.orc_patch_generate_point.restype = None
        L.orc_time_frames.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_time_frames.restype = C.c_int64
        _lib = L
    return _lib


def _arr(ptr, n, dtype, cols=None):
    if not ptr or n == 0:
        shape = (0,) if cols is None else (0, cols)
        return np.zeros(shape, dtype=dtype)
    count = n if cols is None else n * cols
    buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    a = np.frombuffer(buf, dtype=dtype, count=count).copy()
    return a if cols is None else a.reshape(n, cols)


def block_to_patch(view, frame: int) -> np.ndarray:
    """src/codec.rs:205-250 on the CPU.  ``view`` is a tmc2-rs_b200.abi.GofView."""
    g = view.c
    res = g.params.occupancy_resolution
    out = np.zeros((g.width // res) * (g.height // res), dtype=np.uint64)
    st = lib().orc_generate_block_to_patch_from_occupancy_map_video(C.addressof(g), frame, out.ctypes.data)
    if st:
        raise OracleError(st)
    return out


class OracleError(RuntimeError):
    def __init__(self, status):
        self.status = int(status)
        super().__init__(f"oracle status {status}")


def reconstruct_frame(view, frame: int) -> Dict[str, np.ndarray]:
    """The per-frame body of src/decoder.rs:188-311 on the CPU; returns every intermediate as numpy arrays."""
    p = C.POINTER(OrcFrame)()
    st = lib().orc_reconstruct_frame(C.addressof(view.c), frame, C.byref(p))
    if st:
        raise OracleError(st)
    f = p.contents
    n = int(f.point_count)
    g = view.c
    out = {
        "point_count": n,
        "positions": _arr(f.positions, n, np.uint16, 3),
        "colors": _arr(f.colors, n if f.with_colors else 0, np.uint8, 3),
        "colors16bit": _arr(f.colors16bit, n if f.with_colors else 0, np.uint16, 3),
        "point_patch_indexes": _arr(f.point_patch_indexes, n, np.uint64, 2),
        "partition": _arr(f.partition, n, np.uint64),
        "point_to_pixel": _arr(f.point_to_pixel, n, np.uint64, 3),
        "occupancy_map": _arr(f.occupancy_map, g.width * g.height, np.uint8).reshape(g.height, g.width),
        "block_to_patch": _arr(f.block_to_patch, int(f.block_count), np.uint64),
        "boundary_type": _arr(f.boundary_type, n, np.uint8),
        "positions_presmooth": _arr(f.positions_presmooth, n, np.uint16, 3),
        "colors16bit_presmooth": _arr(f.colors16bit_presmooth, n, np.uint16, 3),
        "smoothed_positions": int(f.smoothed_positions),
        "smoothed_colors": int(f.smoothed_colors),
        "with_colors": bool(f.with_colors),
    }
    lib().orc_frame_free(p)
    return out


def convert_yuv10_to_rgb8(yuv) -> np.ndarray:
    yuv = np.ascontiguousarray(yuv, dtype=np.uint16).reshape(-1, 3)
    out = np.zeros_like(yuv, dtype=np.uint8)
    L = lib()
    for i in range(len(yuv)):
        L.orc_convert_yuv10_to_rgb8(yuv[i].ctypes.data, out[i].ctypes.data)
    return out


def time_frames(view, first: int, count: int) -> int:
    r = lib().orc_time_frames(C.addressof(view.c), first, count)
    if r < 0:
        raise OracleError(-r)
    return int(r)
