"""Loader of the CUDA library ``libtmc2gpu.so`` (the C ABI of ``include/tmc2gpu.h``).

Fails loudly when the library is missing or cannot be loaded: there is no CPU fallback and nothing here ever touches
``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TMC2_LIB") or os.path.join(_HERE, "libtmc2gpu.so")   # TMC2_LIB: tuning variants

# every symbol include/tmc2gpu.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("tmc2gpu_abi_version", C.c_uint32, []),
    ("tmc2gpu_device_count", C.c_int, []),
    ("tmc2gpu_create", C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(abi.CLimits), C.POINTER(_P)]),
    ("tmc2gpu_destroy", None, [_P]),
    ("tmc2gpu_last_error", C.c_char_p, [_P]),
    ("tmc2gpu_status_string", C.c_char_p, [C.c_int]),
    ("tmc2gpu_alloc_pinned", _P, [C.c_size_t]),
    ("tmc2gpu_free_pinned", None, [_P]),
    ("tmc2gpu_submit_gof", C.c_int, [_P, C.POINTER(abi.CGof)]),
    ("tmc2gpu_wait_inputs", C.c_int, [_P]),
    ("tmc2gpu_next_frame", C.c_int, [_P, C.POINTER(abi.CFrameOut)]),
    ("tmc2gpu_release_frame", C.c_int, [_P, C.POINTER(abi.CFrameOut)]),
    ("tmc2gpu_frame_to_ply", C.c_int, [_P, C.POINTER(abi.CFrameOut), C.c_uint32, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    ("tmc2gpu_upload_gof", C.c_int, [_P, C.POINTER(abi.CGof), C.POINTER(_P)]),
    ("tmc2gpu_reconstruct_resident", C.c_int, [_P, _P, _P]),
    ("tmc2gpu_reconstruct_resident_ex", C.c_int, [_P, _P, _P, C.c_uint32]),
    ("tmc2gpu_resident_counts", C.c_int, [_P, _P, C.POINTER(C.c_uint64)]),
    ("tmc2gpu_resident_fetch", C.c_int, [_P, _P, C.c_uint32, _P, _P, C.c_uint64]),
    ("tmc2gpu_free_resident", C.c_int, [_P, _P]),
    ("tmc2gpu_last_launch_info", C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("tmc2gpu_last_unpack_ms", C.c_int, [_P, C.POINTER(C.c_float)]),
    ("tmc2gpu_last_stage_ms", C.c_int, [_P, C.POINTER(C.c_float)]),
    ("tmc2gpu_generate_block_to_patch_from_occupancy_map_video", C.c_int, [_P, C.POINTER(abi.CGof), C.c_uint32, _P]),
    ("tmc2gpu_generate_point_cloud", C.c_int, [_P, C.POINTER(abi.CGof), C.c_uint32, C.POINTER(abi.CPointCloudOut)]),
    ("tmc2gpu_convert_yuv16_to_rgb8", C.c_int, [_P, _P, C.c_uint64, _P]),
]

_lib = None


def load(build_if_missing: bool = False) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build
            build.build()
        else:
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                               " -- the reconstruction path is CUDA-only, there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError here == ABI drift
        fn.restype = res
        fn.argtypes = args
    if lib.tmc2gpu_abi_version() != abi.ABI_VERSION:
        raise RuntimeError("libtmc2gpu.so ABI version mismatch")
    _lib = lib
    return lib
