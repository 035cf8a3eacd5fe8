"""ctypes mirror of ``include/tmc2gpu.h`` (the C ABI of the reconstruction path).

Every structure here has the same field order, types and size as its C counterpart; ``tests/test_abi.py``
checks sizes/offsets against a table emitted by the C compiler.  The numpy ``PATCH_DTYPE`` is the same
40-byte record as ``tmc2_patch`` so patch lists travel as plain arrays.

Reference counterparts: ``Patch`` src/decoder.rs:711-783, ``GeneratePointCloudParams`` src/codec.rs:140-170,
``PointSet3`` src/codec.rs:20-36, ``Video``/``Image`` src/decoder.rs:913-1021.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

ABI_VERSION = 1

# ---- status codes (tmc2_status) ---------------------------------------------------------------------------------
OK, END, ERR_INVALID_ARG, ERR_PATCH_OUT_OF_CANVAS, ERR_SHORT_VIDEO, ERR_MAP_COUNT, ERR_UNSUPPORTED, \
    ERR_CAPACITY, ERR_STATE, ERR_NO_DEVICE, ERR_CUDA, ERR_INTERNAL = range(12)
STATUS_NAMES = ["OK", "END", "ERR_INVALID_ARG", "ERR_PATCH_OUT_OF_CANVAS", "ERR_SHORT_VIDEO", "ERR_MAP_COUNT",
                "ERR_UNSUPPORTED", "ERR_CAPACITY", "ERR_STATE", "ERR_NO_DEVICE", "ERR_CUDA", "ERR_INTERNAL"]

# ---- PatchOrientation (src/decoder.rs:694-707) ------------------------------------------------------------------
ORIENT_DEFAULT, ORIENT_SWAP, ORIENT_ROT90, ORIENT_ROT180, ORIENT_ROT270, ORIENT_MIRROR, ORIENT_MROT90, \
    ORIENT_MROT180, ORIENT_MROT270 = range(9)
ORIENTATION_REFERENCE, ORIENTATION_SPEC = 0, 1

CTX_TWO_PASS_SCAN = 1
CTX_DEVICE_OUTPUT = 2
LAUNCH_TIMED = 1
PLY_ASCII, PLY_BINARY_LE = 0, 1          # tmc2gpu_frame_to_ply formats (src/writer.rs:8-12)


class Tmc2Error(RuntimeError):
    """Raised where the reference would panic; carries the C status code."""

    def __init__(self, status: int, where: str = "", detail: str = ""):
        self.status = int(status)
        name = STATUS_NAMES[status] if 0 <= status < len(STATUS_NAMES) else str(status)
        super().__init__(f"{where}: {name}" + (f" ({detail})" if detail else ""))


class CPatch(C.Structure):
    _fields_ = [("u0", C.c_uint32), ("v0", C.c_uint32), ("size_u0", C.c_uint32), ("size_v0", C.c_uint32),
                ("u1", C.c_uint32), ("v1", C.c_uint32), ("d1", C.c_uint32),
                ("lod_x", C.c_uint16), ("lod_y", C.c_uint16),
                ("normal_axis", C.c_uint8), ("tangent_axis", C.c_uint8), ("bitangent_axis", C.c_uint8),
                ("projection_mode", C.c_uint8), ("patch_orientation", C.c_uint8),
                ("axis_of_additional_plane", C.c_uint8), ("_reserved", C.c_uint8 * 2)]


PATCH_DTYPE = np.dtype([("u0", "<u4"), ("v0", "<u4"), ("size_u0", "<u4"), ("size_v0", "<u4"),
                        ("u1", "<u4"), ("v1", "<u4"), ("d1", "<u4"), ("lod_x", "<u2"), ("lod_y", "<u2"),
                        ("normal_axis", "u1"), ("tangent_axis", "u1"), ("bitangent_axis", "u1"),
                        ("projection_mode", "u1"), ("patch_orientation", "u1"),
                        ("axis_of_additional_plane", "u1"), ("_reserved", "u1", (2,))])
assert PATCH_DTYPE.itemsize == C.sizeof(CPatch) == 40


class CParams(C.Structure):
    _fields_ = [("occupancy_resolution", C.c_uint32), ("occupancy_precision", C.c_uint32),
                ("map_count_minus1", C.c_uint8), ("absolute_d1", C.c_uint8), ("geometry_bitdepth_3d", C.c_uint8),
                ("attribute_count", C.c_uint8), ("orientation_mode", C.c_uint8),
                ("enable_size_quantization", C.c_uint8), ("multiple_streams", C.c_uint8), ("pbf_enabled", C.c_uint8),
                ("enhanced_occupancy_map", C.c_uint8), ("point_local_reconstruction", C.c_uint8),
                ("single_map_pixel_interleaving", C.c_uint8), ("use_additional_points_patch", C.c_uint8),
                ("geometry_smoothing", C.c_uint8), ("color_smoothing", C.c_uint8), ("attribute_bitdepth", C.c_uint8),
                ("_reserved0", C.c_uint8),
                ("grid_size", C.c_uint16), ("threshold_smoothing", C.c_uint16), ("cgrid_size", C.c_uint16),
                ("threshold_color_smoothing", C.c_uint16), ("threshold_color_difference", C.c_uint16),
                ("threshold_color_variation", C.c_uint16)]


class CFrame(C.Structure):
    _fields_ = [("occ", C.c_void_p), ("geo", C.c_void_p * 2), ("attr_y", C.c_void_p * 2),
                ("attr_u", C.c_void_p * 2), ("attr_v", C.c_void_p * 2), ("patches", C.c_void_p),
                ("patch_count", C.c_uint32), ("occ_stride", C.c_uint32), ("geo_stride", C.c_uint32),
                ("attr_stride_y", C.c_uint32), ("attr_stride_c", C.c_uint32), ("_reserved", C.c_uint32)]


class CGof(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("occ_width", C.c_uint32), ("occ_height", C.c_uint32),
                ("frame_count", C.c_uint32), ("geo_video_frames", C.c_uint32), ("attr_video_frames", C.c_uint32),
                ("_reserved", C.c_uint32), ("frames", C.POINTER(CFrame)), ("params", CParams)]


class CFrameOut(C.Structure):
    _fields_ = [("frame_index", C.c_uint64), ("point_count", C.c_uint64), ("positions", C.c_void_p),
                ("colors", C.c_void_p), ("with_colors", C.c_uint8), ("memory_space", C.c_uint8), ("device", C.c_uint8),
                ("_reserved", C.c_uint8 * 5),
                ("smoothed_positions", C.c_uint64), ("smoothed_colors", C.c_uint64), ("_handle", C.c_void_p)]


class CLimits(C.Structure):
    _fields_ = [("max_width", C.c_uint32), ("max_height", C.c_uint32), ("max_frames", C.c_uint32),
                ("max_patches_per_frame", C.c_uint32), ("gofs_in_flight", C.c_uint32), ("flags", C.c_uint32)]


class CPointCloudOut(C.Structure):
    _fields_ = [("capacity_points", C.c_uint64), ("point_count", C.c_uint64), ("positions", C.c_void_p),
                ("colors", C.c_void_p), ("colors16bit", C.c_void_p), ("partition", C.c_void_p),
                ("point_to_pixel", C.c_void_p), ("occupancy_map", C.c_void_p), ("block_to_patch", C.c_void_p),
                ("boundary_type", C.c_void_p), ("positions_presmooth", C.c_void_p),
                ("colors16bit_presmooth", C.c_void_p), ("smoothed_positions", C.c_uint64),
                ("smoothed_colors", C.c_uint64)]


# ---- host-side descriptions -------------------------------------------------------------------------------------
@dataclass
class Params:
    """Mirror of ``GeneratePointCloudParams`` (src/codec.rs:140-170) + smoothing parameter surface."""
    occupancy_resolution: int = 16
    occupancy_precision: int = 4
    map_count_minus1: int = 1
    absolute_d1: bool = True
    geometry_bitdepth_3d: int = 10
    attribute_count: int = 1
    orientation_mode: int = ORIENTATION_REFERENCE
    enable_size_quantization: bool = False
    multiple_streams: bool = False
    pbf_enabled: bool = False
    enhanced_occupancy_map: bool = False
    point_local_reconstruction: bool = False
    single_map_pixel_interleaving: bool = False
    use_additional_points_patch: bool = False
    geometry_smoothing: bool = False
    color_smoothing: bool = False
    attribute_bitdepth: int = 10
    grid_size: int = 8
    threshold_smoothing: int = 64
    cgrid_size: int = 4
    threshold_color_smoothing: int = 10
    threshold_color_difference: int = 10
    threshold_color_variation: int = 6

    def to_c(self) -> CParams:
        c = CParams()
        for name, _ in CParams._fields_:
            if name.startswith("_"):
                continue
            setattr(c, name, int(getattr(self, name)))
        return c


@dataclass
class Gof:
    """Decoded planes + patch lists of one group of frames, as numpy arrays (tight, C-contiguous).

    occ    [F, occH, occW] u8      reference ``atlas.occ_frames``           (Video<u8>)
    geo    [F, 2, H, W]   u16      reference ``atlas.geo_frames[0]``  frame f*2+m, channel 0
    attr_y [F, 2, H, W]   u16      reference ``atlas.attr_frames[0]`` frame f*2+m, channel 0
    attr_u/attr_v [F, 2, H/2, W/2] u16  channels 1, 2 (4:2:0)
    patches: one PATCH_DTYPE array per frame (``tile.patches``)
    """
    width: int
    height: int
    occ: np.ndarray
    geo: np.ndarray
    attr_y: Optional[np.ndarray]
    attr_u: Optional[np.ndarray]
    attr_v: Optional[np.ndarray]
    patches: List[np.ndarray]
    params: Params = field(default_factory=Params)
    geo_video_frames: Optional[int] = None
    attr_video_frames: Optional[int] = None

    @property
    def frame_count(self) -> int:
        return int(self.occ.shape[0])

    def input_bytes(self) -> int:
        n = self.occ.nbytes + self.geo.nbytes
        for a in (self.attr_y, self.attr_u, self.attr_v):
            if a is not None:
                n += a.nbytes
        return n + sum(p.nbytes for p in self.patches)

    def subset(self, frames: Sequence[int]) -> "Gof":
        idx = list(frames)
        sel = lambda a: None if a is None else np.ascontiguousarray(a[idx])
        return Gof(self.width, self.height, sel(self.occ), sel(self.geo), sel(self.attr_y), sel(self.attr_u),
                   sel(self.attr_v), [self.patches[i] for i in idx], self.params)


class GofView:
    """Builds (and keeps alive) the ``tmc2_gof`` C view of a :class:`Gof`."""

    def __init__(self, gof: Gof):
        self.gof = gof
        F = gof.frame_count
        # planes may be row-padded views (ffmpeg's linesize > width): samples of a row contiguous, row pitch a multiple of
        # the sample size -- the pitch travels in the *_stride fields (the reference itself assumes tight planes)
        def pitch(name):
            a = getattr(gof, name)
            if a.strides[-1] != a.itemsize or a.strides[-2] % a.itemsize or a.strides[-2] < a.shape[-1] * a.itemsize:
                raise ValueError(f"{name}: rows must be contiguous runs of samples (row pitch >= width)")
            return a.strides[-2] // a.itemsize
        assert gof.occ.dtype == np.uint8 and gof.geo.dtype == np.uint16
        self._frames = (CFrame * max(F, 1))()
        self._keep = []
        for f in range(F):
            fr = self._frames[f]
            fr.occ = gof.occ[f].ctypes.data
            fr.occ_stride = pitch("occ")
            for m in range(2):
                fr.geo[m] = gof.geo[f, m].ctypes.data if gof.geo.shape[1] > m else None
                if gof.attr_y is not None:
                    fr.attr_y[m] = gof.attr_y[f, m].ctypes.data
                    fr.attr_u[m] = gof.attr_u[f, m].ctypes.data
                    fr.attr_v[m] = gof.attr_v[f, m].ctypes.data
            fr.geo_stride = pitch("geo")
            if gof.attr_y is not None:
                fr.attr_stride_y = pitch("attr_y")
                fr.attr_stride_c = pitch("attr_u")
                if pitch("attr_v") != fr.attr_stride_c:
                    raise ValueError("attr_u and attr_v must share one row pitch")
            p = np.ascontiguousarray(gof.patches[f], dtype=PATCH_DTYPE)
            self._keep.append(p)
            fr.patches = p.ctypes.data if len(p) else None
            fr.patch_count = len(p)
        g = CGof()
        g.width, g.height = gof.width, gof.height
        g.occ_height, g.occ_width = gof.occ.shape[1], gof.occ.shape[2]
        g.frame_count = F
        g.geo_video_frames = gof.geo_video_frames if gof.geo_video_frames is not None else F * gof.geo.shape[1]
        if gof.attr_video_frames is not None:
            g.attr_video_frames = gof.attr_video_frames
        else:
            g.attr_video_frames = 0 if gof.attr_y is None else F * gof.attr_y.shape[1]
        g.frames = C.cast(self._frames, C.POINTER(CFrame))
        g.params = gof.params.to_c()
        self.c = g

    def ref(self):
        return C.byref(self.c)
