"""The C ABI from plain C++ (examples/stream_decode.cpp): worker thread + bounded(1) channel like the reference's
src/lib.rs, no Python and no CUDA headers on the consumer side.  The GPU test feeds it a GOF file and checks every frame
against the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

import tmc2rs_b200  # noqa: F401
from oracle import oracle
from tmc2rs_b200 import abi, build, synth

MAGIC = 0x32434D54


def write_gof(path, g):
    view = abi.GofView(g)
    with open(path, "wb") as f:
        f.write(struct.pack("<6I", MAGIC, g.width, g.height, g.occ.shape[2], g.occ.shape[1], g.frame_count))
        f.write(bytes(view.c.params))
        for k in range(g.frame_count):
            patches = view._keep[k]                                  # structured array with the C layout of tmc2_patch
            assert patches.dtype.itemsize == C.sizeof(abi.CPatch)
            f.write(struct.pack("<I", len(patches)))
            f.write(patches.tobytes())
            f.write(np.ascontiguousarray(g.occ[k]).tobytes())
            for arr in (g.geo, g.attr_y, g.attr_u, g.attr_v):
                for m in range(2):
                    f.write(np.ascontiguousarray(arr[k, m]).tobytes())
    return view


import ctypes as C  # noqa: E402


def test_cpp_driver_builds_and_fails_loudly_without_gpu(tmp_path):
    exe = build.build_example()
    assert os.path.exists(exe)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    g = synth.make_gof(synth.config("tiny"))
    write_gof(tmp_path / "in.gof", g)
    r = subprocess.run([exe, str(tmp_path / "in.gof"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr          # no CPU fallback


@pytest.mark.gpu
def test_cpp_driver_matches_oracle(tmp_path):
    exe = build.build_example()
    g = synth.make_gof(synth.config("small"))
    g.params.geometry_smoothing = True
    g.params.color_smoothing = True
    view = write_gof(tmp_path / "in.gof", g)
    r = subprocess.run([exe, str(tmp_path / "in.gof"), str(tmp_path / "out.bin"), "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert f"frames {3 * g.frame_count} " in r.stdout
    raw = open(tmp_path / "out.bin", "rb").read()
    off = 0
    for f in range(g.frame_count):
        n = struct.unpack_from("<Q", raw, off)[0]; off += 8
        pos = np.frombuffer(raw, np.uint16, 3 * n, off).reshape(n, 3); off += 6 * n
        col = np.frombuffer(raw, np.uint8, 3 * n, off).reshape(n, 3); off += 3 * n
        want = oracle.reconstruct_frame(view, f)
        assert n == want["point_count"]
        assert np.array_equal(pos, want["positions"]) and np.array_equal(col, want["colors"])
    assert off == len(raw)
