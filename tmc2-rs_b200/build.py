"""nvcc recipe for ``libtmc2gpu.so`` (sm_100a only; in-tree so the built library travels with gpurun snapshots)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtmc2gpu.so")
SOURCES = ["kernels.cu", "ply.cu", "tmc2gpu.cu"]
HEADERS = ["device_types.h", os.path.join("..", "..", "include", "tmc2gpu.h")]

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "--shared", "-cudart", "static"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")
    return exe


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    files = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(f) and os.path.getmtime(f) > t for f in files)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    if not force and not stale() and out == LIB:
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in defines] + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building libtmc2gpu.so")
    return out


EXAMPLE_SRC = os.path.join(os.path.dirname(HERE), "examples", "stream_decode.cpp")
EXAMPLE_BIN = os.path.join(os.path.dirname(HERE), "examples", "stream_decode")


def build_example(force: bool = False) -> str:
    """The plain C++ streaming driver (examples/stream_decode.cpp): g++ only, links the C ABI library, no CUDA headers."""
    if not force and os.path.exists(EXAMPLE_BIN) and os.path.getmtime(EXAMPLE_BIN) >= max(
            os.path.getmtime(EXAMPLE_SRC), os.path.getmtime(LIB) if os.path.exists(LIB) else 0):
        return EXAMPLE_BIN
    cxx = shutil.which("g++") or "g++"
    cmd = [cxx, "-std=c++17", "-O2", "-pthread", EXAMPLE_SRC, "-o", EXAMPLE_BIN, "-L" + HERE, "-l:libtmc2gpu.so",
           "-Wl,-rpath," + HERE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building examples/stream_decode")
    return EXAMPLE_BIN


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
