"""tmc2-rs_b200 -- B200-native V-PCC rec0 reconstruction (the hot path of benclmnt/tmc2-rs).

Only what the path needs: ``csrc/`` (sm_100a CUDA kernels + the C ABI ``libtmc2gpu.so``), ``abi`` (ctypes mirror of
``include/tmc2gpu.h``), ``codec`` (host-side mirror of the reference's ``src/codec.rs`` interface, calling the C ABI),
``decoder`` (streaming GOF driver mirroring ``src/decoder.rs:188-314`` / ``src/lib.rs``), ``synth`` (seeded synthetic
planes at the codec boundary), ``build`` (nvcc recipe).  There is no CPU fallback: importing ``codec`` without the
built CUDA library raises.
"""
from . import abi, synth  # noqa: F401

__all__ = ["abi", "synth", "build", "codec", "decoder"]


def __getattr__(name):
    if name in ("codec", "decoder", "build", "_lib"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
