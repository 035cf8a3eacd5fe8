"""tmc2-rs_b200 -- B200-native V-PCC rec0 reconstruction (the hot path of benclmnt/tmc2-rs).

Only what the path needs: ``csrc/`` (sm_100a CUDA kernels + the C ABI ``libtmc2gpu.so``), ``abi`` (ctypes mirror of
``include/tmc2gpu.h``), ``codec`` (host-side mirror of the reference's ``src/codec.rs`` interface and of the frame loop
``src/decoder.rs:188-314``, calling the C ABI), ``synth`` (seeded synthetic planes at the codec boundary), ``shard``
(frame-wise sharding over ranks), ``ply`` (the reference's PLY consumer format, ``src/writer.rs``), ``build`` (nvcc recipe).  There is no CPU fallback: importing ``codec`` without the
built CUDA library raises.
"""
from . import abi, synth  # noqa: F401

__all__ = ["abi", "synth", "build", "codec", "shard", "ply"]


def __getattr__(name):
    if name in ("codec", "build", "_lib", "shard", "ply"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
