"""The C-ABI library loads (no GPU needed), exports every symbol include/tmc2gpu.h declares, and the ctypes mirror has
the same struct layout as the C compiler's."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

import tmc2rs_b200  # noqa: F401
from tmc2rs_b200 import _lib, abi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tmc2gpu.h")


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"TMC2_API\s+[\w\s\*]+?\b(tmc2gpu_\w+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tmc2gpu.h but not exported by libtmc2gpu.so"
    assert sorted(s[0] for s in _lib.SYMBOLS) == names, "ctypes symbol table out of sync with the header"


def test_abi_version_and_no_device_behaviour(lib):
    assert lib.tmc2gpu_abi_version() == abi.ABI_VERSION
    assert lib.tmc2gpu_status_string(abi.ERR_NO_DEVICE) == b"TMC2_ERR_NO_DEVICE"
    if lib.tmc2gpu_device_count() == 0:
        # the product path must fail loudly without a GPU: no CPU fallback
        h = C.c_void_p()
        assert lib.tmc2gpu_create(None, 0, None, C.byref(h)) == abi.ERR_NO_DEVICE
        assert not h.value
        from tmc2rs_b200 import codec
        with pytest.raises(abi.Tmc2Error) as e:
            codec.Context()
        assert e.value.status == abi.ERR_NO_DEVICE


def test_struct_layout_matches_c(tmp_path):
    src = tmp_path / "layout.c"
    structs = {"tmc2_patch": abi.CPatch, "tmc2_params": abi.CParams, "tmc2_frame": abi.CFrame, "tmc2_gof": abi.CGof,
               "tmc2_frame_out": abi.CFrameOut, "tmc2_limits": abi.CLimits, "tmc2_point_cloud_out": abi.CPointCloudOut}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_product_path_never_touches_oracle():
    """The oracle is test infrastructure: nothing under tmc2-rs_b200/ may import, link or load it."""
    pkg = os.path.join(ROOT, "tmc2-rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "tmc2_oracle" not in text, f
                if f.endswith(".py"):
                    assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f


def test_rust_ffi_crate_mirrors_header():
    """integration/tmc2gpu-sys cannot be compiled here (no cargo/rustc): keep its #[repr(C)] structs textually in step with
    the ctypes mirror (which test_struct_layout_matches_c pins to the C header) -- same field names in the same order --
    and make sure every extern fn it declares exists in the header."""
    text = open(os.path.join(ROOT, "integration", "tmc2gpu-sys", "src", "lib.rs")).read()
    for rust_name, cls in {"tmc2_patch": abi.CPatch, "tmc2_params": abi.CParams, "tmc2_frame": abi.CFrame,
                           "tmc2_gof": abi.CGof, "tmc2_frame_out": abi.CFrameOut, "tmc2_limits": abi.CLimits}.items():
        m = re.search(r"pub struct %s \{(.*?)\n\}" % rust_name, text, re.S)
        assert m, rust_name
        fields = re.findall(r"pub (\w+):", m.group(1))
        assert fields == [f for f, _ in cls._fields_], rust_name
    header = open(HEADER).read()
    for fn in re.findall(r"pub fn (tmc2gpu_\w+)\(", text):
        assert re.search(r"\b%s\(" % fn, header), fn
