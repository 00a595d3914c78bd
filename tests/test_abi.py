"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes as C
import os
import re

import libh264_b200 as L
from libh264_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, macro):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = "\n".join(l for l in txt.splitlines() if not l.lstrip().startswith("#"))
    return sorted(set(re.findall(macro + r"[^;{]*?\b(\w+)\s*\(", txt)))


def test_gpu_library_exports_header_symbols():
    if not os.path.exists(build.GPU_SO):
        build.build_gpu()
    lib = L.load_gpu_lib()
    names = declared("h264gpu.h", "H264GPU_API")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(L.GPU_SYMBOLS) == names
    assert b"sm_100a" in lib.h264gpu_version()
    slice_names = declared("h264gpu_slice.h", "H264GPU_API")
    assert sorted(L.SLICE_SYMBOLS) == slice_names
    for n in slice_names:
        assert hasattr(lib, n), "missing export " + n


def test_no_device_is_an_error_not_a_fallback():
    lib = L.load_gpu_lib()
    if lib.h264gpu_device_count() > 0:
        return  # on the GPU box the gpu-marked tests cover the live path
    h = C.c_void_p()
    assert lib.h264gpu_create(0, C.byref(h)) < 0
    assert not h.value
