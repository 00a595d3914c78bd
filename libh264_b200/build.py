"""In-tree build recipes (explicit nvcc / gcc, no JIT cache): the shared
libraries land next to this file so they travel with the repo snapshot.

  libh264gpu.so    CUDA stages + C-ABI (include/h264gpu.h), sm_100a only
  libh264synth.so  synthetic workload generator (host C)
  libh264.so       libh264-compatible host library (C) on top of libh264gpu.so
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
INC = os.path.join(ROOT, "include")

GPU_SO = os.path.join(HERE, "libh264gpu.so")
SYNTH_SO = os.path.join(HERE, "libh264synth.so")
H264_SO = os.path.join(HERE, "libh264.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(d, exts):
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts)) \
        if os.path.isdir(d) else []


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return p if os.path.exists(p) else None


def build_gpu(force=False, verbose=False):
    cu = _sources(CSRC, (".cu",))
    deps = cu + _sources(CSRC, (".cuh", ".h")) + _sources(INC, (".h",))
    if not force and not _stale(GPU_SO, deps):
        return GPU_SO
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libh264gpu.so (no CPU fallback exists)")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-I" + INC, "-I" + CSRC, "-o", GPU_SO] + cu
    subprocess.check_call(cmd)
    return GPU_SO


def build_synth(force=False):
    """Workload generators: synth.c / synth_video.c (C) + synth_cabac.cpp (C++, shares the
    CABAC syntax walker header with the GPU kernel)."""
    csrcs = [os.path.join(CSRC, "synth.c"), os.path.join(CSRC, "synth_video.c")]
    cxx = os.path.join(CSRC, "synth_cabac.cpp")
    deps = csrcs + [cxx] + [os.path.join(CSRC, h) for h in
                            ("cavlc_luts.h", "cabac_syntax.h", "cabac_engine.h", "cabac_tables.h",
                             "mb_syntax.h")] + [os.path.join(INC, "h264gpu_slice.h")]
    if force or _stale(SYNTH_SO, deps):
        objdir = os.path.join(HERE, "_obj")
        os.makedirs(objdir, exist_ok=True)
        objs = []
        for src in csrcs:
            o = os.path.join(objdir, os.path.basename(src) + ".o")
            subprocess.check_call(["gcc", "-O2", "-std=gnu99", "-fPIC", "-pthread", "-Wall",
                                   "-I" + CSRC, "-c", "-o", o, src])
            objs.append(o)
        o = os.path.join(objdir, "synth_cabac.o")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-Wall", "-Wno-unused-function",
                               "-I" + CSRC, "-I" + INC, "-c", "-o", o, cxx])
        objs.append(o)
        subprocess.check_call(["g++", "-shared", "-pthread", "-o", SYNTH_SO] + objs + ["-lm"])
    return SYNTH_SO


def build_host(force=False):
    srcs = _sources(HOST, (".c",))
    if not srcs:
        return None
    deps = srcs + _sources(HOST, (".h",)) + _sources(INC, (".h",)) + \
        _sources(os.path.join(INC, "h264"), (".h",))
    if force or _stale(H264_SO, deps + [GPU_SO]):
        subprocess.check_call(["gcc", "-O2", "-std=gnu99", "-D_GNU_SOURCE", "-DH264_API_EXPORTS",
                               "-fPIC", "-fvisibility=hidden", "-shared", "-Wall",
                               "-I" + INC, "-I" + HOST, "-I" + CSRC, "-o", H264_SO] + srcs +
                              ["-L" + HERE, "-lh264gpu", "-Wl,-rpath,$ORIGIN"])
    return H264_SO


def build_all(force=False, verbose=False):
    build_gpu(force, verbose)
    build_synth(force)
    build_host(force)
