"""The host library (libh264_b200/libh264.so: reader, writer, context, bitstream) against the
compiled reference (oracle/_ref/libh264_ref.so) through the PUBLIC API only — the same
harness binary (tests/host/host_harness.c) dlopens either library.  CPU-only: header syntax,
callback order, writer bytes.  Slice data (GPU) is covered in test_host_lib_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import support as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "libh264_b200", "libh264.so")
REF = os.path.join(ROOT, "oracle", "_ref", "libh264_ref.so")
HARNESS_SRC = os.path.join(ROOT, "tests", "host", "host_harness.c")
HARNESS_SO = os.path.join(ROOT, "tests", "host", "libhost_harness.so")

pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="compiled reference not present")


def harness():
    if not os.path.exists(HARNESS_SO) or os.path.getmtime(HARNESS_SO) < os.path.getmtime(HARNESS_SRC):
        subprocess.check_call(["gcc", "-O1", "-std=gnu99", "-fPIC", "-shared", "-Wall",
                               "-I" + os.path.join(ROOT, "include"), "-o", HARNESS_SO, HARNESS_SRC, "-ldl"])
    import libh264_b200.build as B
    B.build_all()
    lib = C.CDLL(HARNESS_SO)
    lib.hh_gen.restype = C.c_long
    lib.hh_gen.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    lib.hh_trace.restype = C.c_int
    lib.hh_trace.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_void_p,
                             C.c_size_t, C.POINTER(C.c_size_t)]
    return lib


def gen(lib, path, seed, rounds=6, conceal=0):
    out = np.zeros(1 << 20, np.uint8)
    n = lib.hh_gen(path.encode(), seed, rounds, conceal, out.ctypes.data, len(out))
    assert n > 0, (path, seed, n)
    return out[:n].copy()


def trace(lib, path, stream, flags=0, mode=1):
    log = np.zeros(64 << 20, np.uint8)
    used = C.c_size_t(0)
    stream = np.ascontiguousarray(stream)
    rc = lib.hh_trace(path.encode(), stream.ctypes.data, len(stream), flags, mode, log.ctypes.data,
                      len(log), C.byref(used))
    assert rc == 0, (path, rc)
    return log[:used.value].copy()


def split_log(log):
    out, p = [], 0
    b = log.tobytes()
    while p < len(b):
        tag, n = np.frombuffer(b, "<u4", 2, p)
        out.append((int(tag), b[p + 8:p + 8 + int(n)]))
        p += 8 + int(n)
    return out


def assert_same_trace(a, b, what):
    if a.tobytes() == b.tobytes():
        return
    ea, eb = split_log(a), split_log(b)
    for i, (x, y) in enumerate(zip(ea, eb)):
        if x != y:
            d = next((k for k in range(min(len(x[1]), len(y[1]))) if x[1][k] != y[1][k]), -1)
            raise AssertionError("%s: event %d differs: tag %d/%d len %d/%d first byte %d" %
                                 (what, i, x[0], y[0], len(x[1]), len(y[1]), d))
    raise AssertionError("%s: %d vs %d events" % (what, len(ea), len(eb)))


@pytest.mark.parametrize("seed", range(12))
def test_writer_bytes_identical(seed):
    """SPS/PPS/AUD/SEI/filler/slice-header NAL units written from the same seeded structs."""
    lib = harness()
    assert np.array_equal(gen(lib, OURS, seed), gen(lib, REF, seed))


@pytest.mark.parametrize("seed", range(12))
def test_reader_trace_identical(seed):
    """Every callback, in order, with the parsed structs byte for byte (per-NAL entry point)."""
    lib = harness()
    stream = gen(lib, REF, 100 + seed, rounds=8)
    assert_same_trace(trace(lib, OURS, stream), trace(lib, REF, stream), "seed %d" % seed)


@pytest.mark.parametrize("seed", range(6))
def test_concealment_slices_identical(seed):
    """h264_write_grey_i_slice / h264_write_skipped_p_slice, CAVLC and CABAC."""
    lib = harness()
    a, b = gen(lib, OURS, 200 + seed, rounds=6, conceal=1), gen(lib, REF, 200 + seed, rounds=6, conceal=1)
    assert np.array_equal(a, b)
    assert_same_trace(trace(lib, OURS, a), trace(lib, REF, a), "conceal %d" % seed)


def test_reader_on_damaged_streams():
    """Truncated and bit-flipped NAL units: same errors, same callbacks up to the error."""
    lib = harness()
    rng = np.random.default_rng(5)
    base = gen(lib, REF, 77, rounds=5)
    for it in range(30):
        s = base.copy()
        if it % 2:
            s = s[:int(rng.integers(8, len(s)))]
        else:
            for p in rng.integers(5, len(s), 6):
                s[p] ^= 1 << int(rng.integers(0, 8))
        assert_same_trace(trace(lib, OURS, s), trace(lib, REF, s), "damaged %d" % it)


def test_synthetic_video_headers():
    """The CAVLC workload generator's streams: SPS/PPS/slice headers and AU boundaries."""
    import libh264_b200 as L
    lib = harness()
    stream, _, _ = L.synth_video(frames=6, width_mbs=20, height_mbs=12, slices_per_frame=3, profile_idc=100,
                                 transform_8x8=1, b_frames=1, num_ref_frames=2, idr_period=4, pct_skip=20,
                                 coef_density=40, seed=3)
    assert_same_trace(trace(lib, OURS, stream), trace(lib, REF, stream), "synth video")
