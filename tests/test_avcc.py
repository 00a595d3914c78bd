"""N1: Annex-B (4-byte start codes) -> AVCC lengths in place (h264_byte_stream_to_avcc,
src/h264.c:184-246).  The compiled reference is the checker: CPU tests run the kernel sources
on the SIMT emulator, the GPU test goes through the C-ABI (h264gpu_byte_stream_to_avcc_host)."""
import ctypes as C

import numpy as np
import pytest

import libh264_b200 as L
import support as S

pytestmark = pytest.mark.skipif(not S.have_ref(), reason="oracle/_ref not built")


def ref_avcc(buf):
    out = np.ascontiguousarray(buf, dtype=np.uint8).copy()
    lib = S.ref()
    lib.h264_byte_stream_to_avcc.restype = C.c_int
    lib.h264_byte_stream_to_avcc.argtypes = [C.c_void_p, C.c_size_t]
    rc = lib.h264_byte_stream_to_avcc(out.ctypes.data, len(out))
    assert rc == 0
    return out


def emu_avcc(buf, rows):
    lib = S.emu()
    lib.emu_avcc.restype = C.c_int
    lib.emu_avcc.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
    out = np.ascontiguousarray(buf, dtype=np.uint8).copy()
    cap = len(out) // 4 + 2
    pos = np.zeros(cap, np.uint64)
    n = C.c_uint64(0)
    assert lib.emu_avcc(out.ctypes.data, len(out), cap, pos.ctypes.data, C.byref(n), rows) == 0
    return out, pos[:n.value]


def streams(rng):
    for it in range(10):
        n = int(rng.integers(1, 40))
        payloads = S.gen_payloads(rng, n, lo=1, hi=int(rng.choice([40, 700, 6000])))
        esc = [S.oracle_insert(p) for p in payloads]
        parts = []
        for e in esc:
            parts += [np.array([0, 0, 0, 1], np.uint8), e]
        if it % 3 == 0:
            parts.append(np.array([0, 0, 0, 1][:int(rng.integers(1, 5))], np.uint8))  # (partial) code at the end
        lead = rng.integers(1, 255, int(rng.integers(0, 9))).astype(np.uint8)
        yield np.concatenate([lead] + parts)
    alpha = np.array([0, 0, 0, 0, 1, 1, 3, 9], np.uint8)
    for it in range(10):
        yield rng.choice(alpha, int(rng.integers(1, 20000)))  # code-dense garbage, 3-byte codes, zero runs
    yield np.tile(np.array([0, 0, 0, 1], np.uint8), 300)
    yield np.array([7, 7, 7], np.uint8)


@pytest.mark.parametrize("rows", [1, 2, 8])
def test_emu_avcc_matches_reference(rows):
    rng = np.random.default_rng(40 + rows)
    for k, b in enumerate(streams(rng)):
        got, pos = emu_avcc(b, rows)
        assert np.array_equal(got, ref_avcc(b)), (rows, k)
        assert (np.diff(pos.astype(np.int64)) >= 4).all()


@pytest.mark.gpu
def test_gpu_avcc_matches_reference():
    rng = np.random.default_rng(77)
    g = L.Gpu(0)
    try:
        g.lib.h264gpu_byte_stream_to_avcc_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        cases = list(streams(rng))
        offs = L.synth_offsets(9, 48 << 20)
        rbsp = L.synth_payloads(9, offs)
        big, _ = L.synth_annexb(9, rbsp, offs, mixed_sc=False, trailing=False)  # 4-byte codes, config-5 shape
        mixed, _ = L.synth_annexb(9, rbsp, offs)                                # 3- and 4-byte codes mixed
        for k, b in enumerate(cases + [big, mixed]):
            out = np.ascontiguousarray(b).copy()
            n = C.c_uint64(0)
            L._check(g.lib.h264gpu_byte_stream_to_avcc_host(g.h, out.ctypes.data, len(out), C.byref(n)), "avcc")
            assert np.array_equal(out, ref_avcc(b)), k
        assert n.value > 0
    finally:
        g.close()
