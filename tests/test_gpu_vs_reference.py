"""GPU stages against the COMPILED REFERENCE itself (oracle/_ref/libh264_ref.so: the unmodified
libh264 sources), not against the restatement: the h264_find_nalu loop (src/h264_reader.c:133-140,
src/h264_bitstream.c:159-184), h264_bs_read_bits(8) per NAL (include/h264/h264_bitstream.h:168-218)
and the h264_bs_write_bits(8) writer (src/h264_bitstream.c:211-239, 54-81), on a 64 MiB sample of
the BASELINE config-2 / config-5 workloads.  The built reference travels to the GPU box."""
import numpy as np
import pytest

import libh264_b200 as L
import support as S

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not S.have_ref(), reason="oracle/_ref not built")]


@pytest.fixture(scope="module")
def gpu():
    g = L.Gpu(0)
    yield g
    g.close()


@pytest.fixture(scope="module")
def sample():
    seed = 0x264
    offs = L.synth_offsets(seed, 64 << 20)
    rbsp = L.synth_payloads(seed, offs)
    stream, _ = L.synth_annexb(seed, rbsp, offs)
    return stream, rbsp, offs


def test_inplace_scan_strip_against_reference_find_nalu_and_read_bits(gpu, sample):
    stream, rbsp, offs = sample
    rs, re_, roff = S.ref_scan(stream)
    got = gpu.split_strip_inplace(stream, cap=len(rs) + 64)
    assert got["res"].n_nal == len(rs)
    assert np.array_equal(got["start"], rs) and np.array_equal(got["end"], re_)
    total = 0
    for k in range(len(rs)):
        want, _ = S.ref_strip(stream[int(rs[k]):int(re_[k])])
        assert int(got["rbsp_len"][k]) == len(want), k
        assert np.array_equal(got["nal_rbsp"][k], want), k
        total += len(want)
    assert got["res"].rbsp_bytes == total
    # scan only (the launch of h264_reader_parse) and the reader session: same table, same *off
    only = gpu.split_strip_inplace(stream, want_rbsp=False, cap=len(rs) + 64)
    assert np.array_equal(only["start"], rs) and np.array_equal(only["end"], re_)
    import ctypes as C
    st, en = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint64)()
    n, off = C.c_uint64(0), C.c_uint64(0)
    gpu.lib.h264gpu_reader_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p]
    L._check(gpu.lib.h264gpu_reader_scan(gpu.h, stream.ctypes.data, len(stream), C.byref(st), C.byref(en),
                                         C.byref(n), C.byref(off)), "h264gpu_reader_scan")
    assert n.value == len(rs) and off.value == roff
    assert np.array_equal(np.ctypeslib.as_array(st, (n.value,)), rs)
    assert np.array_equal(np.ctypeslib.as_array(en, (n.value,)), re_)


def test_packed_host_pipeline_against_reference(gpu, sample):
    stream, rbsp, offs = sample
    rs, re_, roff = S.ref_scan(stream)
    got = gpu.split_strip_host(stream)
    assert got["n_nal"] == len(rs) and got["final_off"] == roff
    assert np.array_equal(got["start"], rs) and np.array_equal(got["end"], re_)
    for k in (0, 1, len(rs) // 2, len(rs) - 2, len(rs) - 1):
        want, _ = S.ref_strip(stream[int(rs[k]):int(re_[k])])
        o = int(got["rbsp_off"][k])
        assert np.array_equal(got["rbsp"][o:o + len(want)], want), k


def test_writer_against_reference_writer_and_through_reference_reader(gpu, sample):
    """The GPU writer's NALs are what the reference's bit writer emits for the same payloads, and
    the REFERENCE reader (find_nalu loop + read_bits) returns the payloads from the GPU's stream."""
    _, rbsp, offs = sample
    out, oo = gpu.frame_host(rbsp, offs, sc_len=4)
    n = len(offs) - 1
    for k in list(range(0, n, max(1, n // 64))) + [n - 1]:
        a, b = int(offs[k]), int(offs[k + 1])
        want = S.ref_insert(rbsp[a:b])
        o = int(oo[k])
        assert tuple(out[o:o + 4]) == (0, 0, 0, 1), k
        assert np.array_equal(out[o + 4:int(oo[k + 1])], want), k
    rs, re_, roff = S.ref_scan(out)
    assert len(rs) == n and roff == len(out)
    assert np.array_equal(rs, oo[:-1] + 4) and np.array_equal(re_, oo[1:])
    for k in range(n):
        back, _ = S.ref_strip(out[int(rs[k]):int(re_[k])])
        assert np.array_equal(back, rbsp[int(offs[k]):int(offs[k + 1])]), k
