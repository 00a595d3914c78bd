"""N4: bulk synthesis of concealment slices (libh264_b200/csrc/conceal.cuh) against the
reference's h264_write_grey_i_slice / h264_write_skipped_p_slice (src/h264_writer.c:49-219):
the reference library writes every slice (random SPS / PPS / slice headers, CAVLC and CABAC,
grey I and skipped P / B); the kernel gets the header bits + parameters and must produce the same
NAL bytes.  CPU: the kernel source on the emulator + the oracle's EPB insert; GPU: the C-ABI
(h264gpu_conceal_slices_host = synthesis + the K3 framing kernel)."""
import ctypes as C

import numpy as np
import pytest

import libh264_b200 as L
import support as S
import test_host_lib as T

pytestmark = T.pytestmark

CONCEAL_REC = np.dtype([("hdr_off", "<u8"), ("hdr_bits", "<u4"), ("mb_count", "<u4"), ("first_mb_in_slice", "<u4"),
                        ("pic_width_in_mbs", "<u2"), ("kind", "u1"), ("entropy_coding_mode_flag", "u1"),
                        ("slice_type", "u1"), ("cabac_init_idc", "u1"), ("slice_qp", "i1"), ("reserved", "u1", 5),
                        ("ref_off", "<u8"), ("ref_len", "<u8")])
CONCEAL_PARAMS = np.dtype([(n, CONCEAL_REC.fields[n][0]) for n in CONCEAL_REC.names[:11]])
assert CONCEAL_PARAMS.itemsize == 32 and CONCEAL_REC.itemsize == 48


def reference_slices(seed, rounds=12):
    """(stream, recs, hdr): concealment slices written by the REFERENCE library."""
    lib = T.harness()
    lib.hh_gen_conceal.restype = C.c_long
    lib.hh_gen_conceal.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                   C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t]
    out = np.zeros(8 << 20, np.uint8)
    recs = np.zeros(256, CONCEAL_REC)
    hdr = np.zeros(1 << 16, np.uint8)
    n = C.c_size_t(0)
    used = lib.hh_gen_conceal(T.REF.encode(), seed, rounds, out.ctypes.data, len(out), recs.ctypes.data, len(recs),
                              C.byref(n), hdr.ctypes.data, len(hdr))
    assert used > 0 and n.value > 0, (used, n.value)
    return out[:used].copy(), recs[:n.value].copy(), hdr


def params_of(recs):
    p = np.zeros(len(recs), CONCEAL_PARAMS)
    for name in CONCEAL_PARAMS.names:
        p[name] = recs[name]
    return p


def expected(stream, recs):
    return [stream[int(r["ref_off"]):int(r["ref_off"] + r["ref_len"])] for r in recs]


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_emu_concealment_slices_match_the_reference_writer(seed):
    stream, recs, hdr = reference_slices(seed)
    p = params_of(recs)
    lib = S.emu()
    lib.emu_conceal.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    cap = int((recs["hdr_bits"] // 8 + 2 * recs["mb_count"] + 16).sum())
    out = np.full(cap, 0xEE, np.uint8)
    off = np.zeros(len(p) + 1, np.uint64)
    lib.emu_conceal(p.ctypes.data, len(p), hdr.ctypes.data, out.ctypes.data, cap, off.ctypes.data)
    assert int(off[-1]) <= cap
    # payloads are unescaped: the oracle's EPB insert + 4-byte start code gives the NAL bytes
    framed, oo = S.oracle_frame(out[:int(off[-1])], off, 4)
    kinds = set()
    for k, want in enumerate(expected(stream, recs)):
        got = framed[int(oo[k]):int(oo[k + 1])]
        assert np.array_equal(got, want), (seed, k, recs[k])
        kinds.add((int(recs["kind"][k]), int(recs["entropy_coding_mode_flag"][k])))
    if seed == 6:
        return
    assert len(kinds) >= 2


def test_every_kind_is_covered_by_the_seeds():
    kinds = set()
    for seed in (1, 2, 3, 4, 5, 6):
        _, recs, _ = reference_slices(seed)
        kinds |= {(int(a), int(b)) for a, b in zip(recs["kind"], recs["entropy_coding_mode_flag"])}
    assert kinds == {(0, 0), (0, 1), (1, 0), (1, 1)}


@pytest.mark.gpu
def test_gpu_concealment_slices_match_the_reference_writer(gpu):
    total_slices = 0
    for seed in (11, 12, 13, 14):
        stream, recs, hdr = reference_slices(seed, rounds=20)
        p = params_of(recs)
        out = np.zeros(int(stream.nbytes) + 4096, np.uint8)
        oo = np.zeros(len(p) + 1, np.uint64)
        tot = C.c_uint64(0)
        L._check(gpu.lib.h264gpu_conceal_slices_host(gpu.h, C.c_void_p(p.ctypes.data), C.c_uint32(len(p)),
                                                     C.c_void_p(hdr.ctypes.data), C.c_uint64(len(hdr)), C.c_int(4),
                                                     C.c_void_p(out.ctypes.data), C.c_uint64(len(out)),
                                                     C.c_void_p(oo.ctypes.data), C.byref(tot)),
                 "h264gpu_conceal_slices_host")
        assert int(oo[-1]) == tot.value == int(recs["ref_len"].sum())
        for k, want in enumerate(expected(stream, recs)):
            assert np.array_equal(out[int(oo[k]):int(oo[k + 1])], want), (seed, k, recs[k])
        total_slices += len(p)
    assert total_slices > 50
