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
    lib.hh_rewrite.restype = C.c_long
    lib.hh_rewrite.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.hh_rewrite_patches.restype = C.c_long
    lib.hh_rewrite_patches.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32,
                                       C.POINTER(C.c_uint32)]
    lib.hh_time_parse.restype = C.c_double
    lib.hh_time_parse.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_void_p]
    return lib


def time_parse(lib, path, stream, flags=1, reps=3):
    """Best wall time of h264_reader_parse(stream) through `path`'s public API with counting
    callbacks, and the counts [nalu, slices, macroblocks, sps + pps, macroblock checksum]."""
    stream = np.ascontiguousarray(stream)
    counts = np.zeros(5, np.uint64)
    dt = lib.hh_time_parse(path.encode(), stream.ctypes.data, len(stream), flags, reps, counts.ctypes.data)
    assert dt > 0, (path, dt)
    return dt, counts


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


def test_abi_struct_sizes_and_exports():
    """x86-64 sizes measured on the reference and the 64 exported functions
    (Appendix C)."""
    lib = harness()
    out = (C.c_uint32 * 32)()
    n = lib.hh_sizes(out, 32)
    assert list(out[:n]) == [48, 12, 4184, 3276, 952, 412, 2064, 3968, 520, 2056, 1292, 4, 544, 88, 92, 144]
    names = """h264_reader_new h264_reader_destroy h264_reader_get_ctx h264_reader_stop h264_reader_parse
    h264_reader_parse_nalu h264_parse_nalu_header h264_parse_sps h264_parse_pps h264_find_nalu h264_bs_write_bits
    h264_bs_read_bits_ue h264_bs_write_bits_ue h264_bs_read_bits_ff_coded h264_bs_write_bits_ff_coded
    h264_bs_more_rbsp_data h264_bs_next_bits h264_bs_read_rbsp_trailing_bits h264_bs_write_rbsp_trailing_bits
    h264_bs_read_raw_bytes h264_bs_write_raw_bytes h264_bs_acquire_buf h264_ctx_new h264_ctx_destroy h264_ctx_clear
    h264_ctx_clear_nalu h264_ctx_set_nalu_header h264_ctx_is_nalu_unknown h264_ctx_set_aud h264_ctx_set_sps
    h264_ctx_set_pps h264_ctx_set_filler h264_ctx_get_sps h264_ctx_get_pps h264_ctx_add_sei h264_ctx_get_sei_count
    h264_ctx_sei_pic_timing_to_ts h264_ctx_sei_pic_timing_to_us h264_ctx_set_slice_header h264_ctx_get_info
    h264_write_nalu h264_write_grey_i_slice h264_write_skipped_p_slice h264_rewrite_slice_header
    h264_get_sps_derived h264_get_info h264_sar_to_aspect_ratio_idc h264_byte_stream_to_avcc
    h264_avcc_to_byte_stream h264_nalu_type_str h264_slice_type_str h264_mb_type_str h264_mb_type_is_intra
    h264_mb_type_is_inter h264_profile_str h264_color_format_str h264_aspect_ratio_str_alloc h264_sei_type_str
    h264_dump_new h264_dump_destroy h264_dump_clear h264_dump_get_json_object h264_dump_get_json_str
    h264_dump_nalu""".split()
    assert len(names) == 64
    ours = C.CDLL(OURS)
    for nm in names:
        assert hasattr(ours, nm), nm


def test_small_utilities_match_reference():
    """find_nalu, Exp-Golomb, AVCC conversion, enum strings, info: same answers from both libraries."""
    ours, ref = C.CDLL(OURS), C.CDLL(REF)
    rng = np.random.default_rng(12)
    alpha = np.array([0, 0, 0, 1, 2, 3, 0xFF, 0x65], np.uint8)
    for lib_ in (ours, ref):
        lib_.h264_find_nalu.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    for it in range(300):
        b = rng.choice(alpha, int(rng.integers(1, 60)))
        got = []
        for lib_ in (ours, ref):
            s, e = C.c_size_t(7777), C.c_size_t(7777)
            rc = lib_.h264_find_nalu(b.ctypes.data, len(b), C.byref(s), C.byref(e))
            got.append((rc, s.value if rc in (0, -11) else None, e.value if rc in (0, -11) else None))
        assert got[0] == got[1], (bytes(b).hex(), got)
    for fn, rng_ in (("h264_nalu_type_str", range(-1, 40)), ("h264_slice_type_str", range(-2, 8)),
                     ("h264_mb_type_str", range(-1, 20)), ("h264_profile_str", range(0, 256)),
                     ("h264_color_format_str", range(-1, 6)), ("h264_sei_type_str", range(-1, 60))):
        for lib_ in (ours, ref):
            getattr(lib_, fn).restype = C.c_char_p
        for v in rng_:
            assert getattr(ours, fn)(v) == getattr(ref, fn)(v), (fn, v)
    for v in range(0, 24):
        assert ours.h264_mb_type_is_intra(v) == ref.h264_mb_type_is_intra(v)
        assert ours.h264_mb_type_is_inter(v) == ref.h264_mb_type_is_inter(v)
    for w, h in ((1, 1), (12, 11), (4, 3), (5, 7), (160, 99), (2, 1)):
        assert ours.h264_sar_to_aspect_ratio_idc(w, h) == ref.h264_sar_to_aspect_ratio_idc(w, h)
    # Annex-B <-> AVCC in place
    hl = harness()
    base = gen(hl, REF, 5, rounds=3)
    a, b = base.copy(), base.copy()
    assert ours.h264_byte_stream_to_avcc(a.ctypes.data, len(a)) == ref.h264_byte_stream_to_avcc(b.ctypes.data, len(b))
    assert np.array_equal(a, b) and not np.array_equal(a, base)
    assert ours.h264_avcc_to_byte_stream(a.ctypes.data, len(a)) == ref.h264_avcc_to_byte_stream(b.ctypes.data, len(b))
    assert np.array_equal(a, b) and np.array_equal(a, base)


PATCH_DT = np.dtype([("nal_off", "<u8"), ("nbytes", "<u4"), ("tail_bits", "<u4"), ("bytes", "u1", 64)])


def rewrite_patches(lib, stream, mode):
    """Patch records of this library's h264_rewrite_slice_header_patch for every slice of the stream
    (same edits as hh_rewrite), and the per-slice return codes."""
    stream = np.ascontiguousarray(stream)
    patches = np.zeros(4096, PATCH_DT)
    rc = np.full(4096, 99, np.int32)
    npatch = C.c_uint32(0)
    n = lib.hh_rewrite_patches(OURS.encode(), stream.ctypes.data, len(stream), mode, patches.ctypes.data,
                               rc.ctypes.data, len(rc), C.byref(npatch))
    assert n >= 0
    return patches[:npatch.value].copy(), rc[:n].copy()


def rewrite_streams(lib):
    import libh264_b200 as L
    streams = [gen(lib, REF, 400 + k, rounds=6) for k in range(3)]
    streams.append(L.synth_video(frames=4, width_mbs=20, height_mbs=12, slices_per_frame=3, profile_idc=100,
                                 transform_8x8=1, b_frames=1, num_ref_frames=2, idr_period=3, pct_skip=30,
                                 coef_density=50, seed=21)[0])
    return [np.ascontiguousarray(s) for s in streams]


def test_rewrite_slice_header_patches_in_bulk():
    """N4: h264_rewrite_slice_header_patch + the bulk patch kernel (here on the CPU emulator) give the
    bytes of the REFERENCE's h264_rewrite_slice_header applied NAL by NAL, with its return codes."""
    import support as S
    lib = harness()
    emu = S.emu()
    applied = 0
    for s in rewrite_streams(lib):
        for mode in (0, 1):
            ref_out = np.zeros(len(s), np.uint8)
            ref_rc = np.full(4096, 99, np.int32)
            n = lib.hh_rewrite(REF.encode(), s.ctypes.data, len(s), mode, ref_out.ctypes.data, ref_rc.ctypes.data, len(ref_rc))
            patches, rc = rewrite_patches(lib, s, mode)
            assert np.array_equal(rc, ref_rc[:n])
            assert len(patches) == int((rc == 0).sum())
            out = s.copy()
            assert emu.emu_patch_headers(C.c_void_p(out.ctypes.data), C.c_uint64(len(out)),
                                         C.c_void_p(patches.ctypes.data), C.c_uint32(len(patches))) == 0
            assert np.array_equal(out, ref_out), mode
            applied += len(patches)
    assert applied > 10
    # a patch that does not fit the stream is left out
    s = rewrite_streams(lib)[0]
    patches, _ = rewrite_patches(lib, s, 0)
    bad = patches[:1].copy()
    bad["nal_off"] = len(s) - 1
    out = s.copy()
    emu.emu_patch_headers(C.c_void_p(out.ctypes.data), C.c_uint64(len(out)), C.c_void_p(bad.ctypes.data), C.c_uint32(1))
    assert np.array_equal(out, s)


def test_rewrite_slice_header_behaviour():
    """h264_rewrite_slice_header (src/h264_writer.c:311-370) on every slice of generated streams:
    an edit that keeps the bit length patches the NAL in place exactly like the reference does
    (whole header bytes + the bits of the byte shared with the slice data), one that changes the
    length fails with -EPROTO and leaves the bytes alone."""
    lib = harness()
    import libh264_b200 as L
    streams = [gen(lib, REF, 400 + k, rounds=6) for k in range(3)]
    streams.append(L.synth_video(frames=4, width_mbs=20, height_mbs=12, slices_per_frame=3, profile_idc=100,
                                 transform_8x8=1, b_frames=1, num_ref_frames=2, idr_period=3, pct_skip=30,
                                 coef_density=50, seed=21)[0])
    patched = 0
    for s in streams:
        s = np.ascontiguousarray(s)
        for mode in (0, 1):
            outs, rcs = [], []
            for path in (OURS, REF):
                out = np.zeros(len(s), np.uint8)
                rc = np.full(4096, 99, np.int32)
                n = lib.hh_rewrite(path.encode(), s.ctypes.data, len(s), mode, out.ctypes.data, rc.ctypes.data, len(rc))
                assert n >= 0
                outs.append(out)
                rcs.append(rc[:n].copy())
            assert np.array_equal(rcs[0], rcs[1]), (mode, rcs)
            assert np.array_equal(outs[0], outs[1]), mode
            if mode == 0:
                ok = rcs[0] == 0
                patched += int(ok.sum())
                if ok.any():
                    assert not np.array_equal(outs[0], s)  # frame_num really changed in the bytes
            else:
                changed = rcs[0] == 0  # slice_qp_delta + 17 may keep its length for some values
                if not changed.any():
                    assert np.array_equal(outs[0], s)
                assert ((rcs[0] == 0) | (rcs[0] == -71)).all(), rcs[0]  # -EPROTO
    assert patched > 10


@pytest.mark.parametrize("map_type", range(7))
def test_slice_group_map_matches_the_reference(map_type):
    """Row A12: the macroblock -> slice group map the host library builds for every slice
    (h264_ctx_get_slice_group_map, libh264_b200/host/h264_fmo.c) is the one the reference holds in
    ctx->slice.group_map (h264_gen_slice_group_map, src/h264_fmo.c:244-291) for the same slice: all
    seven slice_group_map_type values, maps that move with slice_group_change_cycle (types 3-5)."""
    import libh264_b200 as L
    lib = harness()
    lib.hh_group_maps.restype = C.c_int
    lib.hh_group_maps.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    for groups, seed in ((2, 3), (3, 4), (5, 9)):
        stream, nmb, nsl = L.synth_video(width_mbs=13, height_mbs=7, frames=9, slices_per_frame=1, idr_period=5,
                                         seed=seed, fmo=groups | map_type << 4)
        ev, mbs, off = S.ref_trace(stream)
        want = np.concatenate([np.frombuffer(bytes(p), np.uint8) for t, p in ev if t == S.TR_GROUP_MAP])
        assert len(want) == nsl * 91 and len(np.unique(want)) == (2 if 3 <= map_type <= 5 else groups)
        got = np.zeros(len(want) + 91, np.uint8)
        used = C.c_size_t(0)
        stream = np.ascontiguousarray(stream)
        rc = lib.hh_group_maps(OURS.encode(), stream.ctypes.data, len(stream), got.ctypes.data, len(got), C.byref(used))
        assert rc == 0 and used.value == len(want)
        assert np.array_equal(got[:used.value], want), (map_type, groups)
