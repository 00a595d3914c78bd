"""K5 CABAC slice-data parse.

The reference has NO CABAC decoder, so parity is pinned in layers:

 1. engine: the generator's arithmetic encoder produces the same bytes as the reference's
    h264_bac_encode_bin/bypass/terminate (src/h264_bac.c:150-358) on random bin scripts with
    the reference's own context initialisation; the decoder inverts them.
 2. known answers: CABAC concealment slices written by the REFERENCE
    (h264_write_grey_i_slice / h264_write_skipped_p_slice, src/h264_writer.c:49-219) decode to
    N x I_16x16 (raw mb_type 3) / N x P_Skip.
 3. twin streams: a CAVLC stream is parsed by the REFERENCE, its per-macroblock syntax
    elements (ctx->mb) are re-coded as CABAC, and the CABAC parse must return the reference's
    own records (mb_addr, mb_type, syntax checksum) for every macroblock.
 4. generator round trip: what the CABAC generator intended is what the parse returns.

CPU tests run the walker on the host (libh264synth: checker only); GPU tests run the kernel
through the C-ABI and compare with the same expectations.
"""
import ctypes as C
import os

import numpy as np
import pytest

import libh264_b200 as L
import support as S

HERE = os.path.dirname(os.path.abspath(__file__))


def synth():
    lib = L.load_synth_lib()
    vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
    lib.synth_cabac_decode.restype = C.c_int
    lib.synth_cabac_decode.argtypes = [vp, u64, vp, u32, vp, u64, vp]
    lib.synth_cabac_ops.restype = u64
    lib.synth_cabac_ops.argtypes = [vp, u64, u32, C.c_int32, vp, u64]
    lib.synth_cabac_ops_decode.restype = C.c_int
    lib.synth_cabac_ops_decode.argtypes = [vp, u64, u32, C.c_int32, vp, u32, vp]
    lib.synth_cabac_transcode.restype = u64
    lib.synth_cabac_transcode.argtypes = [vp, vp, u32, vp, u64]
    lib.synth_cabac_record_sink.restype = None
    lib.synth_cabac_record_sink.argtypes = [vp, u64]
    lib.synth_cabac_records_written.restype = u64
    lib.synth_cabac_sizeof_mb_syntax.restype = u32
    return lib


def cpu_decode(stream, params, n_records):
    lib = synth()
    stream = np.ascontiguousarray(stream, np.uint8)
    params = np.ascontiguousarray(params, np.uint8)
    n = len(params) // S.PARAMS_SIZE
    recs = np.zeros(max(n_records, 1), S.MB_RECORD)
    res = np.zeros(max(n, 1), S.SLICE_RESULT)
    assert lib.synth_cabac_decode(S.ptr(stream), len(stream), S.ptr(params), n, S.ptr(recs), n_records,
                                  S.ptr(res)) == 0
    return recs[:n_records], res[:n]


def gen_cabac(**cfg):
    """CABAC stream + params + the records the generator intended."""
    lib = synth()
    # sizing pass first (no sink), then the real one with the sink armed
    _, nmb, nsl = L.synth_video(entropy_cabac=1, **cfg)
    want = np.zeros(nmb, S.MB_RECORD)
    lib.synth_cabac_record_sink(S.ptr(want), nmb)
    try:
        out = np.empty(64 << 20, np.uint8)
        stream, nmb2, nsl2, params = L.synth_video(entropy_cabac=1, want_params=True, out=out, **cfg)
        assert lib.synth_cabac_records_written() == nmb == nmb2
    finally:
        lib.synth_cabac_record_sink(None, 0)
    return stream.copy(), params, want


CFGS = [
    dict(width_mbs=20, height_mbs=12, frames=6, slices_per_frame=3, profile_idc=77, transform_8x8=0,
         b_frames=0, num_ref_frames=1, idr_period=3, pct_skip=30, coef_density=50, seed=21),
    dict(width_mbs=22, height_mbs=10, frames=8, slices_per_frame=4, profile_idc=100, transform_8x8=1,
         b_frames=1, num_ref_frames=3, idr_period=4, pct_skip=25, coef_density=70, pct_pcm=20, seed=22),
    dict(width_mbs=11, height_mbs=9, frames=5, slices_per_frame=1, profile_idc=100, transform_8x8=1,
         b_frames=1, num_ref_frames=2, idr_period=5, pct_skip=60, coef_density=100, pct_intra_in_inter=30,
         seed=23),
    dict(width_mbs=16, height_mbs=8, frames=4, slices_per_frame=2, profile_idc=100, chroma_format_idc=2,
         transform_8x8=1, b_frames=1, num_ref_frames=4, idr_period=2, pct_skip=10, coef_density=80, seed=24),
    dict(width_mbs=12, height_mbs=6, frames=3, slices_per_frame=2, profile_idc=100, chroma_format_idc=0,
         transform_8x8=1, b_frames=1, num_ref_frames=2, idr_period=3, pct_skip=20, coef_density=60, seed=25),
]


# ---- 1. engine ----------------------------------------------------------------------------

def bin_script(rng, n):
    kind = rng.choice([0, 0, 0, 0, 1, 2], n).astype(np.uint32)
    ctx = rng.integers(0, 460, n).astype(np.uint32)
    # skewed bins so that states move through the whole table; terminate bins stay 0 until the end
    b = (rng.random(n) < rng.choice([0.05, 0.3, 0.5, 0.9])).astype(np.uint32)
    b[kind == 2] = 0
    ops = kind << 24 | ctx << 8 | b
    return np.append(ops, np.uint32(2 << 24 | 1)).astype(np.uint32)  # terminate(1): flush


@pytest.mark.parametrize("seed", range(10))
def test_engine_matches_reference_encoder(seed):
    rng = np.random.default_rng(seed)
    lib, ref = synth(), S.ref()
    ref.ref_bac_ops.restype = C.c_size_t
    ref.ref_bac_ops.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_int32, C.c_void_p, C.c_size_t]
    for n in (1, 7, 300, 5000):
        ops = bin_script(rng, n)
        table, qp = int(rng.integers(0, 4)), int(rng.integers(0, 52))
        a = np.zeros(len(ops) * 2 + 64, np.uint8)
        b = np.zeros(len(ops) * 2 + 64, np.uint8)
        na = lib.synth_cabac_ops(S.ptr(ops), len(ops), table, qp, S.ptr(a), len(a))
        nb = ref.ref_bac_ops(S.ptr(ops), len(ops), table, qp, S.ptr(b), len(b))
        assert na == nb and np.array_equal(a[:na], b[:nb]), (seed, n)
        bins = np.zeros(len(ops), np.uint8)
        assert lib.synth_cabac_ops_decode(S.ptr(ops), len(ops), table, qp, S.ptr(b), nb, S.ptr(bins)) == 0
        assert np.array_equal(bins, (ops & 1).astype(np.uint8))


# ---- 2. known answers from the reference's writer ---------------------------------------------

CONCEAL_SEEDS = [411, 412, 416, 419]


def conceal_slices(seed, rounds=6):
    """CABAC concealment slices written by the reference: (stream, params, expected records)."""
    import test_host_lib as T
    lib = T.harness()
    stream = T.gen(lib, T.REF, seed, rounds=rounds, conceal=1)
    ev, _, _ = S.ref_trace(stream, flags=0)
    params = np.frombuffer(S.slice_params_from_trace(ev), S.SLICE_PARAMS).copy()
    # ChromaArrayType 1/2 only: the reference's concealment writer codes intra_chroma_pred_mode
    # and one luma DC coded_block_flag unconditionally (src/h264_writer.c:95-117), which is not
    # the macroblock_layer() syntax for monochrome or 4:4:4 pictures
    keep = params[(params["entropy_coding_mode_flag"] == 1) &
                  ((params["chroma_array_type"] == 1) | (params["chroma_array_type"] == 2))]
    off = 0
    for p in keep:  # the reference fires no MB callbacks for CABAC, so size the outputs here
        p["mb_out_off"] = off
        p["mb_out_cap"] = int(p["pic_width_in_mbs"]) * int(p["pic_height_in_mbs"]) - int(p["first_mb_in_slice"])
        off += int(p["mb_out_cap"])
    return stream, keep, off


@pytest.mark.parametrize("seed", CONCEAL_SEEDS)
def test_reference_concealment_slices_decode(seed):
    stream, params, cap = conceal_slices(seed)
    if len(params) == 0:
        pytest.skip("no CABAC concealment slice for this seed")
    recs, res = cpu_decode(stream, params.view(np.uint8), cap)
    check_concealment(stream, params, recs, res)


def check_concealment(stream, params, recs, res):
    i16 = L.mb_hash_term(1, 0, 3) + L.mb_hash_term(29, 0, 2)  # raw mb_type 3, Intra16x16PredMode 2
    n_checked = 0
    for p, r in zip(params, res):
        assert r["status"] == 0, r
        got = recs[int(p["mb_out_off"]):int(p["mb_out_off"]) + int(r["mb_count"])]
        assert len(got) >= 1
        assert np.array_equal(got["mb_addr"], np.arange(len(got)) + int(p["first_mb_in_slice"]))
        if p["slice_type"] == 2:
            assert (got["mb_type"] == 2).all() and (got["hash"] == np.uint64(i16 & (2**64 - 1))).all()
        else:
            assert (got["mb_type"] == 10).all() and (got["hash"] == 0).all()
        # every bit up to the stop bit consumed: only alignment zeros remain in the NAL
        assert int(p["nal_len"]) * 8 - 8 < int(r["end_bit"]) <= int(p["nal_len"]) * 8
        n_checked += len(got)
    assert n_checked > 0


# ---- 3. twin streams ------------------------------------------------------------------------

TWIN_CFGS = [
    dict(width_mbs=20, height_mbs=12, frames=6, slices_per_frame=3, profile_idc=77, transform_8x8=0,
         b_frames=1, num_ref_frames=2, idr_period=3, pct_skip=30, coef_density=50, seed=31),
    dict(width_mbs=22, height_mbs=10, frames=8, slices_per_frame=2, profile_idc=100, transform_8x8=1,
         b_frames=1, num_ref_frames=3, idr_period=4, pct_skip=25, coef_density=70, pct_pcm=30, seed=32),
    dict(width_mbs=9, height_mbs=7, frames=6, slices_per_frame=1, profile_idc=100, chroma_format_idc=2,
         transform_8x8=1, b_frames=1, num_ref_frames=4, idr_period=6, pct_skip=15, coef_density=90,
         pct_intra_in_inter=30, seed=33),
]


def make_twin(cfg):
    """(blob of pseudo NALs with CABAC slice data, params, the REFERENCE's records)."""
    lib = synth()
    stream, nmb, nsl = L.synth_video(cabac_twin=1, **cfg)
    ev, ref_mbs, syn = S.ref_trace_syntax(stream, flags=1)
    assert len(ref_mbs) == nmb and syn.shape[1] == lib.synth_cabac_sizeof_mb_syntax()
    params = np.frombuffer(S.slice_params_from_trace(ev), S.SLICE_PARAMS).copy()
    assert len(params) == nsl
    rng = np.random.default_rng(cfg["seed"])
    blob = [np.zeros(16, np.uint8)]
    pos = 16
    buf = np.zeros(4 << 20, np.uint8)
    for p in params:
        lo, n = int(p["mb_out_off"]), int(p["mb_out_cap"])
        p["cabac_init_idc"] = int(rng.integers(0, 3))
        need = lib.synth_cabac_transcode(p.tobytes(), S.ptr(np.ascontiguousarray(syn[lo:lo + n])), n,
                                         S.ptr(buf), len(buf))
        assert 0 < need <= len(buf), "slice not representable in CABAC"
        esc = S.oracle_insert(buf[:need])  # RBSP -> escaped bytes
        nal = np.concatenate([np.array([0x65], np.uint8), esc])
        p["nal_off"], p["nal_len"], p["data_bit_off"] = pos, len(nal), 8
        p["entropy_coding_mode_flag"] = 1
        blob.append(nal)
        pos += len(nal)
    return np.concatenate(blob), params, ref_mbs


@pytest.mark.parametrize("cfg", TWIN_CFGS, ids=lambda c: "seed%d" % c["seed"])
def test_twin_stream_matches_reference_records(cfg):
    blob, params, ref_mbs = make_twin(cfg)
    recs, res = cpu_decode(blob, params.view(np.uint8), len(ref_mbs))
    assert (res["status"] == 0).all(), res
    assert np.array_equal(res["mb_count"], params["mb_out_cap"])
    assert np.array_equal(recs, ref_mbs)
    assert (res["end_bit"] <= params["nal_len"].astype(np.uint64) * 8).all()
    assert (res["end_bit"] > (params["nal_len"].astype(np.uint64) - 1) * 8).all()


GOLDEN = ["main_b", "high_t8_pcm"]


def load_golden_twin(name):
    g = np.load(os.path.join(HERE, "golden", "cabac_twin_%s.npz" % name))
    return g["blob"], g["params"], g["mbs"]


@pytest.mark.parametrize("name", GOLDEN)
def test_golden_twin_fixture(name):
    """Committed fixture (tests/golden/make_golden.py): CABAC blob + the reference's records."""
    blob, params, ref_mbs = load_golden_twin(name)
    recs, res = cpu_decode(blob, params, len(ref_mbs))
    assert (res["status"] == 0).all() and np.array_equal(recs, ref_mbs)


# ---- 4. generator round trip ------------------------------------------------------------------

@pytest.mark.parametrize("cfg", CFGS, ids=lambda c: "seed%d" % c["seed"])
def test_generator_round_trip(cfg):
    stream, params, want = gen_cabac(**cfg)
    recs, res = cpu_decode(stream, params, len(want))
    P = np.frombuffer(params, S.SLICE_PARAMS)
    assert (res["status"] == 0).all()
    assert np.array_equal(res["mb_count"], P["mb_out_cap"])
    assert np.array_equal(recs, want)
    # the reference reader accepts the stream and (F2) delivers headers but no macroblocks
    ev, mbs, off = S.ref_trace(stream, flags=1)
    assert len(mbs) == 0 and sum(1 for t, _ in ev if t == S.TR_SLICE) == len(P)
    pr = np.frombuffer(S.slice_params_from_trace(ev), S.SLICE_PARAMS)
    for f in ("nal_off", "nal_len", "data_bit_off", "first_mb_in_slice", "slice_type", "cabac_init_idc",
              "slice_qp", "entropy_coding_mode_flag", "transform_8x8_mode_flag", "chroma_array_type",
              "num_ref_idx_l0_active_minus1", "num_ref_idx_l1_active_minus1"):
        assert np.array_equal(pr[f], P[f]), f


def test_unsupported_and_damaged_slices():
    stream, params, want = gen_cabac(**CFGS[0])
    P = np.frombuffer(params, S.SLICE_PARAMS).copy()
    P["mbaff_frame_flag"][0] = 1
    P["entropy_coding_mode_flag"][1] = 0
    P["nal_len"][2] = P["nal_len"][2] // 2  # truncated: must stop with -EIO or a short count
    recs, res = cpu_decode(stream, P.view(np.uint8), len(want))
    assert res["status"][0] == -38 and res["status"][1] == 1
    assert res["status"][2] in (-5, 0) and (res["status"][2] == -5 or res["mb_count"][2] < P["mb_out_cap"][2])
    assert (res["status"][3:] == 0).all()


# ---- GPU ------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def gpu():
    g = L.Gpu(0)
    yield g
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CFGS, ids=lambda c: "seed%d" % c["seed"])
def test_gpu_generator_streams(gpu, cfg):
    stream, params, want = gen_cabac(**cfg)
    recs, res = gpu.cabac_parse_host(stream, params, len(want))
    crecs, cres = cpu_decode(stream, params, len(want))
    assert np.array_equal(res, cres)
    assert np.array_equal(recs, want)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", TWIN_CFGS, ids=lambda c: "seed%d" % c["seed"])
def test_gpu_twin_stream_matches_reference_records(gpu, cfg):
    blob, params, ref_mbs = make_twin(cfg)
    recs, res = gpu.cabac_parse_host(blob, params.view(np.uint8), len(ref_mbs))
    assert (res["status"] == 0).all()
    assert np.array_equal(recs, ref_mbs)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN)
def test_gpu_golden_twin_fixture(gpu, name):
    blob, params, ref_mbs = load_golden_twin(name)
    recs, res = gpu.cabac_parse_host(blob, params, len(ref_mbs))
    assert (res["status"] == 0).all() and np.array_equal(recs, ref_mbs)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", CONCEAL_SEEDS)
def test_gpu_reference_concealment_slices(gpu, seed):
    stream, params, cap = conceal_slices(seed)
    if len(params) == 0:
        pytest.skip("no CABAC concealment slice for this seed")
    recs, res = gpu.cabac_parse_host(stream, params.view(np.uint8), cap)
    check_concealment(stream, params, recs, res)


@pytest.mark.gpu
@pytest.mark.parametrize("lanes", [0, 2, 5])
def test_gpu_many_slices_all_lane_packings(gpu, lanes, monkeypatch):
    cfg = dict(width_mbs=30, height_mbs=17, frames=24, slices_per_frame=17, profile_idc=100, transform_8x8=1,
               b_frames=1, num_ref_frames=2, idr_period=8, pct_skip=30, coef_density=60, seed=77)
    stream, params, want = gen_cabac(**cfg)
    monkeypatch.setenv("H264GPU_CABAC_LANES_LOG2", str(lanes))
    recs, res = gpu.cabac_parse_host(stream, params, len(want))
    assert (res["status"] == 0).all()
    assert np.array_equal(recs, want)
    P = np.frombuffer(params, S.SLICE_PARAMS).copy()
    P["nal_len"][5] //= 3  # a damaged slice must not disturb its neighbours
    recs2, res2 = gpu.cabac_parse_host(stream, P.view(np.uint8), len(want))
    ok = np.ones(len(P), bool)
    ok[5] = False
    assert (res2["status"][ok] == 0).all()
    lo, hi = int(P["mb_out_off"][5]), int(P["mb_out_off"][5] + P["mb_out_cap"][5])
    assert np.array_equal(recs2[:lo], want[:lo]) and np.array_equal(recs2[hi:], want[hi:])


# ---- 5. the independent decoder (oracle/oracle_cabac_spec.c) -----------------------------------
# Written from H.264 9.3 without the shared walker / engine of cabac_syntax.h + cabac_engine.h.
# It must agree with (a) the REFERENCE's records of every twin stream, (b) the reference writer's
# concealment slices, (c) the generator's intended records on every generator stream: any ctxIdx /
# binarisation drift that generator and kernel share would show up as a disagreement here.

def spec_decode(stream, params, n_records):
    lib = S.oracle()
    lib.oracle_cabac_spec_decode.restype = C.c_int
    lib.oracle_cabac_spec_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64,
                                             C.c_void_p]
    stream = np.ascontiguousarray(stream, np.uint8)
    params = np.ascontiguousarray(params, np.uint8)
    n = len(params) // S.PARAMS_SIZE
    recs = np.zeros(max(n_records, 1), S.MB_RECORD)
    res = np.zeros(max(n, 1), S.SLICE_RESULT)
    assert lib.oracle_cabac_spec_decode(S.ptr(stream), len(stream), S.ptr(params), n, S.ptr(recs), n_records,
                                        S.ptr(res)) == 0
    return recs[:n_records], res[:n]


@pytest.mark.parametrize("cfg", TWIN_CFGS, ids=lambda c: "seed%d" % c["seed"])
def test_independent_decoder_on_twin_streams(cfg):
    blob, params, ref_mbs = make_twin(cfg)
    recs, res = spec_decode(blob, params.view(np.uint8), len(ref_mbs))
    assert (res["status"] == 0).all(), res
    assert np.array_equal(res["mb_count"], params["mb_out_cap"])
    bad = np.nonzero(recs != ref_mbs)[0]
    assert len(bad) == 0, (bad[:5], recs[bad[:3]], ref_mbs[bad[:3]])
    k_recs, k_res = cpu_decode(blob, params.view(np.uint8), len(ref_mbs))
    assert np.array_equal(k_res["end_bit"], res["end_bit"])


@pytest.mark.parametrize("name", GOLDEN)
def test_independent_decoder_on_golden_twins(name):
    blob, params, ref_mbs = load_golden_twin(name)
    recs, res = spec_decode(blob, params, len(ref_mbs))
    assert (res["status"] == 0).all() and np.array_equal(recs, ref_mbs)


@pytest.mark.parametrize("seed", CONCEAL_SEEDS)
def test_independent_decoder_on_reference_concealment_slices(seed):
    stream, params, cap = conceal_slices(seed)
    if len(params) == 0:
        pytest.skip("no CABAC concealment slice for this seed")
    recs, res = spec_decode(stream, params.view(np.uint8), cap)
    check_concealment(stream, params, recs, res)


@pytest.mark.parametrize("cfg", CFGS, ids=lambda c: "seed%d" % c["seed"])
def test_independent_decoder_on_generator_streams(cfg):
    stream, params, want = gen_cabac(**cfg)
    recs, res = spec_decode(stream, params, len(want))
    P = np.frombuffer(params, S.SLICE_PARAMS)
    assert (res["status"] == 0).all(), res["status"]
    assert np.array_equal(res["mb_count"], P["mb_out_cap"])
    bad = np.nonzero(recs != want)[0]
    assert len(bad) == 0, (bad[:5], recs[bad[:3]], want[bad[:3]])
    k_recs, k_res = cpu_decode(stream, params, len(want))
    assert np.array_equal(k_res["end_bit"], res["end_bit"])


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CFGS[:3], ids=lambda c: "seed%d" % c["seed"])
def test_gpu_kernel_agrees_with_the_independent_decoder(gpu, cfg):
    stream, params, want = gen_cabac(**cfg)
    recs, res = gpu.cabac_parse_host(stream, params, len(want))
    s_recs, s_res = spec_decode(stream, params, len(want))
    assert (res["status"] == 0).all() and np.array_equal(recs, s_recs)
    assert np.array_equal(res["mb_count"], s_res["mb_count"]) and np.array_equal(res["end_bit"], s_res["end_bit"])
