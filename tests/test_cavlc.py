"""K4 CAVLC slice parse: generator validity, kernel logic on CPU (emu build of the
kernel source) against the compiled reference and the committed golden fixtures."""
import os
import zlib

import numpy as np
import pytest

import libh264_b200 as L
import support as S
from golden.make_golden import CASES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not S.have_ref(), reason="oracle/_ref/libh264_ref.so not built")


def load_golden(name):
    return np.load(os.path.join(GOLD, "cavlc_%s.npz" % name))


def group_maps(g):
    """The reference's macroblock -> slice group maps of a fixture (None without slice groups)."""
    return g["group_maps"] if "group_maps" in g.files and len(g["group_maps"]) else None


@pytest.mark.parametrize("name", sorted(CASES))
def test_emu_parse_matches_golden(name):
    g = load_golden(name)
    stream, nmb, nsl = L.synth_video(**CASES[name])
    assert zlib.crc32(stream.tobytes()) == int(g["crc"]), "generator output changed"
    recs, res = S.emu_cavlc_parse(stream, g["params"], len(g["mbs"]), group_maps=group_maps(g))
    assert (res["status"] == 0).all()
    assert np.array_equal(res["mb_count"], g["mb_counts"])
    assert np.array_equal(recs["mb_addr"], g["mbs"]["mb_addr"])
    assert np.array_equal(recs["mb_type"], g["mbs"]["mb_type"])
    assert np.array_equal(recs["hash"], g["mbs"]["hash"])


@needs_ref
@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_is_what_the_reference_says(name):
    g = load_golden(name)
    stream, nmb, nsl = L.synth_video(**CASES[name])
    ev, mbs, off = S.ref_trace(stream)
    assert len(mbs) == nmb and np.array_equal(mbs, g["mbs"])
    assert np.array_equal(S.group_maps_from_trace(ev)[0], g["params"])
    assert off == int(g["final_off"]) == len(stream)


@needs_ref
def test_generator_streams_are_valid_for_the_reference():
    # every macroblock consumed, every slice ends at its stop bit, across seeds / mixes
    for seed in range(6):
        kw = dict(width_mbs=9, height_mbs=5, frames=5, slices_per_frame=1 + seed % 3, b_frames=seed & 1,
                  num_ref_frames=1 + seed % 4, profile_idc=[66, 77, 100][seed % 3],
                  transform_8x8=int(seed % 3 == 2), pct_skip=10 * seed, coef_density=20 + 15 * seed,
                  pct_pcm=10 * (seed % 2), seed=seed)
        stream, nmb, nsl = L.synth_video(**kw)
        ev, mbs, off = S.ref_trace(stream)
        tags = [t for t, _ in ev]
        assert len(mbs) == nmb and tags.count(S.TR_SLICE) == nsl and tags.count(S.TR_NALU_END) == nsl + 2
        assert sum(int(p.view(np.uint32)[0]) for t, p in ev if t == S.TR_SLICE_DATA_END) == nmb
        recs, res = S.emu_cavlc_parse(stream, S.slice_params_from_trace(ev), nmb)
        assert (res["status"] == 0).all() and np.array_equal(recs, mbs), seed


@needs_ref
def test_emu_parse_1080p_frame():
    stream, nmb, nsl = L.synth_video(width_mbs=120, height_mbs=68, frames=2, idr_period=2, slices_per_frame=4)
    ev, mbs, off = S.ref_trace(stream)
    recs, res = S.emu_cavlc_parse(stream, S.slice_params_from_trace(ev), nmb)
    assert (res["status"] == 0).all() and np.array_equal(recs, mbs)


def test_cabac_slice_is_skipped_like_the_reference():
    g = load_golden("baseline")
    stream, _, _ = L.synth_video(**CASES["baseline"])
    p = g["params"].copy().view(L.SLICE_PARAMS)
    p["entropy_coding_mode_flag"] = 1
    recs, res = S.emu_cavlc_parse(stream, p.view(np.uint8), len(g["mbs"]))
    assert (res["status"] == 1).all() and (res["mb_count"] == 0).all()


def test_truncated_slice_reports_eio_not_garbage():
    g = load_golden("baseline")
    stream, _, _ = L.synth_video(**CASES["baseline"])
    p = g["params"].copy().view(L.SLICE_PARAMS)
    p["nal_len"] = p["nal_len"] // 2
    recs, res = S.emu_cavlc_parse(stream, p.view(np.uint8), len(g["mbs"]))
    assert ((res["status"] == -5) | (res["status"] == 0)).all() and (res["status"] == -5).any()
    assert (res["mb_count"] <= g["mb_counts"]).all()


# ---- GPU ---------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_parse_matches_golden(gpu, name):
    import ctypes as C
    g = load_golden(name)
    stream, nmb, nsl = L.synth_video(**CASES[name])
    maps = group_maps(g)
    if maps is not None:
        maps = np.ascontiguousarray(maps)
        L._check(gpu.lib.h264gpu_reader_set_group_maps(gpu.h, C.c_void_p(maps.ctypes.data), C.c_uint64(len(maps))),
                 "h264gpu_reader_set_group_maps")
    recs, res = gpu.cavlc_parse_host(stream, g["params"], len(g["mbs"]))
    assert (res["status"] == 0).all() and np.array_equal(res["mb_count"], g["mb_counts"])
    assert np.array_equal(recs, g["mbs"])


@pytest.mark.gpu
def test_gpu_parse_many_slices_vs_reference(gpu):
    if not S.have_ref():
        pytest.skip("oracle/_ref/libh264_ref.so not available")
    # 1080p, 16 slices per frame, P/B/I mix: 640 independent slices
    stream, nmb, nsl = L.synth_video(width_mbs=120, height_mbs=68, frames=40, slices_per_frame=16,
                                     profile_idc=100, transform_8x8=1, b_frames=1, num_ref_frames=2,
                                     idr_period=10, seed=9)
    ev, mbs, off = S.ref_trace(stream)
    recs, res = gpu.cavlc_parse_host(stream, S.slice_params_from_trace(ev), nmb)
    assert (res["status"] == 0).all() and int(res["mb_count"].sum()) == nmb
    assert np.array_equal(recs, mbs)


@needs_ref
@pytest.mark.parametrize("name", sorted(CASES))
def test_emu_full_records_match_the_reference_ctx_mb(name):
    """N2: the full per-macroblock record (struct h264_mb_syntax) the kernel writes is, byte for
    byte, what the reference holds in its private ctx->mb for the same macroblock
    (src/h264_macroblock.h:105-167; dumped by oracle/ref_harness.c mb_syntax_from_ref)."""
    stream, nmb, nsl = L.synth_video(**CASES[name])
    ev, mbs, syn = S.ref_trace_syntax(stream)
    params, maps = S.group_maps_from_trace(ev)
    recs, res, got = S.emu_cavlc_parse_full(stream, params, nmb, group_maps=maps if len(maps) else None)
    assert (res["status"] == 0).all() and len(syn) == nmb
    assert np.array_equal(recs["hash"], mbs["hash"])
    if CASES[name].get("chroma_format_idc", 1) == 3:
        return  # the record has no room for the Cb / Cr planes of 4:4:4 (checksum only)
    bad = np.nonzero((got != syn).any(axis=1))[0]
    assert len(bad) == 0, (name, bad[:5], np.nonzero(got[bad[0]] != syn[bad[0]])[0][:10])


@needs_ref
@pytest.mark.gpu
def test_gpu_full_records_match_the_reference_ctx_mb(gpu):
    import ctypes as C
    cfg = dict(frames=6, width_mbs=40, height_mbs=30, slices_per_frame=5, profile_idc=100, transform_8x8=1,
               b_frames=1, num_ref_frames=2, idr_period=4, pct_skip=20, coef_density=60, seed=12)
    stream, nmb, nsl = L.synth_video(**cfg)
    ev, mbs, syn = S.ref_trace_syntax(stream)
    params = S.slice_params_from_trace(ev)
    sz = syn.shape[1]
    d_s = gpu.alloc(len(stream) + 64)
    d_s.upload(stream)
    d_p = gpu.alloc(len(params))
    d_p.upload(params)
    d_r, d_q, d_y = gpu.alloc(nmb * 16 + 16), gpu.alloc(nsl * 16), gpu.alloc(nmb * sz + 16)
    d_y.upload(np.full(nmb * sz, 0xEE, np.uint8))
    try:
        L._check(gpu.lib.h264gpu_cavlc_parse_full_dev(gpu.h, C.c_void_p(d_s.ptr), C.c_uint64(len(stream)), C.c_void_p(d_p.ptr),
                                                      C.c_uint32(nsl), C.c_void_p(d_r.ptr), C.c_void_p(d_q.ptr),
                                                      C.c_void_p(d_y.ptr), None), "h264gpu_cavlc_parse_full_dev")
        gpu.sync()
        got = d_y.download(nmb * sz).reshape(nmb, sz)
        res = d_q.download(dtype=np.uint8).view(L.SLICE_RESULT)
    finally:
        for d in (d_s, d_p, d_r, d_q, d_y):
            d.free()
    assert (res["status"] == 0).all()
    bad = np.nonzero((got != syn).any(axis=1))[0]
    assert len(bad) == 0, (bad[:5],)


@pytest.mark.parametrize("name", sorted(CASES))
def test_emu_end_position_and_truncated_slices(name):
    """end_bit sits in the slice's last byte (right before the stop bit); a truncated slice gives
    -EIO or a shorter clean parse, never more macroblocks than the slice has."""
    g = load_golden(name)
    stream, nmb, nsl = L.synth_video(**CASES[name])
    r, q = S.emu_cavlc_parse(stream, g["params"], nmb, group_maps=group_maps(g))
    nal_len = g["params"].view(L.SLICE_PARAMS)["nal_len"].astype(np.uint64)
    assert ((q["end_bit"] > (nal_len - 2) * 8) & (q["end_bit"] < nal_len * 8)).all()
    p = g["params"].copy().view(L.SLICE_PARAMS)
    p["nal_len"] = p["nal_len"] * 2 // 3
    r, q = S.emu_cavlc_parse(stream, p.view(np.uint8), nmb, group_maps=group_maps(g))
    assert ((q["status"] == -5) | (q["status"] == 0)).all() and (q["status"] == -5).any()
    assert (q["mb_count"] <= g["mb_counts"]).all()


@needs_ref
@pytest.mark.parametrize("map_type", range(7))
def test_emu_slice_groups_match_the_reference(map_type):
    """N3 / A12 (FMO): streams with 2-3 slice groups of every slice_group_map_type; the kernel walks
    each slice with the macroblock -> slice group map (next macroblock of the group, neighbours only
    from the slice's own group) and delivers what the reference delivers."""
    for groups in (2, 3):
        kw = dict(width_mbs=11, height_mbs=9, frames=6, slices_per_frame=2, b_frames=1, num_ref_frames=2,
                  profile_idc=100, transform_8x8=1, pct_skip=25, seed=5 + map_type, idr_period=4,
                  fmo=groups | map_type << 4)
        stream, nmb, nsl = L.synth_video(**kw)
        ev, mbs, off = S.ref_trace(stream)
        assert len(mbs) == nmb and off == len(stream)
        params, maps = S.group_maps_from_trace(ev)
        assert len(maps) == nsl * 99
        recs, res = S.emu_cavlc_parse(stream, params, nmb, group_maps=maps)
        assert (res["status"] == 0).all(), res["status"]
        assert np.array_equal(recs, mbs), (map_type, groups)
        # without the maps the slices are refused, not misparsed
        recs, res = S.emu_cavlc_parse(stream, params, nmb)
        assert (res["status"] == -38).all()


@needs_ref
@pytest.mark.gpu
def test_gpu_slice_groups_match_the_reference(gpu):
    import ctypes as C
    for map_type in range(7):
        kw = dict(width_mbs=16, height_mbs=15, frames=6, slices_per_frame=3, b_frames=1, num_ref_frames=2,
                  profile_idc=100, transform_8x8=1, pct_skip=25, seed=15 + map_type, idr_period=4,
                  fmo=3 | map_type << 4)
        stream, nmb, nsl = L.synth_video(**kw)
        ev, mbs, off = S.ref_trace(stream)
        params, maps = S.group_maps_from_trace(ev)
        L._check(gpu.lib.h264gpu_reader_set_group_maps(gpu.h, C.c_void_p(maps.ctypes.data), C.c_uint64(len(maps))),
                 "h264gpu_reader_set_group_maps")
        recs, res = gpu.cavlc_parse_host(stream, params, nmb)
        assert (res["status"] == 0).all() and np.array_equal(recs, mbs), map_type
        recs, res = gpu.cavlc_parse_host(stream, params, nmb)  # the maps were for one call
        assert (res["status"] == -38).all()


@needs_ref
@pytest.mark.gpu
def test_gpu_every_lane_packing_gives_the_reference_records(gpu, monkeypatch):
    """Same records / results at every packing of slices into warps and with several slices per lane."""
    stream, nmb, nsl = L.synth_video(width_mbs=40, height_mbs=30, frames=24, slices_per_frame=10,
                                     profile_idc=100, transform_8x8=1, b_frames=1, num_ref_frames=3,
                                     idr_period=8, pct_skip=35, pct_pcm=20, seed=77)
    ev, mbs, off = S.ref_trace(stream)
    params = S.slice_params_from_trace(ev)
    for lanes in ("0", "2", "5"):
        for per_lane in ("1", "3"):
            monkeypatch.setenv("H264GPU_CAVLC_LANES_LOG2", lanes)
            monkeypatch.setenv("H264GPU_CAVLC_PER_LANE", per_lane)
            recs, res = gpu.cavlc_parse_host(stream, params, nmb)
            assert (res["status"] == 0).all() and np.array_equal(recs, mbs), (lanes, per_lane)


MBAFF_CASES = [dict(width_mbs=11, height_mbs=8, frames=6, slices_per_frame=1 + s % 3, b_frames=s & 1,
                    num_ref_frames=1 + s % 3, profile_idc=[77, 100, 100][s % 3], transform_8x8=int(s % 3 == 1),
                    pct_skip=[0, 25, 50, 70][s % 4], pct_pcm=10 * (s % 2), seed=50 + s, fmo=0x100, idr_period=4,
                    chroma_format_idc=[1, 1, 2, 3, 0][s % 5]) for s in range(10)]


@needs_ref
@pytest.mark.parametrize("kw", MBAFF_CASES, ids=lambda k: "seed%d" % k["seed"])
def test_emu_mbaff_matches_the_reference(kw):
    """N3 (MBAFF): frames coded as macroblock pairs with a field / frame decision per pair
    (mb_field_decoding_flag read, inherited or inferred; 6.4.12.2 neighbours for nC; ref_idx over
    fields): records -- addresses, types and the checksum that includes mb_field_decoding_flag as
    the reference holds it at callback time -- identical to the reference's."""
    stream, nmb, nsl = L.synth_video(**kw)
    ev, mbs, off = S.ref_trace(stream)
    assert len(mbs) == nmb and off == len(stream)
    params = S.slice_params_from_trace(ev)
    assert (params.view(L.SLICE_PARAMS)["mbaff_frame_flag"] == 1).all()
    recs, res = S.emu_cavlc_parse(stream, params, nmb)
    assert (res["status"] == 0).all(), res["status"]
    bad = np.nonzero(recs != mbs)[0]
    assert len(bad) == 0, (bad[:5], recs[bad[:3]], mbs[bad[:3]])


@needs_ref
@pytest.mark.gpu
def test_gpu_mbaff_and_field_pictures_match_the_reference(gpu):
    for fmo in (0x100, 0x200):
        for seed in (71, 72, 73):
            stream, nmb, nsl = L.synth_video(width_mbs=30, height_mbs=16, frames=8, slices_per_frame=4, b_frames=seed & 1,
                                             num_ref_frames=1 + seed % 3, profile_idc=100, transform_8x8=1,
                                             pct_skip=20 * (seed % 4), pct_pcm=10, seed=seed, idr_period=4, fmo=fmo)
            ev, mbs, off = S.ref_trace(stream)
            recs, res = gpu.cavlc_parse_host(stream, S.slice_params_from_trace(ev), nmb)
            assert (res["status"] == 0).all() and np.array_equal(recs, mbs), (fmo, seed)
