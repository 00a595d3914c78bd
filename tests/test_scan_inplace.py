"""K1/K2: Annex-B scan + EPB strip with every NAL's RBSP written in place
(h264gpu_split_strip_inplace_dev, annexb_scan7.cuh: scan7_kernel + second pass + fin7_*).

Per NAL the result must be what the reference yields: (start, end) of the h264_find_nalu loop
(src/h264_reader.c:133-140) and the byte sequence h264_bs_read_bits(8) reads from the NAL
(include/h264/h264_bitstream.h:168-218) — compared against the oracle (which tests/test_oracle.py
pins to the compiled reference).  CPU tests run the kernel sources on the SIMT emulator, GPU
tests go through the C-ABI."""
import ctypes as C

import numpy as np
import pytest

import libh264_b200 as L
import support as S

ALPHA = np.array([0, 0, 0, 1, 2, 3, 3, 0x41, 0xFF], np.uint8)


def check_per_nal(got, exp, what):
    n = len(exp["start"])
    assert got["res"].n_nal == n, (what, got["res"].n_nal, n)
    assert np.array_equal(got["start"], exp["start"]), what
    assert np.array_equal(got["end"], exp["end"]), what
    assert np.array_equal(got["rbsp_off"], exp["start"]), what  # in place: where the NAL starts
    ro = exp["rbsp_off"]
    assert np.array_equal(got["rbsp_len"], np.diff(ro).astype(np.uint64)), what
    for k in range(n):
        assert np.array_equal(got["nal_rbsp"][k], exp["rbsp"][int(ro[k]):int(ro[k + 1])]), (what, k)
    assert got["res"].rbsp_bytes == len(exp["rbsp"]), what
    assert got["res"].reserved == 0


def run_emu(buf, cpt, what):
    exp = S.oracle_split_strip(buf)
    check_per_nal(S.emu_split_strip_inplace(buf, cpt=cpt), exp, what)
    # scan only: same table, no RBSP
    g = S.emu_split_strip_inplace(buf, strip=False, cpt=cpt)
    assert np.array_equal(g["start"], exp["start"]) and np.array_equal(g["end"], exp["end"])


@pytest.mark.parametrize("cpt", [81, 82, 86, 88])
def test_emu_random_streams(cpt):
    rng = np.random.default_rng(100 + cpt)
    for it in range(12):
        n = int(rng.integers(1, 40))
        hi = int(rng.choice([64, 600, 5000, 40000]))
        buf = S.gen_annexb(rng, n, lo=1, hi=hi, p_zero=float(rng.choice([3 / 16, 0.5, 0.02])))
        run_emu(buf, cpt, ("random", cpt, it))


@pytest.mark.parametrize("cpt", [81, 82, 86, 88])
def test_emu_pathological_streams(cpt):
    rng = np.random.default_rng(7)
    for it in range(20):
        n = int(rng.integers(1, 30000))
        kind = it % 5
        if kind == 0:
            buf = np.tile(np.array([0, 0, 1], np.uint8), n // 3 + 1)[:n]  # nothing but start codes
        elif kind == 1:
            buf = np.zeros(n, np.uint8)
        elif kind == 2:
            buf = rng.choice(np.array([0, 0, 0, 1, 3], np.uint8), n)
        elif kind == 3:
            buf = rng.choice(ALPHA, n)
        else:
            buf = np.tile(np.array([0, 0, 3], np.uint8), n // 3 + 1)[:n]  # nothing but EPBs
        run_emu(buf, cpt, ("patho", cpt, it))


@pytest.mark.parametrize("cpt", [81, 88])
def test_emu_long_nal_crosses_many_tiles(cpt):
    """One NAL over > 32 tiles: the look-back has to walk more than one window, and tiles
    without a start code chain their shifts."""
    rng = np.random.default_rng(3)
    body = rng.choice(np.array([0, 0, 3, 7, 9], np.uint8), 4096 * 70)
    buf = np.concatenate([np.array([9, 9, 0, 0, 1, 0x65], np.uint8), S.oracle_insert(body),
                          np.array([0, 0, 0, 1, 0x41, 1, 2, 3], np.uint8)])
    run_emu(buf, cpt, "long")


@pytest.mark.parametrize("cpt", [81, 88])
def test_emu_more_tiles_than_one_finalize_block(cpt):
    """> 1024 tiles: the tile prefix of the finalize step spans several blocks."""
    rng = np.random.default_rng(9)
    buf = S.gen_annexb(rng, 1500, lo=64, hi=20000)
    assert len(buf) > 1100 * 4096
    run_emu(buf, cpt, "finalize blocks")


@pytest.mark.parametrize("cpt", [81, 82, 88])
def test_emu_seams_at_every_offset(cpt):
    """A start code / EPB / terminator sliding over a tile seam (4 KiB tiles; for gen 6 also a
    warp-span seam and, with 62, a row seam inside a span)."""
    for pat in ([0, 0, 1, 0x65, 0, 0, 3, 1], [0, 0, 0, 1, 0x41], [0, 0, 3, 0, 0, 3], [0, 0, 0, 0, 0, 1, 5]):
        for shift in range(0, 12):
            buf = np.full(4096 * 2 + 64, 0x55, np.uint8)
            buf[:4] = [0, 0, 1, 0x67]
            pos = 4096 - 8 + shift
            buf[pos:pos + len(pat)] = pat
            run_emu(buf, cpt, (pat, shift))


@pytest.mark.parametrize("cpt", [81])
def test_emu_event_buffer_overflow_is_reported(cpt):
    buf = np.tile(np.array([0, 0, 1, 7], np.uint8), 400)
    g = S.emu_split_strip_inplace(buf, cpt=cpt, ev_cap=16)
    assert g["res"].n_nal == 400 and g["res"].reserved == 1


def _merge_shards(run, b, cuts):
    """Byte-range shards through `run` (emulator or GPU) + the product's in-place merge."""
    lib = L.load_gpu_lib()
    n = len(b)
    cap = n // 3 + 2
    st, en = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
    rl = np.zeros(cap, np.uint64)
    pieces = [[] for _ in range(cap)]
    m = L.Merge()
    lib.h264gpu_merge_init(C.byref(m))
    bounds = [0] + list(cuts) + [n]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        e = S.ShardEdge()
        if lo >= 2:
            e.has_left, e.left[0], e.left[1] = 1, int(b[lo - 2]), int(b[lo - 1])
        if hi < n:
            e.has_right = 1
            e.right[0] = int(b[hi])
            e.right[1] = int(b[hi + 1]) if hi + 1 < n else 0xFF
        e.assume_in = 1 if lo > 0 else 0
        r = run(b[lo:hi], e, lo)
        k = int(r["res"].n_nal)
        first = int(m.n_nal)
        st[first:first + k] = r["start"]
        en[first:first + k] = r["end"]
        rl[first:first + k] = r["rbsp_len"]
        for i in range(k):
            pieces[first + i].append(r["nal_rbsp"][i])
        res = L.ScanResult.from_buffer_copy(bytes(r["res"]))
        carry = C.c_uint64(0)
        open_nal = first - 1
        lib.h264gpu_merge_shard_inplace(C.byref(m), C.byref(res), S.ptr(en), S.ptr(rl), cap, C.byref(carry))
        if carry.value:  # continues at this shard's start (after the tail of a straddling start code)
            at = max(0, int(st[open_nal]) - lo)
            pieces[open_nal].append(r["body"][at:at + carry.value])
    off = C.c_uint64(0)
    lib.h264gpu_merge_finish(C.byref(m), n, S.ptr(en), cap, C.byref(off))
    k = int(m.n_nal)
    rbsp = [np.concatenate(p) if p else np.zeros(0, np.uint8) for p in pieces[:k]]
    return dict(start=st[:k], end=en[:k], rbsp_len=rl[:k], nal_rbsp=rbsp, final_off=off.value,
                rbsp_bytes=int(m.rbsp_bytes))


def check_merged(g, o, what):
    assert np.array_equal(g["start"], o["start"]), what
    assert np.array_equal(g["end"], o["end"]), what
    ro = o["rbsp_off"]
    assert np.array_equal(g["rbsp_len"], np.diff(ro).astype(np.uint64)), what
    for k in range(len(o["start"])):
        assert np.array_equal(g["nal_rbsp"][k], o["rbsp"][int(ro[k]):int(ro[k + 1])]), (what, k)
    assert g["final_off"] == o["final_off"] and g["rbsp_bytes"] == len(o["rbsp"]), what


@pytest.mark.parametrize("cpt", [81, 88])
def test_emu_sharded_scan_merges_to_whole(cpt):
    rng = np.random.default_rng(5)
    run = lambda buf, e, lo: S.emu_split_strip_inplace(buf, cpt=cpt, edge=e, base=lo)
    for it in range(12):
        b = S.gen_annexb(rng, 30, 1, 3000) if it % 2 else rng.choice(ALPHA, int(rng.integers(200, 20000)))
        n = len(b)
        cuts = sorted(set(int(c) // 16 * 16 for c in rng.integers(16, max(n, 17), int(rng.integers(1, 5)))) - {0})
        cuts = [c for c in cuts if c < n]
        check_merged(_merge_shards(run, b, cuts), S.oracle_split_strip(b), (it, cuts))
    # a NAL running through a whole shard without any event in it
    body = rng.choice(np.array([0, 0, 3, 7], np.uint8), 9000)
    b = np.concatenate([np.array([0, 0, 1, 0x65], np.uint8), S.oracle_insert(body), np.array([0, 0, 1, 9], np.uint8)])
    check_merged(_merge_shards(run, b, [2048, 4096, 6000 // 16 * 16]), S.oracle_split_strip(b), "through")


@pytest.mark.parametrize("gen", [8])
def test_emu_gen6_row_and_span_seams_and_shard_ends(gen):
    """Gen 6 / gen 7 (cpt 8x) specifics: patterns sliding over a row seam inside a span (62: rows of 512 B, spans of
    1 KiB, tiles of 8 KiB), streams ending within +-3 bytes of a tile multiple (the launch covers
    len + 2 bytes because events are owned by their third byte), and a shard cut exactly at a
    tile multiple with the start code straddling it."""
    pats = ([0, 0, 1, 0x65, 0, 0, 3, 1], [0, 0, 0, 1, 0x41], [0, 0, 3, 0, 0, 3, 0, 0, 3, 0, 0, 3], [0, 0, 0, 0, 0, 1, 5])
    for pat in pats:
        for seam in (512, 1024):
            for shift in range(0, 14):
                buf = np.full(8192 + 100, 0x55, np.uint8)
                buf[:4] = [0, 0, 1, 0x67]
                buf[100:106] = [0, 0, 3, 0, 0, 3]  # a non-zero shift before the seam
                pos = seam - 10 + shift
                buf[pos:pos + len(pat)] = pat
                run_emu(buf, gen * 10 + 2, (pat, seam, shift))
    rng = np.random.default_rng(21)
    for tail in ([0, 0, 1], [0, 0, 0], [0, 0, 3], [0, 0, 1, 9], [0, 0], [7, 0, 0, 3, 0]):
        for d in range(-3, 4):
            n = 4096 * 2 + d
            buf = S.oracle_insert(rng.choice(np.array([0, 0, 3, 7, 9], np.uint8), n))[:n - len(tail) - 4]
            buf = np.concatenate([np.array([0, 0, 1, 0x65], np.uint8), buf, np.array(tail, np.uint8)])
            assert len(buf) == n
            run_emu(buf, gen * 10 + 1, (tail, d))
    run = lambda b, e, lo: S.emu_split_strip_inplace(b, cpt=gen * 10 + 1, edge=e, base=lo)
    for shift in range(0, 8):
        b = np.full(4096 * 3, 0x33, np.uint8)
        b[:4] = [0, 0, 1, 0x67]
        b[200:206] = [0, 0, 3, 0, 0, 3]
        b[4096 - 4 + shift:4096 - 4 + shift + 5] = [0, 0, 0, 1, 0x41]
        b[8192 - 3 + shift % 4:8192 - 3 + shift % 4 + 3] = [0, 0, 3]
        check_merged(_merge_shards(run, b, [4096, 8192]), S.oracle_split_strip(b), ("cut at tile", shift))


# ---- GPU ------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def gpu():
    g = L.Gpu(0)
    yield g
    g.close()


@pytest.mark.gpu
def test_gpu_random_and_pathological_streams(gpu):
    rng = np.random.default_rng(11)
    for it in range(10):
        buf = S.gen_annexb(rng, int(rng.integers(1, 300)), lo=1, hi=int(rng.choice([600, 40000, 300000])),
                           p_zero=float(rng.choice([3 / 16, 0.5, 0.02])))
        check_per_nal(gpu.split_strip_inplace(buf), S.oracle_split_strip(buf), ("gpu random", it))
    for it in range(10):
        n = int(rng.integers(1, 300000))
        kind = it % 5
        buf = (np.tile(np.array([0, 0, 1], np.uint8), n // 3 + 1)[:n] if kind == 0 else
               np.zeros(n, np.uint8) if kind == 1 else
               rng.choice(np.array([0, 0, 0, 1, 3], np.uint8), n) if kind == 2 else
               rng.choice(ALPHA, n) if kind == 3 else
               np.tile(np.array([0, 0, 3], np.uint8), n // 3 + 1)[:n])
        check_per_nal(gpu.split_strip_inplace(buf), S.oracle_split_strip(buf), ("gpu patho", it))


@pytest.mark.gpu
def test_gpu_config2_shape_against_oracle_and_packed_kernel(gpu):
    """64 MiB of the BASELINE config-2 workload: oracle parity per NAL, and the same NAL table
    and RBSP bytes as the packed kernel."""
    offs = L.synth_offsets(5, 64 << 20)
    rbsp = L.synth_payloads(5, offs)
    stream, _ = L.synth_annexb(5, rbsp, offs)
    got = gpu.split_strip_inplace(stream, cap=len(offs) + 64)
    exp = S.oracle_split_strip(stream)
    check_per_nal(got, exp, "config2")
    packed = gpu.split_strip_host(stream)
    assert np.array_equal(packed["start"], got["start"]) and np.array_equal(packed["end"], got["end"])
    assert np.array_equal(np.concatenate(got["nal_rbsp"]), packed["rbsp"])
    # round trip: the RBSPs are the generator's payloads (trailing zero bytes at the very end of
    # the stream belong to the last NAL, as for the reference)
    allr = np.concatenate(got["nal_rbsp"])
    assert np.array_equal(allr[:len(rbsp)], rbsp) and len(allr) - len(rbsp) <= 2 and not allr[len(rbsp):].any()


@pytest.mark.gpu
def test_gpu_config2_full_size_properties(gpu):
    """BASELINE config 2 at its full size (4 GiB of synthetic Annex-B): too big for the oracle in
    the test budget, so size-independent properties: strip(escape(x)) == x NAL by NAL against the
    generator's payloads, the NAL table against the generator's offsets, ordering, nothing written
    outside the output buffer (checked by the wrapper)."""
    seed = 0x264
    offs = L.synth_offsets(seed, int((4 << 30) * 0.992))
    rbsp = L.synth_payloads(seed, offs)
    stream, nal_off = L.synth_annexb(seed, rbsp, offs)
    n = len(offs) - 1
    got = gpu.split_strip_inplace(stream, cap=n + 64)
    assert got["res"].n_nal == n
    # the generator's own table: NAL k starts right after its 3- or 4-byte start code
    sc = got["start"].astype(np.int64) - np.asarray(nal_off[:n], np.int64)
    assert ((sc == 3) | (sc == 4)).all()
    assert (got["end"] >= got["start"]).all() and (got["start"][1:] >= got["end"][:-1] + 3).all()
    assert np.array_equal(got["rbsp_off"], got["start"])  # in place: the RBSP begins where the NAL does
    want_len = np.diff(offs.astype(np.int64))
    tz = int(got["rbsp_len"][-1]) - int(want_len[-1])  # zero bytes after the last NAL stay with it
    assert 0 <= tz <= 2
    assert np.array_equal(got["rbsp_len"][:-1].astype(np.int64), want_len[:-1])
    body = got["body"]
    for k in range(n):
        a, b = int(offs[k]), int(offs[k + 1])
        o = int(got["rbsp_off"][k])
        assert np.array_equal(body[o:o + (b - a)], rbsp[a:b]), k
    assert not body[int(got["rbsp_off"][-1]) + int(want_len[-1]):][:tz].any()


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"H264GPU_SCAN7_TICKS": "1"}, {"H264GPU_SCAN7_TICKS": "2"},
                                 {"H264GPU_SCAN7_TICKS": "15"}, {"H264GPU_SCAN7_ROWS": "6"},
                                 {"H264GPU_SCAN7_TICKS": "15", "H264GPU_SCAN7_REGIONS": "7"}])
def test_gpu_ticket_counters_and_span_size(gpu, env, monkeypatch):
    """The span tickets of scan7_kernel come from 8 counters by default; any number of counters
    (1 = the single word of the first builds) and the 3 KiB-span instantiation give the same NAL
    table and RBSP bytes (the launcher reads the knobs at every launch).  Streams from one span to
    hundreds of regions, so that small grids (fewer warps than counters) are covered too."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(23)
    for it, hi in enumerate([40, 3000, 40000, 300000]):
        buf = S.gen_annexb(rng, int(rng.integers(1, 300)), lo=1, hi=hi, p_zero=3 / 16)
        check_per_nal(gpu.split_strip_inplace(buf), S.oracle_split_strip(buf), ("ticket counters", env, it))
    offs = L.synth_offsets(7, 24 << 20)
    stream, _ = L.synth_annexb(7, L.synth_payloads(7, offs), offs)
    check_per_nal(gpu.split_strip_inplace(stream, cap=len(offs) + 64), S.oracle_split_strip(stream), ("config2 24 MiB", env))


@pytest.mark.gpu
def test_gpu_sharded_merge(gpu):
    rng = np.random.default_rng(17)
    run = lambda buf, e, lo: gpu.split_strip_inplace(buf, edge=e, base=lo)
    for it in range(4):
        b = S.gen_annexb(rng, 200, 1, 20000)
        n = len(b)
        cuts = sorted(set(int(c) // 16 * 16 for c in rng.integers(16, n, 3)) - {0})
        check_merged(_merge_shards(run, b, cuts), S.oracle_split_strip(b), (it, cuts))


@pytest.mark.gpu
def test_gpu_one_context_streams_of_changing_length(gpu):
    """The workspace is one buffer whose arrays move with the stream length: a long stream, then
    shorter and again longer ones on the SAME context (a dense start-code pattern, so that the
    span words a shorter launch leaves behind have small non-zero top bits, like launch epochs)."""
    rng = np.random.default_rng(5)
    L._check(gpu.lib.h264gpu_scan_reserve(gpu.h, 48 << 20, 1 << 20), "h264gpu_scan_reserve")
    for it, mb in enumerate([24, 1, 9, 2, 17, 3, 12, 5]):
        buf = S.gen_annexb(rng, 1100 * mb, lo=1, hi=2000, p_zero=3 / 16)[:mb << 20]
        check_per_nal(gpu.split_strip_inplace(buf), S.oracle_split_strip(buf), ("changing length", it, mb))
