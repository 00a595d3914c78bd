"""Kernel LOGIC on the CPU SIMT emulator (tests/emu/) against the oracle: the very
source the GPU runs (libh264_b200/csrc/*.cuh) compiled with -DH264_EMU.  Covers
tile/warp/lane seams, the look-back fold, shard edges + host merge."""
import ctypes as C

import numpy as np
import pytest

import known_answers as KA
import libh264_b200 as L
import support as S

ALPHA = np.array([0, 0, 0, 1, 2, 3, 0xFF, 0x65, 0x77], np.uint8)


def check_scan(b, items=4, tag=""):
    o = S.oracle_split_strip(b)
    em = S.emu_split_strip(b, items=items)
    assert em["res"].n_nal == len(o["start"]), tag
    assert np.array_equal(em["start"], o["start"]), tag
    assert np.array_equal(em["end"], o["end"]), tag
    assert np.array_equal(em["rbsp_off"], o["rbsp_off"][:-1]), tag
    assert np.array_equal(em["rbsp"], o["rbsp"]), tag
    assert (em["guard"] == 0xAA).all(), tag  # nothing written past the RBSP end
    em2 = S.emu_split_strip(b, strip=False, items=items)
    assert np.array_equal(em2["start"], o["start"]) and np.array_equal(em2["end"], o["end"]), tag


@pytest.mark.parametrize("case", KA.SCAN)
def test_emu_scan_known_answers(case):
    hexs, exp, _ = case
    b = KA.hx(hexs)
    em = S.emu_split_strip(b)
    assert list(zip(em["start"].tolist(), em["end"].tolist())) == exp
    check_scan(b, 1)


def test_emu_scan_random_streams():
    rng = np.random.default_rng(1)
    for it in range(4):
        b = S.gen_annexb(rng, 40, 1, 6000)
        for items in (1, 2, 4):
            check_scan(b, items, ("valid", it, items))
    for it in range(12):
        b = rng.choice(ALPHA, int(rng.integers(1, 30000)))
        check_scan(b, int(rng.choice([1, 2, 4])), ("adv", it))


def test_emu_scan_events_on_every_seam():
    # a start code / EPB / zero run planted at every offset around lane (16 B),
    # warp item (512 B), warp region and tile (4096 B at ITEMS=1) boundaries
    rng = np.random.default_rng(3)
    base = rng.integers(4, 256, 3 * 4096 + 40, dtype=np.uint8)
    pats = [[0, 0, 1], [0, 0, 0, 1], [0, 0, 3, 0, 0, 3], [0, 0, 0], [0, 0, 0, 0, 0, 1, 0x65, 0, 0, 3]]
    for seam in (16, 512, 4096, 8192):
        for d in range(-6, 3):
            b = base.copy()
            b[0:4] = [0, 0, 1, 0x65]
            for i, p in enumerate(pats):
                pos = seam + d + (i * 1024 if seam < 4096 else 0)
                b[pos:pos + len(p)] = p
            check_scan(b, 1, ("seam", seam, d))


def test_emu_scan_lookback_fold():
    lib = S.emu()
    rng = np.random.default_rng(4)
    try:
        for every in (3, 40):
            lib.emu_set_prefix_every(every)
            check_scan(S.gen_annexb(rng, 300, 1, 3000), 1, ("fold", every))
            check_scan(rng.choice(ALPHA, 200000), 1, ("foldadv", every))
        # more than one 128-tile look-back window without any published prefix
        lib.emu_set_prefix_every(300)
        check_scan(S.gen_annexb(rng, 500, 1, 5000), 1, ("fold", 300))
    finally:
        lib.emu_set_prefix_every(1)


def _merge_shards(b, cuts, strip=True, items=1):
    """Run byte-range shards through the emulated kernel + the product's host merge."""
    lib = L.load_gpu_lib()
    n = len(b)
    cap = n // 3 + 2
    st = np.zeros(cap, np.uint64)
    en = np.zeros(cap, np.uint64)
    ro = np.zeros(cap, np.uint64)
    rbsp = []
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
        r = S.emu_split_strip(b[lo:hi], strip=strip, items=items, edge=e, base=lo)
        k = int(r["res"].n_nal)
        st[m.n_nal:m.n_nal + k] = r["start"]
        en[m.n_nal:m.n_nal + k] = r["end"]
        if strip:
            ro[m.n_nal:m.n_nal + k] = r["rbsp_off"]
        res = L.ScanResult.from_buffer_copy(bytes(r["res"]))
        skip, take = C.c_uint64(0), C.c_uint64(0)
        lib.h264gpu_merge_shard(C.byref(m), C.byref(res), S.ptr(st), S.ptr(en),
                                S.ptr(ro) if strip else None, cap, k, C.byref(skip), C.byref(take))
        if strip:
            rbsp.append(r["rbsp"][skip.value:skip.value + take.value])
    off = C.c_uint64(0)
    lib.h264gpu_merge_finish(C.byref(m), n, S.ptr(en), cap, C.byref(off))
    k = int(m.n_nal)
    return dict(start=st[:k], end=en[:k], rbsp_off=ro[:k],
                rbsp=np.concatenate(rbsp) if rbsp else np.zeros(0, np.uint8), final_off=off.value)


def test_emu_sharded_scan_merges_to_whole():
    rng = np.random.default_rng(5)
    for it in range(10):
        b = S.gen_annexb(rng, 30, 1, 3000) if it % 2 else rng.choice(ALPHA, int(rng.integers(200, 20000)))
        n = len(b)
        ncut = int(rng.integers(1, 5))
        cuts = sorted(set(int(c) // 16 * 16 for c in rng.integers(16, max(n, 17), ncut)) - {0})
        cuts = [c for c in cuts if c < n]
        o = S.oracle_split_strip(b)
        g = _merge_shards(b, cuts)
        assert np.array_equal(g["start"], o["start"]), (it, cuts)
        assert np.array_equal(g["end"], o["end"]), (it, cuts)
        assert np.array_equal(g["rbsp"], o["rbsp"]), (it, cuts)
        assert np.array_equal(g["rbsp_off"], o["rbsp_off"][:-1]), (it, cuts)
        assert g["final_off"] == o["final_off"], (it, cuts)


def test_emu_shard_cut_inside_start_code_and_gap():
    # cuts landing inside a start code, inside trailing zeros, and in inter-NAL garbage
    b = np.array([0x41] * 13 + [0, 0, 0, 1, 0x65, 9, 9, 0, 0, 0, 7, 7, 7, 7, 7, 0, 0, 1, 0x41] + [5] * 29, np.uint8)
    o = S.oracle_split_strip(b)
    for cut in (16, 32, 48):
        g = _merge_shards(b, [cut])
        assert np.array_equal(g["start"], o["start"]) and np.array_equal(g["end"], o["end"])
        assert np.array_equal(g["rbsp"], o["rbsp"]) and g["final_off"] == o["final_off"]


# ---- K3: EPB insert + framing ----------------------------------------------

def emu_frame(rbsp, offs, sc_len, items):
    lib = S.emu()
    lib.emu_frame.restype = C.c_int
    lib.emu_frame.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p,
                              C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
    rbsp = np.ascontiguousarray(rbsp, np.uint8)
    offs = np.ascontiguousarray(offs, np.uint64)
    n = len(offs) - 1
    cap = len(rbsp) * 3 // 2 + 4 * n + 64
    out = np.full(cap, 0xAA, np.uint8)
    oo = np.full(n + 1, S.NONE64, np.uint64)
    tot = C.c_uint64(0)
    lib.emu_frame(S.ptr(rbsp), len(rbsp), S.ptr(offs), n, sc_len, S.ptr(out), cap, S.ptr(oo),
                  C.byref(tot), items)
    return out, oo, tot.value


def expected_frame(rbsp, offs, sc_len):
    if sc_len:
        return S.oracle_frame(rbsp, offs, sc_len)
    parts = [S.oracle_insert(rbsp[int(offs[k]):int(offs[k + 1])]) for k in range(len(offs) - 1)]
    exp = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    eoo = np.zeros(len(offs), np.uint64)
    eoo[1:] = np.cumsum([len(p) for p in parts])
    return exp, eoo


def check_frame(rbsp, offs, sc_len, items, tag=""):
    exp, eoo = expected_frame(rbsp, offs, sc_len)
    out, oo, tot = emu_frame(rbsp, offs, sc_len, items)
    assert tot == len(exp), tag
    assert np.array_equal(out[:tot], exp), tag
    assert (out[tot:] == 0xAA).all(), tag
    assert np.array_equal(oo, eoo), tag


def test_emu_frame_known_answer():
    p = KA.hx(KA.INSERT_IN)
    out, oo, tot = emu_frame(p, np.array([0, len(p)], np.uint64), 0, 1)
    assert bytes(out[:tot]) == bytes(KA.hx(KA.INSERT_OUT))


def test_emu_frame_random_and_adversarial():
    lib = S.emu()
    rng = np.random.default_rng(5)
    try:
        for every in (1, 3, 50):
            lib.emu_set_prefix_every(every)
            for it in range(4):
                data, offs = S.gen_payloads(rng, int(rng.integers(1, 60)), 1, 5000)
                check_frame(data, offs, 4 if it % 2 else 3, int(rng.choice([1, 2, 4])), ("rand", every, it))
            for it in range(8):
                tot = int(rng.integers(0, 40000))
                data = rng.choice(np.array([0, 0, 0, 0, 1, 2, 3, 4, 0xFF], np.uint8), tot)
                n = int(rng.integers(0, 30))
                cuts = np.sort(rng.integers(0, tot + 1, n)) if n else np.zeros(0, np.int64)
                offs = np.concatenate([[0], cuts, [tot]]).astype(np.uint64)
                check_frame(data, offs, int(rng.choice([0, 3, 4])), int(rng.choice([1, 2, 4])), ("adv", every, it))
            for it in range(3):  # zero runs spanning whole tiles, unstaged (50 % growth) tiles
                tot = int(rng.integers(9000, 30000))
                data = np.zeros(tot, np.uint8)
                for p in rng.integers(0, tot, 3):
                    data[p] = rng.choice([1, 3, 7])
                check_frame(data, np.array([0, tot // 3, tot // 3, tot], np.uint64), 4, 1, ("zeros", every, it))
    finally:
        lib.emu_set_prefix_every(1)
    check_frame(np.zeros(0, np.uint8), np.array([0, 0, 0], np.uint64), 4, 1, "empty payloads only")
    check_frame(np.zeros(0, np.uint8), np.array([0], np.uint64), 4, 1, "no payloads")


# ---- K3 gen 6 (frame6_kernel): items 61/62/64/68 = 1/2/4/8 rows per warp -------------------

def test_emu_frame6_known_answer():
    p = KA.hx(KA.INSERT_IN)
    for items in (61, 68):
        out, oo, tot = emu_frame(p, np.array([0, len(p)], np.uint64), 0, items)
        assert bytes(out[:tot]) == bytes(KA.hx(KA.INSERT_OUT))


def test_emu_frame6_random_and_adversarial():
    lib = S.emu()
    rng = np.random.default_rng(55)
    g6 = [61, 62, 64, 68]
    try:
        for every in (1, 3, 50):
            lib.emu_set_prefix_every(every)
            for it in range(4):
                data, offs = S.gen_payloads(rng, int(rng.integers(1, 60)), 1, 5000)
                check_frame(data, offs, 4 if it % 2 else 3, int(rng.choice(g6)), ("rand6", every, it))
            for it in range(8):
                tot = int(rng.integers(0, 40000))
                data = rng.choice(np.array([0, 0, 0, 0, 1, 2, 3, 4, 0xFF], np.uint8), tot)
                n = int(rng.integers(0, 30))
                cuts = np.sort(rng.integers(0, tot + 1, n)) if n else np.zeros(0, np.int64)
                offs = np.concatenate([[0], cuts, [tot]]).astype(np.uint64)
                check_frame(data, offs, int(rng.choice([0, 3, 4])), int(rng.choice(g6)), ("adv6", every, it))
            for it in range(3):  # zero runs spanning whole rows, spans and tiles
                tot = int(rng.integers(9000, 70000))
                data = np.zeros(tot, np.uint8)
                for p in rng.integers(0, tot, 3):
                    data[p] = rng.choice([1, 3, 7])
                check_frame(data, np.array([0, tot // 3, tot // 3, tot], np.uint64), 4, int(rng.choice([61, 68])),
                            ("zeros6", every, it))
    finally:
        lib.emu_set_prefix_every(1)
    for items in (61, 68):
        check_frame(np.zeros(0, np.uint8), np.array([0, 0, 0], np.uint64), 4, items, "empty payloads only")
        check_frame(np.zeros(0, np.uint8), np.array([0], np.uint64), 4, items, "no payloads")


def test_emu_frame6_stream_shaped_and_dense_starts():
    rng = np.random.default_rng(56)
    for it in range(3):  # the bench's byte statistics, a handful of payloads over several 32 KiB tiles
        tot = int(rng.integers(100000, 250000))
        data = rng.integers(0, 256, tot).astype(np.uint8)
        data[rng.random(tot) < 0.1875] = 0
        cuts = np.sort(rng.integers(0, tot + 1, int(rng.integers(1, 8))))
        offs = np.concatenate([[0], cuts, [tot]]).astype(np.uint64)
        check_frame(data, offs, 4, 68, ("stream6", it))
        check_frame(data, offs, 3, 62, ("stream6", it))
    for it in range(3):  # payloads of a few bytes: every row goes byte by byte
        tot = int(rng.integers(3000, 20000))
        data = rng.choice(np.array([0, 0, 0, 1, 3, 0xFF], np.uint8), tot)
        cuts = np.sort(rng.integers(0, tot + 1, tot // 3))
        offs = np.concatenate([[0], cuts, [tot]]).astype(np.uint64)
        check_frame(data, offs, 4, int(rng.choice([61, 64])), ("dense6", it))


def test_emu_frame6_output_capacity_is_respected():
    """Nothing is written past out_cap (the C-ABI reports -ENOBUFS from the total)."""
    lib = S.emu()
    rng = np.random.default_rng(57)
    data = rng.choice(np.array([0, 0, 0, 1, 3, 0xFF], np.uint8), 20000)
    offs = np.array([0, 7000, 20000], np.uint64)
    exp, eoo = S.oracle_frame(data, offs, 4)
    for cap in (0, 5, 4099, len(exp) - 1):
        out = np.full(len(exp) + 64, 0xAA, np.uint8)
        oo = np.full(3, S.NONE64, np.uint64)
        tot = C.c_uint64(0)
        lib.emu_frame(S.ptr(data), len(data), S.ptr(offs), 2, 4, S.ptr(out), cap, S.ptr(oo), C.byref(tot), 62)
        assert tot.value == len(exp)
        assert np.array_equal(out[:cap], exp[:cap]) and (out[cap:] == 0xAA).all(), cap


# ---- K3 gen 7 (frame7_kernel): items 71/72/74/76/78 = 1/2/4/6/8 rows per warp-owned span, bytes
# staged twice (one buffer per warp); 81/.../88: staged once (two buffers per warp) ----

G7 = [71, 72, 74, 76, 78, 81, 82, 84, 86, 88]


def test_emu_frame7_known_answer():
    p = KA.hx(KA.INSERT_IN)
    for items in G7:
        out, oo, tot = emu_frame(p, np.array([0, len(p)], np.uint64), 0, items)
        assert bytes(out[:tot]) == bytes(KA.hx(KA.INSERT_OUT))


def test_emu_frame7_random_and_adversarial():
    rng = np.random.default_rng(75)
    for it in range(10):
        data, offs = S.gen_payloads(rng, int(rng.integers(1, 60)), 1, 5000)
        check_frame(data, offs, 4 if it % 2 else 3, int(rng.choice(G7)), ("rand7", it))
    for it in range(20):
        tot = int(rng.integers(0, 40000))
        data = rng.choice(np.array([0, 0, 0, 0, 1, 2, 3, 4, 0xFF], np.uint8), tot)
        n = int(rng.integers(0, 30))
        cuts = np.sort(rng.integers(0, tot + 1, n)) if n else np.zeros(0, np.int64)
        offs = np.concatenate([[0], cuts, [tot]]).astype(np.uint64)
        check_frame(data, offs, int(rng.choice([0, 3, 4])), int(rng.choice(G7)), ("adv7", it))
    for it in range(8):  # zero runs spanning whole rows and spans (the run entering a span is longer than its 16-byte halo)
        tot = int(rng.integers(9000, 70000))
        data = np.zeros(tot, np.uint8)
        for p in rng.integers(0, tot, 3):
            data[p] = rng.choice([1, 3, 7])
        check_frame(data, np.array([0, tot // 3, tot // 3, tot], np.uint64), 4, int(rng.choice(G7)), ("zeros7", it))
    for items in (71, 78, 81, 88):
        check_frame(np.zeros(0, np.uint8), np.array([0, 0, 0], np.uint64), 4, items, "empty payloads only")
        check_frame(np.zeros(0, np.uint8), np.array([0], np.uint64), 4, items, "no payloads")
        check_frame(np.array([0, 0, 1], np.uint8), np.array([0, 3], np.uint64), 3, items, "three bytes")


def test_emu_frame7_groups_and_supergroups():
    """More than 1024 spans of 512 bytes: every level of the chain (span words, group sums, supergroup
    sums and prefixes) carries inserts; payload starts at span seams."""
    rng = np.random.default_rng(76)
    tot = 512 * 2100 + 77
    data = rng.choice(np.array([0, 0, 0, 1, 3, 0x55, 0xFF], np.uint8), tot)
    cuts = np.sort(np.concatenate([rng.integers(0, tot + 1, 40), [512 * 1024, 512 * 1024 + 1, 512 * 33]]))
    offs = np.concatenate([[0], cuts, [tot]]).astype(np.uint64)
    check_frame(data, offs, 4, 71, "chain7")
    check_frame(data, offs, 4, 81, "chain7")
    data = rng.integers(0, 256, tot).astype(np.uint8)
    data[rng.random(tot) < 0.1875] = 0
    check_frame(data, np.array([0, tot], np.uint64), 3, 71, "chain7 one payload")
    check_frame(data, np.array([0, tot], np.uint64), 3, 81, "chain7 one payload")


def test_emu_frame7_stream_shaped_and_dense_starts():
    rng = np.random.default_rng(77)
    for it in range(3):  # the bench's byte statistics, a handful of payloads
        tot = int(rng.integers(100000, 250000))
        data = rng.integers(0, 256, tot).astype(np.uint8)
        data[rng.random(tot) < 0.1875] = 0
        cuts = np.sort(rng.integers(0, tot + 1, int(rng.integers(1, 8))))
        offs = np.concatenate([[0], cuts, [tot]]).astype(np.uint64)
        check_frame(data, offs, 4, 78, ("stream7", it))
        check_frame(data, offs, 3, 76, ("stream7", it))
        check_frame(data, offs, 4, 88, ("stream7", it))
    for it in range(4):  # payloads of a few bytes: every row goes byte by byte, more starts than staged
        tot = int(rng.integers(3000, 20000))
        data = rng.choice(np.array([0, 0, 0, 1, 3, 0xFF], np.uint8), tot)
        cuts = np.sort(rng.integers(0, tot + 1, tot // 3))
        offs = np.concatenate([[0], cuts, [tot]]).astype(np.uint64)
        check_frame(data, offs, 4, int(rng.choice(G7)), ("dense7", it))


def test_emu_frame7_output_capacity_is_respected():
    lib = S.emu()
    rng = np.random.default_rng(78)
    data = rng.choice(np.array([0, 0, 0, 1, 3, 0xFF], np.uint8), 20000)
    offs = np.array([0, 7000, 20000], np.uint64)
    exp, eoo = S.oracle_frame(data, offs, 4)
    for cap in (0, 5, 4099, len(exp) - 1):
        out = np.full(len(exp) + 64, 0xAA, np.uint8)
        oo = np.full(3, S.NONE64, np.uint64)
        tot = C.c_uint64(0)
        lib.emu_frame(S.ptr(data), len(data), S.ptr(offs), 2, 4, S.ptr(out), cap, S.ptr(oo), C.byref(tot), 72)
        assert tot.value == len(exp)
        assert np.array_equal(out[:cap], exp[:cap]) and (out[cap:] == 0xAA).all(), cap


# ---- K3 gen 8 (frame8_kernel): the tiles of gen 6 with the chain of gen 7; items 91/92/94/98 ----

G8 = [91, 92, 94, 98]


def test_emu_frame8_known_answer_and_edges():
    p = KA.hx(KA.INSERT_IN)
    for items in G8:
        out, oo, tot = emu_frame(p, np.array([0, len(p)], np.uint64), 0, items)
        assert bytes(out[:tot]) == bytes(KA.hx(KA.INSERT_OUT))
    for items in (91, 98):
        check_frame(np.zeros(0, np.uint8), np.array([0, 0, 0], np.uint64), 4, items, "empty payloads only")
        check_frame(np.zeros(0, np.uint8), np.array([0], np.uint64), 4, items, "no payloads")
        check_frame(np.array([0, 0, 1], np.uint8), np.array([0, 3], np.uint64), 3, items, "three bytes")


def test_emu_frame8_random_and_adversarial():
    rng = np.random.default_rng(85)
    for it in range(8):
        data, offs = S.gen_payloads(rng, int(rng.integers(1, 60)), 1, 5000)
        check_frame(data, offs, 4 if it % 2 else 3, int(rng.choice(G8)), ("rand8", it))
    for it in range(16):
        tot = int(rng.integers(0, 60000))
        data = rng.choice(np.array([0, 0, 0, 0, 1, 2, 3, 4, 0xFF], np.uint8), tot)
        n = int(rng.integers(0, 30))
        cuts = np.sort(rng.integers(0, tot + 1, n)) if n else np.zeros(0, np.int64)
        offs = np.concatenate([[0], cuts, [tot]]).astype(np.uint64)
        check_frame(data, offs, int(rng.choice([0, 3, 4])), int(rng.choice(G8)), ("adv8", it))
    for it in range(6):  # zero runs spanning whole rows, spans and tiles
        tot = int(rng.integers(9000, 90000))
        data = np.zeros(tot, np.uint8)
        for p in rng.integers(0, tot, 3):
            data[p] = rng.choice([1, 3, 7])
        check_frame(data, np.array([0, tot // 3, tot // 3, tot], np.uint64), 4, int(rng.choice(G8)), ("zeros8", it))


def test_emu_frame8_chain_levels_and_stream_shape():
    """More than 1024 tiles of 4 KiB (every level of the chain), the bench's byte statistics over
    32 KiB tiles, payloads of a few bytes."""
    rng = np.random.default_rng(86)
    tot = 4096 * 1100 + 77
    data = rng.integers(0, 256, tot).astype(np.uint8)
    data[rng.random(tot) < 0.1875] = 0
    cuts = np.sort(np.concatenate([rng.integers(0, tot + 1, 40), [4096 * 1024, 4096 * 1024 + 1, 4096 * 33]]))
    offs = np.concatenate([[0], cuts, [tot]]).astype(np.uint64)
    check_frame(data, offs, 4, 91, "chain8")
    for it in range(2):
        tot = int(rng.integers(100000, 250000))
        data = rng.integers(0, 256, tot).astype(np.uint8)
        data[rng.random(tot) < 0.1875] = 0
        cuts = np.sort(rng.integers(0, tot + 1, int(rng.integers(1, 8))))
        offs = np.concatenate([[0], cuts, [tot]]).astype(np.uint64)
        check_frame(data, offs, 4, 98, ("stream8", it))
        check_frame(data, offs, 3, 92, ("stream8", it))
    for it in range(2):
        tot = int(rng.integers(3000, 20000))
        data = rng.choice(np.array([0, 0, 0, 1, 3, 0xFF], np.uint8), tot)
        cuts = np.sort(rng.integers(0, tot + 1, tot // 3))
        offs = np.concatenate([[0], cuts, [tot]]).astype(np.uint64)
        check_frame(data, offs, 4, int(rng.choice([91, 94])), ("dense8", it))


def test_emu_frame8_output_capacity_is_respected():
    lib = S.emu()
    rng = np.random.default_rng(88)
    data = rng.choice(np.array([0, 0, 0, 1, 3, 0xFF], np.uint8), 20000)
    offs = np.array([0, 7000, 20000], np.uint64)
    exp, eoo = S.oracle_frame(data, offs, 4)
    for cap in (0, 5, 4099, len(exp) - 1):
        out = np.full(len(exp) + 64, 0xAA, np.uint8)
        oo = np.full(3, S.NONE64, np.uint64)
        tot = C.c_uint64(0)
        lib.emu_frame(S.ptr(data), len(data), S.ptr(offs), 2, 4, S.ptr(out), cap, S.ptr(oo), C.byref(tot), 92)
        assert tot.value == len(exp)
        assert np.array_equal(out[:cap], exp[:cap]) and (out[cap:] == 0xAA).all(), cap


def test_emu_frame_then_scan_round_trip():
    rng = np.random.default_rng(8)
    data, offs = S.gen_payloads(rng, 50, 2, 3000)
    data[offs[1:].astype(np.int64) - 1] = 0x80  # RBSP never ends in 00
    out, oo, tot = emu_frame(data, offs, 4, 2)
    back = S.emu_split_strip(out[:tot], items=2)
    assert np.array_equal(back["rbsp"], data)
    assert np.array_equal(back["rbsp_off"], offs[:-1])
    assert np.array_equal(back["start"], oo[:-1] + 4)


# ---- gen-2 scan kernel: dense / degenerate inputs (list overflow, unlisted paths) -------

def test_emu_scan2_dense_and_degenerate_inputs():
    rng = np.random.default_rng(7)
    cases = []
    b = np.tile(np.array([0, 0, 3], np.uint8), 30000)
    b[:4] = [0, 0, 1, 0x65]
    cases.append(("dense epb", b, (1, 4)))          # every chunk deletes: lists overflow
    cases.append(("dense sc", np.tile(np.array([0, 0, 1], np.uint8), 30000), (1, 4)))
    cases.append(("dense sc2", np.tile(np.array([0, 0, 1, 0x41, 0, 0, 0, 9], np.uint8), 12000), (1, 4)))
    cases.append(("zeros", np.zeros(100000, np.uint8), (1, 4)))
    b = np.zeros(100000, np.uint8)
    b[5000:5004] = [0, 0, 1, 0x65]
    b[5004:9000] = rng.integers(1, 256, 3996)
    b[70000:70003] = [0, 0, 1]
    b[70003:70100] = 7
    cases.append(("islands", b, (1, 2, 4)))
    b = rng.integers(4, 256, 150000).astype(np.uint8)  # long inter-NAL garbage over tiles
    b[:4] = [0, 0, 1, 0x65]
    b[1000:1003] = 0
    b[120000:120004] = [0, 0, 0, 1]
    cases.append(("garbage", b, (1, 2, 4)))
    cases.append(("no start code", rng.integers(1, 256, 70000).astype(np.uint8), (4,)))
    for it in range(2):
        cases.append(("big%d" % it, S.gen_annexb(rng, 60, 1, 20000), (2, 4)))
    for tag, b, items in cases:
        for it in items:
            check_scan(b, it, (tag, it))
