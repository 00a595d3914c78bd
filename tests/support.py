"""Shared test plumbing: ctypes bindings for the oracle (CPU restatement), the
compiled reference (when present), the SIMT-emulated kernels, and the product's
C-ABI library; plus seeded synthetic Annex-B generators.

Only tests/ (and bench.py's cpu_baseline leg / smoke()) may touch oracle/.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libh264_ref.so")
PORT_SO = os.path.join(ORACLE_DIR, "_ref", "liboracle.so")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libemu_kernels.so")

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)
NONE64 = 0xFFFFFFFFFFFFFFFF


class ShardEdge(C.Structure):
    _fields_ = [("left", C.c_uint8 * 2), ("right", C.c_uint8 * 2), ("has_left", C.c_uint8),
                ("has_right", C.c_uint8), ("assume_in", C.c_uint8), ("pad", C.c_uint8)]


class ScanResult(C.Structure):
    _fields_ = [("n_nal", C.c_uint64), ("rbsp_bytes", C.c_uint64),
                ("first_event_pos", C.c_uint64), ("head_bytes", C.c_uint64),
                ("first_event_is_sc", C.c_uint32), ("any_event", C.c_uint32),
                ("end_open", C.c_uint32), ("reserved", C.c_uint32)]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_oracle():
    """make port (always) and ref (when /root/reference is here)."""
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "port"])
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


def build_emu():
    srcs = [os.path.join(EMU_DIR, "emu_kernels.cpp"), os.path.join(EMU_DIR, "cuda_emu.h")]
    csrc = os.path.join(ROOT, "libh264_b200", "csrc")
    srcs += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    srcs.append(os.path.join(ROOT, "include", "h264gpu.h"))
    if _newer(EMU_SO, srcs):
        return
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DH264_EMU", "-fPIC", "-shared",
                           "-I" + EMU_DIR, "-I" + csrc, "-I" + os.path.join(ROOT, "include"),
                           "-o", EMU_SO, srcs[0]])


_cache = {}


def oracle():
    if "port" not in _cache:
        if not os.path.exists(PORT_SO) or os.path.isdir("/root/reference/src"):
            build_oracle()
        lib = C.CDLL(PORT_SO)
        lib.oracle_scan.restype = C.c_size_t
        lib.oracle_scan.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                    u64p, C.POINTER(C.c_int)]
        lib.oracle_strip.restype = C.c_size_t
        lib.oracle_strip.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        lib.oracle_insert.restype = C.c_size_t
        lib.oracle_insert.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        lib.oracle_insert_count.restype = C.c_size_t
        lib.oracle_insert_count.argtypes = [C.c_void_p, C.c_size_t]
        lib.oracle_split_strip.restype = C.c_size_t
        lib.oracle_split_strip.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_size_t, C.c_void_p, u64p, u64p]
        lib.oracle_frame.restype = C.c_size_t
        lib.oracle_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                     C.c_void_p]
        _cache["port"] = lib
    return _cache["port"]


def have_ref():
    if not os.path.exists(REF_SO) and os.path.isdir("/root/reference/src"):
        build_oracle()
    return os.path.exists(REF_SO)


def ref():
    if "ref" not in _cache:
        assert have_ref()
        lib = C.CDLL(REF_SO)
        lib.ref_scan.restype = C.c_size_t
        lib.ref_scan.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, u64p]
        lib.ref_strip.restype = C.c_size_t
        lib.ref_strip.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, u64p]
        lib.ref_insert.restype = C.c_size_t
        lib.ref_insert.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        lib.ref_mt_split_strip.restype = C.c_double
        lib.ref_mt_split_strip.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p,
                                           u64p, u64p]
        lib.ref_mt_insert.restype = C.c_double
        lib.ref_mt_insert.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, u64p]
        _cache["ref"] = lib
    return _cache["ref"]


def emu():
    if "emu" not in _cache:
        build_emu()
        lib = C.CDLL(EMU_SO)
        lib.emu_split_strip.restype = C.c_int
        lib.emu_split_strip.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                        C.c_void_p, C.c_int]
        _cache["emu"] = lib
    return _cache["emu"]


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ----------------------------------------------------------------------------
# oracle / reference wrappers returning numpy

def oracle_scan(buf):
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    cap = len(buf) // 3 + 2
    s = np.zeros(cap, np.uint64)
    e = np.zeros(cap, np.uint64)
    off = C.c_uint64(0)
    op = C.c_int(0)
    n = oracle().oracle_scan(ptr(buf), len(buf), ptr(s), ptr(e), cap, C.byref(off), C.byref(op))
    return s[:n].copy(), e[:n].copy(), off.value, op.value


def ref_scan(buf):
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    cap = len(buf) // 3 + 2
    s = np.zeros(cap, np.uint64)
    e = np.zeros(cap, np.uint64)
    off = C.c_uint64(0)
    n = ref().ref_scan(ptr(buf), len(buf), ptr(s), ptr(e), cap, C.byref(off))
    return s[:n].copy(), e[:n].copy(), off.value


def oracle_strip(nal):
    nal = np.ascontiguousarray(nal, dtype=np.uint8)
    out = np.zeros(len(nal) + 1, np.uint8)
    n = oracle().oracle_strip(ptr(nal), len(nal), ptr(out))
    return out[:n].copy()


def ref_strip(nal):
    nal = np.ascontiguousarray(nal, dtype=np.uint8)
    out = np.zeros(len(nal) + 1, np.uint8)
    off = C.c_uint64(0)
    n = ref().ref_strip(ptr(nal), len(nal), ptr(out), C.byref(off))
    return out[:n].copy(), off.value


def oracle_insert(rbsp):
    rbsp = np.ascontiguousarray(rbsp, dtype=np.uint8)
    out = np.zeros(len(rbsp) * 3 // 2 + 2, np.uint8)
    n = oracle().oracle_insert(ptr(rbsp), len(rbsp), ptr(out))
    return out[:n].copy()


def ref_insert(rbsp):
    rbsp = np.ascontiguousarray(rbsp, dtype=np.uint8)
    out = np.zeros(len(rbsp) * 3 // 2 + 2, np.uint8)
    n = ref().ref_insert(ptr(rbsp), len(rbsp), ptr(out), len(out))
    return out[:n].copy()


def oracle_split_strip(buf):
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    cap = len(buf) // 3 + 2
    s = np.zeros(cap, np.uint64)
    e = np.zeros(cap, np.uint64)
    r = np.zeros(cap + 1, np.uint64)
    out = np.zeros(len(buf) + 1, np.uint8)
    tot = C.c_uint64(0)
    off = C.c_uint64(0)
    n = oracle().oracle_split_strip(ptr(buf), len(buf), ptr(s), ptr(e), ptr(r), cap, ptr(out),
                                    C.byref(tot), C.byref(off))
    return dict(start=s[:n].copy(), end=e[:n].copy(), rbsp_off=r[:n + 1].copy(),
                rbsp=out[:tot.value].copy(), final_off=off.value)


def oracle_frame(rbsp, offs, sc_len=4):
    rbsp = np.ascontiguousarray(rbsp, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    n = len(offs) - 1
    out = np.zeros(len(rbsp) * 3 // 2 + 4 * n + 16, np.uint8)
    oo = np.zeros(n + 1, np.uint64)
    tot = oracle().oracle_frame(ptr(rbsp), ptr(offs), n, sc_len, ptr(out), ptr(oo))
    return out[:tot].copy(), oo


SCAN_GEN = 2  # which scan kernel generation the emulator tests exercise


def emu_split_strip(buf, strip=True, items=4, edge=None, base=0, gen=None):
    """items = 16-byte chunks per thread: 1, 2, 4 (gen 1) / 1, 2, 8 (gen 2; 4 maps to 8)."""
    g = gen or SCAN_GEN
    if g == 2 and items < 100:
        items = 100 + (8 if items == 4 else items)
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    cap = len(buf) // 3 + 2
    s = np.full(cap, NONE64, np.uint64)
    e = np.full(cap, NONE64, np.uint64)
    r = np.full(cap, NONE64, np.uint64)
    out = np.full(len(buf) + 16, 0xAA, np.uint8) if strip else None
    res = ScanResult()
    rc = emu().emu_split_strip(ptr(buf), len(buf), base, C.byref(edge) if edge else None, ptr(out),
                               ptr(s), ptr(e), ptr(r), cap, C.byref(res), items)
    assert rc == 0
    n = res.n_nal
    return dict(start=s[:n].copy(), end=e[:n].copy(), rbsp_off=r[:n].copy(),
                rbsp=out[:res.rbsp_bytes].copy() if strip else None,
                guard=out[res.rbsp_bytes:].copy() if strip else None, res=res)


# ----------------------------------------------------------------------------
# full-reader trace through the compiled reference (oracle/ref_harness.c)

TR_NALU_BEGIN, TR_NALU_END, TR_AU_END, TR_SPS, TR_PPS, TR_SLICE = 1, 2, 3, 4, 5, 6
TR_SLICE_DATA_BEGIN, TR_SLICE_DATA_END, TR_AUD, TR_SEI, TR_SLICE_PARAMS, TR_GROUP_MAP = 7, 8, 9, 10, 11, 12
MB_RECORD = np.dtype([("mb_addr", "<u4"), ("mb_type", "<u4"), ("hash", "<u8")])
SLICE_RESULT = np.dtype([("status", "<i4"), ("mb_count", "<u4"), ("end_bit", "<u8")])
PARAMS_SIZE = 56
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from libh264_b200 import SLICE_PARAMS  # noqa: E402


def ref_trace(stream, flags=1, mb_cap=None, log_cap=None):
    """h264_reader_parse(flags) of the UNMODIFIED reference with recording callbacks.
    Returns (events [(tag, payload bytes)], mb records, final off)."""
    lib = ref()
    lib.ref_trace_parse.restype = C.c_int
    lib.ref_trace_parse.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    mb_cap = mb_cap or max(1024, len(stream))
    log_cap = log_cap or (len(stream) * 4 + (32 << 20))
    log = np.zeros(log_cap, np.uint8)
    mbs = np.zeros(mb_cap, MB_RECORD)
    ll, mn, off = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    rc = lib.ref_trace_parse(ptr(stream), len(stream), flags, ptr(log), len(log), C.byref(ll),
                             ptr(mbs), mb_cap, C.byref(mn), C.byref(off))
    assert rc == 0, "ref_trace_parse: buffers too small"
    log = log[:ll.value]
    ev, i = [], 0
    while i < len(log):
        tag = int(log[i:i + 4].view(np.uint32)[0])
        n = int(log[i + 4:i + 8].view(np.uint32)[0])
        ev.append((tag, log[i + 8:i + 8 + n]))
        i += 8 + ((n + 7) & ~7)
    return ev, mbs[:mn.value].copy(), off.value


def ref_trace_syntax(stream, flags=1):
    """ref_trace plus the reference's full per-macroblock syntax elements (ctx->mb) as raw
    struct h264_mb_syntax blobs (libh264_b200/csrc/mb_syntax.h).  Returns (ev, mbs, syn bytes
    [n_mb, sizeof])."""
    lib = ref()
    lib.ref_trace_parse_syntax.restype = C.c_int
    lib.ref_trace_parse_syntax.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_size_t,
                                           C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t,
                                           C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_void_p]
    lib.ref_sizeof_mb_syntax.restype = C.c_uint32
    sz = lib.ref_sizeof_mb_syntax()
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    mb_cap = max(1024, len(stream))
    log = np.zeros(len(stream) * 4 + (32 << 20), np.uint8)
    mbs = np.zeros(mb_cap, MB_RECORD)
    syn = np.zeros((mb_cap, sz), np.uint8)
    ll, mn, off = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    rc = lib.ref_trace_parse_syntax(ptr(stream), len(stream), flags, ptr(log), len(log), C.byref(ll),
                                    ptr(mbs), mb_cap, C.byref(mn), C.byref(off), ptr(syn))
    assert rc == 0
    log = log[:ll.value]
    ev, i = [], 0
    while i < len(log):
        tag = int(log[i:i + 4].view(np.uint32)[0])
        n = int(log[i + 4:i + 8].view(np.uint32)[0])
        ev.append((tag, log[i + 8:i + 8 + n]))
        i += 8 + ((n + 7) & ~7)
    return ev, mbs[:mn.value].copy(), syn[:mn.value].copy()


def group_maps_from_trace(ev):
    """For streams with several slice groups: (params, maps) where maps is the concatenation of
    the reference's macroblock -> slice group map of every slice and params[i].row_state_off is
    slice i's offset into it (the group_maps argument of the CAVLC kernel)."""
    params = slice_params_from_trace(ev).view(SLICE_PARAMS).copy()
    maps, offs, k = [], 0, 0
    for t, p in ev:
        if t == TR_SLICE_PARAMS:
            k += 1
        elif t == TR_GROUP_MAP:
            params["row_state_off"][k - 1] = offs
            maps.append(np.frombuffer(bytes(p), np.uint8))
            offs += len(p)
    return params.view(np.uint8), (np.concatenate(maps) if maps else np.zeros(0, np.uint8))


def slice_params_from_trace(ev):
    """The packed h264gpu_slice_params blocks the reference's ctx held for every slice."""
    parts = [bytes(p[:PARAMS_SIZE]) for t, p in ev if t == TR_SLICE_PARAMS]
    return np.frombuffer(b"".join(parts), dtype=np.uint8).copy()


def emu_split_strip_inplace(buf, strip=True, cpt=8, edge=None, base=0, ev_cap=None):
    """Gen-5 kernel + finalize on the emulator.  Returns the per-NAL view: start, end, rbsp_off,
    rbsp_len, nal_rbsp (list of per-NAL byte arrays), the raw output buffer and the result."""
    lib = emu()
    lib.emu_split_strip_inplace.restype = C.c_int
    lib.emu_split_strip_inplace.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                            C.c_void_p, C.c_int, C.c_uint64]
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    cap = len(buf) // 3 + 2
    s = np.full(cap, NONE64, np.uint64)
    e = np.full(cap, NONE64, np.uint64)
    r = np.full(cap, NONE64, np.uint64)
    rl = np.full(cap, NONE64, np.uint64)
    guard = 64
    out = np.full(len(buf) + 2 * guard, 0xAA, np.uint8) if strip else None
    res = ScanResult()
    if strip:  # the kernel wants a 16-byte aligned output base: carve one out of the guarded buffer
        addr = out.ctypes.data
        lead = (-(addr + guard)) % 16 + guard
        out_ptr = C.c_void_p(addr + lead)
    else:
        lead, out_ptr = 0, None
    rc = lib.emu_split_strip_inplace(ptr(buf), len(buf), base, C.byref(edge) if edge else None, out_ptr,
                                     ptr(s), ptr(e), ptr(r), ptr(rl), cap, C.byref(res), cpt,
                                     cap if ev_cap is None else ev_cap)
    assert rc == 0
    n = res.n_nal
    body = out[lead:lead + len(buf)] if strip else None
    if strip:  # nothing may be written outside [0, len)
        assert (out[:lead] == 0xAA).all() and (out[lead + len(buf):] == 0xAA).all(), "write outside d_rbsp"
    per_nal = None
    if strip:
        per_nal = [body[int(r[k]):int(r[k]) + int(rl[k])].copy() for k in range(min(n, cap))]
    return dict(start=s[:n].copy(), end=e[:n].copy(), rbsp_off=r[:n].copy(), rbsp_len=rl[:n].copy(),
                nal_rbsp=per_nal, body=body, res=res)


def _emu_cavlc_steps(stream, params, n_records, full, group_maps):
    """K4 (cavlc_steps.cuh): the lane step function run serially per slice."""
    lib = emu()
    lib.emu_cavlc_steps.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    n = len(params) // PARAMS_SIZE
    recs = np.zeros(max(n_records, 1), MB_RECORD)
    res = np.zeros(max(n, 1), SLICE_RESULT)
    syn = np.full((max(n_records, 1), ref().ref_sizeof_mb_syntax()), 0xEE, np.uint8) if full else None
    gm = np.ascontiguousarray(group_maps, dtype=np.uint8) if group_maps is not None else None
    lib.emu_cavlc_steps(ptr(stream), len(stream), ptr(params), n, ptr(recs), ptr(res),
                        ptr(syn) if full else None, ptr(gm) if gm is not None else None)
    if full:
        return recs[:n_records], res[:n], syn[:n_records]
    return recs[:n_records], res[:n]


def emu_cavlc_parse_full(stream, params, n_records, group_maps=None):
    """The CAVLC parse with full per-macroblock records (struct h264_mb_syntax blobs)."""
    return _emu_cavlc_steps(stream, params, n_records, True, group_maps)


def emu_cavlc_parse(stream, params, n_records, group_maps=None):
    return _emu_cavlc_steps(stream, params, n_records, False, group_maps)


# ----------------------------------------------------------------------------
# synthetic streams (SURVEY.md §8d config 2 shape, small)

def gen_payloads(rng, n, lo=1, hi=4096, p_zero=3 / 16):
    """n random RBSP payloads: first byte a legal NAL header, rest iid with P(00)=p_zero."""
    sizes = np.exp(rng.uniform(np.log(lo), np.log(hi), n)).astype(np.int64)
    sizes = np.maximum(sizes, 1)
    offs = np.zeros(n + 1, np.uint64)
    offs[1:] = np.cumsum(sizes)
    tot = int(offs[-1])
    data = rng.integers(1, 256, tot, dtype=np.uint8)
    data[rng.random(tot) < p_zero] = 0
    # NAL header: forbidden_zero_bit 0, nal_ref_idc random, type 1..23
    hdr = (rng.integers(0, 4, n) << 5 | rng.integers(1, 24, n)).astype(np.uint8)
    data[offs[:-1].astype(np.int64)] = hdr
    return data, offs


def gen_annexb(rng, n, lo=1, hi=4096, p_zero=3 / 16, mixed_sc=True, trailing=True, lead=b""):
    """Valid Annex-B stream: payloads escaped by the oracle writer, 3/4-byte codes,
    optional 0-2 trailing zero bytes after a NAL."""
    data, offs = gen_payloads(rng, n, lo, hi, p_zero)
    parts = [np.frombuffer(lead, np.uint8)] if lead else []
    for k in range(n):
        p = data[int(offs[k]):int(offs[k + 1])]
        esc = oracle_insert(p)
        sc = 3 if (mixed_sc and rng.random() < 0.5) else 4
        parts.append(np.array([0] * (sc - 1) + [1], np.uint8))
        parts.append(esc)
        if trailing:
            tz = int(rng.integers(0, 3)) if rng.random() < 0.3 else 0
            if tz:
                parts.append(np.zeros(tz, np.uint8))
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)
