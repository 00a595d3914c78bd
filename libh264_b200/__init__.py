"""libh264_b200 — B200-native H.264 bitstream front end (libh264-compatible).

The product is two shared libraries built in-tree by libh264_b200.build:

  libh264gpu.so  hand-written sm_100a CUDA stages behind a C-ABI (include/h264gpu.h)
  libh264.so     libh264's C API (include/h264/*.h) in host C on top of it

This Python package is only the harness-side binding (ctypes) used by bench.py,
the tests and __graft_entry__; it adds no compute of its own and has NO CPU
fallback: loading fails loudly when libh264gpu.so is missing, and every stage
call fails when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))

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


class Merge(C.Structure):
    _fields_ = [("n_nal", C.c_uint64), ("rbsp_bytes", C.c_uint64), ("open", C.c_uint32),
                ("shards", C.c_uint32)]


class SliceParams(C.Structure):
    """struct h264gpu_slice_params (include/h264gpu_slice.h)."""
    _fields_ = [("nal_off", C.c_uint64), ("nal_len", C.c_uint32), ("data_bit_off", C.c_uint32),
                ("first_mb_in_slice", C.c_uint32), ("mb_out_off", C.c_uint32),
                ("mb_out_cap", C.c_uint32), ("row_state_off", C.c_uint32),
                ("pic_width_in_mbs", C.c_uint16), ("pic_height_in_mbs", C.c_uint16),
                ("slice_type", C.c_uint8), ("chroma_array_type", C.c_uint8),
                ("bit_depth_luma", C.c_uint8), ("bit_depth_chroma", C.c_uint8),
                ("transform_8x8_mode_flag", C.c_uint8), ("direct_8x8_inference_flag", C.c_uint8),
                ("num_ref_idx_l0_active_minus1", C.c_uint8), ("num_ref_idx_l1_active_minus1", C.c_uint8),
                ("field_pic_flag", C.c_uint8), ("mbaff_frame_flag", C.c_uint8),
                ("entropy_coding_mode_flag", C.c_uint8), ("num_slice_groups_minus1", C.c_uint8),
                ("cabac_init_idc", C.c_uint8), ("slice_qp", C.c_int8), ("reserved", C.c_uint8 * 2)]


MB_RECORD = np.dtype([("mb_addr", "<u4"), ("mb_type", "<u4"), ("hash", "<u8")])
SLICE_RESULT = np.dtype([("status", "<i4"), ("mb_count", "<u4"), ("end_bit", "<u8")])
SLICE_PARAMS = np.dtype([("nal_off", "<u8"), ("nal_len", "<u4"), ("data_bit_off", "<u4"),
                         ("first_mb_in_slice", "<u4"), ("mb_out_off", "<u4"), ("mb_out_cap", "<u4"),
                         ("row_state_off", "<u4"), ("pic_width_in_mbs", "<u2"),
                         ("pic_height_in_mbs", "<u2"), ("slice_type", "u1"), ("chroma_array_type", "u1"),
                         ("bit_depth_luma", "u1"), ("bit_depth_chroma", "u1"),
                         ("transform_8x8_mode_flag", "u1"), ("direct_8x8_inference_flag", "u1"),
                         ("num_ref_idx_l0_active_minus1", "u1"), ("num_ref_idx_l1_active_minus1", "u1"),
                         ("field_pic_flag", "u1"), ("mbaff_frame_flag", "u1"),
                         ("entropy_coding_mode_flag", "u1"), ("num_slice_groups_minus1", "u1"),
                         ("cabac_init_idc", "u1"), ("slice_qp", "i1"), ("reserved", "u1", (2,)),
                         ("pad", "u1", (4,))])
assert SLICE_PARAMS.itemsize == C.sizeof(SliceParams) == 56

SLICE_SYMBOLS = ["h264gpu_cavlc_parse_dev", "h264gpu_cavlc_parse_host",
                 "h264gpu_cabac_parse_dev", "h264gpu_cabac_parse_host",
                 "h264gpu_reader_parse_cavlc", "h264gpu_reader_parse_cabac",
                 "h264gpu_reader_parse_slices", "h264gpu_cavlc_parse_full_dev",
                 "h264gpu_cavlc_parse_fmo_dev", "h264gpu_fmo_mb_map", "h264gpu_reader_set_group_maps",
                 "h264gpu_conceal_slices_dev", "h264gpu_conceal_slices_host",
                 "h264gpu_patch_slice_headers_dev", "h264gpu_patch_slice_headers"]

# every symbol include/h264gpu.h declares (checked by tests/test_abi.py)
GPU_SYMBOLS = [
    "h264gpu_device_count", "h264gpu_create", "h264gpu_destroy", "h264gpu_device",
    "h264gpu_version", "h264gpu_launch_count", "h264gpu_malloc", "h264gpu_free",
    "h264gpu_host_alloc", "h264gpu_host_free", "h264gpu_memcpy_h2d", "h264gpu_memcpy_d2h",
    "h264gpu_sync", "h264gpu_split_strip_dev", "h264gpu_split_strip_inplace_dev", "h264gpu_scan_reserve", "h264gpu_merge_init", "h264gpu_merge_shard", "h264gpu_merge_shard_inplace",
    "h264gpu_merge_finish", "h264gpu_split_strip_host", "h264gpu_frame_dev",
    "h264gpu_frame_host", "h264gpu_timer_create", "h264gpu_timer_destroy",
    "h264gpu_timer_start", "h264gpu_timer_stop", "h264gpu_timer_elapsed_ms",
    "h264gpu_reader_scan", "h264gpu_reader_upload", "h264gpu_reader_resident",
    "h264gpu_stream_create", "h264gpu_stream_destroy",
    "h264gpu_byte_stream_to_avcc_dev", "h264gpu_byte_stream_to_avcc_host",
]

_libs = {}


def load_gpu_lib():
    """dlopen libh264gpu.so (the CUDA extension).  Raises if it is not built."""
    if "gpu" not in _libs:
        if not os.path.exists(_build.GPU_SO):
            raise RuntimeError(
                "libh264gpu.so is not built (run `python -c 'import __graft_entry__ as g; "
                "g.build()'`); there is no CPU fallback for the CUDA stages")
        lib = C.CDLL(_build.GPU_SO)
        vp, u64, i = C.c_void_p, C.c_uint64, C.c_int
        lib.h264gpu_version.restype = C.c_char_p
        lib.h264gpu_launch_count.restype = u64
        lib.h264gpu_launch_count.argtypes = [vp]
        lib.h264gpu_create.argtypes = [i, C.POINTER(vp)]
        lib.h264gpu_destroy.argtypes = [vp]
        lib.h264gpu_malloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        lib.h264gpu_free.argtypes = [vp, vp]
        lib.h264gpu_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        lib.h264gpu_host_free.argtypes = [vp, vp]
        lib.h264gpu_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t, vp]
        lib.h264gpu_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t, vp]
        lib.h264gpu_sync.argtypes = [vp, vp]
        lib.h264gpu_split_strip_dev.argtypes = [vp, vp, u64, u64, vp, vp, vp, vp, vp, u64, vp, vp]
        lib.h264gpu_split_strip_host.argtypes = [vp, vp, u64, vp, vp, vp, vp, u64p, u64p, u64p]
        lib.h264gpu_split_strip_inplace_dev.argtypes = [vp, vp, u64, u64, vp, vp, vp, vp, vp, vp, u64, vp, vp]
        lib.h264gpu_scan_reserve.argtypes = [vp, u64, u64]
        lib.h264gpu_merge_init.restype = None
        lib.h264gpu_merge_init.argtypes = [vp]
        lib.h264gpu_merge_shard.argtypes = [vp, vp, vp, vp, vp, u64, u64, u64p, u64p]
        lib.h264gpu_merge_shard_inplace.argtypes = [vp, vp, vp, vp, u64, u64p]
        lib.h264gpu_merge_finish.argtypes = [vp, u64, vp, u64, u64p]
        lib.h264gpu_frame_dev.argtypes = [vp, vp, vp, u64, i, vp, u64, vp, vp, vp]
        lib.h264gpu_frame_host.argtypes = [vp, vp, vp, u64, i, vp, u64, vp, u64p]
        lib.h264gpu_timer_create.argtypes = [vp, C.POINTER(vp)]
        lib.h264gpu_timer_destroy.argtypes = [vp, vp]
        lib.h264gpu_timer_start.argtypes = [vp, vp, vp]
        lib.h264gpu_timer_stop.argtypes = [vp, vp, vp]
        lib.h264gpu_timer_elapsed_ms.argtypes = [vp, vp, C.POINTER(C.c_float)]
        lib.h264gpu_cavlc_parse_dev.argtypes = [vp, vp, u64, vp, C.c_uint32, vp, vp, vp]
        lib.h264gpu_cavlc_parse_host.argtypes = [vp, vp, u64, vp, C.c_uint32, vp, u64, vp]
        lib.h264gpu_cabac_parse_dev.argtypes = [vp, vp, u64, vp, C.c_uint32, vp, vp, vp]
        lib.h264gpu_cabac_parse_host.argtypes = [vp, vp, u64, vp, C.c_uint32, vp, u64, vp]
        _libs["gpu"] = lib
    return _libs["gpu"]


def load_synth_lib():
    if "synth" not in _libs:
        _build.build_synth()
        lib = C.CDLL(_build.SYNTH_SO)
        vp, u64, i = C.c_void_p, C.c_uint64, C.c_int
        lib.synth_sizes.restype = u64
        lib.synth_sizes.argtypes = [u64, u64, C.c_uint32, C.c_uint32, vp, u64]
        lib.synth_fill.restype = None
        lib.synth_fill.argtypes = [u64, vp, u64, vp, i]
        lib.synth_frame.restype = u64
        lib.synth_frame.argtypes = [u64, vp, vp, u64, i, i, vp, u64, vp, i]
        lib.synth_video.restype = u64
        lib.synth_video.argtypes = [vp, vp, u64, u64p, u64p, vp, u64]
        _libs["synth"] = lib
    return _libs["synth"]


def _check(rc, what):
    if rc != 0:
        raise OSError(-rc, "%s failed: %s" % (what, os.strerror(-rc) if rc < 0 else rc))


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a)


class DeviceBuffer:
    """Device allocation owned by a Gpu context (plain cudaMalloc through the C-ABI)."""

    def __init__(self, gpu, nbytes):
        self.gpu, self.nbytes = gpu, int(nbytes)
        p = C.c_void_p()
        _check(gpu.lib.h264gpu_malloc(gpu.h, self.nbytes, C.byref(p)), "h264gpu_malloc")
        self.ptr = p.value

    def free(self):
        if self.ptr:
            self.gpu.lib.h264gpu_free(self.gpu.h, C.c_void_p(self.ptr))
            self.ptr = None

    def upload(self, arr, offset=0, stream=None):
        arr = np.ascontiguousarray(arr)
        assert offset + arr.nbytes <= self.nbytes
        _check(self.gpu.lib.h264gpu_memcpy_h2d(self.gpu.h, C.c_void_p(self.ptr + offset), _ptr(arr),
                                               arr.nbytes, stream), "h2d")
        self.gpu.sync(stream)

    def download(self, nbytes=None, dtype=np.uint8, offset=0, stream=None):
        nbytes = self.nbytes - offset if nbytes is None else int(nbytes)
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype)
        if nbytes:
            _check(self.gpu.lib.h264gpu_memcpy_d2h(self.gpu.h, _ptr(out), C.c_void_p(self.ptr + offset),
                                                   nbytes, stream), "d2h")
            self.gpu.sync(stream)
        return out


class PinnedArray:
    """Page-locked host byte buffer exposed as a numpy array."""

    def __init__(self, gpu, nbytes):
        self.gpu, self.nbytes = gpu, int(nbytes)
        p = C.c_void_p()
        _check(gpu.lib.h264gpu_host_alloc(gpu.h, max(self.nbytes, 16), C.byref(p)), "host_alloc")
        self.ptr = p.value
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(self.nbytes, 16)).from_address(self.ptr))[:self.nbytes]

    def view(self, dtype):
        return self.array.view(dtype)

    def free(self):
        if self.ptr:
            self.array = None
            self.gpu.lib.h264gpu_host_free(self.gpu.h, C.c_void_p(self.ptr))
            self.ptr = None


class Gpu:
    """One h264gpu context (one CUDA device).  Mirrors include/h264gpu.h 1:1."""

    def __init__(self, device=0):
        self.lib = load_gpu_lib()
        h = C.c_void_p()
        _check(self.lib.h264gpu_create(device, C.byref(h)), "h264gpu_create")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.lib.h264gpu_destroy(self.h)
            self.h = None

    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def pinned(self, nbytes):
        return PinnedArray(self, nbytes)

    def sync(self, stream=None):
        _check(self.lib.h264gpu_sync(self.h, stream), "sync")

    def launch_count(self):
        return int(self.lib.h264gpu_launch_count(self.h))

    # ---- stage calls ---------------------------------------------------
    def split_strip_dev(self, d_in, length, d_rbsp, d_start, d_end, d_rb, cap, d_result,
                        base=0, edge=None, stream=None):
        _check(self.lib.h264gpu_split_strip_dev(
            self.h, C.c_void_p(d_in), length, base, C.byref(edge) if edge is not None else None,
            C.c_void_p(d_rbsp) if d_rbsp else None, C.c_void_p(d_start), C.c_void_p(d_end),
            C.c_void_p(d_rb) if d_rb else None, cap, C.c_void_p(d_result), stream),
            "h264gpu_split_strip_dev")

    def split_strip_inplace_dev(self, d_in, length, d_rbsp, d_start, d_end, d_rb, d_rblen, cap, d_result,
                                base=0, edge=None, stream=None):
        """Scan + strip with every NAL's RBSP written in place (include/h264gpu.h)."""
        _check(self.lib.h264gpu_split_strip_inplace_dev(
            self.h, C.c_void_p(d_in), length, base, C.byref(edge) if edge is not None else None,
            C.c_void_p(d_rbsp) if d_rbsp else None, C.c_void_p(d_start), C.c_void_p(d_end),
            C.c_void_p(d_rb) if d_rb else None, C.c_void_p(d_rblen) if d_rblen else None, cap,
            C.c_void_p(d_result), stream), "h264gpu_split_strip_inplace_dev")

    def split_strip_inplace(self, buf, want_rbsp=True, cap=None, edge=None, base=0):
        """Host convenience over split_strip_inplace_dev (upload, run, download): the per-NAL
        view dict(start, end, rbsp_off, rbsp_len, nal_rbsp=[bytes per NAL], body, res)."""
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        n = len(buf)
        cap = (n // 3 + 2) if cap is None else cap
        d_in = self.alloc(n + 16)
        d_out = self.alloc(n + 16 + 128) if want_rbsp else None
        d_tab = self.alloc(cap * 32)
        d_res = self.alloc(C.sizeof(ScanResult))
        try:
            d_in.upload(buf)
            if want_rbsp:
                d_out.upload(np.full(n + 16 + 128, 0xAA, np.uint8))
            self.split_strip_inplace_dev(d_in.ptr, n, d_out.ptr + 64 if want_rbsp else None, d_tab.ptr,
                                         d_tab.ptr + cap * 8, d_tab.ptr + cap * 16, d_tab.ptr + cap * 24,
                                         cap, d_res.ptr, base=base, edge=edge)
            self.sync()
            res = ScanResult.from_buffer_copy(d_res.download().tobytes())
            k = min(int(res.n_nal), cap)
            tab = d_tab.download(dtype=np.uint64)
            start, end = tab[:k].copy(), tab[cap:cap + k].copy()
            rb, rl = tab[2 * cap:2 * cap + k].copy(), tab[3 * cap:3 * cap + k].copy()
            body, per_nal = None, None
            if want_rbsp:
                raw = d_out.download()
                assert (raw[:64] == 0xAA).all() and (raw[64 + n:] == 0xAA).all(), "write outside d_rbsp"
                body = raw[64:64 + n]
                per_nal = [body[int(rb[i]):int(rb[i]) + int(rl[i])] for i in range(k)]
            return dict(start=start, end=end, rbsp_off=rb, rbsp_len=rl, nal_rbsp=per_nal, body=body, res=res)
        finally:
            for d in (d_in, d_out, d_tab, d_res):
                if d is not None:
                    d.free()

    def split_strip_host(self, buf, want_rbsp=True, cap=None, out=None):
        """buf: uint8 numpy array (host).  Returns dict(start,end,rbsp_off,rbsp,final_off)."""
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        n = len(buf)
        cap = (n // 3 + 2) if cap is None else cap
        if out is None:
            out = dict(start=np.empty(cap, np.uint64), end=np.empty(cap, np.uint64),
                       rbsp_off=np.empty(cap, np.uint64) if want_rbsp else None,
                       rbsp=np.empty(n + 16, np.uint8) if want_rbsp else None)
        nn, rb, off = C.c_uint64(cap), C.c_uint64(0), C.c_uint64(0)
        rc = self.lib.h264gpu_split_strip_host(self.h, _ptr(buf), n, _ptr(out["rbsp"]),
                                               _ptr(out["start"]), _ptr(out["end"]),
                                               _ptr(out["rbsp_off"]), C.byref(nn), C.byref(rb),
                                               C.byref(off))
        _check(rc, "h264gpu_split_strip_host")
        k = nn.value
        return dict(start=out["start"][:k], end=out["end"][:k],
                    rbsp_off=out["rbsp_off"][:k] if want_rbsp else None,
                    rbsp=out["rbsp"][:rb.value] if want_rbsp else None,
                    final_off=off.value, n_nal=k, rbsp_bytes=rb.value)

    def frame_dev(self, d_rbsp, d_off, n, sc_len, d_out, out_cap, d_out_off, d_total, stream=None):
        _check(self.lib.h264gpu_frame_dev(self.h, C.c_void_p(d_rbsp), C.c_void_p(d_off), n, sc_len,
                                          C.c_void_p(d_out), out_cap, C.c_void_p(d_out_off),
                                          C.c_void_p(d_total), stream), "h264gpu_frame_dev")

    def frame_host(self, rbsp, offs, sc_len=4, out=None, out_off=None):
        rbsp = np.ascontiguousarray(rbsp, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        if out is None:
            out = np.empty(len(rbsp) * 3 // 2 + n * max(sc_len, 0) + 64, np.uint8)
        if out_off is None:
            out_off = np.empty(n + 1, np.uint64)
        tot = C.c_uint64(0)
        rc = self.lib.h264gpu_frame_host(self.h, _ptr(rbsp), _ptr(offs), n, sc_len, _ptr(out),
                                         len(out), _ptr(out_off), C.byref(tot))
        _check(rc, "h264gpu_frame_host")
        return out[:tot.value], out_off

    def cabac_parse_host(self, stream, params, n_records):
        """K5: like cavlc_parse_host for CABAC slices."""
        return self.cavlc_parse_host(stream, params, n_records, fn="h264gpu_cabac_parse_host")

    def cabac_parse_dev(self, d_stream, stream_len, d_params, n_slices, d_records, d_results, stream=None):
        _check(self.lib.h264gpu_cabac_parse_dev(self.h, C.c_void_p(d_stream), stream_len,
                                                C.c_void_p(d_params), n_slices, C.c_void_p(d_records),
                                                C.c_void_p(d_results), stream), "h264gpu_cabac_parse_dev")

    def cavlc_parse_host(self, stream, params, n_records, fn="h264gpu_cavlc_parse_host"):
        """params: uint8 array of packed struct h264gpu_slice_params.  Returns (records, results)."""
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        params = np.ascontiguousarray(params, dtype=np.uint8)
        n = len(params) // C.sizeof(SliceParams)
        recs = np.zeros(max(int(n_records), 1), dtype=MB_RECORD)
        res = np.zeros(max(n, 1), dtype=SLICE_RESULT)
        _check(getattr(self.lib, fn)(self.h, _ptr(stream), len(stream), _ptr(params), n,
                                      _ptr(recs), int(n_records), _ptr(res)), fn)
        return recs[:int(n_records)], res[:n]

    def cavlc_parse_dev(self, d_stream, stream_len, d_params, n_slices, d_records, d_results, stream=None):
        _check(self.lib.h264gpu_cavlc_parse_dev(self.h, C.c_void_p(d_stream), stream_len,
                                                C.c_void_p(d_params), n_slices, C.c_void_p(d_records),
                                                C.c_void_p(d_results), stream), "h264gpu_cavlc_parse_dev")

    # ---- timers ----------------------------------------------------------
    def timer(self):
        t = C.c_void_p()
        _check(self.lib.h264gpu_timer_create(self.h, C.byref(t)), "timer_create")
        return t

    def timer_start(self, t, stream=None):
        _check(self.lib.h264gpu_timer_start(self.h, t, stream), "timer_start")

    def timer_stop(self, t, stream=None):
        _check(self.lib.h264gpu_timer_stop(self.h, t, stream), "timer_stop")

    def timer_ms(self, t):
        ms = C.c_float(0)
        _check(self.lib.h264gpu_timer_elapsed_ms(self.h, t, C.byref(ms)), "timer_elapsed")
        return ms.value


# ---- synthetic workloads (host generator, libh264synth.so) ------------------

def synth_offsets(seed, target_bytes, lo=64, hi=256 * 1024):
    lib = load_synth_lib()
    n = lib.synth_sizes(seed, target_bytes, lo, hi, None, 0)
    offs = np.empty(n + 1, np.uint64)
    lib.synth_sizes(seed, target_bytes, lo, hi, _ptr(offs), n + 1)
    return offs


def synth_payloads(seed, offs, out=None, nthreads=None):
    lib = load_synth_lib()
    nthreads = nthreads or min(os.cpu_count() or 1, 64)
    tot = int(offs[-1])
    if out is None:
        out = np.empty(tot, np.uint8)
    lib.synth_fill(seed, _ptr(offs), len(offs) - 1, _ptr(out), nthreads)
    return out[:tot]


def synth_annexb(seed, rbsp, offs, mixed_sc=True, trailing=True, out=None, nthreads=None):
    """Escape + frame payloads into a valid Annex-B stream.  Returns (stream, nal_offsets)."""
    lib = load_synth_lib()
    nthreads = nthreads or min(os.cpu_count() or 1, 64)
    n = len(offs) - 1
    oo = np.empty(n + 1, np.uint64)
    if out is None:
        need = lib.synth_frame(seed, _ptr(rbsp), _ptr(offs), n, int(mixed_sc), int(trailing),
                               None, 0, _ptr(oo), nthreads)
        out = np.empty(need, np.uint8)
    tot = lib.synth_frame(seed, _ptr(rbsp), _ptr(offs), n, int(mixed_sc), int(trailing),
                          _ptr(out), len(out), _ptr(oo), nthreads)
    if tot > len(out):
        raise ValueError("synth_annexb: output buffer too small (%d > %d)" % (tot, len(out)))
    return out[:tot], oo


def mb_hash_term(field, index, value):
    """h264gpu_mb_hash_term (include/h264gpu_slice.h) as a Python int modulo 2**64."""
    key = (field << 16) | index
    w = (((key + 1) * 0x9E3779B97F4A7C15) | 1) & (2**64 - 1)
    return (value * w) & (2**64 - 1)


class VideoCfg(C.Structure):
    """struct synth_video_cfg (libh264_b200/csrc/synth_video.c)."""
    _fields_ = [("width_mbs", C.c_uint32), ("height_mbs", C.c_uint32), ("frames", C.c_uint32),
                ("slices_per_frame", C.c_uint32), ("profile_idc", C.c_uint32),
                ("chroma_format_idc", C.c_uint32), ("transform_8x8", C.c_uint32),
                ("idr_period", C.c_uint32), ("b_frames", C.c_uint32), ("num_ref_frames", C.c_uint32),
                ("entropy_cabac", C.c_uint32), ("pct_skip", C.c_uint32),
                ("pct_intra_in_inter", C.c_uint32), ("pct_pcm", C.c_uint32),
                ("coef_density", C.c_uint32), ("seed", C.c_uint64),
                ("cabac_twin", C.c_uint32), ("fmo", C.c_uint32)]


def synth_video(width_mbs, height_mbs, frames, slices_per_frame=1, profile_idc=66, chroma_format_idc=1,
                transform_8x8=0, idr_period=30, b_frames=0, num_ref_frames=1, pct_skip=30,
                pct_intra_in_inter=10, pct_pcm=5, coef_density=60, seed=0x264, out=None,
                want_params=False, entropy_cabac=0, cabac_twin=0, fmo=0):
    """Synthetic elementary stream (Annex-B), CAVLC or (entropy_cabac=1) CABAC slice data.
    Returns (stream, total_mbs, total_slices) and, with want_params, the packed
    h264gpu_slice_params block of every slice.  fmo = slice groups | slice_group_map_type << 4
    (CAVLC only; every group of a frame is cut into slices_per_frame slices; the parameter blocks
    carry num_slice_groups_minus1 but no map: build it with h264gpu_fmo / the host library)."""
    lib = load_synth_lib()
    cfg = VideoCfg(width_mbs, height_mbs, frames, slices_per_frame, profile_idc, chroma_format_idc,
                   transform_8x8, idr_period, b_frames, num_ref_frames, entropy_cabac, pct_skip,
                   pct_intra_in_inter, pct_pcm, coef_density, seed, cabac_twin, fmo)
    mbs, sl = C.c_uint64(0), C.c_uint64(0)
    if out is None:
        need = lib.synth_video(C.byref(cfg), None, 0, C.byref(mbs), C.byref(sl), None, 0)
        out = np.empty(need, np.uint8)
    nsl = frames * max(1, slices_per_frame) * max(1, fmo & 15)
    params = np.zeros(nsl, SLICE_PARAMS) if want_params else None
    n = lib.synth_video(C.byref(cfg), _ptr(out), len(out), C.byref(mbs), C.byref(sl),
                        _ptr(params.view(np.uint8)) if want_params else None, nsl)
    if n > len(out):
        raise ValueError("synth_video: buffer too small")
    if want_params:
        return out[:n], mbs.value, sl.value, params[:sl.value].view(np.uint8).copy()
    return out[:n], mbs.value, sl.value
