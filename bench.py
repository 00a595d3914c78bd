#!/usr/bin/env python
"""bench.py — headline benchmark: Annex-B NAL split + EPB strip (BASELINE.json config 2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size-mb M]

A step = one pass of the hot path (fused scan + strip, annexb7::scan7_kernel + finalize) over one
4 GiB synthetic Annex-B buffer per GPU.  With N > 1 the ranks hold consecutive byte-range shards
of ONE 4N GiB stream whose seams fall inside NAL units; there is no data-path collective: after
the timed region the per-shard tables are gathered to rank 0, merged on the host
(h264gpu_merge_shard_inplace) and checked against the generator's table.

value        input GB/s with the buffer resident in HBM (CUDA events around K steps, max over ranks)
scan_only    the same buffer through the table-only kernel (what h264_reader_parse launches)
e2e          same metric through h264gpu_split_strip_host: pinned HOST input, H2D, kernel, D2H of
             NAL tables + RBSP inside the timed region
e2e_reader   h264_reader_parse(SLICE_DATA) of this repo's libh264.so against the reference's on
             the same CAVLC stream through the public libh264 API (counting callbacks)
roofline     algorithmic bytes (N_in + N_rbsp + 32 B per NAL) / step time vs the measured HBM peak
cpu_baseline / --impl reference: the UNMODIFIED reference (oracle/_ref/libh264_ref.so) on the
             box's host cores
extra        macroblock parse (CAVLC 16000 slices, config 1 on the GPU, CABAC config 3, config 4
             share), writer (config 5) with its own cpu_baseline and e2e, strong scaling (N > 1)
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "annexb_split_strip_throughput"
UNIT = "GB/s"
SEED = 0x264
SCAN_KERNEL = ("annexb7::scan7_kernel<8> (warp-autonomous 4 KiB spans, tickets from 8 counters, round robin over regions, "
               "RBSP in place per NAL) + second pass over deferred region heads + fin7_spans/fin7_order/fin7_table")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(n_in):
    """DRAM bytes (read + write) of the dominant kernel per launch from the committed ncu capture
    of this very workload (profiles/r02_traffic.json, scripts/gpu_session.sh); None when the
    capture is of another input size."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f)
        if t.get("bytes_in") == int(n_in):
            return int(t["dram_bytes_read"] + t["dram_bytes_write"])
    except (OSError, ValueError, KeyError):
        pass
    return None


def make_workload(L, size_bytes, seed, out=None, nthreads=None):
    """Config-2 stream: NAL sizes log-uniform 64 B..256 KiB, P(00)=3/16, escaped, mixed
    3/4-byte start codes, 0-2 trailing zero bytes on a quarter of the NALs."""
    target = int(size_bytes * 0.992)  # escaping + start codes add ~0.7 %
    offs = L.synth_offsets(seed, target)
    rbsp = L.synth_payloads(seed, offs, nthreads=nthreads)
    stream, nal_off = L.synth_annexb(seed, rbsp, offs, out=out, nthreads=nthreads)
    return stream, rbsp, offs


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled while the timed regions run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev_index):
        self.dev, self.rows, self.proc = dev_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 8 and f[0] == str(self.dev):
                self.rows.append(f)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in self.rows)]
        top = sorted(sm)[len(sm) // 2:] if sm else []  # upper half = samples under load
        return {"sm_mhz": statistics.median(top) if top else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def pin_to_gpu_numa_node(local_rank):
    """Best effort: run this rank on the CPUs of its GPU's NUMA node (pinned-memory copies of 8
    ranks through one node were the e2e limiter of round 1).  Returns a note for the JSON line."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = "0000:" + bus[9:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return "numa node unknown"
        cpus = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, ids)
        return "node %d (%d cpus)" % (node, len(ids))
    except Exception as e:  # noqa: BLE001 - diagnostics only
        return "not pinned (%s)" % type(e).__name__


def ref_lib():
    so = os.path.join(ROOT, "oracle", "_ref", "libh264_ref.so")
    if not os.path.exists(so):
        return None
    lib = C.CDLL(so)
    lib.ref_mt_split_strip.restype = C.c_double
    lib.ref_mt_split_strip.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p,
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.ref_mt_insert.restype = C.c_double
    lib.ref_mt_insert.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_uint64)]
    lib.ref_mt_parse.restype = C.c_double
    lib.ref_mt_parse.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int,
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    return lib


def host_cores():
    return min(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1), 256)


def cpu_reference_run(stream, sample_bytes, cores, reps=1, strip=1):
    """Reference scan (+ strip) of stream[:sample_bytes] on `cores` threads; best of reps."""
    lib = ref_lib()
    if lib is None:
        return None
    sample = stream[:sample_bytes]
    scratch = np.empty(len(sample) + 64, np.uint8)
    nn, rb = C.c_uint64(0), C.c_uint64(0)
    best = None
    for _ in range(reps):
        dt = lib.ref_mt_split_strip(sample.ctypes.data_as(C.c_void_p), len(sample), cores, strip,
                                    scratch.ctypes.data_as(C.c_void_p), C.byref(nn), C.byref(rb))
        best = dt if best is None else min(best, dt)
    return {"value": len(sample) / best / 1e9, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "first %.0f MiB of the same stream, reference h264_find_nalu loop%s, %d threads over byte "
                      "ranges, best of %d" % (len(sample) / 2**20, " + h264_bs_read_bits(8) strip loop" if strip else "",
                                              cores, reps),
            "seconds": best, "nals": nn.value}


MB_CFG = dict(width_mbs=120, height_mbs=68, slices_per_frame=16, profile_idc=100, transform_8x8=1,
              b_frames=1, num_ref_frames=2, idr_period=30, pct_skip=30, coef_density=60, seed=SEED)
CFG1 = dict(width_mbs=120, height_mbs=68, slices_per_frame=1, profile_idc=66, transform_8x8=0, b_frames=0,
            num_ref_frames=1, idr_period=30, pct_skip=30, coef_density=60, seed=SEED)
CFG4 = dict(width_mbs=240, height_mbs=135, slices_per_frame=8, profile_idc=100, transform_8x8=1, b_frames=1,
            num_ref_frames=2, idr_period=30, pct_skip=30, coef_density=60, seed=SEED, entropy_cabac=1)


def replicate(L, stream, params, nmb, nsl, reps):
    """`reps` copies of a generated stream laid end to end (separate memory, separate slices)."""
    if reps <= 1:
        return stream, params, nmb, nsl
    P = np.frombuffer(params, L.SLICE_PARAMS)
    pad = (-len(stream)) % 16
    one = np.concatenate([stream, np.zeros(pad, np.uint8)])
    allp = []
    for k in range(reps):
        q = P.copy()
        q["nal_off"] += k * len(one)
        q["mb_out_off"] += k * nmb
        allp.append(q)
    return np.tile(one, reps), np.concatenate(allp).view(np.uint8), nmb * reps, nsl * reps


def mb_parse_leg(g, L, cfg, label, steps, warmup, with_cpu=False, cabac=False, reps=1, e2e_reps=3):
    """Slice-parallel macroblock parse: macroblocks/s device-resident (CUDA events), end to end from
    host buffers (warm mean of e2e_reps), and (with_cpu) the CPU baselines."""
    gen_cfg = dict(cfg)
    stream, nmb, nsl, params = L.synth_video(want_params=True, **gen_cfg)
    one_stream, one_nmb, one_params = stream, nmb, params
    stream, params, nmb, nsl = replicate(L, stream, params, nmb, nsl, reps)
    d_stream = g.alloc(len(stream) + 64)
    d_stream.upload(stream)
    d_params = g.alloc(len(params))
    d_params.upload(params)
    d_rec = g.alloc(nmb * 16 + 16)
    d_res = g.alloc(nsl * 16)
    parse_dev = g.cabac_parse_dev if cabac else g.cavlc_parse_dev
    parse_host = g.cabac_parse_host if cabac else g.cavlc_parse_host

    def step():
        parse_dev(d_stream.ptr, len(stream), d_params.ptr, nsl, d_rec.ptr, d_res.ptr)

    for _ in range(warmup):
        step()
    g.sync()
    tm = g.timer()
    g.timer_start(tm)
    for _ in range(steps):
        step()
    g.timer_stop(tm)
    ms = g.timer_ms(tm) / steps
    res = d_res.download(dtype=np.uint8).view(L.SLICE_RESULT)
    ok = bool((res["status"] == 0).all() and int(res["mb_count"].sum()) == nmb)
    for d in (d_stream, d_params, d_rec, d_res):
        d.free()
    out = {"workload": "%s%s (%d slices, %d MBs, %.1f MB stream)"
                       % (label, " x %d copies laid end to end" % reps if reps > 1 else "", nsl, nmb, len(stream) / 1e6),
           "kernel": "cabac::cabac_parse_kernel (persistent grid, slices pulled longest first from a ticket counter, one fused table entry per bin) + order_kernel" if cabac else "cavlc2::cavlc_steps_kernel (per-lane state machine, one syntax element per lane per step) + order_kernel (longest slices first)",
           "macroblocks_per_s": nmb / (ms / 1e3), "ms_per_step": ms, "parity_counts_ok": ok}
    if e2e_reps > 0:
        parse_host(stream, params, nmb)  # warm-up: pools
        ts = []
        for _ in range(e2e_reps):
            t1 = time.perf_counter()
            parse_host(stream, params, nmb)
            ts.append(time.perf_counter() - t1)
        out["e2e_macroblocks_per_s"] = nmb / (sum(ts) / len(ts))
        out["e2e_note"] = ("h264gpu_%s_parse_host: H2D stream + params, kernel, D2H 16 B/MB records, pooled buffers; "
                           "warm mean of %d" % ("cabac" if cabac else "cavlc", e2e_reps))
    if with_cpu:
        lib = ref_lib()
        cores = host_cores()
        if lib is not None:
            # the reference on every core: one reader per thread, each parsing a copy of the generated stream
            bufs = (C.c_void_p * cores)(*[one_stream.ctypes.data] * cores)
            lens = (C.c_size_t * cores)(*[len(one_stream)] * cores)
            mbs, sl = C.c_uint64(0), C.c_uint64(0)
            best = None
            for _ in range(2):
                dt = lib.ref_mt_parse(bufs, lens, cores, 1, cores, C.byref(mbs), C.byref(sl))
                best = dt if best is None else min(best, dt)
            if not cabac:
                out["cpu_baseline"] = {"value": mbs.value / best, "unit": "macroblocks/s", "cores": cores,
                                       "kind": "reference",
                                       "sample": "reference h264_reader_parse(SLICE_DATA) of the generated stream, one "
                                                 "reader per thread (%d threads, %d MBs total), best of 2" % (cores, mbs.value)}
            else:
                # the reference parses no CABAC slice data -> (a) its time on the same
                # stream with the slice data skipped, (b) this repo's own from-spec CPU decoder, labelled
                rec = np.zeros(one_nmb + 1, L.MB_RECORD)
                P1 = np.frombuffer(one_params, L.SLICE_PARAMS).copy()
                rs = np.zeros(len(P1), L.SLICE_RESULT)
                sl_lib = L.load_synth_lib()
                sl_lib.synth_cabac_decode.restype = C.c_int
                sl_lib.synth_cabac_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p,
                                                      C.c_uint64, C.c_void_p]
                t1 = time.perf_counter()
                sl_lib.synth_cabac_decode(one_stream.ctypes.data, len(one_stream), P1.ctypes.data, len(P1),
                                          rec.ctypes.data, one_nmb, rs.ctypes.data)
                dt_port = time.perf_counter() - t1
                out["cpu_baseline"] = {
                    "value": one_nmb / dt_port, "unit": "macroblocks/s", "cores": 1, "kind": "port",
                    "sample": "this repo's own CPU CABAC decoder (synth_cabac_decode: the cabac_syntax.h walker with "
                              "the decoder, NOT the reference) on the %d generated frames, 1 thread" % gen_cfg["frames"],
                    "reference_skip": {"value": len(one_stream) * cores / best / 1e9, "unit": "GB/s",
                                       "note": "reference h264_reader_parse of the same stream on %d threads: it "
                                               "returns before CABAC slice data (src/h264_syntax_slice_data.h:715-717), "
                                               "so this is NAL split + header parse only" % cores}}
    return out


def frame_leg(g, L, size, steps, warmup, rank, with_cpu):
    """Writer side (BASELINE config 5): EPB insertion + 4-byte start-code framing of `size` bytes
    of synthetic RBSP payloads: device-resident GB/s of payload + roofline, e2e through
    h264gpu_frame_host, and the reference writer on the host cores."""
    offs = L.synth_offsets(SEED + 100 + rank, size)
    rbsp = L.synth_payloads(SEED + 100 + rank, offs)
    n = len(offs) - 1
    cap = len(rbsp) + len(rbsp) // 2 + 4 * n + 64
    d_r = g.alloc(len(rbsp) + 16)
    d_r.upload(rbsp)
    d_o = g.alloc(len(offs) * 8)
    d_o.upload(offs)
    d_out = g.alloc(cap + 16)
    d_oo = g.alloc((n + 2) * 8)

    def step():
        g.frame_dev(d_r.ptr, d_o.ptr, n, 4, d_out.ptr, cap, d_oo.ptr, d_oo.ptr + (n + 1) * 8)

    for _ in range(warmup):
        step()
    g.sync()
    tm = g.timer()
    g.timer_start(tm)
    for _ in range(steps):
        step()
    g.timer_stop(tm)
    ms = g.timer_ms(tm) / steps
    total = int(d_oo.download(dtype=np.uint64)[n + 1])
    for d in (d_r, d_o, d_out, d_oo):
        d.free()
    peak, _ = measured_peak()
    alg = len(rbsp) + total + 8 * (n + 1)
    out = {"workload": "EPB insert + framing of %.0f MiB RBSP in %d payloads (BASELINE config 5)" % (size / 2**20, n),
           "kernel": ("frame7::frame7_prepass<8> + frame7::frame7_kernel<8> (warp-autonomous 4 KiB spans, counts published "
                      "ahead of the look-back)" if os.environ.get("H264GPU_FRAME_GEN") == "7" else
                      "frame::frame_prepass<8> + frame6::frame6_kernel<8> (persistent CTAs, 32 KiB tiles)"),
           "gb_per_s": len(rbsp) / (ms / 1e3) / 1e9, "ms_per_step": ms, "out_bytes": total,
           "roofline": {"bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg / (ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_launch": int(alg)},
           "roofline_frac": alg / (ms / 1e3) / 1e9 / peak}
    # e2e: host payloads in, framed stream out, through h264gpu_frame_host (chunks of whole payloads,
    # upload / kernel / download overlapped); input from pinned and from pageable memory
    pin_out = g.pinned(cap + 64)
    pin_in = g.pinned(len(rbsp) + 64)
    pin_in.array[:len(rbsp)] = rbsp
    oo = np.zeros(n + 1, np.uint64)

    def e2e_run(src):
        g.frame_host(src, offs, 4, out=pin_out.array[:cap], out_off=oo)  # warm-up (pools grow)
        ts = []
        for _ in range(3):
            t1 = time.perf_counter()
            g.frame_host(src, offs, 4, out=pin_out.array[:cap], out_off=oo)
            ts.append(time.perf_counter() - t1)
        return sum(ts) / len(ts)

    t_pin = e2e_run(pin_in.array[:len(rbsp)])
    t_page = e2e_run(rbsp)
    out["e2e"] = {"value": len(rbsp) / t_pin / 1e9, "unit": "GB/s of payload",
                  "h2d_bytes_per_step": int(len(rbsp) + 8 * (n + 1)), "d2h_bytes_per_step": int(total + 8 * (n + 1)),
                  "pageable_input": {"value": len(rbsp) / t_page / 1e9, "unit": "GB/s of payload"},
                  "api": "h264gpu_frame_host (pinned host payloads in, pinned host stream out; chunks of whole payloads, "
                         "upload / kernel / download overlapped), warm mean of 3"}
    pin_in.free()
    pin_out.free()
    if with_cpu:
        lib = ref_lib()
        if lib is not None:
            cores = host_cores()
            # bounded sample: ~0.08 GB/s per core, aim at ~4 s
            k = int(np.searchsorted(offs, min(int(offs[-1]), int(0.08e9 * cores * 4))))
            k = max(1, min(k, n))
            ob = C.c_uint64(0)
            best = None
            for _ in range(2):
                dt = lib.ref_mt_insert(rbsp.ctypes.data, offs.ctypes.data, k, cores, C.byref(ob))
                best = dt if best is None else min(best, dt)
            out["cpu_baseline"] = {"value": int(offs[k]) / best / 1e9, "unit": "GB/s of payload", "cores": cores,
                                   "kind": "reference",
                                   "sample": "reference h264_bs_write_bits(8) writer over the first %d payloads "
                                             "(%.0f MiB), %d threads, best of 2" % (k, int(offs[k]) / 2**20, cores)}
    return out


def reader_leg(L, cfg, label, reps=3):
    """h264_reader_parse(SLICE_DATA) through the public libh264 API with counting callbacks: this
    repo's libh264.so (GPU) vs the reference's, same CAVLC stream, pageable and pinned input."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_host_lib as T
    lib = T.harness()
    stream, nmb, nsl = L.synth_video(**cfg)
    out = {"workload": "%s (%d slices, %d MBs, %.1f MB)" % (label, nsl, nmb, len(stream) / 1e6),
           "api": "h264_reader_parse(H264_READER_FLAGS_SLICE_DATA), callbacks nalu_begin/sps/pps/slice/slice_data_mb counted"}
    t_ref, c_ref = T.time_parse(lib, T.REF, stream, 1, 1)
    t_our, c_our = T.time_parse(lib, T.OURS, stream, 1, reps)
    out["callbacks_identical"] = bool((c_ref == c_our).all())
    out["reference_1_core"] = {"seconds": t_ref, "macroblocks_per_s": nmb / t_ref}
    out["ours_pageable"] = {"seconds": t_our, "macroblocks_per_s": nmb / t_our}
    # pinned input
    g = L.Gpu(int(os.environ.get("LOCAL_RANK", "0")))
    pin = g.pinned(len(stream) + 64)
    pin.array[:len(stream)] = stream
    t_pin, c_pin = T.time_parse(lib, T.OURS, pin.array[:len(stream)], 1, reps)
    out["ours_pinned"] = {"seconds": t_pin, "macroblocks_per_s": nmb / t_pin}
    pin.free()
    g.close()
    # the reference on every core: one reader per thread over copies of the stream
    rl = ref_lib()
    if rl is not None:
        cores = host_cores()
        bufs = (C.c_void_p * cores)(*[stream.ctypes.data] * cores)
        lens = (C.c_size_t * cores)(*[len(stream)] * cores)
        mbs, sl = C.c_uint64(0), C.c_uint64(0)
        dt = rl.ref_mt_parse(bufs, lens, cores, 1, cores, C.byref(mbs), C.byref(sl))
        out["reference_all_cores"] = {"cores": cores, "macroblocks_per_s": mbs.value / dt,
                                      "note": "one h264_reader per thread, each parsing a copy of the stream"}
        out["speedup_vs_all_cores"] = (nmb / min(t_our, t_pin)) / (mbs.value / dt)
    out["speedup_vs_1_core"] = t_ref / min(t_our, t_pin)
    out["headers_only"] = {}
    t_ref0, _ = T.time_parse(lib, T.REF, stream, 0, 1)
    t_our0, _ = T.time_parse(lib, T.OURS, stream, 0, reps)
    out["headers_only"] = {"reference_1_core_s": t_ref0, "ours_s": t_our0, "gb_per_s_ours": len(stream) / t_our0 / 1e9}
    return out


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    import libh264_b200 as L
    cores = host_cores()
    size = args.size_mb << 20
    # bounded sample: ~0.12 GB/s per core, aim at ~3 s per step
    sample = min(size, max(64 << 20, int(0.12e9 * cores * 3)))
    stream, _, _ = make_workload(L, sample, SEED)
    times = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_run(stream, len(stream), cores)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libh264_ref.so missing"}))
            return
        if i >= args.warmup:
            times.append(r["seconds"])
    ms = 1e3 * sum(times) / len(times)
    val = len(stream) / (ms / 1e3) / 1e9
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "annexb_scan_strip (BASELINE config 2)", "sample_mib": len(stream) >> 20,
                   "full_size_mib_per_gpu": args.size_mb, "nal_bytes": "loguniform[64,262144]",
                   "p_zero": 0.1875},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": "each step = reference scan+strip of a %d MiB prefix-shaped stream "
                                   "(same generator/seed as the GPU arm) on %d threads" % (len(stream) >> 20, cores)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--size-mb", type=int, default=4096, help="input MiB per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--mb-frames", type=int, default=1000, help="frames of the CAVLC macroblock-parse workload")
    ap.add_argument("--cabac-frames", type=int, default=1000,
                    help="frames of the CABAC macroblock-parse workload, config 3 (0 = skip)")
    ap.add_argument("--cfg4-streams", type=int, default=32,
                    help="4K High CABAC streams per GPU for the config-4 leg (256 over 8 GPUs; 0 = skip)")
    ap.add_argument("--frame-mb", type=int, default=1024,
                    help="MiB of RBSP payloads for the writer leg (BASELINE config 5; 0 = skip)")
    ap.add_argument("--no-reader", action="store_true", help="skip the e2e_reader leg")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling line (N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import libh264_b200 as L
    numa_note = pin_to_gpu_numa_node(local_rank) if world > 1 else "single rank: not pinned"
    dist = None
    torch = None

    def barrier():
        if dist is not None:
            dist.barrier()

    def reduce_ranks(x, op):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.MAX) if dist is not None else x

    def sum_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.SUM) if dist is not None else x

    def gather_bytes(arr):
        """Variable-length uint8 arrays of every rank, on every rank (setup / verification only)."""
        if dist is None:
            return [np.ascontiguousarray(arr)]
        n = torch.tensor([len(arr)], dtype=torch.int64, device="cuda")
        ns = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(ns, n)
        mx = int(max(int(x.item()) for x in ns))
        buf = torch.zeros(mx, dtype=torch.uint8, device="cuda")
        buf[:len(arr)] = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
        outs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf)
        return [o[:int(k.item())].cpu().numpy() for o, k in zip(outs, ns)]

    g = L.Gpu(local_rank)
    size = args.size_mb << 20
    nthreads = max(1, min(os.cpu_count() or 1, 64) // max(1, min(world, 8)))

    # ---- workload: rank r generates stream r; the job's stream is their concatenation and the
    # shards are cut DELTA bytes into each piece, so that every seam falls inside a NAL ----
    t0 = time.time()
    DELTA = 4096 + 16 * 37 if world > 1 else 0
    pin_in = g.pinned(size + DELTA + 8192)
    stream, rbsp_ref, offs = make_workload(L, size, SEED + rank, out=pin_in.array, nthreads=nthreads)
    n_own = len(stream)
    n_in = n_own + (DELTA if rank < world - 1 else 0) - (DELTA if rank > 0 else 0)
    n_nal_piece = len(offs) - 1
    cap = n_nal_piece + 4096
    # The stream buffers and, right after them, the scan workspace are the FIRST device allocations of
    # the process, before torch / NCCL allocate anything.  (With ONE span-ticket word in the kernel,
    # where the workspace landed decided 2.0 vs 2.8 ms per 4 GiB: include/h264gpu.h,
    # profiles/r02_scan_workspace_placement.txt.  The tickets now come from 8 counters and the
    # placement no longer matters, profiles/r02_scan7_ticket_counters.txt; the order is kept.)
    d_in = g.alloc(n_in + 64)
    d_rbsp = g.alloc(n_in + 64)
    d_tab = g.alloc(cap * 8 * 4)
    d_res = g.alloc(C.sizeof(L.ScanResult))
    L._check(g.lib.h264gpu_scan_reserve(g.h, C.c_uint64(n_in), C.c_uint64(cap)), "h264gpu_scan_reserve")

    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pieces_len = [int(x.view(np.int64)[0]) for x in gather_bytes(np.array([n_own], np.int64).view(np.uint8))]
    heads = gather_bytes(stream[:DELTA + 2] if world > 1 else np.zeros(0, np.uint8))
    tails = gather_bytes(stream[n_own - 2:] if world > 1 else np.zeros(0, np.uint8))
    piece_base = [sum(pieces_len[:r]) for r in range(world + 1)]
    total_len = piece_base[world]
    lo = piece_base[rank] + (DELTA if rank > 0 else 0)
    hi = piece_base[rank + 1] + (DELTA if rank < world - 1 else 0)
    if world > 1:
        if rank < world - 1:  # append the head of the next piece
            pin_in.array[n_own:n_own + DELTA] = heads[rank + 1][:DELTA]
        shard = pin_in.array[(DELTA if rank > 0 else 0):n_own + (DELTA if rank < world - 1 else 0)]
    else:
        shard = stream
    assert n_in == len(shard) == hi - lo
    # generator's NAL table of this piece in job coordinates (for the merged-table check on rank 0)
    tz = 0
    while tz < n_own and stream[n_own - 1 - tz] == 0:
        tz += 1
    gen_s = time.time() - t0

    # the shard must start 16-byte aligned in device memory: copy from the (possibly unaligned) host view
    stage = g.pinned(n_in + 64)
    stage.array[:n_in] = shard
    L._check(g.lib.h264gpu_memcpy_h2d(g.h, C.c_void_p(d_in.ptr), stage.array.ctypes.data_as(C.c_void_p), n_in, None), "h2d")
    g.sync()

    edge = None
    if world > 1:
        edge = L.ShardEdge()
        if rank > 0:
            edge.has_left = 1
            edge.left[0], edge.left[1] = int(stream[DELTA - 2]), int(stream[DELTA - 1])
            edge.assume_in = 1
        if rank < world - 1:
            nxt = heads[rank + 1]
            edge.has_right, edge.right[0], edge.right[1] = 1, int(nxt[DELTA]), int(nxt[DELTA + 1])

    def step(strip=True):
        g.split_strip_inplace_dev(d_in.ptr, n_in, d_rbsp.ptr if strip else 0, d_tab.ptr, d_tab.ptr + cap * 8,
                                  d_tab.ptr + cap * 16, d_tab.ptr + cap * 24, cap, d_res.ptr, base=lo, edge=edge)

    sampler = ClockSampler(local_rank)
    if not os.environ.get("H264BENCH_NO_SAMPLER"):
        sampler.start()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        g.sync()
        barrier()
        tm = g.timer()
        g.timer_start(tm)
        for _ in range(steps):
            fn()
        g.timer_stop(tm)
        ms = g.timer_ms(tm) / steps
        g.sync()
        barrier()
        return ms

    # ---- device-resident leg -------------------------------------------------
    launches0 = g.launch_count()
    ms_local = timed(step, args.steps, args.warmup)
    launches = (g.launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    ms_step = max_over_ranks(ms_local)
    res = L.ScanResult.from_buffer_copy(d_res.download().tobytes())
    if res.reserved:
        print("bench: event buffer overflow", file=sys.stderr)
        sys.exit(1)
    n_nal_shard, n_rbsp_shard = int(res.n_nal), int(res.rbsp_bytes)
    tab = d_tab.download(dtype=np.uint64)
    sh_start, sh_end, sh_len = tab[:n_nal_shard].copy(), tab[cap:cap + n_nal_shard].copy(), tab[3 * cap:3 * cap + n_nal_shard].copy()

    # ---- merge on rank 0 and check against the generator's table (not in the timed region) ----
    merge_ms = None
    res_all = gather_bytes(np.frombuffer(bytes(res), np.uint8))
    st_all = gather_bytes(sh_start.view(np.uint8))
    en_all = gather_bytes(sh_end.view(np.uint8))
    ln_all = gather_bytes(sh_len.view(np.uint8))
    # generator truth: NAL k of piece r starts after its 3- or 4-byte start code at piece_base[r] + nal_off
    truth = gather_bytes(np.diff(offs.astype(np.int64)).view(np.uint8))
    tz_all = gather_bytes(np.array([tz], np.int64).view(np.uint8))
    merged_ok = None
    if rank == 0:
        t1 = time.perf_counter()
        tot = sum(len(x) // 8 for x in st_all)
        m_st, m_en, m_ln = np.zeros(tot + 1, np.uint64), np.zeros(tot + 1, np.uint64), np.zeros(tot + 1, np.uint64)
        mg = L.Merge()
        g.lib.h264gpu_merge_init(C.byref(mg))
        for r in range(world):
            k = len(st_all[r]) // 8
            first = int(mg.n_nal)
            m_st[first:first + k] = st_all[r].view(np.uint64)
            m_en[first:first + k] = en_all[r].view(np.uint64)
            m_ln[first:first + k] = ln_all[r].view(np.uint64)
            rr = L.ScanResult.from_buffer_copy(res_all[r].tobytes())
            carry = C.c_uint64(0)
            g.lib.h264gpu_merge_shard_inplace(C.byref(mg), C.byref(rr), m_en.ctypes.data_as(C.c_void_p),
                                              m_ln.ctypes.data_as(C.c_void_p), C.c_uint64(tot + 1), C.byref(carry))
        off = C.c_uint64(0)
        g.lib.h264gpu_merge_finish(C.byref(mg), C.c_uint64(total_len), m_en.ctypes.data_as(C.c_void_p),
                                   C.c_uint64(tot + 1), C.byref(off))
        merge_ms = (time.perf_counter() - t1) * 1e3
        want_len = np.concatenate([x.view(np.int64) for x in truth])
        # zero bytes after a NAL are cut off before a start code but stay with the stream's last NAL
        want_len[-1] += int(tz_all[-1].view(np.int64)[0])
        n_total = int(mg.n_nal)
        merged_ok = bool(n_total == len(want_len) and np.array_equal(m_ln[:n_total].astype(np.int64), want_len)
                         and (m_en[:n_total] >= m_st[:n_total]).all() and (m_st[1:n_total] >= m_en[:n_total - 1] + 3).all()
                         and int(off.value) == total_len and int(mg.rbsp_bytes) == int(want_len.sum()))
        if not merged_ok:
            print("bench: merged NAL table differs from the generator's (n %d want %d, rbsp %d want %d)" %
                  (n_total, len(want_len), int(mg.rbsp_bytes), int(want_len.sum())), file=sys.stderr)
            sys.exit(1)

    total_in = sum_over_ranks(float(n_in))
    value = total_in / (ms_step / 1e3) / 1e9
    alg_bytes = n_in + n_rbsp_shard + 32 * n_nal_shard
    achieved = alg_bytes / (ms_local / 1e3) / 1e9
    peak, peak_src = measured_peak()

    # ---- scan only (the launch h264_reader_parse makes): 1.0 B of traffic per input byte ----
    ms_scan = timed(lambda: step(False), max(5, args.steps // 2), 3)
    ms_scan_max = max_over_ranks(ms_scan)
    scan_only = {"value": total_in / (ms_scan_max / 1e3) / 1e9, "unit": UNIT, "ms_per_step": ms_scan_max,
                 "kernel": "annexb7::scan7_only_kernel<8> (16-byte loads into registers, no shared memory) + finalize",
                 "roofline": {"bound": "hbm", "achieved": (n_in + 32 * n_nal_shard) / (ms_scan / 1e3) / 1e9, "peak": peak,
                              "unit": "GB/s", "frac": (n_in + 32 * n_nal_shard) / (ms_scan / 1e3) / 1e9 / peak,
                              "algorithmic_bytes_per_launch": int(n_in + 32 * n_nal_shard)}}

    # ---- strong scaling: ONE 4 GiB stream cut into N byte-range shards (N > 1) ----
    strong = None
    if world > 1 and not args.no_strong:
        s_stream, _, s_offs = make_workload(L, size, SEED, nthreads=nthreads)  # same stream on every rank
        cut = [(len(s_stream) * r // world) // 16 * 16 for r in range(world)] + [len(s_stream)]
        a, b = cut[rank], cut[rank + 1]
        e2 = L.ShardEdge()
        if rank > 0:
            e2.has_left, e2.left[0], e2.left[1], e2.assume_in = 1, int(s_stream[a - 2]), int(s_stream[a - 1]), 1
        if rank < world - 1:
            e2.has_right, e2.right[0], e2.right[1] = 1, int(s_stream[b]), int(s_stream[b + 1])
        stage.array[:b - a] = s_stream[a:b]
        L._check(g.lib.h264gpu_memcpy_h2d(g.h, C.c_void_p(d_in.ptr), stage.array.ctypes.data_as(C.c_void_p), b - a, None), "h2d")
        g.sync()

        def sstep():
            g.split_strip_inplace_dev(d_in.ptr, b - a, d_rbsp.ptr, d_tab.ptr, d_tab.ptr + cap * 8, d_tab.ptr + cap * 16,
                                      d_tab.ptr + cap * 24, cap, d_res.ptr, base=a, edge=e2)
        ms_s = max_over_ranks(timed(sstep, args.steps, 3))
        r2 = L.ScanResult.from_buffer_copy(d_res.download().tobytes())
        nn = sum_over_ranks(float(r2.n_nal))
        strong = {"value": len(s_stream) / (ms_s / 1e3) / 1e9, "unit": UNIT, "ms_per_step": ms_s, "scaling": "strong",
                  "workload": "one %.1f GiB stream cut into %d byte-range shards at 16-byte multiples" % (len(s_stream) / 2**30, world),
                  "nals_found": int(nn), "nals_expected": len(s_offs) - 1, "ok": bool(int(nn) == len(s_offs) - 1)}
        del s_stream
        # restore the weak-scaling shard for the legs below
        stage.array[:n_in] = shard
        L._check(g.lib.h264gpu_memcpy_h2d(g.h, C.c_void_p(d_in.ptr), stage.array.ctypes.data_as(C.c_void_p), n_in, None), "h2d")
        g.sync()

    # ---- end-to-end leg: pinned host in -> tables + RBSP in pinned host out ----
    e2e_in = stage.array[:n_in]
    exp = np.concatenate([x.view(np.int64) for x in truth]) if world == 1 else None
    pin_rbsp = g.pinned(n_in + 4096)
    tabs = g.pinned(cap * 8 * 3)
    tv = tabs.view(np.uint64)
    out = dict(start=tv[:cap], end=tv[cap:2 * cap], rbsp_off=tv[2 * cap:3 * cap], rbsp=pin_rbsp.array)
    h = g.split_strip_host(e2e_in, cap=cap, out=out)  # warm-up (allocates the chunk pipeline)
    if world == 1 and (h["n_nal"] != n_nal_piece or h["rbsp_bytes"] != int(exp.sum()) + tz):
        print("bench: e2e result mismatch", file=sys.stderr)
        sys.exit(1)
    barrier()
    e2e_t = []
    for _ in range(args.e2e_steps):
        t1 = time.perf_counter()
        h = g.split_strip_host(e2e_in, cap=cap, out=out)
        e2e_t.append(time.perf_counter() - t1)
    barrier()
    e2e_s = max_over_ranks(sum(e2e_t) / len(e2e_t))
    e2e_val = total_in / e2e_s / 1e9
    e2e_d2h = int(h["rbsp_bytes"] + 24 * h["n_nal"])
    # ceiling of this platform for the same bytes: one H2D and one D2H copy of the step's sizes, concurrently
    s1, s2 = C.c_void_p(), C.c_void_p()
    ceiling = None
    try:
        L._check(g.lib.h264gpu_stream_create(g.h, C.byref(s1)), "stream")
        L._check(g.lib.h264gpu_stream_create(g.h, C.byref(s2)), "stream")
        ts = []
        for _ in range(2):
            barrier()
            t1 = time.perf_counter()
            g.lib.h264gpu_memcpy_h2d(g.h, C.c_void_p(d_in.ptr), stage.array.ctypes.data_as(C.c_void_p), n_in, s1)
            g.lib.h264gpu_memcpy_d2h(g.h, pin_rbsp.array.ctypes.data_as(C.c_void_p), C.c_void_p(d_rbsp.ptr), min(e2e_d2h, n_in), s2)
            g.lib.h264gpu_sync(g.h, s1)
            g.lib.h264gpu_sync(g.h, s2)
            ts.append(time.perf_counter() - t1)
        barrier()
        ceiling = total_in / max_over_ranks(min(ts)) / 1e9
        g.lib.h264gpu_stream_destroy(g.h, s1)
        g.lib.h264gpu_stream_destroy(g.h, s2)
    except OSError:
        ceiling = None
    clocks = sampler.stop()

    # ---- CPU baseline (rank 0, N=1): the unmodified reference on the host cores ----
    with_cpu = rank == 0 and world == 1 and not args.no_cpu
    cpu = cpu_scan = None
    if with_cpu:
        cores = host_cores()
        sample = min(n_in, max(64 << 20, int(0.12e9 * cores * 4)))
        cpu = cpu_reference_run(stream, sample, cores, reps=3)
        cpu_scan = cpu_reference_run(stream, sample, cores, reps=2, strip=0)
        scan_only["cpu_baseline"] = cpu_scan

    for d in (d_in, d_rbsp, d_tab, d_res):
        d.free()
    for p in (pin_in, stage, pin_rbsp, tabs):
        p.free()

    # ---- second half of the metric: macroblocks/s of the slice-parallel parse ----
    mb = mb_parse_leg(g, L, dict(MB_CFG, frames=args.mb_frames, seed=MB_CFG["seed"] + rank),
                      "CAVLC 1080p High, 16 slices/frame, %d frames" % args.mb_frames,
                      max(3, args.steps // 4), 2, with_cpu=with_cpu)
    mb1 = None
    if world == 1:
        mb1 = mb_parse_leg(g, L, dict(CFG1, frames=300), "BASELINE config 1 on the GPU: CAVLC 1080p Baseline, 1 slice/frame, 300 frames",
                           3, 1, with_cpu=False, e2e_reps=0)
    mbc = None
    if args.cabac_frames > 0:
        fr = min(args.cabac_frames, 250)  # the CABAC generator is single-threaded (~0.08 s per frame)
        mbc = mb_parse_leg(g, L, dict(MB_CFG, frames=fr, profile_idc=77, transform_8x8=0, entropy_cabac=1,
                                      seed=MB_CFG["seed"] + rank),
                           "BASELINE config 3: CABAC 1080p Main, 16 slices/frame, %d generated frames" % fr,
                           3, 1, with_cpu=with_cpu, cabac=True, reps=(args.cabac_frames + fr - 1) // fr)
    mb4 = None
    if args.cfg4_streams > 0:
        mb4 = mb_parse_leg(g, L, dict(CFG4, frames=30, seed=CFG4["seed"] + rank),
                           "BASELINE config 4 share: 4K High CABAC (8x8 transform), 8 slices/frame, 30 frames per stream, "
                           "%d streams on this GPU (256 over 8 GPUs)" % args.cfg4_streams,
                           2, 1, with_cpu=False, cabac=True, reps=args.cfg4_streams, e2e_reps=0)
    fr_leg = frame_leg(g, L, args.frame_mb << 20, 5, 2, rank, with_cpu) if args.frame_mb > 0 else None
    rd = None
    if world == 1 and not args.no_reader and not args.no_cpu:
        rd = {"many_slices": reader_leg(L, dict(MB_CFG, frames=min(args.mb_frames, 250)),
                                        "CAVLC 1080p High, 16 slices/frame, %d frames" % min(args.mb_frames, 250)),
              "config1": reader_leg(L, dict(CFG1, frames=300), "BASELINE config 1: CAVLC 1080p Baseline, 1 slice/frame, 300 frames", reps=2)}
    if world > 1:
        for leg in (mb, mbc, mb4):
            if leg:
                leg["macroblocks_per_s"] = sum_over_ranks(leg["macroblocks_per_s"])
                if "e2e_macroblocks_per_s" in leg:
                    leg["e2e_macroblocks_per_s"] = sum_over_ranks(leg["e2e_macroblocks_per_s"])
        if fr_leg:
            fr_leg["gb_per_s"] = sum_over_ranks(fr_leg["gb_per_s"])

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "annexb_scan_strip (BASELINE config 2)",
                       "bytes_per_gpu": n_in, "nals_per_gpu": n_nal_shard, "rbsp_bytes_per_gpu": n_rbsp_shard,
                       "nal_bytes": "loguniform[64,262144]", "p_zero": 0.1875,
                       "start_codes": "3/4-byte mixed, 0-2 trailing zeros on 1/4 of NALs",
                       "l2": "input %.1f GiB >> 126 MB L2, no flush needed" % (n_in / 2**30),
                       "parallelism": "byte-range shards of ONE stream, 1 per GPU, seams %d bytes into each rank's piece "
                                      "(inside NALs), host merge after the timed region, no collective" % DELTA
                                      if world > 1 else "single GPU",
                       "kernel": SCAN_KERNEL, "gen_seconds": round(gen_s, 2), "numa": numa_note},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": measured_traffic(n_in) if world == 1 else None,
                         "traffic_source": "profiles/r02_traffic.json (ncu dram__bytes of this workload's launch; a committed "
                                           "measurement, not re-measured in this run; final build at 1 GiB: "
                                           "profiles/r02g_traffic_1gib.json, 0.97 x algorithmic)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg_bytes),
                         "note": "duration = CUDA-event step time on the launch stream: both scan7_kernel launches "
                                 "(main pass + deferred region heads) + the three finalize kernels; no memset"},
            "scan_only": scan_only,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(n_in), "d2h_bytes_per_step": e2e_d2h,
                    "ms_per_step": e2e_s * 1e3, "api": "h264gpu_split_strip_host (pinned host buffers)",
                    "kernel_note": "the host pipeline writes PACKED RBSP (annexb2::scan2_kernel per 128 MiB chunk), "
                                   "`value` times the in-place kernel; PCIe hides the difference",
                    "platform_ceiling": {"value": ceiling, "unit": UNIT,
                                         "note": "same run: one cudaMemcpyAsync H2D of the input and one D2H of the "
                                                 "output size, concurrently, per rank (max over ranks)"}},
            "e2e_reader": rd,
            "merged_table": {"verified": merged_ok, "merge_ms_host": merge_ms,
                             "note": "per-rank tables gathered to rank 0, h264gpu_merge_shard_inplace + merge_finish, "
                                     "RBSP lengths / order / totals compared with the generator's"},
            "strong_scaling": strong,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "cpu_baseline": cpu,
            "extra": {"macroblocks_per_s": mb["macroblocks_per_s"], "mb_parse": mb, "mb_parse_config1_gpu": mb1,
                      "mb_parse_cabac": mbc, "mb_parse_config4": mb4, "writer_frame": fr_leg},
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    g.close()


if __name__ == "__main__":
    main()
