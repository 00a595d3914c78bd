#!/usr/bin/env python
"""bench.py — headline benchmark: Annex-B NAL split + EPB strip (BASELINE.json config 2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size-mb M]

A step = one pass of the hot path (fused scan + strip kernel) over one 4 GiB synthetic
Annex-B buffer per GPU (weak scaling: rank r holds byte-range shard r of a 4N GiB
stream; no data-path collective — per-shard results are merged on the host).

value  : input GB/s with the buffer resident in HBM (CUDA events around K steps).
e2e    : same metric through h264gpu_split_strip_host — pinned HOST input, H2D, kernel,
         D2H of NAL tables + RBSP inside the timed region.
roofline: algorithmic bytes (N_in + N_rbsp + 24 B per NAL) / step time vs measured HBM peak.
cpu_baseline / --impl reference: the UNMODIFIED reference (oracle/_ref/libh264_ref.so:
         h264_find_nalu loop + h264_bs_read_bits(8) loop) on the box's host cores.
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


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(kernel_name, n_in):
    """DRAM bytes (read + write) of the dominant kernel per launch from the committed ncu
    capture of this very workload (profiles/r01_traffic.json, written by scripts/gpu_session.sh);
    None when no capture matches the kernel and input size."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)
        if t.get("bytes_in") == int(n_in) and t.get("kernel", "").split("<")[0] in kernel_name:
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


def ref_lib():
    so = os.path.join(ROOT, "oracle", "_ref", "libh264_ref.so")
    kind = "reference"
    if not os.path.exists(so):
        return None, None
    lib = C.CDLL(so)
    lib.ref_mt_split_strip.restype = C.c_double
    lib.ref_mt_split_strip.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p,
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    return lib, kind


def cpu_reference_run(stream, sample_bytes, cores, reps=1):
    """Reference scan + strip of stream[:sample_bytes] on `cores` threads; best of reps."""
    lib, kind = ref_lib()
    if lib is None:
        return None
    sample = stream[:sample_bytes]
    scratch = np.empty(len(sample) + 64, np.uint8)
    nn, rb = C.c_uint64(0), C.c_uint64(0)
    best = None
    for _ in range(reps):
        dt = lib.ref_mt_split_strip(sample.ctypes.data_as(C.c_void_p), len(sample), cores, 1,
                                    scratch.ctypes.data_as(C.c_void_p), C.byref(nn), C.byref(rb))
        best = dt if best is None else min(best, dt)
    return {"value": len(sample) / best / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "first %.0f MiB of the same stream, reference h264_find_nalu loop + "
                      "h264_bs_read_bits(8) strip loop, %d threads over byte ranges, best of %d"
                      % (len(sample) / 2**20, cores, reps),
            "seconds": best, "nals": nn.value}


MB_CFG = dict(width_mbs=120, height_mbs=68, slices_per_frame=16, profile_idc=100, transform_8x8=1,
              b_frames=1, num_ref_frames=2, idr_period=30, pct_skip=30, coef_density=60, seed=SEED)


def SCAN_KERNEL_NAME(gen=2):
    if gen == 7:
        return "annexb6::scan6p_kernel<4,strip> + fin_tiles/fin_order/fin_head/fin_table (RBSP in place per NAL, two tiles per CTA in flight)"
    if gen == 6:
        return "annexb6::scan6_kernel<8,strip> + fin_tiles/fin_order/fin_head/fin_table (RBSP in place per NAL, warp-autonomous tiles)"
    if gen == 5:
        return "annexb5::scan5_kernel<8,strip> + fin_tiles/fin_order/fin_head/fin_table (RBSP in place per NAL)"
    it = int(os.environ.get("H264GPU_SCAN_ITEMS", "108"))
    return ("annexb2::scan2_kernel<%d,strip>" % (it % 10)) if it >= 100 else \
        ("annexb::scan_kernel<%d,strip>" % it)


def mb_parse_leg(g, L, frames, steps, warmup, rank, with_cpu, cabac=False):
    """Slice-parallel macroblock parse (1080p, 16 slices/frame; CAVLC High or, cabac=True, the
    CABAC kernel on BASELINE config 3's shape): macroblocks/s device-resident, end-to-end from
    host buffers, and the reference reader on the host cores (CAVLC only: the reference parses
    no CABAC slice data)."""
    cfg = dict(MB_CFG, seed=MB_CFG["seed"] + rank)
    if cabac:
        cfg.update(profile_idc=77, transform_8x8=0, entropy_cabac=1)
    reps = 1
    if cabac and frames > 250:
        # the CABAC generator is single-threaded (~0.08 s per frame): generate 250 frames and lay
        # `reps` copies of them end to end (separate memory, separate slices; said in "workload")
        reps, frames = (frames + 249) // 250, 250
    stream, nmb, nsl, params = L.synth_video(frames=frames, want_params=True, **cfg)
    if reps > 1:
        P = np.frombuffer(params, L.SLICE_PARAMS)
        pad = (-len(stream)) % 16
        one = np.concatenate([stream, np.zeros(pad, np.uint8)])
        allp = []
        for k in range(reps):
            q = P.copy()
            q["nal_off"] += k * len(one)
            q["mb_out_off"] += k * nmb
            allp.append(q)
        stream = np.tile(one, reps)
        params = np.concatenate(allp).view(np.uint8)
        nmb, nsl, frames = nmb * reps, nsl * reps, frames * reps
    d_stream = g.alloc(len(stream) + 16)
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
    t1 = time.perf_counter()
    recs, res2 = parse_host(stream, params, nmb)
    e2e_s = time.perf_counter() - t1
    out = {"workload": "%s, 16 slices/frame, %d frames%s (%d slices, %d MBs, %.1f MB stream)"
                       % ("CABAC 1080p Main (BASELINE config 3 shape)" if cabac else "CAVLC 1080p High",
                          frames, " = %d copies of 250 generated frames" % reps if reps > 1 else "",
                          nsl, nmb, len(stream) / 1e6),
           "kernel": "cabac::cabac_parse_kernel" if cabac else "cavlc::cavlc_parse_kernel",
           "macroblocks_per_s": nmb / (ms / 1e3), "ms_per_step": ms, "parity_counts_ok": ok,
           "e2e_macroblocks_per_s": nmb / e2e_s,
           "e2e_note": "h264gpu_%s_parse_host: H2D stream+params, kernel, D2H 16 B/MB records"
                       % ("cabac" if cabac else "cavlc")}
    if cabac:
        out["cpu_baseline"] = None  # the reference has no CABAC slice-data parser
    if with_cpu and not cabac:
        lib, kind = ref_lib()
        if lib is not None:
            lib.ref_mt_parse.restype = C.c_double
            lib.ref_mt_parse.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int,
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
            cores = min(os.cpu_count() or 1, 256)
            bufs = (C.c_void_p * cores)(*[stream.ctypes.data] * cores)
            lens = (C.c_size_t * cores)(*[len(stream)] * cores)
            mbs, sl = C.c_uint64(0), C.c_uint64(0)
            best = None
            for _ in range(2):
                dt = lib.ref_mt_parse(bufs, lens, cores, 1, cores, C.byref(mbs), C.byref(sl))
                best = dt if best is None else min(best, dt)
            out["cpu_baseline"] = {"value": mbs.value / best, "unit": "macroblocks/s", "cores": cores,
                                   "kind": kind,
                                   "sample": "reference h264_reader_parse(SLICE_DATA) of the same stream, one "
                                             "reader per thread (%d threads, %d MBs total), best of 2" %
                                             (cores, mbs.value)}
    for d in (d_stream, d_params, d_rec, d_res):
        d.free()
    return out


def frame_leg(g, L, size, steps, warmup, rank):
    """Writer side (BASELINE config 5): EPB insertion + 4-byte start-code framing of `size` bytes
    of synthetic RBSP payloads, device-resident; GB/s of payload, roofline on N_rbsp + N_out."""
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
    peak, _ = measured_peak()
    out = {"workload": "EPB insert + framing of %.0f MiB RBSP in %d payloads (BASELINE config 5)" % (size / 2**20, n),
           "kernel": ("frame::frame_prepass + frame::frame_kernel<4>" if os.environ.get("H264GPU_FRAME_GEN") == "1"
                      else "frame::frame_prepass<8> + frame6::frame6_kernel<8> (persistent CTAs, 32 KiB tiles)"),
           "gb_per_s": len(rbsp) / (ms / 1e3) / 1e9, "ms_per_step": ms, "out_bytes": total,
           "roofline_frac": (len(rbsp) + total + 8 * (n + 1)) / (ms / 1e3) / 1e9 / peak}
    for d in (d_r, d_o, d_out, d_oo):
        d.free()
    return out


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    import libh264_b200 as L
    cores = min(os.cpu_count() or 1, 256)
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
    ap.add_argument("--scan-gen", type=int, default=6, choices=[2, 5, 6, 7],
                    help="6 / 5: RBSP in place per NAL (h264gpu_split_strip_inplace_dev; 6 = warp-autonomous "
                         "tiles, 5 = block-wide unit binning); 2: packed RBSP")
    ap.add_argument("--mb-frames", type=int, default=1000, help="frames of the macroblock-parse workload")
    ap.add_argument("--cabac-frames", type=int, default=1000,
                    help="frames of the CABAC macroblock-parse workload (0 = skip)")
    ap.add_argument("--frame-mb", type=int, default=1024,
                    help="MiB of RBSP payloads for the writer leg (BASELINE config 5; 0 = skip)")
    args = ap.parse_args()
    if args.scan_gen >= 5:
        os.environ["H264GPU_INPLACE_GEN"] = str(args.scan_gen)
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import libh264_b200 as L
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    g = L.Gpu(local_rank)
    size = args.size_mb << 20
    nthreads = max(1, min(os.cpu_count() or 1, 64) // max(1, min(world, 8)))

    # ---- workload: rank r's shard, generated straight into pinned host memory ----
    t0 = time.time()
    pin_in = g.pinned(size + 4096)
    stream, rbsp_ref, offs = make_workload(L, size, SEED + rank, out=pin_in.array, nthreads=nthreads)
    n_in = len(stream)
    n_nal_expected = len(offs) - 1
    # The generator puts 0-2 zero bytes after a quarter of the NALs.  In front of a start code they
    # are cut off (the boundary is the first 00 00 00); at the very end of the input the reference
    # keeps them in the last NAL (h264_find_nalu ends it at `len`), so they count as RBSP bytes
    # where the shard has no right neighbour (the last rank; every rank in the stand-alone e2e leg).
    tz = 0
    while tz < len(stream) and stream[len(stream) - 1 - tz] == 0:
        tz += 1
    n_rbsp_e2e_expected = len(rbsp_ref) + tz
    n_rbsp_expected = len(rbsp_ref) + (tz if rank == world - 1 else 0)
    del rbsp_ref
    gen_s = time.time() - t0

    cap = n_nal_expected + 1024
    d_in = g.alloc(n_in + 16)
    d_rbsp = g.alloc(n_in + 16)
    d_tab = g.alloc(cap * 8 * 4)
    d_res = g.alloc(C.sizeof(L.ScanResult))
    L._check(g.lib.h264gpu_memcpy_h2d(g.h, C.c_void_p(d_in.ptr), pin_in.array.ctypes.data_as(C.c_void_p),
                                     n_in, None), "h2d")
    g.sync()

    edge = None
    if world > 1:  # shard r of a 4N GiB stream: neighbours' bytes are setup, not data path
        import torch
        mine = torch.tensor([int(stream[0]), int(stream[1]), int(stream[-2]), int(stream[-1])],
                            dtype=torch.int32, device="cuda")
        allh = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        edge = L.ShardEdge()
        if rank > 0:
            edge.has_left, edge.left[0], edge.left[1] = 1, int(allh[rank - 1][2]), int(allh[rank - 1][3])
            edge.assume_in = 1
        if rank < world - 1:
            edge.has_right, edge.right[0], edge.right[1] = 1, int(allh[rank + 1][0]), int(allh[rank + 1][1])

    def step():
        if args.scan_gen >= 5:
            g.split_strip_inplace_dev(d_in.ptr, n_in, d_rbsp.ptr, d_tab.ptr, d_tab.ptr + cap * 8,
                                      d_tab.ptr + cap * 16, d_tab.ptr + cap * 24, cap, d_res.ptr,
                                      base=rank * n_in, edge=edge)
        else:
            g.split_strip_dev(d_in.ptr, n_in, d_rbsp.ptr, d_tab.ptr, d_tab.ptr + cap * 8, d_tab.ptr + cap * 16,
                              cap, d_res.ptr, base=rank * n_in, edge=edge)

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- device-resident leg -------------------------------------------------
    for _ in range(args.warmup):
        step()
    g.sync()
    barrier()
    launches0 = g.launch_count()
    tm = g.timer()
    g.timer_start(tm)
    for _ in range(args.steps):
        step()
    g.timer_stop(tm)
    ms_total = g.timer_ms(tm)
    g.sync()
    barrier()
    launches = g.launch_count() - launches0
    ms_step = max_over_ranks(ms_total / args.steps)

    res = L.ScanResult.from_buffer_copy(d_res.download().tobytes())
    ok = (res.n_nal == n_nal_expected and res.rbsp_bytes == n_rbsp_expected)
    if not ok:
        print("bench: result mismatch: n_nal %d (want %d) rbsp %d (want %d)" %
              (res.n_nal, n_nal_expected, res.rbsp_bytes, n_rbsp_expected), file=sys.stderr)
        sys.exit(1)

    total_in = sum_over_ranks(float(n_in))
    value = total_in / (ms_step / 1e3) / 1e9
    alg_bytes = n_in + res.rbsp_bytes + 24 * res.n_nal
    achieved = alg_bytes / (ms_total / args.steps / 1e3) / 1e9
    peak, peak_src = measured_peak()

    # ---- end-to-end leg: pinned host in -> tables + RBSP in pinned host out ----
    pin_rbsp = g.pinned(n_in + 4096)
    tabs = g.pinned(cap * 8 * 3)
    tv = tabs.view(np.uint64)
    out = dict(start=tv[:cap], end=tv[cap:2 * cap], rbsp_off=tv[2 * cap:3 * cap], rbsp=pin_rbsp.array)
    h = g.split_strip_host(stream, cap=cap, out=out)  # warm-up (allocates the chunk pipeline)
    if h["n_nal"] != n_nal_expected or h["rbsp_bytes"] != n_rbsp_e2e_expected:
        print("bench: e2e result mismatch", file=sys.stderr)
        sys.exit(1)
    barrier()
    e2e_t = []
    for _ in range(args.e2e_steps):
        t1 = time.perf_counter()
        h = g.split_strip_host(stream, cap=cap, out=out)
        e2e_t.append(time.perf_counter() - t1)
    barrier()
    e2e_s = max_over_ranks(sum(e2e_t) / len(e2e_t))
    e2e_val = total_in / e2e_s / 1e9
    clocks = sampler.stop()

    # ---- CPU baseline (rank 0, N=1): the unmodified reference on the host cores ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = min(os.cpu_count() or 1, 256)
        sample = min(n_in, max(64 << 20, int(0.12e9 * cores * 4)))
        cpu = cpu_reference_run(stream, sample, cores, reps=3)

    # ---- second half of the metric: macroblocks/s of the slice-parallel CAVLC parse ----
    for d in (d_in, d_rbsp, d_tab, d_res):
        d.free()
    mb = mb_parse_leg(g, L, args.mb_frames, max(3, args.steps // 4), 2, rank,
                      with_cpu=(rank == 0 and world == 1 and not args.no_cpu))
    mbc = mb_parse_leg(g, L, args.cabac_frames, 3, 1, rank, with_cpu=False, cabac=True) \
        if args.cabac_frames > 0 else None
    fr = frame_leg(g, L, args.frame_mb << 20, 5, 2, rank) if args.frame_mb > 0 else None
    if world > 1:
        mb["macroblocks_per_s"] = sum_over_ranks(mb["macroblocks_per_s"])
        mb["e2e_macroblocks_per_s"] = sum_over_ranks(mb["e2e_macroblocks_per_s"])
        if mbc:
            mbc["macroblocks_per_s"] = sum_over_ranks(mbc["macroblocks_per_s"])
        if fr:
            fr["gb_per_s"] = sum_over_ranks(fr["gb_per_s"])

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "annexb_scan_strip (BASELINE config 2)",
                       "bytes_per_gpu": n_in, "nals_per_gpu": int(res.n_nal),
                       "rbsp_bytes_per_gpu": int(res.rbsp_bytes),
                       "nal_bytes": "loguniform[64,262144]", "p_zero": 0.1875,
                       "start_codes": "3/4-byte mixed, 0-2 trailing zeros on 1/4 of NALs",
                       "l2": "input %.1f GiB >> 126 MB L2, no flush needed" % (n_in / 2**30),
                       "parallelism": "byte-range shards, 1 per GPU, host merge, no collective",
                       "kernel": SCAN_KERNEL_NAME(args.scan_gen),
                       "gen_seconds": round(gen_s, 2)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": measured_traffic(SCAN_KERNEL_NAME(args.scan_gen), n_in) if world == 1 else None,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg_bytes),
                         "note": "duration = CUDA-event step time (kernel + two tiny memsets) on the launch stream"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(n_in),
                    "d2h_bytes_per_step": int(res.rbsp_bytes + 24 * res.n_nal),
                    "ms_per_step": e2e_s * 1e3, "api": "h264gpu_split_strip_host (pinned host buffers)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "cpu_baseline": cpu,
            "extra": {"macroblocks_per_s": mb["macroblocks_per_s"], "mb_parse": mb,
                      "mb_parse_cabac": mbc, "writer_frame": fr},
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    g.close()


if __name__ == "__main__":
    main()
