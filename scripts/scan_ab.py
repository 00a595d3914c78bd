#!/usr/bin/env python
"""The in-place scan + strip kernel and the scan-only kernel on one GPU: same buffer, device
resident, CUDA events (tuning knobs: H264GPU_SCAN7_REGIONS / _PF / _NAP).
   python scripts/scan_ab.py [--size-mb 4096] [--steps 10]"""
import argparse, ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libh264_b200 as L
from bench import make_workload, SEED

ap = argparse.ArgumentParser()
ap.add_argument("--size-mb", type=int, default=4096)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--cap-extra", type=int, default=1024)
ap.add_argument("--alloc-extra", type=int, default=16)
ap.add_argument("--stage", type=int, default=0, help="1: copy through a second pinned buffer like bench.py")
ap.add_argument("--sampler", type=int, default=0, help="1: nvidia-smi clock sampler running like bench.py")
ap.add_argument("--order", default="strip,only")
ap.add_argument("--dev", type=int, default=0, help="CUDA device index")
ap.add_argument("--regrow", type=int, default=0, help="after timing, grow the scan workspace N times (free + allocate) and time again")
ap.add_argument("--threads", type=int, default=0, help="generator threads (0 = all)")
args = ap.parse_args()
g = L.Gpu(args.dev)
size = args.size_mb << 20
pin = g.pinned(size + 4096)
stream, rbsp_ref, offs = make_workload(L, size, SEED, out=pin.array, nthreads=args.threads or min(os.cpu_count() or 1, 64))
n_in, n_nal = len(stream), len(offs) - 1
tz = 0
while stream[n_in - 1 - tz] == 0:
    tz += 1
n_rbsp = len(rbsp_ref) + tz
cap = n_nal + args.cap_extra
d_in, d_rbsp = g.alloc(n_in + args.alloc_extra), g.alloc(n_in + args.alloc_extra)
d_tab, d_res = g.alloc(cap * 32), g.alloc(C.sizeof(L.ScanResult))
def one_call():
    g.split_strip_inplace_dev(d_in.ptr, n_in, d_rbsp.ptr, d_tab.ptr, d_tab.ptr + cap * 8,
                              d_tab.ptr + cap * 16, d_tab.ptr + cap * 24, cap, d_res.ptr)
    g.sync()
if args.stage == 3:  # the scan workspace exists before the second pinned buffer does
    one_call()
if args.stage == 6:  # the library's own answer: reserve the workspace right after the stream buffers
    L._check(g.lib.h264gpu_scan_reserve(g.h, C.c_uint64(n_in), C.c_uint64(cap)), "reserve")
if args.stage == 5:  # a first small launch (kernel loaded, small workspace) before the second pinned buffer
    g.split_strip_inplace_dev(d_in.ptr, 1 << 20, d_rbsp.ptr, d_tab.ptr, d_tab.ptr + cap * 8,
                              d_tab.ptr + cap * 16, d_tab.ptr + cap * 24, cap, d_res.ptr)
    g.sync()
print("d_in %x d_rbsp %x d_tab %x" % (d_in.ptr, d_rbsp.ptr, d_tab.ptr), file=sys.stderr)
if args.stage:
    stage = g.pinned(n_in + 64)
    print("stage %x" % stage.array.ctypes.data, file=sys.stderr)
    if args.stage != 4:
        stage.array[:n_in] = stream
        L._check(g.lib.h264gpu_memcpy_h2d(g.h, C.c_void_p(d_in.ptr), stage.array.ctypes.data_as(C.c_void_p), n_in, None), "h2d")
    else:  # 4: the buffer exists but is never touched or copied from
        d_in.upload(stream)
    g.sync()
    if args.stage == 2:
        stage.free()
        stage = None
else:
    d_in.upload(stream)
g.sync()
if args.sampler:
    from bench import ClockSampler
    smp = ClockSampler(0)
    smp.start()
out = {}
for strip in [x == "strip" for x in args.order.split(",")]:
    for gen in (7,):
        def step():
            g.split_strip_inplace_dev(d_in.ptr, n_in, d_rbsp.ptr if strip else 0, d_tab.ptr, d_tab.ptr + cap * 8,
                                      d_tab.ptr + cap * 16, d_tab.ptr + cap * 24, cap, d_res.ptr)
        for _ in range(args.warmup):
            step()
        g.sync()
        tm = g.timer()
        g.timer_start(tm)
        for _ in range(args.steps):
            step()
        g.timer_stop(tm)
        ms = g.timer_ms(tm) / args.steps
        res = L.ScanResult.from_buffer_copy(d_res.download().tobytes())
        ok = res.n_nal == n_nal and (not strip or res.rbsp_bytes == n_rbsp) and res.reserved == 0
        key = "gen%d_%s" % (gen, "strip" if strip else "scan_only")
        out[key] = {"ms": ms, "GBps": n_in / ms / 1e6, "ok": bool(ok), "n_nal": int(res.n_nal),
                    "rbsp_bytes": int(res.rbsp_bytes)}
        print(key, out[key], flush=True)
for k in range(args.regrow):
    # a larger reservation frees the workspace and allocates a new one somewhere else
    L._check(g.lib.h264gpu_scan_reserve(g.h, C.c_uint64(n_in + (k + 1) * (n_in >> 2)), C.c_uint64(cap)), "reserve")
    def step():
        g.split_strip_inplace_dev(d_in.ptr, n_in, d_rbsp.ptr, d_tab.ptr, d_tab.ptr + cap * 8,
                                  d_tab.ptr + cap * 16, d_tab.ptr + cap * 24, cap, d_res.ptr)
    for _ in range(args.warmup):
        step()
    g.sync()
    tm = g.timer()
    g.timer_start(tm)
    for _ in range(args.steps):
        step()
    g.timer_stop(tm)
    out["regrow%d" % (k + 1)] = {"ms": g.timer_ms(tm) / args.steps}
    print("dev", args.dev, "regrow", k + 1, out["regrow%d" % (k + 1)], flush=True)
out["dev"] = args.dev
print(json.dumps(out))
