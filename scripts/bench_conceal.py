#!/usr/bin/env python
"""N4 throughput probe: concealment slices per second through h264gpu_conceal_slices_dev
(device-resident descriptors, payloads out) + the K3 framing of the result.  Inputs: slice
descriptors of slices the REFERENCE wrote (tests/test_conceal.py reference_slices), tiled.
   python scripts/bench_conceal.py [--slices 65536]"""
import argparse, ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libh264_b200 as L
import test_conceal as TC

ap = argparse.ArgumentParser()
ap.add_argument("--slices", type=int, default=65536)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
stream, recs, hdr = TC.reference_slices(11, rounds=20)
p = TC.params_of(recs)
reps = (args.slices + len(p) - 1) // len(p)
P = np.tile(p, reps)[:args.slices]
n = len(P)
bound = int((P["hdr_bits"] // 8 + 2 * P["mb_count"] + 16).sum())
g = L.Gpu(0)
d_p, d_h = g.alloc(P.nbytes), g.alloc(len(hdr))
d_pay, d_off = g.alloc(bound + 64), g.alloc((n + 2) * 8)
d_out, d_oo = g.alloc(bound * 3 // 2 + 4 * n + 64), g.alloc((n + 3) * 8)
d_p.upload(P.view(np.uint8)); d_h.upload(hdr)
lib = g.lib
def step():
    L._check(lib.h264gpu_conceal_slices_dev(g.h, C.c_void_p(d_p.ptr), C.c_uint32(n), C.c_void_p(d_h.ptr),
                                            C.c_void_p(d_pay.ptr), C.c_uint64(bound), C.c_void_p(d_off.ptr), None), "conceal")
    L._check(lib.h264gpu_frame_dev(g.h, C.c_void_p(d_pay.ptr), C.c_void_p(d_off.ptr), C.c_uint64(n), C.c_int(4),
                                   C.c_void_p(d_out.ptr), C.c_uint64(d_out.nbytes), C.c_void_p(d_oo.ptr),
                                   C.c_void_p(d_oo.ptr + (n + 1) * 8), None), "frame")
step(); g.sync()
tm = g.timer(); g.timer_start(tm)
for _ in range(args.reps):
    step()
g.timer_stop(tm); ms = g.timer_ms(tm) / args.reps
tot = int(d_oo.download(dtype=np.uint64)[n])
mbs = int(P["mb_count"].sum())
want = int(np.tile(recs["ref_len"], reps)[:n].sum())
print("conceal: %d slices (%d MBs, %.1f MB out) %.3f ms  %.2f M slices/s  %.1f M MB/s  bytes ok=%s" %
      (n, mbs, tot / 1e6, ms, n / ms / 1e3, mbs / ms / 1e3, tot == want))
