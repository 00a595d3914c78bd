#!/usr/bin/env python
"""The writer kernel (EPB insert + framing) on one GPU, device resident, CUDA events; prints
GB/s of payload, the roofline fraction and a hash of the output (to compare builds).
Knobs: H264GPU_FRAME_GEN (6 | 7), H264GPU_FRAME7_NBUF (2 | 1), H264GPU_FRAME7_ROWS (1,2,4,6,8).
   python scripts/frame_ab.py [--size-mb 1024] [--steps 10] [--hash 1]"""
import argparse, hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libh264_b200 as L
from bench import SEED, measured_peak

ap = argparse.ArgumentParser()
ap.add_argument("--size-mb", type=int, default=1024)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--hash", type=int, default=1)
args = ap.parse_args()
g = L.Gpu(0)
offs = L.synth_offsets(SEED + 100, args.size_mb << 20)
rbsp = L.synth_payloads(SEED + 100, offs)
n = len(offs) - 1
cap = len(rbsp) + len(rbsp) // 2 + 4 * n + 64
d_r = g.alloc(len(rbsp) + 16); d_r.upload(rbsp)
d_o = g.alloc(len(offs) * 8); d_o.upload(offs)
d_out = g.alloc(cap + 16)
d_oo = g.alloc((n + 2) * 8)
def step():
    g.frame_dev(d_r.ptr, d_o.ptr, n, 4, d_out.ptr, cap, d_oo.ptr, d_oo.ptr + (n + 1) * 8)
for _ in range(args.warmup):
    step()
g.sync()
tm = g.timer()
g.timer_start(tm)
for _ in range(args.steps):
    step()
g.timer_stop(tm)
ms = g.timer_ms(tm) / args.steps
oo = d_oo.download(dtype=np.uint64)
total = int(oo[n + 1])
peak, _ = measured_peak()
alg = len(rbsp) + total + 8 * (n + 1)
res = {"env": {k: v for k, v in os.environ.items() if k.startswith("H264GPU_FRAME")}, "size_mb": args.size_mb,
       "payloads": n, "ms": ms, "GBps_payload": len(rbsp) / ms / 1e6, "frac": alg / (ms / 1e3) / 1e9 / peak, "total": total}
if args.hash:
    out = d_out.download(dtype=np.uint8)[:total]
    h = hashlib.blake2b(digest_size=8)
    h.update(out.tobytes()); h.update(oo[:n + 1].tobytes())
    res["hash"] = h.hexdigest()
print(json.dumps(res))
