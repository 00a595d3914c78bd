#!/usr/bin/env python
"""Phase trace of the gen-7 writer kernel: runs it once with H264GPU_FRAME_TRACE and prints where
a span's life goes.  Per span: [0] ticket known  [1] bulk copy issued  [2] classified
[3] published  [4] look-back begins  [5] look-back done  [6] emitted  [7] look-back misses
(level 1 | level 2 << 10 | super_p polls << 20 | resolved itself << 31).
   python scripts/frame7_trace.py [--size-mb 1024]"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libh264_b200 as L
from bench import SEED

ap = argparse.ArgumentParser()
ap.add_argument("--size-mb", type=int, default=1024)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "frame7_trace.bin"))
args = ap.parse_args()
os.environ.setdefault("H264GPU_FRAME_GEN", "7")
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
for _ in range(3):
    step()
g.sync()
os.environ["H264GPU_FRAME_TRACE"] = args.out
step()
g.sync()
os.environ.pop("H264GPU_FRAME_TRACE")
tr = np.fromfile(args.out, np.uint64).reshape(-1, 8)
ns = len(tr)
t = tr[:, :7].astype(np.int64)
t0 = t[:, 1][t[:, 1] > 0].min()
end = t[:, 6].max()
print("spans %d, kernel span %.1f us" % (ns, (end - t0) / 1e3))
ok = (t[:, 0] > 0)
names = ["ticket -> copy issued", "copy issued -> classified", "classified -> published", "published -> look-back begins",
         "look-back", "emit"]
for k in range(6):
    sel = ok & (t[:, k] > 0) & (t[:, k + 1] > 0)
    d = (t[sel, k + 1] - t[sel, k]) / 1e3
    print("%-32s mean %8.2f us  median %8.2f  p90 %8.2f  p99 %8.2f  max %8.2f" %
          (names[k], d.mean(), np.median(d), np.percentile(d, 90), np.percentile(d, 99), d.max()))
life = (t[ok, 6] - t[ok, 0]) / 1e3
print("life (ticket -> emitted)         mean %8.2f us  median %8.2f" % (life.mean(), np.median(life)))
m = tr[:, 7]
print("look-back misses per span: level 1 %.2f, level 2 %.2f, super_p polls %.2f, resolved itself %.3f" %
      ((m & 1023).mean(), ((m >> 10) & 1023).mean(), ((m >> 20) & 1023).mean(), ((m >> 31) & 1).mean()))
# how far ahead of its use is a span's word published: the look-back of span s needs all spans < s
pub = t[:, 3]
cummax = np.maximum.accumulate(pub)
need = np.concatenate([[0], cummax[:-1]])  # latest publication among the predecessors
slack = (t[:, 4] - need)[1:] / 1e3
print("look-back begins - last predecessor published: mean %.2f us, median %.2f, p10 %.2f, p1 %.2f, min %.2f" %
      (slack.mean(), np.median(slack), np.percentile(slack, 10), np.percentile(slack, 1), slack.min()))
# the same by position in the stream (eighths)
for i in range(8):
    a, b = ns * i // 8, ns * (i + 1) // 8
    sel = slice(max(a, 1), b)
    lb = (t[sel, 5] - t[sel, 4]) / 1e3
    it = (t[sel, 6] - t[sel, 1]) / 1e3
    print("  spans %7d-%7d: look-back mean %7.2f us, copy issued -> emitted mean %7.2f us, publish lag of the slowest predecessor %7.2f us" %
          (a, b, lb.mean(), it.mean(), -np.minimum(slack[a:b - 1], 0).mean()))
# which spans publish late relative to their ticket
lag = (t[:, 3] - t[:, 0]) / 1e3
sel = ok & (t[:, 3] > 0)
print("ticket -> published: mean %.2f us, median %.2f, p99 %.2f, max %.2f" %
      (lag[sel].mean(), np.median(lag[sel]), np.percentile(lag[sel], 99), lag[sel].max()))
