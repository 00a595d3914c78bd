#!/usr/bin/env python
"""Per-span phase trace of scan7_kernel (H264GPU_SCAN_TRACE): where a span's life goes.
   python scripts/scan7_trace.py [--size-mb 1024]"""
import argparse, ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libh264_b200 as L
from bench import make_workload, SEED
ap = argparse.ArgumentParser()
ap.add_argument("--size-mb", type=int, default=1024)
args = ap.parse_args()
g = L.Gpu(0)
size = args.size_mb << 20
stream, rbsp_ref, offs = make_workload(L, size, SEED, nthreads=min(os.cpu_count() or 1, 64))
n_in, cap = len(stream), len(offs) + 1024
d_in, d_rbsp = g.alloc(n_in + 16), g.alloc(n_in + 16)
L._check(g.lib.h264gpu_scan_reserve(g.h, C.c_uint64(n_in), C.c_uint64(cap)), "reserve")  # workspace next to the stream buffers
d_tab, d_res = g.alloc(cap * 32), g.alloc(C.sizeof(L.ScanResult))
d_in.upload(stream)
def step():
    g.split_strip_inplace_dev(d_in.ptr, n_in, d_rbsp.ptr, d_tab.ptr, d_tab.ptr + cap * 8, d_tab.ptr + cap * 16,
                              d_tab.ptr + cap * 24, cap, d_res.ptr)
for _ in range(3):
    step()
g.sync()
path = "/tmp/scan7_trace.bin"
os.environ["H264GPU_SCAN_TRACE"] = path
step()
g.sync()
del os.environ["H264GPU_SCAN_TRACE"]
tr = np.fromfile(path, np.uint64).reshape(-1, 8).astype(np.int64)
tr = tr[tr[:, 6] > 0][:, :7]
np.save(os.path.join(ROOT, "gpurun_out", "scan7_trace.npy"), (tr - tr.min()).astype(np.int32))
names = ["load wait + clear", "classify -> chain word", "prefix + events", "look-back",
         "event records + rows", "byte-exact pass"]
life = (tr[:, 6] - tr[:, 0]) / 1e3
print("spans %d, kernel span %.1f us, span life mean %.2f us p50 %.2f p99 %.2f" %
      (len(tr), (tr[:, 6].max() - tr[:, 0].min()) / 1e3, life.mean(), np.median(life), np.percentile(life, 99)))
for k, n in enumerate(names):
    d = (tr[:, k + 1] - tr[:, k]) / 1e3
    print("  %-26s mean %6.2f us  p50 %6.2f  p90 %6.2f  p99 %6.2f  (%4.1f %% of life)" %
          (n, d.mean(), np.median(d), np.percentile(d, 90), np.percentile(d, 99), 100 * d.sum() / life.sum()))
