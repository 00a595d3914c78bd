"""Per-tile phase timing of the scan kernel (H264GPU_SCAN_TRACE): where a tile's life goes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = "gpurun_out/scan_trace.bin"
os.environ["H264GPU_SCAN_TRACE"] = path
import libh264_b200 as L
import ctypes as C

size = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 1024 << 20
gen = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = L.Gpu(0)
offs = L.synth_offsets(5, size)
rbsp = L.synth_payloads(5, offs)
stream, _ = L.synth_annexb(5, rbsp, offs)
n = len(stream); cap = len(offs) + 1024
d_in = g.alloc(n + 16); d_in.upload(stream)
d_rbsp = g.alloc(n + 16); d_tab = g.alloc(cap * 32); d_res = g.alloc(C.sizeof(L.ScanResult))
for rep in range(3):
    if gen >= 5:
        os.environ["H264GPU_INPLACE_GEN"] = str(gen)
        g.split_strip_inplace_dev(d_in.ptr, n, d_rbsp.ptr, d_tab.ptr, d_tab.ptr + cap * 8, d_tab.ptr + cap * 16,
                                  d_tab.ptr + cap * 24, cap, d_res.ptr)
    else:
        g.split_strip_dev(d_in.ptr, n, d_rbsp.ptr, d_tab.ptr, d_tab.ptr + cap * 8, d_tab.ptr + cap * 16, cap, d_res.ptr)
    g.sync()
t = np.fromfile(path, np.uint64).reshape(-1, 8).astype(np.int64)
t0 = t[:, 0].min()
print("tiles", len(t), "kernel span us", (t[:, 4].max() - t0) / 1e3)
names = ["ticket->loaded", "classify..count (P1-P2)", "look-back (gen5: + tail segments)", "emit (gen5: head segment)"]
for k in range(4):
    d = (t[:, k + 1] - t[:, k]) / 1e3
    print("%-26s mean %7.2f  p10 %7.2f  p50 %7.2f  p90 %7.2f  p99 %7.2f us" %
          (names[k], d.mean(), *np.percentile(d, [10, 50, 90, 99])))
if gen >= 6:
    # marks: 0 ticket, 1 loaded, 5 classified, 6 prefix + event walk, 2 after barrier 2, 3 look-back
    # done, 7 row loop done (warp 0), 4 end
    for nm, x, y in (("load wait", 0, 1), ("classify (warp 0)", 1, 5), ("prefix + walk (warp 0)", 5, 6), ("emit (warp 0)", 3, 4),
                     ("barrier 2 wait", 6, 2), ("aggregate + look-back", 2, 3)):
        d = (t[:, y] - t[:, x]) / 1e3
        print("  %-26s mean %7.2f  p50 %7.2f  p90 %7.2f us" % (nm, d.mean(), *np.percentile(d, [50, 90])))
if gen == 5:
    for nm, x, y in (("seg table+premark+bin", 3, 5), ("T scan", 5, 6), ("clean copy (tid 0)", 6, 7), ("dirty+edges (tid 0)", 7, 4)):
        d = (t[:, y] - t[:, x]) / 1e3
        print("  %-26s mean %7.2f  p50 %7.2f  p90 %7.2f us" % (nm, d.mean(), *np.percentile(d, [50, 90])))
life = (t[:, 4] - t[:, 0]) / 1e3
print("tile life mean %.2f us; tiles in flight ~ %.0f" % (life.mean(), life.sum() / ((t[:, 4].max() - t0) / 1e3)))
# convoy: how far behind the ticket order do look-backs complete
order = np.argsort(t[:, 3])
print("look-back completion order inversions:", int((np.diff(order) < 0).sum()))
np.save("gpurun_out/scan_trace_sample.npy", t[::97])
