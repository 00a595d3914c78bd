#!/usr/bin/env python
"""Phase times of this repo's h264_reader_parse(SLICE_DATA) (H264_READER_TIMING=1) on the bench's
two reader workloads.   python scripts/reader_phases.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["H264_READER_TIMING"] = "1"
import libh264_b200 as L
import test_host_lib as T
from bench import MB_CFG, CFG1
lib = T.harness()
for name, cfg in (("4000 slices", dict(MB_CFG, frames=250)), ("config 1, 300 slices", dict(CFG1, frames=300))):
    stream, nmb, nsl = L.synth_video(**cfg)
    print("==", name, nmb, "macroblocks", file=sys.stderr, flush=True)
    t, c = T.time_parse(lib, T.OURS, stream, 1, 3)
    print("   best of runs: %.2f ms, %.1f M MB/s" % (t * 1e3, nmb / t / 1e6), file=sys.stderr, flush=True)
