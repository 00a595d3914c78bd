#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_lib_gpu.py tests/test_cavlc.py tests/test_cabac.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python - <<'PY' 2>&1 | tail -12
import sys, time; sys.path.insert(0, 'tests')
import numpy as np
import test_host_lib as T, libh264_b200 as L
lib = T.harness()
for name, cfg in (("config1 (1080p Baseline, 1 slice/frame, 300 frames)", dict(frames=300, width_mbs=120, height_mbs=68, slices_per_frame=1, profile_idc=66, transform_8x8=0, b_frames=0, num_ref_frames=1, idr_period=30, pct_skip=30, coef_density=60, seed=0x264)),
                  ("1080p High CAVLC, 16 slices/frame, 250 frames", dict(frames=250, width_mbs=120, height_mbs=68, slices_per_frame=16, profile_idc=100, transform_8x8=1, b_frames=1, num_ref_frames=2, idr_period=30, pct_skip=30, coef_density=60, seed=7))):
    s, nmb, nsl = L.synth_video(**cfg)
    for fl in (1, 0):
        t_ref, c_ref = T.time_parse(lib, T.REF, s, fl, 2)
        t_our, c_our = T.time_parse(lib, T.OURS, s, fl, 3)
        assert (c_ref == c_our).all(), (c_ref, c_our)
        print("%s flags=%d: %d bytes %d MBs: reference 1 core %.1f ms (%.2f M MB/s), ours %.1f ms (%.2f M MB/s) = %.1fx" %
              (name, fl, len(s), nmb, t_ref * 1e3, nmb / t_ref / 1e6, t_our * 1e3, nmb / t_our / 1e6, t_ref / t_our), flush=True)
PY
