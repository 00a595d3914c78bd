"""h264_reader_parse of this repo's libh264.so (GPU scan + GPU slice-data parse + callback
replay) against the compiled reference, callback for callback, through the public API."""
import numpy as np
import pytest

import libh264_b200 as L
import test_host_lib as T

pytestmark = [pytest.mark.gpu, T.pytestmark]

CFGS = [
    dict(frames=5, width_mbs=20, height_mbs=12, slices_per_frame=3, profile_idc=66, transform_8x8=0,
         b_frames=0, num_ref_frames=1, idr_period=3, pct_skip=30, coef_density=50, seed=11),
    dict(frames=6, width_mbs=40, height_mbs=30, slices_per_frame=5, profile_idc=100, transform_8x8=1,
         b_frames=1, num_ref_frames=2, idr_period=4, pct_skip=20, coef_density=60, seed=12),
    dict(frames=3, width_mbs=120, height_mbs=68, slices_per_frame=16, profile_idc=100, transform_8x8=1,
         b_frames=1, num_ref_frames=2, idr_period=30, pct_skip=30, coef_density=60, seed=13),
]


@pytest.mark.parametrize("cfg", CFGS, ids=lambda c: "seed%d" % c["seed"])
def test_bulk_parse_with_slice_data(cfg):
    lib = T.harness()
    stream, nmb, nsl = L.synth_video(**cfg)
    ours = T.trace(lib, T.OURS, stream, flags=1, mode=0)
    ref = T.trace(lib, T.REF, stream, flags=1, mode=0)
    T.assert_same_trace(ours, ref, "bulk slice data")
    assert sum(1 for t, _ in T.split_log(ours) if t == 9) == nmb  # every macroblock delivered


def test_bulk_parse_headers_only_and_per_nalu_slice_data():
    lib = T.harness()
    stream, nmb, nsl = L.synth_video(**CFGS[1])
    T.assert_same_trace(T.trace(lib, T.OURS, stream, 0, 0), T.trace(lib, T.REF, stream, 0, 0), "flags=0")
    # single-NAL entry point with slice data: one-slice launches
    T.assert_same_trace(T.trace(lib, T.OURS, stream, 1, 1), T.trace(lib, T.REF, stream, 1, 1), "per NAL")


def test_bulk_parse_mixed_and_damaged():
    lib = T.harness()
    rng = np.random.default_rng(9)
    video, _, _ = L.synth_video(**CFGS[0])
    hdrs = T.gen(lib, T.REF, 321, rounds=4)            # SEI/AUD/filler/CABAC/FMO headers in between
    s = np.concatenate([hdrs, video, hdrs[:300], video])
    T.assert_same_trace(T.trace(lib, T.OURS, s, 1, 0), T.trace(lib, T.REF, s, 1, 0), "mixed")
    for it in range(6):
        d = video.copy()
        for p in rng.integers(40, len(d), 4):
            d[p] ^= 1 << int(rng.integers(0, 8))
        a, b = T.trace(lib, T.OURS, d, 1, 0), T.trace(lib, T.REF, d, 1, 0)
        T.assert_same_trace(a, b, "damaged %d" % it)
