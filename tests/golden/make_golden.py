#!/usr/bin/env python
"""Generate tests/golden/cavlc_*.npz from the COMPILED REFERENCE (dev container only).

For a few tiny synthetic CAVLC streams this records what the unmodified reference
(oracle/_ref/libh264_ref.so, h264_reader_parse with H264_READER_FLAGS_SLICE_DATA)
delivers: per-macroblock (mb_addr, mb_type), the checksum of ctx->mb
(include/h264gpu_slice.h), slice_data_end counts, and the slice parameter blocks
its ctx held.  The streams themselves are regenerated from their configs by the
deterministic generator (libh264_b200/csrc/synth_video.c); a CRC pins them.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import libh264_b200 as L  # noqa: E402
import support as S  # noqa: E402

CASES = {
    "baseline": dict(width_mbs=8, height_mbs=6, frames=6, idr_period=3),
    "main_b_multislice": dict(width_mbs=11, height_mbs=9, frames=8, slices_per_frame=3, b_frames=1,
                              num_ref_frames=3, profile_idc=77),
    "high_t8": dict(width_mbs=10, height_mbs=7, frames=6, slices_per_frame=2, profile_idc=100,
                    transform_8x8=1, b_frames=1, num_ref_frames=2),
    "high_422": dict(width_mbs=6, height_mbs=5, frames=4, profile_idc=100, chroma_format_idc=2),
    "high_444": dict(width_mbs=6, height_mbs=5, frames=4, profile_idc=100, chroma_format_idc=3,
                     transform_8x8=1),
    "high_mono": dict(width_mbs=6, height_mbs=5, frames=4, profile_idc=100, chroma_format_idc=0),
    # round 2 (N3): MBAFF frames, field pictures, slice groups (fmo = groups | map type << 4)
    "mbaff_b": dict(width_mbs=9, height_mbs=6, frames=6, slices_per_frame=2, b_frames=1, num_ref_frames=2,
                    profile_idc=100, transform_8x8=1, pct_skip=35, idr_period=3, seed=91, fmo=0x100),
    "field_pictures": dict(width_mbs=8, height_mbs=6, frames=8, slices_per_frame=2, b_frames=1, num_ref_frames=2,
                           profile_idc=77, pct_skip=30, idr_period=4, seed=92, fmo=0x200),
    "fmo_dispersed": dict(width_mbs=9, height_mbs=7, frames=5, slices_per_frame=2, num_ref_frames=2,
                          profile_idc=66, pct_skip=25, idr_period=3, seed=93, fmo=3 | 1 << 4),
    "fmo_box_out": dict(width_mbs=9, height_mbs=7, frames=7, slices_per_frame=1, num_ref_frames=1,
                        profile_idc=66, pct_skip=25, idr_period=4, seed=94, fmo=2 | 3 << 4),
}


def main():
    for name, kw in CASES.items():
        stream, nmb, nsl = L.synth_video(**kw)
        ev, mbs, off = S.ref_trace(stream)
        params, maps = S.group_maps_from_trace(ev)  # maps: the reference's, empty without slice groups
        counts = np.array([int(p.view(np.uint32)[0]) for t, p in ev if t == S.TR_SLICE_DATA_END], np.uint32)
        order = np.array([t for t, _ in ev], np.uint8)
        np.savez_compressed(os.path.join(HERE, "cavlc_%s.npz" % name), crc=np.uint32(zlib.crc32(stream.tobytes())),
                            mbs=mbs, params=params, mb_counts=counts, callback_order=order,
                            final_off=np.uint64(off), group_maps=maps)
        print(name, len(stream), "bytes", nmb, "mbs", nsl, "slices")


def main_cabac():
    """CABAC twins: a CAVLC stream parsed by the reference, its ctx->mb syntax elements re-coded
    as CABAC (tests/test_cabac.py make_twin); the fixture holds the CABAC blob, the parameter
    blocks and the REFERENCE's records, so the parse is checked against reference output even
    where oracle/_ref is not available."""
    import test_cabac as TC
    cases = {
        "main_b": dict(width_mbs=8, height_mbs=6, frames=6, slices_per_frame=2, profile_idc=77, transform_8x8=0,
                       b_frames=1, num_ref_frames=2, idr_period=3, pct_skip=30, coef_density=40, seed=41),
        "high_t8_pcm": dict(width_mbs=7, height_mbs=5, frames=6, slices_per_frame=1, profile_idc=100,
                            transform_8x8=1, b_frames=1, num_ref_frames=3, idr_period=3, pct_skip=20,
                            coef_density=50, pct_pcm=40, seed=42),
    }
    for name, cfg in cases.items():
        blob, params, ref_mbs = TC.make_twin(cfg)
        np.savez_compressed(os.path.join(HERE, "cabac_twin_%s.npz" % name), blob=blob,
                            params=params.view(np.uint8), mbs=ref_mbs)
        print("cabac twin", name, len(blob), "bytes", len(ref_mbs), "mbs", len(params), "slices")


if __name__ == "__main__":
    main()
    main_cabac()
