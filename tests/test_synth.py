"""The synthetic workload generator emits valid Annex-B that matches the oracle writer."""
import numpy as np

import libh264_b200 as L
import support as S


def test_synth_stream_is_oracle_framing():
    seed = 0x264
    offs = L.synth_offsets(seed, 300_000, lo=64, hi=20_000)
    rbsp = L.synth_payloads(seed, offs, nthreads=3)
    assert np.array_equal(rbsp, L.synth_payloads(seed, offs, nthreads=1))  # thread-count independent
    stream, nal_off = L.synth_annexb(seed, rbsp, offs, mixed_sc=False, trailing=False, nthreads=2)
    exp, eoo = S.oracle_frame(rbsp, offs, 4)
    assert np.array_equal(stream, exp) and np.array_equal(nal_off, eoo)
    # zero fraction ~ 3/16
    z = float((rbsp == 0).mean())
    assert 0.17 < z < 0.205


def test_synth_mixed_round_trips_through_oracle():
    seed = 11
    offs = L.synth_offsets(seed, 200_000, lo=64, hi=9000)
    rbsp = L.synth_payloads(seed, offs)
    stream, _ = L.synth_annexb(seed, rbsp, offs, mixed_sc=True, trailing=True)
    o = S.oracle_split_strip(stream)
    assert len(o["start"]) == len(offs) - 1
    assert np.array_equal(o["rbsp"], rbsp)
    assert np.array_equal(o["rbsp_off"], offs)
