"""Quick CABAC / CAVLC parse throughput probe (device-resident), several slice counts."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libh264_b200 as L

g = L.Gpu(0)
for cabac in (1, 0):
    for frames in ((250,) if cabac else (1000,)):
        cfg = dict(width_mbs=120, height_mbs=68, frames=frames, slices_per_frame=16, profile_idc=100,
                   transform_8x8=1, b_frames=1, num_ref_frames=2, idr_period=30, pct_skip=30, coef_density=60,
                   seed=7)
        t0 = time.time()
        stream, nmb, nsl, params = L.synth_video(want_params=True, entropy_cabac=cabac, **cfg)
        tg = time.time() - t0
        d_s = g.alloc(len(stream) + 16); d_s.upload(stream)
        d_p = g.alloc(len(params)); d_p.upload(params)
        d_r = g.alloc(nmb * 16 + 16); d_q = g.alloc(nsl * 16)
        fn = g.cabac_parse_dev if cabac else g.cavlc_parse_dev
        for lanes in ("auto", "0", "1", "2", "3"):
            if lanes == "auto":
                os.environ.pop("H264GPU_CABAC_LANES_LOG2", None); os.environ.pop("H264GPU_CAVLC_LANES_LOG2", None)
            else:
                os.environ["H264GPU_CABAC_LANES_LOG2"] = lanes; os.environ["H264GPU_CAVLC_LANES_LOG2"] = lanes
            fn(d_s.ptr, len(stream), d_p.ptr, nsl, d_r.ptr, d_q.ptr); g.sync()
            tm = g.timer(); g.timer_start(tm)
            for _ in range(3):
                fn(d_s.ptr, len(stream), d_p.ptr, nsl, d_r.ptr, d_q.ptr)
            g.timer_stop(tm); ms = g.timer_ms(tm) / 3
            res = np.frombuffer(d_q.download().tobytes(), L.SLICE_RESULT)
            ok = (res["status"] == 0).all() and int(res["mb_count"].sum()) == nmb
            print("cabac=%d frames=%4d slices=%6d MBs=%8d bytes=%9d lanes=%s: %8.2f ms %7.2f M MB/s ok=%s gen %.1fs" %
                  (cabac, frames, nsl, nmb, len(stream), lanes, ms, nmb / ms / 1e3, ok, tg), flush=True)
        for d in (d_s, d_p, d_r, d_q):
            d.free()
