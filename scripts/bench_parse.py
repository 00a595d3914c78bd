"""CAVLC / CABAC slice-parse throughput probe (device-resident) vs slices per warp.
   python scripts/bench_parse.py [--cabac 0|1] [--frames N] [--lanes 0,2,3,4,5] [--slices-per-frame 16]"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libh264_b200 as L

ap = argparse.ArgumentParser()
ap.add_argument("--cabac", type=int, default=0)
ap.add_argument("--frames", type=int, default=1000)
ap.add_argument("--lanes", default="auto,0,2,3,4,5")
ap.add_argument("--slices-per-frame", type=int, default=16)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--copies", type=int, default=1, help="lay the generated stream end to end this many times")
ap.add_argument("--sweep", default="", help="ENV=v1,v2;ENV2=w1,w2 : cartesian product of launch knobs")
args = ap.parse_args()
g = L.Gpu(0)
cfg = dict(width_mbs=120, height_mbs=68, frames=args.frames, slices_per_frame=args.slices_per_frame, profile_idc=100,
           transform_8x8=1, b_frames=1, num_ref_frames=2, idr_period=30, pct_skip=30, coef_density=60, seed=7)
t0 = time.time()
stream, nmb, nsl, params = L.synth_video(want_params=True, entropy_cabac=args.cabac, **cfg)
tg = time.time() - t0
if args.copies > 1:  # as bench.py's replicate(): separate memory, separate slices
    P = np.frombuffer(params, L.SLICE_PARAMS)
    one = np.concatenate([stream, np.zeros((-len(stream)) % 16, np.uint8)])
    allp = []
    for k in range(args.copies):
        q = P.copy(); q["nal_off"] += k * len(one); q["mb_out_off"] += k * nmb; allp.append(q)
    stream, params = np.tile(one, args.copies), np.concatenate(allp).view(np.uint8)
    nmb, nsl = nmb * args.copies, nsl * args.copies
d_s = g.alloc(len(stream) + 16); d_s.upload(stream)
d_p = g.alloc(len(params)); d_p.upload(params)
d_r = g.alloc(nmb * 16 + 16); d_q = g.alloc(nsl * 16)
fn = g.cabac_parse_dev if args.cabac else g.cavlc_parse_dev
import itertools
sweep = [kv.split("=") for kv in args.sweep.split(";") if kv]
combos = list(itertools.product(*[[(k, v) for v in vs.split(",")] for k, vs in sweep])) or [()]
for combo, lanes in itertools.product(combos, args.lanes.split(",")):
    for k, v in combo:
        os.environ[k] = v
    for k in ("H264GPU_CABAC_LANES_LOG2", "H264GPU_CAVLC_LANES_LOG2"):
        os.environ.pop(k, None)
        if lanes != "auto":
            os.environ[k] = lanes
    fn(d_s.ptr, len(stream), d_p.ptr, nsl, d_r.ptr, d_q.ptr); g.sync()
    tm = g.timer(); g.timer_start(tm)
    for _ in range(args.reps):
        fn(d_s.ptr, len(stream), d_p.ptr, nsl, d_r.ptr, d_q.ptr)
    g.timer_stop(tm); ms = g.timer_ms(tm) / args.reps
    res = np.frombuffer(d_q.download().tobytes(), L.SLICE_RESULT)
    ok = (res["status"] == 0).all() and int(res["mb_count"].sum()) == nmb
    print("cabac=%d frames=%4d slices=%6d MBs=%8d bytes=%9d lanes=%s %s: %8.2f ms %7.2f M MB/s ok=%s gen %.1fs" %
          (args.cabac, args.frames, nsl, nmb, len(stream), lanes, " ".join("%s=%s" % kv for kv in combo), ms,
           nmb / ms / 1e3, ok, tg), flush=True)
