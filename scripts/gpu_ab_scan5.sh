#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-ab5}
export H264GPU_DEBUG_WS=1
run() { echo "== $*" | tee -a gpurun_out/${T}.log; timeout 300 python scripts/scan_ab.py --size-mb 4096 --steps 10 --order strip "$@" 2>&1 | grep "^gen7_\|^h264gpu: ws7\|^d_in\|^stage" | cut -c1-80 | tee -a gpurun_out/${T}.log; }
run --stage 1
run
H264GPU_WS7_PREALLOC_MB=0 run --stage 1
unset H264GPU_DEBUG_WS
S="--steps 10 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2 --cabac-frames 0 --cfg4-streams 0 --frame-mb 0 --no-reader"
echo "== bench"; timeout 600 python bench.py $S 2>/dev/null | cut -c1-140 | tee -a gpurun_out/${T}.log
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.log
echo "== K4"
timeout 600 python scripts/bench_parse.py --cabac 0 --frames 1000 --lanes auto 2>&1 | tee -a gpurun_out/${T}.log
timeout 600 python scripts/bench_parse.py --cabac 0 --frames 300 --slices-per-frame 1 --lanes auto 2>&1 | tee -a gpurun_out/${T}.log
