#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-ab6}
export H264GPU_DEBUG_WS=1
run() { echo "== $*" | tee -a gpurun_out/${T}.log; timeout 300 python scripts/scan_ab.py --size-mb 4096 --steps 10 --order strip "$@" 2>&1 | grep "^gen7_\|^h264gpu: ws7\|^d_in\|^stage" | cut -c1-80 | tee -a gpurun_out/${T}.log; }
run --stage 6
run --stage 1
unset H264GPU_DEBUG_WS
S="--steps 10 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2 --cabac-frames 0 --cfg4-streams 0 --frame-mb 0 --no-reader"
echo "== bench"; timeout 600 python bench.py $S 2>/dev/null | cut -c1-140 | tee -a gpurun_out/${T}.log
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.log
echo "== conceal"; timeout 300 python scripts/bench_conceal.py 2>&1 | tail -2 | tee -a gpurun_out/${T}.log
