#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-ab4}
export H264GPU_DEBUG_WS=1
run() { echo "== $*" | tee -a gpurun_out/${T}.log; timeout 300 python scripts/scan_ab.py --size-mb 4096 --steps 10 --order strip "$@" 2>&1 | grep "^gen7_\|^h264gpu: ws7\|^d_in\|^stage" | cut -c1-80 | tee -a gpurun_out/${T}.log; }
run --stage 1
run --stage 5
H264GPU_WS7_MIN_MB=1024 run --stage 1
run --stage 3
echo "== K4 sorted slices" | tee -a gpurun_out/${T}.log
timeout 600 python scripts/bench_parse.py --cabac 0 --frames 1000 --lanes 2,3,4,5 2>&1 | tee -a gpurun_out/${T}.log
H264GPU_CAVLC_SORT=0 timeout 600 python scripts/bench_parse.py --cabac 0 --frames 1000 --lanes 3,5 2>&1 | sed "s/^/nosort /" | tee -a gpurun_out/${T}.log
H264GPU_CAVLC_PER_LANE=2 timeout 600 python scripts/bench_parse.py --cabac 0 --frames 1000 --lanes 4,5 2>&1 | sed "s/^/per_lane=2 /" | tee -a gpurun_out/${T}.log
echo "== gpu tests"; timeout 900 python -m pytest tests/test_cavlc.py tests/test_host_lib_gpu.py tests/test_conceal.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.log
