#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-ab3}
run() { echo "== $*" | tee -a gpurun_out/${T}.log; timeout 300 python scripts/scan_ab.py --size-mb 4096 --steps 10 --order strip "$@" 2>&1 | grep "^gen7_" | cut -c1-60 | tee -a gpurun_out/${T}.log; }
run --stage 1
run --stage 2
run --stage 3
run --stage 4
run
echo "== K4 sweep (two-phase loop)" | tee -a gpurun_out/${T}.log
timeout 600 python scripts/bench_parse.py --cabac 0 --frames 1000 --lanes 2,3,4,5 2>&1 | tee -a gpurun_out/${T}.log
timeout 600 python scripts/bench_parse.py --cabac 0 --frames 300 --slices-per-frame 1 --lanes 0 2>&1 | tee -a gpurun_out/${T}.log
echo "== gpu tests"; timeout 900 python -m pytest tests/test_cavlc.py tests/test_host_lib_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.log
echo "== ncu K4 lanes=4"
H264GPU_CAVLC_LANES_LOG2=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:cavlc_steps -s 1 -c 1 -f -o gpurun_out/${T}_prof_cavlc python scripts/bench_parse.py --cabac 0 --frames 250 --lanes 4 --reps 1 > gpurun_out/${T}_ncu_cavlc.log 2>&1
tail -2 gpurun_out/${T}_ncu_cavlc.log
