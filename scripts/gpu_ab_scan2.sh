#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-ab2}
run() { echo "== $*"; timeout 300 python scripts/scan_ab.py --size-mb 4096 --steps 10 "$@" 2>&1 | grep "^gen7_" | cut -c1-60 | tee -a gpurun_out/${T}.log; }
run
run --cap-extra 4096
run --alloc-extra 64
run --stage 1
run --sampler 1
run --order only,strip
run --cap-extra 4096 --alloc-extra 64 --stage 1 --sampler 1
S="--steps 10 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2 --cabac-frames 0 --cfg4-streams 0 --frame-mb 0 --no-reader"
echo "== bench"; timeout 600 python bench.py $S 2>/dev/null | cut -c1-140 | tee -a gpurun_out/${T}.log
echo "== bench no sampler"; H264BENCH_NO_SAMPLER=1 timeout 600 python bench.py $S 2>/dev/null | cut -c1-140 | tee -a gpurun_out/${T}.log
echo "== gpu tests"; timeout 900 python -m pytest tests/test_cavlc.py tests/test_host_lib_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.log
