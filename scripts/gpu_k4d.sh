#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k4d}
for b in 1 4 2 8; do
H264GPU_CAVLC_REPEAT=$b timeout 600 python scripts/bench_parse.py --cabac 0 --frames 1000 --lanes auto,3 2>&1 | sed "s/^/repeat=$b /" | tee -a gpurun_out/${T}_parse_cavlc.log
done
for b in 1 4 8; do
H264GPU_CAVLC_REPEAT=$b timeout 600 python scripts/bench_parse.py --cabac 0 --frames 300 --slices-per-frame 1 --lanes auto 2>&1 | sed "s/^/repeat=$b /" | tee -a gpurun_out/${T}_parse_cavlc.log
H264GPU_CAVLC_REPEAT=$b timeout 600 python scripts/bench_parse.py --cabac 0 --frames 250 --lanes auto 2>&1 | sed "s/^/repeat=$b /" | tee -a gpurun_out/${T}_parse_cavlc.log
done
echo "== gpu tests"; timeout 1500 python -m pytest tests/test_cavlc.py tests/test_host_lib_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.log
