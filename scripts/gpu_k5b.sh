#!/bin/bash
# K5 persistent grid: parity tests, then the launch knobs at 4000 and 16000 slices
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k5b}
timeout 900 python -m pytest tests/test_cabac.py tests/test_host_lib_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.log
timeout 600 python scripts/bench_parse.py --cabac 1 --frames 250 --lanes 0,1 --sweep "H264GPU_CABAC_SORT=1,0" 2>&1 | tee gpurun_out/${T}_parse_cabac.log
timeout 1200 python scripts/bench_parse.py --cabac 1 --frames 1000 --lanes 0,1 --sweep "H264GPU_CABAC_SORT=1,0;H264GPU_CABAC_PER_SM=32,24,16" 2>&1 | tee -a gpurun_out/${T}_parse_cabac.log
