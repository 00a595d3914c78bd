#!/bin/bash
# K5 with the register-resident residual decoder: parity, then knobs at 16000 slices (250 frames x 4)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k5d}
timeout 900 python -m pytest tests/test_cabac.py tests/test_host_lib_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
timeout 600 python scripts/bench_parse.py --cabac 1 --frames 250 --copies 4 --lanes 0,1,2,3 2>&1 | tee gpurun_out/${T}_parse_cabac.log
timeout 600 python scripts/bench_parse.py --cabac 1 --frames 250 --copies 1 --lanes 0,1 2>&1 | tee -a gpurun_out/${T}_parse_cabac.log
