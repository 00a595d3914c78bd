#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2r}
timeout 600 python -m pytest tests/test_cavlc.py tests/test_host_lib_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python scripts/bench_parse.py --cabac 0 --frames 1000 --lanes auto,5,4,3,2,0 2>&1 | tail -7 | tee gpurun_out/${TAG}_cavlc.log
