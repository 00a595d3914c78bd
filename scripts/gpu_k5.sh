#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k5}
for w in 1 2 4; do
H264GPU_CABAC_WARPS=$w timeout 900 python scripts/bench_parse.py --cabac 1 --frames 1000 --lanes 0,1,2 2>&1 | sed "s/^/warps=$w /" | tee -a gpurun_out/${T}_parse_cabac.log
done
