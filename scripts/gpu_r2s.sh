#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2s}
CMD="python scripts/bench_parse.py --cabac 0 --frames 250 --lanes 5 --reps 1"
timeout 300 $CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cavlc_parse -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_cavlc $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_plain.log; tail -1 gpurun_out/${TAG}_ncu.log
