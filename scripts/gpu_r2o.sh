#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2o}
CMD="python scripts/scan_ab.py --size-mb 2048 --gens 7 --steps 2 --warmup 1"
timeout 300 $CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan7_kernel -s 2 -c 2 -f -o gpurun_out/${TAG}_prof_scan7 $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_ncu.log; grep gen7 gpurun_out/${TAG}_plain.log | head -2
timeout 300 python scripts/scan7_trace.py --size-mb 2048 2>&1 | tail -8 | tee gpurun_out/${TAG}_trace.txt
