#!/bin/bash
# Phase trace + ncu full capture of the gen-6 scan kernel.  Output in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-p6}
echo "== phase trace gen 6"; timeout 300 python scripts/scan_trace.py 1024 ${GEN:-6} 2>&1 | tee gpurun_out/${TAG}_trace6.txt
SHORT="--size-mb 1024 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2 --cabac-frames 0 --frame-mb 0"
timeout 300 python bench.py $SHORT --scan-gen ${GEN:-6} > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan6p?_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_scan6 python bench.py $SHORT --scan-gen ${GEN:-6} > gpurun_out/${TAG}_ncu_scan6.log 2>&1
ls -la gpurun_out/${TAG}_*
