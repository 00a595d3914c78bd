#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2m}
timeout 600 python -m pytest tests/test_scan_inplace.py tests/test_gpu_annexb.py -m gpu -x -q 2>&1 | tail -3
echo "== A/B"; timeout 600 python scripts/scan_ab.py --size-mb 4096 --gens 6,7 --steps 10 2>&1 | grep -E "gen._" | head -4 | tee -a gpurun_out/${TAG}_ab.log
timeout 300 python scripts/scan7_trace.py --size-mb 1024 2>&1 | tail -8 | tee gpurun_out/${TAG}_trace.txt
CMD="python scripts/scan_ab.py --size-mb 1024 --gens 7 --steps 2 --warmup 1"
timeout 300 $CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan7_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_scan7 $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_ncu.log
