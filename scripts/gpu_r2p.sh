#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2p}
timeout 600 python -m pytest tests/test_scan_inplace.py tests/test_gpu_annexb.py -m gpu -x -q 2>&1 | tail -3
for cnt in 1 4 16; do for k in 1472 2960 5920; do
echo "== cnt $cnt regions $k"; H264GPU_SCAN7_CNT=$cnt H264GPU_SCAN7_REGIONS=$k timeout 600 python scripts/scan_ab.py --size-mb 4096 --gens 7 --steps 10 2>&1 | grep -E "gen7_strip" | head -1 | tee -a gpurun_out/${TAG}_sweep.log
done; done
timeout 300 python scripts/scan7_trace.py --size-mb 4096 2>&1 | tail -8 | tee gpurun_out/${TAG}_trace.txt
