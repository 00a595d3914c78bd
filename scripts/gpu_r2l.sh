#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2l}
timeout 600 python -m pytest tests/test_scan_inplace.py tests/test_gpu_annexb.py -m gpu -x -q 2>&1 | tail -3
for nap in 128 512; do
echo "== nap $nap"; H264GPU_SCAN7_NAP=$nap timeout 600 python scripts/scan_ab.py --size-mb 4096 --gens 7 --steps 10 2>&1 | grep gen7_strip | head -1 | tee -a gpurun_out/${TAG}_sweep.log
done
echo "== two-phase"; H264GPU_SCAN7_TWO=1 timeout 600 python scripts/scan_ab.py --size-mb 4096 --gens 7 --steps 10 2>&1 | grep gen7_strip | head -1 | tee -a gpurun_out/${TAG}_sweep.log
timeout 300 python scripts/scan7_trace.py --size-mb 1024 2>&1 | tail -9 | tee gpurun_out/${TAG}_trace.txt
