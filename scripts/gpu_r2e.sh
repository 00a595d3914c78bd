#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2e}
for rows in 8 6; do for nap in 256 512 1024 2048; do
echo "== rows $rows nap $nap"; H264GPU_SCAN7_ROWS=$rows H264GPU_SCAN7_NAP=$nap timeout 600 python scripts/scan_ab.py --size-mb 4096 --gens 7 --steps 10 2>&1 | grep gen7_strip | head -1 | tee -a gpurun_out/${TAG}_sweep.log
done; done
H264GPU_SCAN7_ROWS=6 timeout 300 python -m pytest tests/test_scan_inplace.py -m gpu -x -q 2>&1 | tail -2
