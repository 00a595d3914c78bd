#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2h}
export H264GPU_SCAN7_NAP=512 H264GPU_SCAN7_PF=2
for v in 0 1 2 3; do
echo "== V $v"; H264GPU_SCAN7_V=$v timeout 600 python scripts/scan_ab.py --size-mb 4096 --gens 7 --steps 10 2>&1 | grep gen7_strip | head -1 | tee -a gpurun_out/${TAG}_sweep.log
done
H264GPU_SCAN7_V=3 timeout 300 python -m pytest tests/test_scan_inplace.py -m gpu -x -q 2>&1 | tail -2
