#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2g}
for pf in 0 1 2 4 8 16; do
echo "== pf $pf/4 waves"; H264GPU_SCAN7_NAP=512 H264GPU_SCAN7_PF=$pf timeout 600 python scripts/scan_ab.py --size-mb 4096 --gens 7 --steps 10 2>&1 | grep gen7_strip | head -1 | tee -a gpurun_out/${TAG}_sweep.log
done
