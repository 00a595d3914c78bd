#!/bin/bash
# span-ticket counters of scan7_kernel (H264GPU_SCAN7_TICKS) in the fast (--stage 6) and slow (--stage 1) workspace placement
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-tk}
O=gpurun_out/${TAG}_ticks.log
: > $O
echo "== pytest gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_gpu.log
A="--order strip --steps 10"
run() { st=$1; shift; echo "== stage $st $*" >> $O; env "$@" timeout 100 python scripts/scan_ab.py $A --stage $st 2>&1 | grep -E "gen7_strip \{" | cut -c1-100 >> $O; }
for t in 1 2 4 8 15; do run 1 H264GPU_SCAN7_TICKS=$t; done
for t in 1 4 8 15; do run 6 H264GPU_SCAN7_TICKS=$t; done
cat $O
