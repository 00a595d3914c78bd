#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2q}
timeout 600 python -m pytest tests/test_scan_inplace.py tests/test_gpu_annexb.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 600 python scripts/scan_ab.py --size-mb 4096 --gens 7 --steps 20 2>&1 | grep -E "gen7_" | head -2 | tee -a gpurun_out/${TAG}_ab.log
done
CMD="python scripts/scan_ab.py --size-mb 4096 --gens 7 --steps 2 --warmup 2"
timeout 300 $CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan7_kernel -s 4 -c 2 -f -o gpurun_out/${TAG}_prof_scan7 $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_ncu.log
