#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2d}
echo "== pytest gpu (scan)"; timeout 900 python -m pytest tests/test_scan_inplace.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
echo "== A/B"; timeout 600 python scripts/scan_ab.py --size-mb 4096 --gens 7 --steps 10 2>&1 | tail -3 | tee gpurun_out/${TAG}_ab.log
echo "== trace"; timeout 300 python scripts/scan7_trace.py --size-mb 1024 2>&1 | tail -12 | tee gpurun_out/${TAG}_trace.txt
