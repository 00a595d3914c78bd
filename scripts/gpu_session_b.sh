#!/bin/bash
# GPU session: full gpu test suite with details, CABAC micro-bench
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-b}
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== cabac bench"; timeout 600 python scripts/bench_cabac.py 2>&1 | tail -20 | tee gpurun_out/${TAG}_cabac.log
