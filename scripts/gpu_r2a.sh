#!/bin/bash
# round 2, session A: gen-7 scan parity + A/B against gen 6 + scan-only
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt 2>&1
echo "== pytest gpu (scan)"; timeout 900 python -m pytest tests/test_scan_inplace.py tests/test_gpu_annexb.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2a_pytest.log
echo "== A/B"; timeout 600 python scripts/scan_ab.py --size-mb 4096 --gens 6,7 --steps 10 2>&1 | tail -8 | tee gpurun_out/r2a_ab.log
