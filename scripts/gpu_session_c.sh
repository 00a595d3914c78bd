#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest host gpu"; timeout 600 python -m pytest tests/test_host_lib_gpu.py -m gpu -q 2>&1 | tail -15
echo "== scan trace"; timeout 300 python scripts/scan_trace.py 1024 2>&1 | tail -12
