#!/bin/bash
# GPU session: parity tests + bench after a kernel change (no ncu).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --no-cpu --e2e-steps 2 2> gpurun_out/bench.err | tee gpurun_out/bench_run.json
tail -5 gpurun_out/bench.err
