#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest inplace"; timeout 900 python -m pytest tests/test_scan_inplace.py -m gpu -q 2>&1 | tail -3
SHORT="--size-mb 4096 --steps 10 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2 --cabac-frames 0 --frame-mb 0"
for shape in 85 84; do
echo "== bench gen 5 shape $shape"
H264GPU_SCAN5_SHAPE=$shape timeout 600 python bench.py $SHORT --scan-gen 5 2> gpurun_out/e_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
echo "== trace"; timeout 300 python scripts/scan_trace.py 1024 5 2>&1 | tail -12
