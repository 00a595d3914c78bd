#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest inplace"; timeout 900 python -m pytest tests/test_scan_inplace.py tests/test_host_lib_gpu.py -m gpu -q 2>&1 | tail -30
echo "== scan trace gen5"; timeout 300 python scripts/scan_trace.py 1024 5 2>&1 | tail -14
SHORT="--size-mb 4096 --steps 10 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2"
for gen in 5 2; do
  echo "== bench gen $gen"
  timeout 600 python bench.py $SHORT --scan-gen $gen 2> gpurun_out/d_bench_$gen.err | tee gpurun_out/d_bench_$gen.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
  tail -2 gpurun_out/d_bench_$gen.err
done
