#!/bin/bash
# Round-1 GPU session 1: parity tests, bench (tile-size variants), ncu launch list + full capture.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench items=4" ; timeout 900 python bench.py 2> gpurun_out/bench4.err | tee gpurun_out/bench4.json
for it in 2 1; do
  echo "== bench items=$it"
  H264GPU_SCAN_ITEMS=$it timeout 600 python bench.py --no-cpu --e2e-steps 1 2> gpurun_out/bench$it.err | tee gpurun_out/bench$it.json
done
echo "== ncu"
CMD="python bench.py --size-mb 512 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 2 -o gpurun_out/prof_scan $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
