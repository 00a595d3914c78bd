#!/bin/bash
# GPU session 4: second-generation scan kernel: parity, bench (gen2 vs gen1), ncu.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest gpu" ; H264GPU_SCAN_ITEMS=408 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
SHORT="--size-mb 2048 --steps 10 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2"
for it in 408 418 108 118; do
  echo "== bench scan items=$it"
  H264GPU_SCAN_ITEMS=$it timeout 600 python bench.py $SHORT 2> gpurun_out/bench_$it.err | tee gpurun_out/bench_$it.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
  tail -2 gpurun_out/bench_$it.err
done
echo "== ncu scan2"
CMD="env H264GPU_SCAN_ITEMS=408 python bench.py --size-mb 512 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan4 -s 3 -c 1 -o gpurun_out/prof_scan4 $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
