#!/bin/bash
# GPU session 3: parity tests (scan/strip/frame + CAVLC), full bench, ncu of the CAVLC kernel.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench" ; timeout 1500 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench_run.json
tail -5 gpurun_out/bench.err
echo "== ncu cavlc"
CMD="python bench.py --size-mb 256 --steps 4 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 60"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches3.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cavlc -c 1 -o gpurun_out/prof_cavlc $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
