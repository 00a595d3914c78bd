#!/bin/bash
# K5: timing of the current build at 16000 slices, then one full ncu capture of the same launch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k5c}
timeout 900 python -m pytest tests/test_cabac.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
timeout 600 python scripts/bench_parse.py --cabac 1 --frames 1000 --lanes 0,1,2 2>&1 | tee gpurun_out/${T}_parse_cabac.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:cabac_parse_kernel -s 1 -c 1 -o gpurun_out/${T}_prof_cabac -f \
  python scripts/bench_parse.py --cabac 1 --frames 1000 --lanes 0 --reps 1 > gpurun_out/${T}_ncu.log 2>&1
