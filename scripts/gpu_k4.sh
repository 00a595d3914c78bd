#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k4}
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.log
timeout 600 python scripts/bench_parse.py --cabac 0 --frames 1000 --lanes auto,2,3 2>&1 | tee gpurun_out/${T}_parse_cavlc.log
timeout 600 python scripts/bench_parse.py --cabac 0 --frames 300 --slices-per-frame 1 --lanes auto 2>&1 | tee -a gpurun_out/${T}_parse_cavlc.log
echo "== ncu K4 (16000 slices would take long under ncu: 4000 slices, launcher's packing)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cavlc_steps -s 2 -c 1 -f -o gpurun_out/${T}_prof_cavlc python scripts/bench_parse.py --cabac 0 --frames 250 --lanes auto --reps 1 > gpurun_out/${T}_ncu_cavlc.log 2>&1
tail -2 gpurun_out/${T}_ncu_cavlc.log
